"""FABind+ weight layout (LayerNorm MLPs, propagated pair embedding) on the GPU through the C ABI against
 (a) the golden vectors produced by the unmodified FABind+ reference (X, H, pair embedding, per-sub-layer trace), and
 (b) the CPU oracle at the published width / depth (hidden 512, 5 layers).
fp32 mode tolerance: 1e-4 relative (north star); bf16 mode has its own, looser bound."""
import json
import os

import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import fabind_plus_oracle as orcp
from oracle import ref_shims
from oracle.det_weights import det_state_dict
from fabind_b200.plus import EfficientMCAttModel
from fabind_b200.synthetic import make_batch
from helpers import plus_golden_files, load_golden, rel_err

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _model(hidden, L, IT, sd=None, precision="fp32", seed=33):
    m = EfficientMCAttModel(ref_shims.published_args_plus(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    if sd is None:
        sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    return m, sd


def _run(m, b):
    X, H, pair = m(**b.to("cuda").forward_args())
    torch.cuda.synchronize()
    return X.cpu(), H.cpu(), pair.cpu()


def _log(name, rec):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=name, **rec)) + "\n")


@pytest.mark.parametrize("path", plus_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_plus_golden_fp32(path):
    g, r, b, sd, cfg = load_golden(path)
    m, _ = _model(r["hidden"], r["n_layers"], r["n_iter"], sd)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(g["shapes"])   # drop-in state_dict layout
    m.debug_trace = True
    X, H, pair = _run(m, b)
    st = m.last_stats
    e_int = st["inter_edges_per_iter"].cpu().tolist()
    rec = dict(case=os.path.basename(path), e_int=e_int, e_int_ref=[int(e[1].shape[1]) for e in g["edges"]],
               x_err=rel_err(X, g["X"]), h_err=rel_err(H, g["H"]), pair_err=rel_err(pair, g["pair"]))
    th, tx = st["trace"]
    for k, (tag, h_ref, x_ref) in enumerate(g["trace_last_iter"]):
        rec[f"{tag}_h"] = rel_err(th[k].cpu(), h_ref)
        rec[f"{tag}_x"] = rel_err(tx[k].cpu(), x_ref.squeeze(1))
    _log("plus_golden_fp32", rec)
    assert e_int == rec["e_int_ref"]
    assert pair.shape == g["pair"].shape
    assert rec["x_err"] < 1e-4 and rec["h_err"] < 1e-4 and rec["pair_err"] < 1e-4, rec


@pytest.mark.parametrize("hidden,L,IT,bkw", [
    (512, 1, 1, dict(n_complexes=1, seed=0, n_c=30, n_p=200)),
    (512, 5, 3, dict(n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200))),     # published depth (5 layers)
])
def test_plus_oracle_fp32(hidden, L, IT, bkw):
    b = make_batch(embed=hidden, **bkw)
    m, sd = _model(hidden, L, IT)
    with torch.no_grad():
        Xo, Ho, Po, edges = orcp.model_forward(sd, orc.make_cfg(n_layers=L, n_iter=IT), b.X, b.H, b.batch_id, b.segment_id,
                                               b.mask, b.is_global, b.compound_edge_index, b.LAS_edge_index, b.X_LAS,
                                               return_edges=True)
    X, H, pair = _run(m, b)
    e_int = m.last_stats["inter_edges_per_iter"].cpu().tolist()
    rec = dict(hidden=hidden, L=L, IT=IT, x_err=rel_err(X, Xo), h_err=rel_err(H, Ho), pair_err=rel_err(pair, Po),
               e_int=e_int, e_int_ref=[int(e[1].shape[1]) for e in edges], moved=float((Xo - b.X).abs().max()))
    _log("plus_oracle_fp32", rec)
    assert e_int == rec["e_int_ref"], rec
    assert rec["x_err"] < 1e-4 and rec["h_err"] < 1e-4 and rec["pair_err"] < 1e-4, rec


@pytest.mark.parametrize("IT,big_heads", [(1, True), (3, False)])
def test_plus_bf16_mode_deviation(IT, big_heads):
    """bf16 production mode of the FABind+ layout (tcgen05 GEMMs with bf16 operands, LayerNorm statistics / softmax /
    coordinates / residual stream in fp32).  The reference has no bf16 mode; the bound is this build's own.
    Case 1: one iteration with the deliberately O(1) test coordinate heads (every layer moves atoms by up to the clamp):
    per-iteration error.  Case 2: three iterations with the reference's own head initialisation (xavier gain 0.001,
    P/models/egnn.py:41,135) - with O(1) heads the refinement loop amplifies a 1 % coordinate deviation chaotically
    (different interface edges in the next iteration), which says nothing about the arithmetic."""
    hidden, L = 512, 5
    b = make_batch(embed=hidden, n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200))
    m, sd = _model(hidden, L, IT, precision="bf16")
    if not big_heads:
        g = torch.Generator().manual_seed(7)
        sd = dict(sd)
        for k in sd:
            if k.endswith("coord_mlp.linear2.weight"):
                bound = 0.001 * (6.0 / (sd[k].shape[0] + sd[k].shape[1])) ** 0.5
                sd[k] = (torch.rand(sd[k].shape, generator=g) * 2 - 1) * bound
        m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        Xo, Ho, Po = orcp.model_forward(sd, orc.make_cfg(n_layers=L, n_iter=IT), b.X, b.H, b.batch_id, b.segment_id, b.mask,
                                        b.is_global, b.compound_edge_index, b.LAS_edge_index, b.X_LAS)
    X, H, pair = _run(m, b)
    # a node whose interface edge sits within the coordinate deviation of the cutoff gains / loses that edge (a discrete
    # effect, observed on 1 node of 346): the max-norm bound is applied to all but the worst 1 % of the rows, the mean
    # deviation to all of them
    row_err = (H - Ho).abs().max(1).values / Ho.abs().max()
    k = max(1, int(0.01 * row_err.numel()))
    rec = dict(IT=IT, big_heads=big_heads, x_abs=float((X - Xo).abs().max()), x_err=rel_err(X, Xo), h_err=rel_err(H, Ho),
               h_err_99=float(row_err.sort().values[-k - 1]), h_mean=float((H - Ho).abs().mean() / Ho.pow(2).mean().sqrt()),
               pair_err=rel_err(pair, Po), pair_mean=float((pair - Po).abs().mean() / Po.pow(2).mean().sqrt()))
    _log("plus_bf16_mode", rec)
    assert rec["x_abs"] < 0.15 and rec["h_err_99"] < 0.03 and rec["h_mean"] < 0.01 and rec["pair_mean"] < 0.01, rec


def test_plus_training_with_autograd_raises():
    """train() with autograd enabled = training, which is not built (train() under no_grad is the sampling mode, see
    test_gpu_plus_sampling.py); the weight containers of the attention block raise when called on their own"""
    m, _ = _model(64, 1, 1)
    with pytest.raises(NotImplementedError):
        m.gnn.att_0.cross_attn_module(None)
    m.train()
    b = make_batch(embed=64, n_complexes=1, seed=0, n_c=5, n_p=10).to("cuda")
    with pytest.raises(NotImplementedError):
        m(**b.forward_args())


def test_plus_moving_rows_subset_is_exact():
    """FABind+ layout: out_layer of the non-final iterations on the context edges INTO the masked rows only == on all edges (fp32 mode:
    identical X, H and pair embedding)"""
    b = make_batch(embed=64, n_complexes=3, seed=21, n_c=14, n_p=60)
    m, _ = _model(64, 2, 3)
    X1, H1, P1 = _run(m, b)
    m.moving_rows = False
    X0, H0, P0 = _run(m, b)
    assert torch.equal(X1, X0) and torch.equal(H1, H0) and torch.equal(P1, P0)
