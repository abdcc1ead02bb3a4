"""EfficientMCAttModel.forward on the GPU (through the C ABI) against
 (a) the golden vectors produced by the unmodified reference, and
 (b) the CPU oracle on larger, ragged batches at the published width.
fp32 mode tolerance: 1e-4 relative (north star); bf16 mode reports its own, looser bound."""
import json
import os

import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import ref_shims
from oracle.det_weights import det_state_dict
from fabind_b200 import EfficientMCAttModel
from fabind_b200.synthetic import make_batch
from helpers import golden_files, load_golden, rel_err

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _model(hidden, L, IT, sd, precision="fp32"):
    m = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    return m


def _run(m, b):
    bc = b.to("cuda")
    X, H = m(**bc.forward_args())
    torch.cuda.synchronize()
    return X.cpu(), H.cpu()


def _log(name, rec):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=name, **rec)) + "\n")


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_golden_fp32(path):
    g, r, b, sd, cfg = load_golden(path)
    m = _model(r["hidden"], r["n_layers"], r["n_iter"], sd)
    m.debug_trace = True
    X, H = _run(m, b)
    st = m.last_stats
    e_int = st["inter_edges_per_iter"].cpu().tolist()
    rec = dict(case=os.path.basename(path), e_int=e_int, e_int_ref=[int(e[1].shape[1]) for e in g["edges"]],
               x_err=rel_err(X, g["X"]), h_err=rel_err(H, g["H"]))
    th, tx = st["trace"]
    for k, (tag, h_ref, x_ref) in enumerate(g["trace_last_iter"]):
        rec[f"{tag}_h"] = rel_err(th[k].cpu(), h_ref)
        rec[f"{tag}_x"] = rel_err(tx[k].cpu(), x_ref.squeeze(1))
    _log("golden_fp32", rec)
    nb = b.compound_edge_index.shape[1]
    assert st["ctx_edges"] == int(g["edges"][0][0].shape[1]) + nb
    assert e_int == rec["e_int_ref"]
    assert rec["x_err"] < 1e-4 and rec["h_err"] < 1e-4, rec


@pytest.mark.parametrize("hidden,L,IT,bkw", [
    (512, 1, 1, dict(n_complexes=1, seed=0, n_c=30, n_p=200)),                                  # BASELINE config 1
    (512, 4, 8, dict(n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200))),          # published depth
    (128, 1, 1, dict(n_complexes=3, seed=6, n_c_range=(10, 40), n_p_range=(150, 600))),         # pocket-stage shape
])
def test_oracle_fp32(hidden, L, IT, bkw):
    b = make_batch(embed=hidden, **bkw)
    m0 = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                             normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m0.state_dict().items()}, 31)
    cfg = orc.make_cfg(n_layers=L, n_iter=IT)
    with torch.no_grad():
        Xo, Ho, edges = orc.model_forward(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
                                          b.compound_edge_index, b.LAS_edge_index, b.X_LAS, return_edges=True)
    m = _model(hidden, L, IT, sd)
    X, H = _run(m, b)
    e_int = m.last_stats["inter_edges_per_iter"].cpu().tolist()
    rec = dict(hidden=hidden, L=L, IT=IT, x_err=rel_err(X, Xo), h_err=rel_err(H, Ho), e_int=e_int,
               e_int_ref=[int(e[1].shape[1]) for e in edges], moved=float((Xo - b.X).abs().max()))
    _log("oracle_fp32", rec)
    assert e_int == rec["e_int_ref"], rec
    assert rec["x_err"] < 1e-4 and rec["h_err"] < 1e-4, rec


def _edge_drift(m, e_ref):
    """interface-edge counts per refinement iteration against the oracle's (coordinate drift moves atoms across the cutoff)"""
    e = m.last_stats["inter_edges_per_iter"].cpu().tolist()
    return e, [abs(a - b) for a, b in zip(e, e_ref)]


def test_tensor_core_parity_at_the_benched_shape():
    """The BENCHED configuration (BASELINE configs[1]: B = 16, n_c = 30, n_p = 200, hidden 512, 4 layers x 8 iterations) against the
    CPU oracle, in every precision mode:
      fp32_tc (tcgen05, six bf16 products per term)  <= 1e-4 on X and H  -- the tensor-core kernel IS a parity kernel;
      fp32    (FFMA)                                 <= 1e-4;
      bf16x3 / bf16: reported (own, looser bounds), with the interface-edge drift per iteration."""
    hidden, L, IT = 512, 4, 8
    b = make_batch(embed=hidden, n_complexes=16, seed=100, n_c=30, n_p=200)
    m0 = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                             normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m0.state_dict().items()}, 31)
    cfg = orc.make_cfg(n_layers=L, n_iter=IT)
    with torch.no_grad():
        Xo, Ho, edges = orc.model_forward(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
                                          b.compound_edge_index, b.LAS_edge_index, b.X_LAS, return_edges=True)
    e_ref = [int(e[1].shape[1]) for e in edges]
    recs = {}
    for prec in ("fp32_tc", "fp32", "bf16x3", "bf16"):
        m = _model(hidden, L, IT, sd, precision=prec)
        X, H = _run(m, b)
        e, drift = _edge_drift(m, e_ref)
        recs[prec] = dict(precision=prec, x_err=rel_err(X, Xo), h_err=rel_err(H, Ho), x_abs=float((X - Xo).abs().max()), e_int=e, e_int_ref=e_ref,
                          edge_drift=drift)
        _log("benched_shape", recs[prec])
    for prec in ("fp32_tc", "fp32"):
        assert recs[prec]["x_err"] < 1e-4 and recs[prec]["h_err"] < 1e-4, recs[prec]
        assert recs[prec]["e_int"] == e_ref, recs[prec]
    assert recs["bf16x3"]["x_err"] < 2e-3 and recs["bf16x3"]["h_err"] < 2e-3, recs["bf16x3"]
    assert recs["bf16"]["x_abs"] < 0.15 and recs["bf16"]["h_err"] < 0.02, recs["bf16"]


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_golden_fp32_tc(path):
    """the tensor-core parity mode against the goldens of the unmodified reference (shapes that do not tile fall to the FFMA kernel)"""
    g, r, b, sd, cfg = load_golden(path)
    m = _model(r["hidden"], r["n_layers"], r["n_iter"], sd, precision="fp32_tc")
    X, H = _run(m, b)
    e_int = m.last_stats["inter_edges_per_iter"].cpu().tolist()
    rec = dict(case=os.path.basename(path), x_err=rel_err(X, g["X"]), h_err=rel_err(H, g["H"]))
    _log("golden_fp32_tc", rec)
    assert e_int == [int(e[1].shape[1]) for e in g["edges"]]
    assert rec["x_err"] < 1e-4 and rec["h_err"] < 1e-4, rec


def test_tcgen05_attention_in_the_stack():
    """bf16 mode with the tcgen05 attention core (module.attention = "tcgen05") against bf16 mode with the SIMT core: same inputs to
    the attention (bf16 projections vs fp32 projections of bf16 operands), so the two forwards agree to bf16 rounding of Q / K / V / P"""
    hidden, L, IT = 512, 2, 2
    b = make_batch(embed=hidden, n_complexes=4, seed=9, n_c_range=(10, 50), n_p_range=(80, 200))
    m0 = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                             normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m0.state_dict().items()}, 31)
    m = _model(hidden, L, IT, sd, precision="bf16")
    Xs, Hs = _run(m, b)
    m.attention = "tcgen05"
    Xt, Ht = _run(m, b)
    rec = dict(x_err=rel_err(Xt, Xs), h_err=rel_err(Ht, Hs))
    _log("tcgen05_attention", rec)
    assert rec["x_err"] < 2e-2 and rec["h_err"] < 2e-2, rec
    assert rec["h_err"] > 0.0          # the other kernel did run


def test_bf16_mode_deviation():
    """bf16 production mode: same path with bf16 GEMM operands.  The reference has no bf16 mode; this bound
    is the build's own: coordinates within 0.15 normalised units (0.75 A) of the fp32 oracle after 8 iterations x 4
    layers with the deliberately large O(1) test coordinate heads (trained heads are ~1000x smaller, egnn.py:52),
    node features within 2% of their scale."""
    hidden, L, IT = 512, 4, 8
    b = make_batch(embed=hidden, n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200))
    m0 = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                             normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m0.state_dict().items()}, 31)
    cfg = orc.make_cfg(n_layers=L, n_iter=IT)
    with torch.no_grad():
        Xo, Ho = orc.model_forward(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
                                   b.compound_edge_index, b.LAS_edge_index, b.X_LAS)
    m = _model(hidden, L, IT, sd, precision="bf16")
    X, H = _run(m, b)
    rec = dict(x_abs=float((X - Xo).abs().max()), x_err=rel_err(X, Xo), h_err=rel_err(H, Ho))
    _log("bf16_mode", rec)
    assert rec["x_abs"] < 0.15 and rec["h_err"] < 0.02, rec


def test_missing_library_fails_loudly(monkeypatch):
    from fabind_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfabind_b200.so")
    with pytest.raises(RuntimeError):
        _lib.lib()


def test_sharded_forward_real_module_single_rank():
    """fabind_b200.shard (the N>1 host logic, gloo-tested on CPU) driving the CUDA module: taking complexes {0,2} out
    of a ragged batch gives those complexes' rows of the full-batch result (complexes are independent)."""
    from fabind_b200 import shard
    hidden, L, IT = 64, 2, 2
    b = make_batch(embed=hidden, n_complexes=3, seed=8, n_c_range=(8, 20), n_p_range=(30, 60))
    m0 = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                             normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m0.state_dict().items()}, 5)
    m = _model(hidden, L, IT, sd)
    fa = b.to("cuda").forward_args()
    sub, idx = shard.take_complexes({k: (v.clone() if torch.is_tensor(v) else v) for k, v in fa.items()}, [0, 2])
    Xs, Hs = m(**sub)
    Xf, Hf = shard.sharded_forward(m, fa)
    assert rel_err(Xs, Xf[idx]) < 1e-5 and rel_err(Hs, Hf[idx]) < 1e-5


def test_derived_weight_slots_match_float64_products():
    """fb_derive_weights: every folded slot [W_b | W_b W_a] / bias W_b b_a + b_b against float64 products of the base slots"""
    from fabind_b200.weights import slots, derive_on_device
    from fabind_b200 import _lib
    H, L = 128, 2
    l = _lib.lib()
    n = l.fb_weight_arena_elems_f(H, L, 0)
    g = torch.Generator().manual_seed(5)
    arena = (torch.randn(n, generator=g) * 0.1).cuda()
    tab = {name: (r, c, off) for name, r, c, off in slots(H, L, 0, derived=True)}
    base = {name for name, *_ in slots(H, L, 0)}
    assert any(k.rpartition(".")[2].startswith("f_") for k in tab) and not any(k.rpartition(".")[2].startswith("f_") for k in base)
    before = arena.clone()
    derive_on_device(arena, H, L, 0)
    torch.cuda.synchronize()
    for name in base:                                   # base slots are not touched
        r, c, off = tab[name]
        assert torch.equal(arena[off:off + r * c], before[off:off + r * c]), name
    A = arena.double().cpu()
    get = lambda name: A[tab[name][2]:tab[name][2] + tab[name][0] * tab[name][1]].view(tab[name][0], tab[name][1])
    HD = 128
    for pre in [f"gcl{i}." for i in range(L)] + ["out."]:          # [e1_b | 0]: bias of the stacked per-node projection GEMM
        got = get(pre + "f_e1b")[0]
        assert torch.equal(got[:H], get(pre + "e1_b")[0]) and float(got[H:].abs().max()) == 0.0
    for i in range(L):
        a, gcl = f"att{i}.", f"gcl{i}."
        checks = [
            (a + "f_cac_w", a + "f_cac_b", get(a + "ca_c_w"), get(a + "ca_c_b")[0], get(gcl + "n2_w"), get(gcl + "n2_b")[0], 0),
            (a + "f_cap_w", a + "f_cap_b", get(a + "ca_p_w"), get(a + "ca_p_b")[0], get(gcl + "n2_w"), get(gcl + "n2_b")[0], 0),
            (a + "f_l3_w", a + "f_l3_b", get(a + "ca_p2_w"), None, get(a + "o_p_w"), get(a + "o_p_b")[0], 0),
            (a + "f_l3_w", a + "f_l3_b", get(a + "tp1_w"), get(a + "tp1_b")[0], get(a + "o_p_w"), get(a + "o_p_b")[0], 2 * HD),
            (a + "f_l5_w", a + "f_l5_b", get(a + "tc1_w"), get(a + "tc1_b")[0], get(a + "o_c_w"), get(a + "o_c_b")[0], 0),
            (a + "f_qkc_w", a + "f_qkc_b", get(a + "qk_w"), get(a + "qk_b")[0], get(a + "tc2_w"), get(a + "tc2_b")[0], 0),
        ]
        for wn, bn, Wb, bb, Wa, ba, row0 in checks:
            R, K = Wb.shape
            want_w = torch.cat([Wb, Wb @ Wa], 1)
            want_b = Wb @ ba + (bb if bb is not None else 0)
            got_w, got_b = get(wn)[row0:row0 + R], get(bn)[0][row0:row0 + R]
            assert float((got_w - want_w).abs().max()) <= 1e-6 * float(want_w.abs().max()) + 1e-9, wn
            assert float((got_b - want_b).abs().max()) <= 1e-6 * float(want_b.abs().max()) + 1e-9, bn


@pytest.mark.parametrize("path", [p for p in golden_files() if "_it" in p and "_it1" not in p][:3], ids=lambda p: p.split("/")[-1][:-3])
def test_moving_rows_subset_is_exact(path):
    """out_layer of the non-final iterations on the context edges INTO the masked rows only (fb_model_params.n_mv) == on all edges:
    identical X and H in fp32 mode (the FFMA kernel computes every row independently of the others)"""
    g, r, b, sd, cfg = load_golden(path)
    m = _model(r["hidden"], r["n_layers"], r["n_iter"], sd, precision="fp32")
    X1, H1 = _run(m, b)
    m.moving_rows = False
    X0, H0 = _run(m, b)
    assert torch.equal(X1, X0) and torch.equal(H1, H0)
