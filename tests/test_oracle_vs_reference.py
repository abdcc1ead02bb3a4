"""Tensor-for-tensor check of the oracle against the LIVE reference (dev container only: skipped
wherever /root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_shims, fabind_oracle as orc
from oracle.det_weights import det_state_dict
from fabind_b200.synthetic import make_batch
from helpers import rel_err

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree absent")


def _build(hidden, L, IT, wseed):
    mods = ref_shims.load_reference("v1")
    args = ref_shims.published_args()
    m = mods.att_model.EfficientMCAttModel(
        args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, n_iter=IT, inter_cutoff=10, intra_cutoff=8,
        normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).eval()
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, wseed)
    m.load_state_dict(sd, strict=True)
    return mods, m, sd


@pytest.mark.parametrize("hidden,L,IT,bkw", [
    (48, 2, 2, dict(n_complexes=4, seed=7, n_c_range=(5, 25), n_p_range=(30, 70))),
    (96, 1, 3, dict(n_complexes=2, seed=8, n_c=20, n_p=100)),
])
def test_full_forward(hidden, L, IT, bkw):
    mods, m, sd = _build(hidden, L, IT, 21)
    b = make_batch(embed=hidden, **bkw)
    cfg = orc.make_cfg(n_layers=L, n_iter=IT)
    with torch.no_grad():
        Xr, Hr = m(**b.clone().forward_args())
        Xo, Ho = orc.model_forward(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
                                   b.compound_edge_index, b.LAS_edge_index, b.X_LAS)
    assert rel_err(Xo, Xr) < 2e-6
    assert rel_err(Ho, Hr) < 2e-5


def test_edges_bit_exact_near_cutoff():
    """Adversarial: residues placed within a few ulp of the 8 A / 10 A cutoffs."""
    mods, m, sd = _build(32, 1, 1, 22)
    b = make_batch(n_complexes=2, seed=9, n_c=12, n_p=60, embed=32)
    X = b.X.clone()
    # snap some protein-protein and ligand-protein distances right onto the cutoffs
    g = torch.Generator().manual_seed(0)
    for k in range(40):
        i = int(torch.randint(14, 74, (1,), generator=g))
        j = int(torch.randint(1, 74, (1,), generator=g))
        if i == j or j in (0, 13):
            continue
        cut = (8.0 if j > 13 else 10.0) / 5.0
        d = X[i, 0] - X[j, 0]
        n = d.norm()
        if n < 1e-3:
            continue
        eps = (int(torch.randint(-3, 4, (1,), generator=g))) * 1.2e-7
        X[i, 0] = X[j, 0] + d / n * (cut * (1 + eps))
    with torch.no_grad():
        ctx, inter, red = m.extract_edges(X, b.batch_id, b.segment_id, b.is_global)
        c2, i2, r2 = orc.build_edges(X, b.batch_id, b.segment_id, b.is_global, 8 / 5.0, 10 / 5.0)
    assert torch.equal(ctx, c2) and torch.equal(inter, i2)
    assert torch.equal(red[0], r2[0]) and torch.equal(red[1], r2[1])


def test_plus_full_forward():
    """FABind+ layout against the live FABind+ modules."""
    from oracle import fabind_plus_oracle as orcp
    hidden, L, IT = 48, 2, 2
    mods = ref_shims.load_reference("plus")
    args = ref_shims.published_args_plus()
    m = mods.att_model.EfficientMCAttModel(
        args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, n_iter=IT, inter_cutoff=10, intra_cutoff=8,
        normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).eval()
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 23)
    m.load_state_dict(sd, strict=True)
    b = make_batch(embed=hidden, n_complexes=3, seed=7, n_c_range=(5, 25), n_p_range=(30, 70))
    with torch.no_grad():
        Xr, Hr, Pr = m(**b.clone().forward_args())
        Xo, Ho, Po = orcp.model_forward(sd, orc.make_cfg(n_layers=L, n_iter=IT), b.X, b.H, b.batch_id, b.segment_id, b.mask,
                                        b.is_global, b.compound_edge_index, b.LAS_edge_index, b.X_LAS)
    assert rel_err(Xo, Xr) < 2e-6 and rel_err(Ho, Hr) < 2e-5 and rel_err(Po, Pr) < 2e-5
