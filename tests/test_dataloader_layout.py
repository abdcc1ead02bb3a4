"""Dataloader-side layout (fabind_b200/dataloader.py): the host-side edge counts of the context graph against the
oracle's edge lists, the layout arrays against the runtime's own, and the registration path a later forward hits."""
import pickle

import numpy as np
import torch

from oracle import fabind_oracle as orc
from fabind_b200 import layout
from fabind_b200.dataloader import layout_hint, attach, prepare_batch, context_degrees
from fabind_b200.synthetic import make_batch

CUT, INTER = 8 / 5.0, 10 / 5.0


def _oracle_counts(b, X):
    ctx, _, _ = orc.build_edges(X, b.batch_id, b.segment_id, b.is_global, CUT, INTER)
    allctx = torch.cat([b.compound_edge_index, ctx], 1)
    deg = torch.bincount(allctx[0], minlength=X.shape[0])
    return deg


def test_counts_match_oracle_edge_lists():
    for kw in (dict(n_complexes=1, seed=0, n_c=30, n_p=200), dict(n_complexes=6, seed=2, n_c_range=(6, 60), n_p_range=(40, 250))):
        b = make_batch(embed=8, **kw)
        deg = _oracle_counts(b, b.X)
        mine = context_degrees(b.X, b.batch_id, b.segment_id, b.is_global, b.compound_edge_index, CUT)
        assert np.array_equal(mine, deg.numpy())
        h = layout_hint(b.X, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, CUT)
        assert h.e_ctx == int(deg.sum()) and h.e_ctx_mv == int(deg[b.mask].sum()) and h.n_bond == b.compound_edge_index.shape[1]


def test_counts_near_cutoff_adversarial():
    """residue pairs snapped to within a few ulp of the cutoff: the host count uses the reference's own CPU expression"""
    b = make_batch(n_complexes=3, seed=9, n_c=12, n_p=60, embed=8)
    X = b.X.clone()
    gen = torch.Generator().manual_seed(0)
    n = X.shape[0]
    for _ in range(400):
        i = int(torch.randint(0, n, (1,), generator=gen))
        j = int(torch.randint(0, n, (1,), generator=gen))
        if i == j or b.batch_id[i] != b.batch_id[j] or b.is_global[i] or b.is_global[j] or not b.segment_id[i] or not b.segment_id[j]:
            continue
        d = X[i, 0] - X[j, 0]
        nn = d.norm()
        if nn < 1e-3:
            continue
        eps = int(torch.randint(-3, 4, (1,), generator=gen)) * 1.2e-7
        X[i, 0] = X[j, 0] + d / nn * (CUT * (1 + eps))
    deg = _oracle_counts(b, X)
    mine = context_degrees(X, b.batch_id, b.segment_id, b.is_global, b.compound_edge_index, CUT)
    assert np.array_equal(mine, deg.numpy())


def test_hint_is_picklable_and_layout_equals_runtime_layout():
    b = make_batch(n_complexes=4, seed=5, n_c_range=(8, 40), n_p_range=(40, 120), embed=8)
    h = pickle.loads(pickle.dumps(layout_hint(b.X, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, CUT)))
    ref = layout._build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    lay = attach(h, b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    assert torch.equal(lay.blob, ref.blob) and torch.equal(lay.flags, ref.flags) and lay.offs == ref.offs
    for k in ("N", "B", "Nc_tot", "P_total", "cap_int", "fb_atom", "fb_res", "max_c", "max_p", "n_mv"):
        assert getattr(lay, k) == getattr(ref, k), k
    # the forward's lookup returns the registered object (with its counts), not a rebuilt one
    got = layout.build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    assert got is lay and got.e_ctx == h.e_ctx and got.hint_n_bond == h.n_bond
    # a different tensor object misses
    other = layout.build_layout(b.batch_id.clone(), b.segment_id, b.is_global, b.mask, "cpu")
    assert other is not lay and other.e_ctx is None


def test_prepare_batch_keeps_host_tensors_when_asked():
    b = make_batch(n_complexes=2, seed=6, n_c=10, n_p=40, embed=8)
    args = b.forward_args()
    out = prepare_batch(args, "cpu", CUT, move=False)
    assert all(out[k] is args[k] for k in args)
    assert layout.build_layout(out["batch_id"], out["segment_id"], out["is_global"], out["mask"], "cpu").e_ctx is not None
