"""ComplexGraph.construct_edges on the GPU against the oracle's edge lists: bit-exact, same order."""
import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import ref_shims
from fabind_b200 import ComplexGraph
from fabind_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu


def _graph():
    return ComplexGraph(ref_shims.published_args(), inter_cutoff=10, intra_cutoff=8, normalize_coord=lambda x: x / 5.0)


def _compare(b, X):
    g = _graph()
    ctx, inter, red = g(X.cuda(), b.batch_id.cuda(), b.segment_id.cuda(), b.is_global.cuda())
    c2, i2, r2 = orc.build_edges(X, b.batch_id, b.segment_id, b.is_global, 8 / 5.0, 10 / 5.0)
    assert ctx.dtype == torch.int64 and inter.dtype == torch.int64
    assert torch.equal(ctx.cpu(), c2), f"ctx differs: {ctx.shape} vs {c2.shape}"
    assert torch.equal(inter.cpu(), i2), f"inter differs: {inter.shape} vs {i2.shape}"
    assert torch.equal(red[0].cpu(), r2[0]) and torch.equal(red[1].cpu(), r2[1])


@pytest.mark.parametrize("kw", [
    dict(n_complexes=1, seed=0, n_c=30, n_p=200),
    dict(n_complexes=5, seed=2, n_c_range=(6, 60), n_p_range=(40, 250)),
    dict(n_complexes=16, seed=3, n_c_range=(10, 80), n_p_range=(80, 250)),
])
def test_edges_match_oracle(kw):
    b = make_batch(embed=8, **kw)
    _compare(b, b.X)


def test_edges_near_cutoff_adversarial():
    """distances snapped to within a few ulp of the 8 A / 10 A cutoffs (bit-exact sets)"""
    b = make_batch(n_complexes=3, seed=9, n_c=12, n_p=60, embed=8)
    X = b.X.clone()
    gen = torch.Generator().manual_seed(0)
    n = X.shape[0]
    for _ in range(400):
        i = int(torch.randint(0, n, (1,), generator=gen))
        j = int(torch.randint(0, n, (1,), generator=gen))
        if i == j or b.batch_id[i] != b.batch_id[j] or b.is_global[i] or b.is_global[j] or not b.segment_id[i]:
            continue
        cut = (8.0 if b.segment_id[j] else 10.0) / 5.0
        d = X[i, 0] - X[j, 0]
        nn = d.norm()
        if nn < 1e-3:
            continue
        eps = int(torch.randint(-3, 4, (1,), generator=gen)) * 1.2e-7
        X[i, 0] = X[j, 0] + d / nn * (cut * (1 + eps))
    _compare(b, X)


def test_edges_zero_inter_fallback():
    b = make_batch(n_complexes=1, seed=4, n_c=5, n_p=24, embed=8)
    X = b.X.clone()
    X[1:6] += 20.0
    _compare(b, X)
