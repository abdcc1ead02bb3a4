"""The CPU oracle against the golden vectors produced by the unmodified reference
(scripts/make_golden.py).  Runs anywhere (no GPU, no /root/reference)."""
import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import fabind_plus_oracle as orcp
from helpers import golden_files, plus_golden_files, load_golden, rel_err


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_oracle_matches_reference_golden(path):
    g, r, b, sd, cfg = load_golden(path)
    trace = []
    with torch.no_grad():
        X, H, edges = orc.model_forward(
            sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
            b.compound_edge_index, b.LAS_edge_index, b.X_LAS, return_edges=True)
    # edge lists: bit-exact, same order; golden stores construct_edges output (before the bond
    # edges are prepended, att_model.py:231)
    assert len(edges) == len(g["edges"])
    nb = b.compound_edge_index.shape[1]
    for (ctx, inter), (gctx, ginter) in zip(edges, g["edges"]):
        assert torch.equal(ctx[:, nb:].to(torch.int32), gctx)
        assert torch.equal(inter.to(torch.int32), ginter)
    assert rel_err(X, g["X"]) < 2e-6
    assert rel_err(H, g["H"]) < 2e-5


def test_golden_present():
    assert len(golden_files()) >= 4


@pytest.mark.parametrize("path", plus_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_plus_oracle_matches_reference_golden(path):
    """FABind+ layout: (X, H, pair_embed) and the last iteration's per-sub-layer trace."""
    g, r, b, sd, cfg = load_golden(path)
    assert r["flavour"] == "plus"
    trace = []
    with torch.no_grad():
        X, H, pair, edges = orcp.model_forward(
            sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
            b.compound_edge_index, b.LAS_edge_index, b.X_LAS, trace=trace, return_edges=True)
    nb = b.compound_edge_index.shape[1]
    assert len(edges) == len(g["edges"])
    for (ctx, inter), (gctx, ginter) in zip(edges, g["edges"]):
        assert torch.equal(ctx[:, nb:].to(torch.int32), gctx)
        assert torch.equal(inter.to(torch.int32), ginter)
    assert rel_err(X, g["X"]) < 2e-6
    assert rel_err(H, g["H"]) < 2e-5
    assert pair.shape == g["pair"].shape and rel_err(pair, g["pair"]) < 2e-5
    last = {tag: (h, x) for tag, h, x in trace[-1][1]}
    for tag, h, x in g["trace_last_iter"]:
        assert rel_err(last[tag][0], h) < 2e-5 and rel_err(last[tag][1], x) < 2e-6, tag


def test_plus_golden_present():
    assert len(plus_golden_files()) >= 3
