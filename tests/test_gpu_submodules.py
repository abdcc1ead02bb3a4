"""Stand-alone sub-module forwards (reference API, egnn.py:130,308,392) through fb_egnn_forward vs the oracle."""
import pytest
import torch

from oracle import fabind_oracle as orc
from oracle.det_weights import det_state_dict
from fabind_b200.config import published_args
from fabind_b200.egnn import MC_E_GCL, MC_Att_L, MCAttEGNN
from fabind_b200.synthetic import make_batch
from helpers import rel_err

pytestmark = pytest.mark.gpu
H = 64


def _graph(b):
    ctx, inter, _ = orc.build_edges(b.X, b.batch_id, b.segment_id, b.is_global, 8 / 5.0, 10 / 5.0)
    return torch.cat([b.compound_edge_index, ctx], 1), inter


def _dense_pair(pairs, layout):
    B, offs, counts, ncp = layout
    npm = max(counts[i] - ncp[i] for i in range(B))
    ncm = max(ncp)
    out = torch.zeros(B, npm, ncm, pairs[0].shape[-1])
    for i, p in enumerate(pairs):
        out[i, :p.shape[0], :p.shape[1]] = p
    return out


def test_gcl_forward():
    b = make_batch(n_complexes=3, seed=21, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    ctx, _ = _graph(b)
    m = MC_E_GCL(published_args(), H, H, H, 1, coord_change_maximum=2.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    h, x = m(b.H.cuda(), ctx.cuda(), b.X.cuda(), batch_id=b.batch_id.cuda())
    ho, xo = orc.gcl_forward(sd, "", b.H, ctx, b.X, b.batch_id, 2.0)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4


def test_att_forward():
    b = make_batch(n_complexes=3, seed=22, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    _, inter = _graph(b)
    m = MC_Att_L(published_args(), H, H, H, 1, coord_change_maximum=2.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 4)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    layout = orc.complex_layout(b.batch_id, b.segment_id)
    g = torch.Generator().manual_seed(0)
    B, offs, counts, ncp = layout
    pairs = [torch.randn(counts[i] - ncp[i], ncp[i], H, generator=g) * 0.3 for i in range(B)]
    h, x, att = m(b.H.cuda(), inter.cuda(), b.X.cuda(), segment_id=b.segment_id.cuda(), batch_id=b.batch_id.cuda(),
                  pair_embed_batched=_dense_pair(pairs, layout).cuda())
    ho, xo, ao, _ = orc.att_forward(sd, "", b.H, inter, b.X, b.batch_id, b.segment_id, pairs, 2.0, layout)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4 and rel_err(att.cpu(), ao) < 1e-4


def test_egnn_forward_with_attention():
    b = make_batch(n_complexes=2, seed=23, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    ctx, inter = _graph(b)
    L = 2
    m = MCAttEGNN(published_args(), H, H, H, 1, n_layers=L, normalize_coord=lambda v: v / 5.0, unnormalize_coord=lambda v: v * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    layout = orc.complex_layout(b.batch_id, b.segment_id)
    g = torch.Generator().manual_seed(1)
    B, offs, counts, ncp = layout
    pairs = [torch.randn(counts[i] - ncp[i], ncp[i], H, generator=g) * 0.3 for i in range(B)]
    h, x, atts = m(b.H.cuda(), b.X.cuda(), ctx.cuda(), inter.cuda(), b.LAS_edge_index.cuda(), b.X_LAS.cuda(),
                   segment_id=b.segment_id.cuda(), batch_id=b.batch_id.cuda(), pair_embed_batched=_dense_pair(pairs, layout).cuda(),
                   return_attention=True)
    cfg = orc.make_cfg(n_layers=L, n_iter=1)
    ho, xo, ao = orc.egnn_forward(sd, "", cfg, b.H, b.X, ctx, inter, b.LAS_edge_index, b.X_LAS, b.batch_id, b.segment_id, pairs, layout)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4
    for a, r in zip(atts, ao):
        assert rel_err(a.cpu(), r) < 1e-4


# ---- FABind+ layout (FABind_plus/fabind/models/egnn.py:100-115,280-300,359-433) ------------------------------------------------
from oracle import fabind_plus_oracle as orcp          # noqa: E402
from fabind_b200.config import published_args_plus     # noqa: E402
from fabind_b200.plus import egnn as pegnn             # noqa: E402


def test_plus_gcl_forward():
    b = make_batch(n_complexes=3, seed=31, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    ctx, _ = _graph(b)
    m = pegnn.MC_E_GCL(published_args_plus(), H, H, H, 1, coord_change_maximum=2.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 6)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    h, x = m(b.H.cuda(), ctx.cuda(), b.X.cuda(), batch_id=b.batch_id.cuda())
    ho, xo = orcp.gcl_forward(sd, "", b.H, ctx, b.X, b.batch_id, 2.0)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4


def test_plus_att_forward_returns_pair():
    b = make_batch(n_complexes=3, seed=32, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    _, inter = _graph(b)
    m = pegnn.MC_Att_L(published_args_plus(), H, H, H, 1, coord_change_maximum=2.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 7)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    layout = orc.complex_layout(b.batch_id, b.segment_id)
    g = torch.Generator().manual_seed(2)
    B, offs, counts, ncp = layout
    pairs = [torch.randn(counts[i] - ncp[i], ncp[i], H, generator=g) * 0.3 for i in range(B)]
    h, x, att, pair = m(b.H.cuda(), inter.cuda(), b.X.cuda(), segment_id=b.segment_id.cuda(), batch_id=b.batch_id.cuda(),
                        pair_embed_batched=_dense_pair(pairs, layout).cuda())
    ho, xo, ao, po = orcp.att_forward(sd, "", b.H, inter, b.X, b.batch_id, b.segment_id, pairs, 2.0, layout)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4 and rel_err(att.cpu(), ao) < 1e-4
    assert rel_err(pair.cpu(), _dense_pair(po, layout)) < 1e-4


def test_plus_egnn_forward():
    b = make_batch(n_complexes=2, seed=33, embed=H, n_c_range=(8, 25), n_p_range=(40, 80))
    ctx, inter = _graph(b)
    L = 2
    m = pegnn.MCAttEGNN(published_args_plus(), H, H, H, 1, n_layers=L, normalize_coord=lambda v: v / 5.0, unnormalize_coord=lambda v: v * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 8)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    layout = orc.complex_layout(b.batch_id, b.segment_id)
    g = torch.Generator().manual_seed(3)
    B, offs, counts, ncp = layout
    pairs = [torch.randn(counts[i] - ncp[i], ncp[i], H, generator=g) * 0.3 for i in range(B)]
    h, x, atts, pair = m(b.H.cuda(), b.X.cuda(), ctx.cuda(), inter.cuda(), b.LAS_edge_index.cuda(), b.X_LAS.cuda(),
                         segment_id=b.segment_id.cuda(), batch_id=b.batch_id.cuda(), pair_embed_batched=_dense_pair(pairs, layout).cuda(),
                         return_attention=True)
    cfg = orc.make_cfg(n_layers=L, n_iter=1)
    ho, xo, ao, po = orcp.egnn_forward(sd, "", cfg, b.H, b.X, ctx, inter, b.LAS_edge_index, b.X_LAS, b.batch_id, b.segment_id, pairs, layout)
    assert rel_err(h.cpu(), ho) < 1e-4 and rel_err(x.cpu(), xo) < 1e-4 and rel_err(pair.cpu(), _dense_pair(po, layout)) < 1e-4
    for a, r in zip(atts, ao):
        assert rel_err(a.cpu(), r) < 1e-4
