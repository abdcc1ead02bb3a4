"""Stand-alone forwards of the stack's sub-modules (`MCAttEGNN`, `MC_E_GCL`, `MC_Att_L`) on caller-supplied graphs,
through `fb_egnn_forward`.  The fused `EfficientMCAttModel.forward` never goes through here; this exists so that
the reference's module-level API (egnn.py:130,308,392) is served by the same kernels."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .layout import build_layout
from .runtime import _scratch_buf, current_stream_ptr
from .weights import pack_state_dict, derive_on_device

_templates = {}


def _template_sd(args, hidden, n_layers, flavour=0):
    """zero-filled state_dict with the full EfficientMCAttModel key set (slots a sub-module does not own stay zero)"""
    key = (hidden, n_layers, flavour)
    if key not in _templates:
        if flavour == 1:
            from .plus.att_model import EfficientMCAttModel
        else:
            from .att_model import EfficientMCAttModel
        m = EfficientMCAttModel(args, hidden, hidden, 1, n_layers=n_layers, n_iter=1, normalize_coord=lambda x: x / 5.0,
                                unnormalize_coord=lambda x: x * 5.0)
        _templates[key] = {k: torch.zeros_like(v) for k, v in m.state_dict().items()}
    return dict(_templates[key])


def packed_arena(module, args, hidden, n_layers, prefix, device, flavour=0):
    params = list(module.parameters())
    key = (device, tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
    cache = getattr(module, "_fb_arena", None)
    if cache is None or cache[0] != key:
        sd = _template_sd(args, hidden, n_layers, flavour)
        for k, v in module.state_dict().items():
            sd[prefix + k] = v.detach().cpu()
        if flavour == 1:      # LayerNorm scales of slots the sub-module does not own: 1 keeps the folded packing finite
            for k in sd:
                if k.endswith("layernorm.weight") and not k.startswith(prefix):
                    sd[k] = torch.ones_like(sd[k])
        w32 = derive_on_device(pack_state_dict(sd, hidden, n_layers, flavour).to(device), hidden, n_layers, flavour)
        cache = (key, w32, None)
        module._fb_arena = cache
    return cache


def _csr(edges, inv, N):
    r = inv[edges[0]]
    c = inv[edges[1]]
    order = np.lexsort((c, r))     # rows ascending, columns ascending inside a row (inter_logit looks mirror edges up by bisection)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=N))]).astype(np.int32)
    return rowptr, r[order].astype(np.int32), c[order].astype(np.int32), order


def egnn_forward(module, args, prefix, hidden, n_layers, steps, h, x, ctx_edges, att_edges, las_edges, x_las, batch_id,
                 segment_id, pair_embed_batched, geom, bf16=False, want_att=False, flavour=0):
    """Returns (h_out [N,H], x_out [N,1,3], atts or None) and, for the FABind+ layout (flavour 1) with an attention step,
    additionally the propagated dense pair embedding.  Inputs are device (or CPU) tensors in the caller's order."""
    l = _lib.lib()
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("fabind_b200 runs on a CUDA device only (no CPU fallback)")
    N = h.shape[0]
    single_side = segment_id is None
    seg = torch.zeros(N, dtype=torch.bool) if single_side else segment_id
    zeros = torch.zeros(N, dtype=torch.bool)
    lay = build_layout(batch_id, seg, zeros, zeros, dev, allow_single_side=single_side)
    o = lay.offs
    blob = lay.blob.cpu().numpy()
    inv = blob[o["inv"]:o["inv"] + N]
    c_off, p_off = blob[o["c_off"]:o["c_off"] + lay.B + 1], blob[o["p_off"]:o["p_off"] + lay.B + 1]
    pair_base, cplx = blob[o["pair_base"]:o["pair_base"] + lay.B + 1], blob[o["node_cplx"]:o["node_cplx"] + N]
    ce = ctx_edges.detach().cpu().numpy() if ctx_edges is not None else np.zeros((2, 0), np.int64)
    crp, crow, ccol, _ = _csr(ce, inv, N)
    if att_edges is not None:
        ae = att_edges.detach().cpu().numpy()
        irp, irow, icol, order = _csr(ae, inv, N)
        b = cplx[irow]
        r_prot = irow >= lay.Nc_tot
        ci = np.where(r_prot, icol, irow)
        pi = np.where(r_prot, irow, icol)
        nc1 = (c_off[1:] - c_off[:-1])
        ipair = (pair_base[b] + (pi - p_off[b]) * nc1[b] + (ci - c_off[b])).astype(np.int32)
    else:
        irp, irow, icol, ipair, order = np.zeros(N + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), None
    E_ctx, E_int = int(crow.shape[0]), int(irow.shape[0])
    las = las_edges.detach().to(dev, torch.int64).contiguous() if las_edges is not None else torch.zeros((2, 0), dtype=torch.int64, device=dev)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    g_dev = [up(a) for a in (crp, crow, ccol, irp, irow, icol, ipair)]
    _, w32, w16 = packed_arena(module, args, hidden, n_layers, prefix, dev, flavour)
    if bf16 and w16 is None:
        w16 = w32.to(torch.bfloat16)
        module._fb_arena = (module._fb_arena[0], w32, w16)
    xv = x.detach().reshape(N, 3).to(dev, torch.float32).contiguous()
    hv = h.detach().to(dev, torch.float32).contiguous()
    xl = (x_las.detach().reshape(N, 3).to(dev, torch.float32).contiguous() if x_las is not None
          else torch.zeros((N, 3), dtype=torch.float32, device=dev))
    H_out = torch.empty((N, hidden), dtype=torch.float32, device=dev)
    X_out = torch.empty((N, 3), dtype=torch.float32, device=dev)
    P0 = None
    if steps & _lib.STEP_ATT:
        nc1 = (c_off[1:] - c_off[:-1])
        np1 = (p_off[1:] - p_off[:-1])
        bi = np.concatenate([np.full(nc1[b] * np1[b], b) for b in range(lay.B)])
        ip = np.concatenate([np.repeat(np.arange(np1[b]), nc1[b]) for b in range(lay.B)])
        jc = np.concatenate([np.tile(np.arange(nc1[b]), np1[b]) for b in range(lay.B)])
        P0 = pair_embed_batched.detach().to(dev, torch.float32)[up(bi), up(ip), up(jc)].contiguous()   # re-pack (data movement)
    att = torch.zeros((max(n_layers, 1), max(E_int, 1)), dtype=torch.float32, device=dev) if want_att else None

    p = _lib.ModelParams()
    p.N, p.B, p.Nc_tot, p.P_total = lay.N, lay.B, lay.Nc_tot, lay.P_total
    p.hidden, p.n_layers, p.n_iter = hidden, n_layers, 1
    p.n_bond, p.n_las = 0, las.shape[1]
    p.E_ctx, p.cap_int, p.bf16_mode = E_ctx, max(lay.cap_int, E_int, 2), 1 if bf16 else 0
    p.max_c, p.max_p, p.fb_atom, p.fb_res = lay.max_c, lay.max_p, max(lay.fb_atom, 0), max(lay.fb_res, 0)
    p.intra_cutoff, p.inter_cutoff = geom["intra_cutoff"], geom["inter_cutoff"]
    p.coord_clamp, p.las_clamp, p.las_step = geom["coord_clamp"], geom["las_clamp"], geom["las_step"]
    p.X_in, p.H_in, p.X_las = xv.data_ptr(), hv.data_ptr(), xl.data_ptr()
    p.bonds, p.las = None, las.data_ptr()
    for k in ("perm", "inv", "node_cplx", "c_off", "p_off", "pair_base"):
        setattr(p, k, lay.ptr(k))
    p.node_flags = lay.flags.data_ptr()
    p.w32, p.w16 = w32.data_ptr(), (w16.data_ptr() if w16 is not None else None)
    p.X_out, p.H_out = X_out.data_ptr(), H_out.data_ptr()
    p.flavour = flavour
    pair_out = None
    if flavour == 1 and (steps & _lib.STEP_ATT):
        pair_out = torch.zeros((lay.B, lay.max_p, lay.max_c, hidden), dtype=torch.float32, device=dev)
        p.pair_out = pair_out.data_ptr()
    st = current_stream_ptr(dev)
    gb = l.fb_graph_workspace_bytes(C.byref(p))
    wsg = _scratch_buf(dev, "graph", gb)
    p.ws_graph, p.ws_graph_bytes = wsg.data_ptr(), wsg.numel()
    _lib.check(l.fb_graph_static(C.byref(p), st), "fb_graph_static")      # LAS CSR (the geometric count is unused here)
    mb = l.fb_model_workspace_bytes(C.byref(p))
    if mb < 0:
        _lib.check(int(mb), "fb_model_workspace_bytes")
    wsm = _scratch_buf(dev, "main", mb)
    p.ws_main, p.ws_main_bytes = wsm.data_ptr(), wsm.numel()
    e = _lib.EgnnExtra()
    e.steps, e.E_int = steps, E_int
    (e.ctx_rowptr, e.ctx_row, e.ctx_col, e.int_rowptr, e.int_row, e.int_col, e.int_pair) = [t.data_ptr() for t in g_dev]
    e.pair0 = P0.data_ptr() if P0 is not None else None
    e.att_out = att.data_ptr() if att is not None else None
    _lib.check(l.fb_egnn_forward(C.byref(p), C.byref(e), st), "fb_egnn_forward")
    atts = None
    if want_att:
        ot = up(order)
        atts = []
        for i in range(n_layers):
            a = torch.empty(E_int, dtype=torch.float32, device=dev)
            a[ot] = att[i, :E_int]
            atts.append(a)
    if flavour == 1:
        return H_out, X_out.view(N, 1, 3), atts, pair_out
    return H_out, X_out.view(N, 1, 3), atts
