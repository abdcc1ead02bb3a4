"""Thin runtime around the C ABI: weight-arena cache, scratch cache, parameter-struct assembly."""
import ctypes as C

import torch

from . import _lib
from .layout import build_layout
from .weights import pack_state_dict, derive_on_device

_scratch = {}


def _scratch_buf(device, tag, nbytes):
    key = (device, tag)
    t = _scratch.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _scratch[key] = t
    return t


_flags = {}


def layout_flag(device):
    """int32[1] on `device`, set to 1 by fb_graph_static when a dataloader-side layout (fb_model_params.layout_flag) claimed edge
    counts that differ from the device's.  Host-buffer forwards check it at their final synchronisation and raise; device-resident
    forwards never synchronise, so a caller that feeds hand-made hints checks `check_layout_flag(device)` at its own sync point."""
    from .layout import _norm_device
    dev = _norm_device(device)
    t = _flags.get(dev)
    if t is None:
        t = _flags[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    return t


def _flag_host(device):
    from .layout import _norm_device
    key = ("host", _norm_device(device))
    t = _flags.get(key)
    if t is None:
        t = _flags[key] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return t


def check_layout_flag(device):
    """Synchronising read of `layout_flag(device)`; raises (and clears the flag) when a hinted forward since the last check ran on
    a wrong hint."""
    t = layout_flag(device)
    if int(t.item()) != 0:
        t.zero_()
        raise RuntimeError("fabind_b200: a dataloader-side layout did not match its batch (context-edge counts differ)")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def current_stream_ptr(device):
    """the caller's current CUDA stream on `device` as the void* the C ABI takes.  Called once per launch by the Python-driven paths
    (1700 per training step): the raw-handle query torch itself uses for its generated kernels costs ~0.1 us, building a
    torch.cuda.Stream object ~1.5 us"""
    if _raw_stream is not None:
        dev = device if isinstance(device, torch.device) else torch.device(device)
        idx = dev.index
        if idx is None:
            idx = torch.cuda.current_device()
        return C.c_void_p(_raw_stream(idx))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def precision_mode(precision):
    """'fp32' | 'bf16' | 'bf16x3' | 'fp32_tc' (or a bool: bf16 yes / no) -> FB_PREC_*"""
    if isinstance(precision, str):
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        return _lib.PRECISIONS[precision]
    return int(precision)


def split_arena(w32, hidden, n_layers, flavour):
    """The weight arena of the split-precision modes: every matrix slot (rows, cols, off) as three bf16 planes [rows, 3*cols] at
    element 3*off (fb_split_rows: w = w0 + w1 + w2, 8 mantissa bits per plane); vector slots are read from the fp32 arena."""
    from .weights import slots
    l = _lib.lib()
    out = torch.zeros(3 * w32.numel(), dtype=torch.bfloat16, device=w32.device)
    st = current_stream_ptr(w32.device)
    for name, r, c, off in slots(hidden, n_layers, flavour, derived=True):
        if r <= 1 or r * c == 0 or c % 8:
            continue
        _lib.check(l.fb_split_rows(w32.data_ptr() + 4 * off, c, r, c, out.data_ptr() + 2 * 3 * off, st), "fb_split_rows")
    return out


class PackedWeights:
    """fp32 (and bf16 / split-bf16) weight arenas on the device, re-packed when any parameter changes.

    What "changes" means: the key holds, per parameter, `_version` and `data_ptr()`.  Optimizer steps, `load_state_dict`, `.to()`,
    `.half()` and in-place ops on the Parameter bump one of them; the owning modules additionally call `invalidate()` from
    `_apply` and a load_state_dict post-hook, which also re-walks the module tree (parameters REPLACED by new Parameter objects).
    Writes through `.data` (`p.data.copy_(ema)`) change neither: call `model.invalidate_packed_weights()` after such edits, or
    set `strict=True`, which folds a content fingerprint (sum of squares of all parameters: a few fused launches) into the key."""

    def __init__(self):
        self.key = None
        self.w32 = None
        self.w16 = {}
        self.params = None
        self.strict = False

    def invalidate(self):
        self.key = None
        self.params = None

    def get(self, module, hidden, n_layers, device, mode, flavour=0):
        # the Parameter OBJECTS of a module persist across load_state_dict / .to() / in-place optimizer updates (which bump
        # _version or change data_ptr), so the module tree is walked once; walking it on every call cost 0.3 ms of host time
        # during which the GPU sat idle at the start of each forward
        mode = precision_mode(mode)
        if self.params is None:
            self.params = list(module.parameters())
        params = self.params
        key = (device, flavour, tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if self.strict:
            key = key + (float(torch.stack(torch._foreach_norm([p.detach() for p in params])).double().square().sum()),)
        if key != self.key:
            arena = pack_state_dict(module.state_dict(), hidden, n_layers, flavour)
            self.w32 = derive_on_device(arena.to(device), hidden, n_layers, flavour)
            self.w16 = {}
            self.key = key
        w16 = None
        if mode == _lib.PREC_BF16:
            w16 = self.w16.get("bf16")
            if w16 is None:
                w16 = self.w16["bf16"] = self.w32.to(torch.bfloat16)
        elif mode in (_lib.PREC_SPLIT3, _lib.PREC_SPLIT6):
            w16 = self.w16.get("split")
            if w16 is None:
                w16 = self.w16["split"] = split_arena(self.w32, hidden, n_layers, flavour)
        return self.w32, w16


def model_forward(module, packed, X, H, batch_id, segment_id, mask, is_global, bonds, las, X_las, cfg, bf16, trace=False,
                  want_pair=None, dropout=None, n_iter=None):
    """Runs fb_graph_static + fb_model_forward.  X is updated in place (reference att_model.py:236,245).
    Returns (H_out, stats[int32 n_iter device tensor], E_ctx, trace) and, for the FABind+ layout (`want_pair` not None),
    additionally the dense pair embedding [B, max_p, max_c, hidden] (or None when want_pair is False)."""
    import os, time
    _T = os.environ.get("FABIND_B200_TIMING") == "1"
    _t = [time.perf_counter()]
    def _mark(tag):
        if _T:
            torch.cuda.synchronize()
            _t.append(time.perf_counter())
            print(f"[timing] {tag}: {(_t[-1] - _t[-2]) * 1e3:.3f} ms")
    l = _lib.lib()
    dev = next(module.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("fabind_b200 runs on a CUDA device only (no CPU fallback): move the module to cuda")
    # HOST-BUFFER mode: every input is a CPU tensor (pinned for async copies); they are copied to the device
    # here, and X / H_out are copied back before returning.  Otherwise inputs are already device tensors.
    host_mode = X.device.type == "cpu"
    N = X.shape[0]
    hidden = cfg["hidden"]
    X_host = X
    if not (X.view(N, 3).is_contiguous() and X.dtype == torch.float32):
        raise ValueError("X must be a contiguous float32 [N,1,3] tensor")
    if tuple(H.shape) != (N, hidden):
        raise ValueError(f"H must be [N, {hidden}]")
    up = lambda t, dt: t.detach().to(device=dev, dtype=dt, non_blocking=True).contiguous()
    xv = up(X.view(N, 3), torch.float32) if host_mode else X.view(N, 3)
    Hc = up(H, torch.float32)
    xl = up(X_las.reshape(N, 3), torch.float32)
    bonds = up(bonds, torch.int64)
    las = up(las, torch.int64)
    _mark("input staging")
    lay = build_layout(batch_id, segment_id, is_global, mask, dev)
    _mark("build_layout")
    flavour = cfg.get("flavour", _lib.FLAVOUR_V1)
    mode = precision_mode(bf16)
    w32, w16 = packed.get(module, hidden, cfg["n_layers"], dev, mode, flavour)
    H_out = torch.empty((N, hidden), dtype=torch.float32, device=dev)
    stats = torch.zeros(cfg["n_iter"], dtype=torch.int32, device=dev)

    p = _lib.ModelParams()
    p.N, p.B, p.Nc_tot, p.P_total = lay.N, lay.B, lay.Nc_tot, lay.P_total
    p.hidden, p.n_layers, p.n_iter = hidden, cfg["n_layers"], cfg["n_iter"]
    p.n_bond, p.n_las = bonds.shape[1], las.shape[1]
    p.E_ctx, p.cap_int, p.bf16_mode = 0, lay.cap_int, mode
    # moving-rows subset of the out_layer (fb_model_params.n_mv); `module.moving_rows = False` evaluates every out_layer on all edges
    p.n_mv = lay.n_mv if getattr(module, "moving_rows", True) else 0
    p.fb_atom, p.fb_res = lay.fb_atom, lay.fb_res
    p.max_c, p.max_p = lay.max_c, lay.max_p
    p.intra_cutoff, p.inter_cutoff = cfg["intra_cutoff"], cfg["inter_cutoff"]
    p.coord_clamp, p.las_clamp, p.las_step = cfg["coord_clamp"], cfg["las_clamp"], cfg["las_step"]
    p.X_in, p.H_in, p.X_las = xv.data_ptr(), Hc.data_ptr(), xl.data_ptr()
    p.bonds, p.las = bonds.data_ptr(), las.data_ptr()
    for k in ("perm", "inv", "node_cplx", "c_off", "p_off", "pair_base"):
        setattr(p, k, lay.ptr(k))
    p.node_flags = lay.flags.data_ptr()
    p.w32 = w32.data_ptr()
    p.w16 = w16.data_ptr() if w16 is not None else None
    p.X_out, p.H_out, p.stats = xv.data_ptr(), H_out.data_ptr(), stats.data_ptr()
    p.flavour = flavour
    if n_iter is not None:                       # training-mode `random_n_iter` draw (att_model.py:210-211)
        p.n_iter = int(n_iter)
    p.attn_tc = 1 if getattr(module, "attention", "simt") == "tcgen05" else 0
    if dropout is not None and dropout[0] > 0:   # (p, seed, colonly): FABind+ sampling mode
        p.dropout_p, p.dropout_seed, p.dropout_colonly = float(dropout[0]), int(dropout[1]) & 0xFFFFFFFF, int(bool(dropout[2]))
    pair = None
    if flavour == _lib.FLAVOUR_PLUS and want_pair:
        pair = torch.zeros((lay.B, lay.max_p, lay.max_c, hidden), dtype=torch.float32, device=dev)
        p.pair_out = pair.data_ptr()

    tr = None
    if trace:
        nl = max(1, 2 * cfg["n_layers"])
        tr = (torch.zeros((nl, N, hidden), dtype=torch.float32, device=dev),
              torch.zeros((nl, N, 3), dtype=torch.float32, device=dev))
        p.trace_h, p.trace_x = tr[0].data_ptr(), tr[1].data_ptr()
    _mark("weights+params")
    st = current_stream_ptr(dev)
    gbytes = l.fb_graph_workspace_bytes(C.byref(p))
    if gbytes < 0:
        _lib.check(int(gbytes), "fb_graph_workspace_bytes")
    wsg = _scratch_buf(dev, "graph", gbytes)
    p.ws_graph, p.ws_graph_bytes = wsg.data_ptr(), wsg.numel()
    # Edge counts of the context graph.  Supplied by a dataloader-side layout (fabind_b200/dataloader.py: counted on the CPU at
    # collate time): the host sizes the edge-level scratch from them and the device only CHECKS the claim (fb_model_params.
    # layout_flag) -- no device->host read between entry and outputs.  Otherwise: one host read per forward after fb_graph_static
    # (the same read fetches the number of context edges into the moving rows, fb_model_params.E_ctx_mv).
    hinted = (lay.e_ctx is not None and lay.hint_n_bond == p.n_bond and lay.hint_cutoff is not None
              and abs(lay.hint_cutoff - cfg["intra_cutoff"]) <= 1e-6 * abs(cfg["intra_cutoff"]))
    if hinted:
        p.E_ctx, p.E_ctx_mv = lay.e_ctx, lay.e_ctx_mv
        p.layout_flag = layout_flag(dev).data_ptr()
        _lib.check(l.fb_graph_static(C.byref(p), st), "fb_graph_static")
        e_ctx = lay.e_ctx
    else:
        _lib.check(l.fb_graph_static(C.byref(p), st), "fb_graph_static")
        cnt_ptr = l.fb_graph_counts_ptr(C.byref(p))
        off = (cnt_ptr - wsg.data_ptr()) // 4
        e_ctx, e_ctx_mv = wsg[: (off + 2) * 4].view(torch.int32)[off:off + 2].tolist()
        p.E_ctx, p.E_ctx_mv = e_ctx, e_ctx_mv
    _mark("graph_static + E_ctx read")
    mbytes = l.fb_model_workspace_bytes(C.byref(p))
    if mbytes < 0:
        _lib.check(int(mbytes), "fb_model_workspace_bytes")
    wsm = _scratch_buf(dev, "main", mbytes)
    p.ws_main, p.ws_main_bytes = wsm.data_ptr(), wsm.numel()
    _t0 = time.perf_counter()
    _lib.check(l.fb_model_forward(C.byref(p), st), "fb_model_forward")
    if _T:
        print(f"[timing] fb_model_forward enqueue (host): {(time.perf_counter() - _t0) * 1e3:.3f} ms")
    _mark("fb_model_forward (enqueue + device)")
    # keep every tensor the enqueued kernels read alive until the stream has consumed them
    for t in (xv, Hc, xl, bonds, las, lay.blob, lay.flags, w32, wsg, wsm):
        t.record_stream(torch.cuda.current_stream(dev))
    if host_mode:
        X_host.view(N, 3).copy_(xv, non_blocking=True)
        H_host = torch.empty((N, hidden), dtype=torch.float32, pin_memory=True)
        H_host.copy_(H_out, non_blocking=True)
        if hinted:
            fl = _flag_host(dev)
            fl.copy_(layout_flag(dev), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        H_out = H_host
        if hinted and int(fl[0]) != 0:
            layout_flag(dev).zero_()
            raise RuntimeError("fabind_b200: the dataloader-side layout of this batch does not match its coordinates / bond list "
                               "(context-edge counts differ from the device's): outputs are invalid; rebuild the hint "
                               "(fabind_b200.dataloader.layout_hint) from the tensors that are passed to the forward")
    if tr is not None:
        perm = lay.blob[lay.offs["perm"]:lay.offs["perm"] + N].long()
        th = torch.empty_like(tr[0]); tx = torch.empty_like(tr[1])
        th[:, perm] = tr[0]; tx[:, perm] = tr[1]
        tr = (th, tx)
    if want_pair is not None:
        if host_mode and pair is not None:
            pair = pair.cpu()
        return H_out, stats, e_ctx, tr, pair
    return H_out, stats, e_ctx, tr
