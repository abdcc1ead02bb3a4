"""Reverse pass of the docking stack on the GPU (training path, fp32): the primitives, the MC_E_GCL / MC_Att_L / LAS sub-layer
reverse passes and the whole last-iteration reverse pass of the v1 stack (`stack_backward_v1`).

The reference trains through torch autograd (main_fabind.py:380-401).  Here the reverse pass is hand-derived for the library's
formulation (hoisted first Linears, collapsed pair-bias vector, pair work on the interface pairs only; specification pinned
against autograd and the unmodified reference's parameter gradients: tests/emulate_backward.py) and runs as launches of
csrc/backward.cu plus fb_gemm for the data-gradient GEMMs (`dX = dY W` on a transposed weight).  torch only allocates / copies
device memory here; there is no torch arithmetic on the path and no fallback: CPU tensors raise.

Status: every kernel and orchestration function here is parity-green on a B200 against the specification's arena gradient and,
assembled into the training step, against the unmodified reference's parameter gradients in both layouts
(tests/test_gpu_train_reverse.py, tests/test_gpu_train_reverse_att.py, tests/test_gpu_train_forward.py); the orchestration is also
validated on the CPU with every kernel wrapper swapped for its torch definition (tests/test_backward_orchestration.py).  The v1
layout carries the reference's training-mode dropout (class Drop); the FABind+ reverse pass does not yet."""
import ctypes as C

import torch

from . import _lib
from .runtime import current_stream_ptr

ACT_NONE, ACT_SILU, ACT_RELU = 0, 1, 2

# GEMMs of the training path: "fp32" = the FFMA parity kernel (what the gradient-parity tests pin); "bf16" = the tcgen05 kernels of
# the inference path (bf16 operands, fp32 accumulation and outputs) for the forward / data-gradient GEMMs, and for the weight
# gradients dW = dY^T X as a GEMM over transposed bf16 copies (reduction over the rows, padded to a multiple of 64).  The casts and
# transposed copies are torch copies (no arithmetic).  Not validated on a GPU yet: default "fp32".
PRECISION = "fp32"
WGRAD_TC_MIN_ROWS = 2048
WGRAD_SPLIT = 4            # row splits of a tensor-core weight gradient = problems of one multi-problem launch (gemm_tc5.cu: up to 4)


# bf16 twins of the weight operands (PRECISION = "bf16"): train.slot_tensors converts the whole arena ONCE per step and registers,
# per fp32 weight view, the bf16 view of the same slot -- no per-GEMM conversion kernels; anything unregistered is converted on the spot
_W16 = {}


def register_bf16(w32, w16):
    _W16[(w32.data_ptr(), tuple(w32.shape))] = w16


def clear_bf16():
    _W16.clear()


def bf16_twin(W):
    if W.dtype == torch.bfloat16:
        return W if W.is_contiguous() else W.contiguous()
    t = _W16.get((W.data_ptr(), tuple(W.shape)))
    return t if t is not None else W.to(torch.bfloat16).contiguous()


def transpose_bf16(X, Mp):
    """X [M, N] fp32 -> bf16 [N, Mp] (Mp >= M, a multiple of 64; the pad columns are zero): the K-major operand of a weight-gradient GEMM"""
    _chk(X)
    M, N = X.shape
    out = torch.empty(N, Mp, dtype=torch.bfloat16, device=X.device)
    _lib.check(_lib.lib().fb_transpose_bf16(X.data_ptr(), N, M, N, out.data_ptr(), Mp, _st(X)), "fb_transpose_bf16")
    return out


class Drop:
    """nn.Dropout of the reference's training step (v1 stack: egnn.py:82,106,236,398,461; cross_att.py:128) for ONE refinement
    iteration, with the library's counter-based masks (csrc/common.cuh, fabind_b200/dropout.py): the forward applies
    keep(seed_it, site, row, col) / (1 - p), the reverse pass applies the same function to the incoming gradient -- no mask is stored."""

    def __init__(self, p, seed, colonly, it):
        from .dropout import iter_seed
        self.p, self.seed, self.colonly = float(p), iter_seed(seed, it), int(bool(colonly))

    def site(self, layer, name):
        from .dropout import site_id
        return site_id(layer, name)

    def apply(self, X, layer, name, row0=0):
        """drop(X) as a new tensor (X [M, N] fp32 contiguous, or a contiguous row slice)"""
        _chk(X)
        Y = torch.empty_like(X)
        M, N = X.shape
        _lib.check(_lib.lib().fb_dropout_apply(X.data_ptr(), Y.data_ptr(), N, M, N, self.p, self.seed, self.site(layer, name), int(row0),
                                               self.colonly, _st(X)), "fb_dropout_apply")
        return Y


def _drop(drop, X, layer, name, row0=0):
    return X if drop is None or drop.p <= 0 else drop.apply(X, layer, name, row0)


def _gemm_call(A, W, bias, act, res, M, N, K, drop=None, A16=None):
    """C[M,N] = drop(act(A W^T + bias)) + res through fb_gemm in the current PRECISION; A [M,K], W [N,K] fp32 in, fp32 out.
    drop = (Drop, layer, site name, row0) or None: dropout in the epilogue, after the activation and before the residual."""
    g = _lib.GemmParams()
    if drop is not None and drop[0] is not None and drop[0].p > 0:
        d, layer, name, row0 = drop
        g.drop_p, g.drop_seed, g.drop_site, g.drop_row0, g.drop_colonly = d.p, d.seed, d.site(layer, name), int(row0), d.colonly
    bf16 = PRECISION == "bf16" and K % 8 == 0
    if bf16:
        # A16: the producer already wrote the bf16 twin of A (fused kernels of the training forward): no conversion pass
        A = A16 if (A16 is not None and A16.dtype == torch.bfloat16 and A16.shape == A.shape) else A.to(torch.bfloat16).contiguous()
        W = bf16_twin(W)
    elif W.dtype != torch.float32:
        raise RuntimeError("fabind_b200.backward: bf16 weight operand outside PRECISION = 'bf16'")
    g.A, g.lda, g.K1 = A.data_ptr(), K, K
    g.W, g.bias, g.act = W.data_ptr(), (bias.data_ptr() if bias is not None else None), act
    g.M, g.N, g.bf16_mode, g.force_simt = M, N, int(bf16), 0
    out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    g.C, g.ldc = out.data_ptr(), N
    if res is not None:
        g.res, g.ldres = res.data_ptr(), res.shape[1]
    _lib.check(_lib.lib().fb_gemm(C.byref(g), _st(A)), "fb_gemm")
    return out


def _chk(t, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError("fabind_b200.backward: CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"fabind_b200.backward: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    return t


def _st(t):
    return current_stream_ptr(t.device)


def act_fwd(Z, act):
    _chk(Z)
    Y = torch.empty_like(Z)
    _lib.check(_lib.lib().fb_act_fwd(Z.data_ptr(), Y.data_ptr(), Z.numel(), act, _st(Z)), "fb_act_fwd")
    return Y


def act_bwd(Z, dY, act):
    _chk(Z), _chk(dY)
    dZ = torch.empty_like(Z)
    _lib.check(_lib.lib().fb_act_bwd(Z.data_ptr(), dY.data_ptr(), dZ.data_ptr(), Z.numel(), act, _st(Z)), "fb_act_bwd")
    return dZ


def outer_act_bwd(Z, u, v, act):
    """dZ[m,n] = u[m] v[n] act'(Z[m,n])"""
    _chk(Z), _chk(u), _chk(v)
    M, N = Z.shape
    dZ = torch.empty_like(Z)
    _lib.check(_lib.lib().fb_outer_act_bwd(Z.data_ptr(), u.data_ptr(), v.data_ptr(), dZ.data_ptr(), M, N, act, _st(Z)), "fb_outer_act_bwd")
    return dZ


def act_bwd_drop(Z, dY, act, drop, layer, name):
    """dZ = drop(dY) * act'(Z) -> (fp32, bf16 twin or None): the gradient of a dropped activation in one pass"""
    if not (FUSED_FORWARD and Z.is_cuda and Z.shape[1] % 4 == 0):
        return act_bwd(Z, _drop(drop, dY, layer, name), act), None
    _chk(Z), _chk(dY)
    M, N = Z.shape
    dZ = torch.empty_like(Z)
    d16 = torch.empty(M, N, dtype=torch.bfloat16, device=Z.device) if PRECISION == "bf16" else None
    on = drop is not None and drop.p > 0
    _lib.check(_lib.lib().fb_act_bwd_drop(Z.data_ptr(), dY.data_ptr(), M, N, act, drop.p if on else 0.0, drop.seed if on else 0,
                                          drop.site(layer, name) if on else 0, 0, drop.colonly if on else 0, dZ.data_ptr(),
                                          d16.data_ptr() if d16 is not None else None, _st(Z)), "fb_act_bwd_drop")
    return dZ, d16


def outer_act_bwd16(Z, u, v, act):
    """outer_act_bwd with the bf16 twin of the result -> (dZ, bf16 twin or None)"""
    if not (FUSED_FORWARD and Z.is_cuda and Z.shape[1] % 4 == 0 and PRECISION == "bf16"):
        return outer_act_bwd(Z, u, v, act), None
    _chk(Z), _chk(u), _chk(v)
    M, N = Z.shape
    dZ = torch.empty_like(Z)
    d16 = torch.empty(M, N, dtype=torch.bfloat16, device=Z.device)
    _lib.check(_lib.lib().fb_outer_act_bwd2(Z.data_ptr(), u.data_ptr(), v.data_ptr(), dZ.data_ptr(), d16.data_ptr(), M, N, act, _st(Z)),
               "fb_outer_act_bwd2")
    return dZ, d16


def colsum(A, w=None, out=None):
    """out[n] += sum_m w[m] A[m,n]"""
    _chk(A)
    M, N = A.shape
    if out is None:
        out = torch.zeros(N, dtype=torch.float32, device=A.device)
    _lib.check(_lib.lib().fb_colsum(A.data_ptr(), N, M, N, _chk(w).data_ptr() if w is not None else None, out.data_ptr(), _st(A)),
               "fb_colsum")
    return out


def rowdot(A, v):
    _chk(A), _chk(v)
    M, N = A.shape
    out = torch.empty(M, dtype=torch.float32, device=A.device)
    _lib.check(_lib.lib().fb_rowdot(A.data_ptr(), N, M, N, v.data_ptr(), out.data_ptr(), _st(A)), "fb_rowdot")
    return out


def scatter_add_rows(src, idx, dst, col0=0, width=None):
    """dst[idx[e], col0:col0+width] += src[e, :]"""
    _chk(src), _chk(dst), _chk(idx, torch.int32)
    E, D = src.shape
    width = D if width is None else width
    _lib.check(_lib.lib().fb_scatter_add_rows(src.data_ptr(), D, idx.data_ptr(), E, width, dst.data_ptr() + 4 * col0, dst.shape[1],
                                              _st(src)), "fb_scatter_add_rows")
    return dst


def gather_add_rows(src, idx, dst, col0=0):
    """dst[e, :] += src[idx[e], col0:col0+dst.shape[1]]"""
    _chk(src), _chk(dst), _chk(idx, torch.int32)
    E, D = dst.shape
    _lib.check(_lib.lib().fb_gather_add_rows(src.data_ptr() + 4 * col0, src.shape[1], idx.data_ptr(), E, D, dst.data_ptr(), D,
                                             _st(src)), "fb_gather_add_rows")
    return dst


def gemm_wgrad(dY, X, out=None):
    """dW[n,k] += sum_m dY[m,n] X[m,k]"""
    _chk(dY), _chk(X)
    M, N = dY.shape
    K = X.shape[1]
    if PRECISION == "bf16" and M >= WGRAD_TC_MIN_ROWS:
        # dW = dY^T X as a GEMM over transposed bf16 operands: [N, K] output tiles (16 for a 512 x 512 weight) with the REDUCTION over
        # the M rows.  As one problem it ran on 16 of 148 SMs for 220 us per edge-level weight (44.9k rows, 702 k-slabs per tile) and
        # 480 us for the pair rows; split over the rows into WGRAD_SPLIT independent problems of ONE multi-problem launch
        # (fb_gemm_multi, gemm_tc5.cu) with a partial output each, summed afterwards.
        S = WGRAD_SPLIT if (K % 128 == 0 and M >= 2 * WGRAD_TC_MIN_ROWS) else 1
        Mp = (M + 64 * S - 1) // (64 * S) * (64 * S)
        At, Wt = transpose_bf16(dY, Mp), transpose_bf16(X, Mp)
        chunk = Mp // S
        r = torch.empty(S, N, K, dtype=torch.float32, device=dY.device)
        arr = (_lib.GemmParams * S)()
        for i in range(S):
            g = arr[i]
            g.A, g.lda, g.K1 = At.data_ptr() + 2 * i * chunk, Mp, chunk
            g.W, g.bias, g.act, g.ldw = Wt.data_ptr() + 2 * i * chunk, None, ACT_NONE, (Mp if S > 1 else 0)
            g.M, g.N, g.bf16_mode, g.force_simt = N, K, 1, 0
            g.C, g.ldc = r.data_ptr() + 4 * i * N * K, K
        if S == 1:
            _lib.check(_lib.lib().fb_gemm(C.byref(arr[0]), _st(dY)), "fb_gemm")
        else:
            # prefetch_w = 0: the "weights" of these problems (X^T) were written by the transpose just before
            _lib.check(_lib.lib().fb_gemm_multi(arr, S, 0, _st(dY)), "fb_gemm_multi")
        r = r[0] if S == 1 else r.sum(0)
        return r if out is None else vec_add_(out, r)
    if out is None:
        out = torch.zeros(N, K, dtype=torch.float32, device=dY.device)
    _lib.check(_lib.lib().fb_gemm_wgrad(dY.data_ptr(), N, X.data_ptr(), K, M, N, K, out.data_ptr(), K, _st(dY)), "fb_gemm_wgrad")
    return out


def gemm_dgrad(dY, Wt, dY16=None):
    """dX = dY W, with Wt = W^T stored [K_in, N_out] contiguous (the 'weight' of the data-gradient GEMM; bf16 under PRECISION = 'bf16');
    dY16 = bf16 twin of dY if its producer wrote one"""
    _chk(dY)
    M, N = dY.shape
    return _gemm_call(dY, Wt, None, ACT_NONE, None, M, Wt.shape[0], N, None, dY16)


def coord_step_bwd(x, row, col, s, step, cnt, cmax, dx_new):
    """returns (dx incl. the identity path, ds)"""
    _chk(x), _chk(s), _chk(step), _chk(dx_new), _chk(row, torch.int32), _chk(col, torch.int32)
    dx = dx_new.clone()
    ds = torch.empty_like(s)
    _lib.check(_lib.lib().fb_coord_step_bwd(x.data_ptr(), row.data_ptr(), col.data_ptr(), row.numel(), s.data_ptr(), step.data_ptr(),
                                            _chk(cnt).data_ptr() if cnt is not None else None, float(cmax), dx_new.data_ptr(),
                                            dx.data_ptr(), ds.data_ptr(), _st(x)), "fb_coord_step_bwd")
    return dx, ds


def radial_bwd(x, row, col, node_cplx, nrm, drn, dx):
    _chk(x), _chk(nrm), _chk(drn), _chk(dx), _chk(node_cplx, torch.int32)
    dot = torch.zeros_like(nrm)
    _lib.check(_lib.lib().fb_radial_bwd(x.data_ptr(), row.data_ptr(), col.data_ptr(), row.numel(), node_cplx.data_ptr(), nrm.data_ptr(),
                                        drn.data_ptr(), dot.data_ptr(), dx.data_ptr(), _st(x)), "fb_radial_bwd")
    return dx


def las_bwd(x, xref, a_idx, b_idx, acc, step_size, lcl, dx_new):
    _chk(x), _chk(xref), _chk(acc), _chk(dx_new), _chk(a_idx, torch.int32), _chk(b_idx, torch.int32)
    dx = dx_new.clone()
    _lib.check(_lib.lib().fb_las_bwd(x.data_ptr(), xref.data_ptr(), a_idx.data_ptr(), b_idx.data_ptr(), a_idx.numel(), acc.data_ptr(),
                                     float(step_size), float(lcl), dx_new.data_ptr(), dx.data_ptr(), _st(x)), "fb_las_bwd")
    return dx


def _linear_bwd(grads, wname, bname, Wt, X, dY, need_dx=True, dY16=None):
    """y = x W^T + b: accumulates dW, db into `grads`, returns dX = dY W"""
    grads[wname] = gemm_wgrad(dY, X, grads.get(wname))
    if bname is not None:
        grads[bname] = colsum(dY, None, grads.get(bname))
    if not need_dx:
        return None
    return gemm_dgrad(dY, Wt, dY16) if dY16 is not None else gemm_dgrad(dY, Wt)


def gcl_backward(w, saved, row, col, node_cplx, cmax, dh_new, dx_new, drop=None, layer=0):
    """Reverse pass of one MC_E_GCL sub-layer (v1 layout; forward: csrc/forward.cu::run_gcl, reference egnn.py:68-144).

    w: dict of fp32 CUDA tensors in the packed-arena naming -- e1_rc [2H,H], e1_rad [H], e2_w, c1_w [H,H], c2_w [H], n1_w [H,2H],
       n2_w [H,H] -- and their transposes `<name>_t` (made once per optimizer step).
    saved (training-mode forward): h [N,H], x [N,3], rn [E], nrm [B], Z1, Z2, Z3 [E,H] (pre-activations of edge_mlp.0 / edge_mlp.2 /
       coord_mlp.0), s [E], deg [N] (float), step [N,3] (unclamped), agg [N,H], Z4 [N,H] (pre-activation of node_mlp.0).
    row, col: int32 context edges (destination, source).  Returns (dh, dx, grads) with grads keyed like `w`."""
    h, x = saved["h"], saved["x"]
    N, H = h.shape
    grads = {}
    # coordinate branch
    dx, ds = coord_step_bwd(x, row, col, saved["s"], saved["step"], saved["deg"], cmax, dx_new)
    T3 = saved["T3"] if "T3" in saved else act_fwd(saved["Z3"], ACT_SILU)
    grads["c2_w"] = colsum(T3, ds)
    dZ3, dZ3h = outer_act_bwd16(saved["Z3"], ds, w["c2_w"], ACT_SILU)        # dZ3[e,f] = ds[e] c2[f] silu'(Z3[e,f])
    # the edge message as the forward used it (egnn.py:82)
    M = saved["M"] if "M" in saved else _drop(drop, act_fwd(saved["Z2"], ACT_SILU), layer, "edge2")
    dM = _linear_bwd(grads, "c1_w", "c1_b", w["c1_w_t"], M, dZ3, dY16=dZ3h)
    # node branch: h_new = h + drop(n2(silu(n1([h | agg]))))
    t1 = saved["t1"] if "t1" in saved else act_fwd(saved["Z4"], ACT_SILU)
    dt1 = _linear_bwd(grads, "n2_w", "n2_b", w["n2_w_t"], t1, _drop(drop, dh_new, layer, "node2"))
    dZ4 = act_bwd(saved["Z4"], dt1, ACT_SILU)
    cat = torch.empty(N, 2 * H, dtype=torch.float32, device=h.device)
    cat[:, :H].copy_(h)
    cat[:, H:].copy_(saved["agg"])
    dcat = _linear_bwd(grads, "n1_w", "n1_b", w["n1_w_t"], cat, dZ4)
    gather_add_rows(dcat, row, dM, col0=H)                    # dM[e] += dagg[row[e]]
    # edge MLP (dM is the gradient of the DROPPED message)
    dZ2, dZ2h = act_bwd_drop(saved["Z2"], dM, ACT_SILU, drop, layer, "edge2")
    A1 = saved["A1"] if "A1" in saved else act_fwd(saved["Z1"], ACT_SILU)
    dA1 = _linear_bwd(grads, "e2_w", "e2_b", w["e2_w_t"], A1, dZ2, dY16=dZ2h)
    dZ1 = act_bwd(saved["Z1"], dA1, ACT_SILU)
    grads["e1_b"] = colsum(dZ1)
    grads["e1_rad"] = colsum(dZ1, saved["rn"])
    drn = rowdot(dZ1, w["e1_rad"])
    dPn = torch.zeros(N, 2 * H, dtype=torch.float32, device=h.device)
    scatter_add_rows(dZ1, row, dPn, col0=0)
    scatter_add_rows(dZ1, col, dPn, col0=H)
    dh_pn = _linear_bwd(grads, "e1_rc", None, w["e1_rc_t"], h, dPn)
    # dh = dh_new + dcat[:, :H] + dh_pn: two row-wise accumulations (identity gather)
    dh = dh_new.clone()
    ident = torch.arange(N, dtype=torch.int32, device=h.device)
    gather_add_rows(dcat, ident, dh, col0=0)
    gather_add_rows(dh_pn, ident, dh, col0=0)
    radial_bwd(x, row, col, node_cplx, saved["nrm"], drn, dx)
    return dh, dx, grads


# ------------------------------------------------------------------------------------------------------------------------
# second group: MC_Att_L reverse and the whole last-iteration reverse pass of the v1 stack.  The orchestration below is also
# validated on the CPU against the pinned specification (tests/test_backward_orchestration.py swaps every wrapper for its torch
# definition), so a kernel regression and an orchestration regression show up in different tests.
# ------------------------------------------------------------------------------------------------------------------------
def rowdot2(A, Bm):
    _chk(A), _chk(Bm)
    M, N = A.shape
    out = torch.empty(M, dtype=torch.float32, device=A.device)
    _lib.check(_lib.lib().fb_rowdot2(A.data_ptr(), N, Bm.data_ptr(), N, M, N, out.data_ptr(), _st(A)), "fb_rowdot2")
    return out


def scale_rows(A, u):
    """A[m,:] *= u[m] (in place)"""
    _chk(A), _chk(u)
    _lib.check(_lib.lib().fb_rows_update(A.data_ptr(), A.shape[1], A.shape[0], A.shape[1], u.data_ptr(), None, 0, _st(A)), "fb_rows_update")
    return A


def rank1_add(A, u, v):
    """A[m,n] += u[m] v[n] (in place)"""
    _chk(A), _chk(u), _chk(v)
    _lib.check(_lib.lib().fb_rows_update(A.data_ptr(), A.shape[1], A.shape[0], A.shape[1], u.data_ptr(), v.data_ptr(), 1, _st(A)),
               "fb_rows_update")
    return A


def vec_mul(a, b):
    _chk(a), _chk(b)
    c = torch.empty_like(a)
    _lib.check(_lib.lib().fb_vec_op(a.data_ptr(), b.data_ptr(), c.data_ptr(), a.numel(), 0, _st(a)), "fb_vec_op")
    return c


def vec_add_(c, a):
    """c += a (in place; c and a contiguous, same numel)"""
    _chk(a), _chk(c)
    _lib.check(_lib.lib().fb_vec_op(c.data_ptr(), a.data_ptr(), c.data_ptr(), c.numel(), 1, _st(a)), "fb_vec_op")
    return c


def gather_rows(src, idx, col0=0, width=None):
    """src[idx, col0:col0+width] as a new contiguous tensor"""
    width = src.shape[1] - col0 if width is None else width
    dst = torch.zeros(idx.numel(), width, dtype=torch.float32, device=src.device)
    return gather_add_rows(src, idx, dst, col0=col0)


def softmax_seg_bwd(alpha, dalpha, row, n_rows):
    _chk(alpha), _chk(dalpha), _chk(row, torch.int32)
    t = torch.zeros(n_rows, dtype=torch.float32, device=alpha.device)
    out = torch.empty_like(alpha)
    _lib.check(_lib.lib().fb_softmax_seg_bwd(alpha.data_ptr(), dalpha.data_ptr(), row.data_ptr(), alpha.numel(), t.data_ptr(),
                                             out.data_ptr(), _st(alpha)), "fb_softmax_seg_bwd")
    return out


def pair_bias_gate_bwd(raw, dPB):
    """raw [P, ld], dPB [P, nblk, 4] -> draw [P, ld]"""
    _chk(raw), _chk(dPB)
    draw = torch.empty_like(raw)
    _lib.check(_lib.lib().fb_pair_bias_gate_bwd(raw.data_ptr(), raw.shape[1], raw.shape[0], dPB.shape[1], dPB.data_ptr(), draw.data_ptr(),
                                                _st(raw)), "fb_pair_bias_gate_bwd")
    return draw


def pair_outer_bwd(douter, pc, geo):
    """douter [P,H], pc [N,H] (both sides, internal order) -> dpc [N,H]"""
    _chk(douter), _chk(pc)
    N, H = pc.shape
    dpc = torch.zeros_like(pc)
    _lib.check(_lib.lib().fb_pair_outer_bwd(douter.data_ptr(), pc.data_ptr(), H, geo["c_off"].data_ptr(), geo["p_off"].data_ptr(),
                                            geo["pair_base"].data_ptr(), geo["node_cplx"].data_ptr(), geo["Nc"], N - geo["Nc"],
                                            dpc.data_ptr(), _st(pc)), "fb_pair_outer_bwd")
    return dpc


def row_attention_bwd(geo, q_is_prot, Q, G, K, V, PB, dO, dQ, dG, dK, dV):
    """Q, G, K, V, dQ, dG, dK, dV: (tensor, first column) pairs -- column slices of the stacked projection buffers; compound-side
    buffers have Nc rows, protein-side buffers N - Nc rows (internal node id minus Nc).  PB [P,4] -> returns dPB [P,4]."""
    Nc = geo["Nc"]

    def ptr(tc, prot_side):
        t, c0 = tc
        _chk(t)
        return t.data_ptr() + 4 * c0 - (4 * Nc * t.shape[1] if prot_side else 0), t.shape[1]
    qs, ks = bool(q_is_prot), not bool(q_is_prot)
    (q, ldq), (g, ldg), (k, ldk), (v, ldv) = ptr(Q, qs), ptr(G, qs), ptr(K, ks), ptr(V, ks)
    (dq, lddq), (dg, lddg), (dk, lddk), (dv, lddv) = ptr(dQ, qs), ptr(dG, qs), ptr(dK, ks), ptr(dV, ks)
    _chk(PB), _chk(dO)
    dPB = torch.zeros_like(PB)
    do = dO.data_ptr() - (4 * Nc * dO.shape[1] if qs else 0)
    _lib.check(_lib.lib().fb_row_attention_bwd(geo["c_off"].data_ptr(), geo["p_off"].data_ptr(), geo["pair_base"].data_ptr(), geo["B"],
                                               int(qs), geo["max_p"] if qs else geo["max_c"], geo["max_c"] if qs else geo["max_p"], q, ldq,
                                               g, ldg, k, ldk, v, ldv,
                                               PB.data_ptr(), do, dO.shape[1], dq, lddq, dg, lddg, dk, lddk, dv, lddv, dPB.data_ptr(),
                                               _st(dO)), "fb_row_attention_bwd")
    return dPB


HD = 128


def att_backward(w, sv, geo, row, col, cmax, dh3, dx_new, dP0, drop=None, layer=0):
    """Reverse pass of one MC_Att_L sub-layer (v1 layout; forward: csrc/forward.cu::run_att; reference egnn.py:186-333,
    cross_att.py:24-54).  Specification: tests/emulate_backward.py::att_bwd.

    w: packed-arena tensors of the layer + `<name>_t` transposes of the matrices.  geo: int32 device tensors c_off, p_off, pair_base,
    node_cplx and ints Nc, B, max_c, max_p.  row, col: int32 interface edges (destination, source).
    sv (training-mode forward): h_in [N,H], x [N,3], CAc [Nc,4*128], CAp [Np,2*128], CAp2 [Np,2*128], PB_p, PB_c [P,4], Op, Oc, hp1, hc1,
      Ttp, Ttc (transition hidden, post-ReLU), h2, QK [N,4H+128], pc32 [N,32], pair [E] int32, u_pair / u_pi / u_ci [U] int32 (pair row,
      protein node, compound node of the lig->prot edges), zcat [U,H+64], Zp [U,H], rn [E], nrm [B], alpha, se [E], zc [E,H], step [N,3].
    dP0 [P,H] is accumulated in place.  Returns (dh, dx, grads, dPB_p, dPB_c)."""
    Nc, N = geo["Nc"], sv["h2"].shape[0]
    H = sv["h2"].shape[1]
    dev = dh3.device
    x, rn, alpha, se, QK = sv["x"], sv["rn"], sv["alpha"], sv["se"], sv["QK"]
    grads = {}
    dQK = torch.zeros_like(QK)
    # interfacial coordinate update
    dx, dw = coord_step_bwd(x, row, col, vec_mul(alpha, se), sv["step"], None, cmax, dx_new)
    dalpha, dse = vec_mul(dw, se), vec_mul(dw, alpha)
    grads["ac2_w"] = colsum(sv["Tzc"] if "Tzc" in sv else act_fwd(sv["zc"], ACT_SILU), dse)
    dzc = outer_act_bwd(sv["zc"], dse, w["ac2_w"], ACT_SILU)
    grads["ac1_b"] = colsum(dzc)
    grads["ac_u"] = colsum(dzc, rn)
    drn = rowdot(dzc, w["ac_u"])
    scatter_add_rows(dzc, col, dQK, col0=3 * H + 128)
    # interfacial aggregation h3 = h2 + sum_e alpha_e (V[col] + rn v_r)
    dh2 = dh3.clone()
    dve = gather_rows(_drop(drop, dh3, layer, "agg"), row)            # h3 = h2 + drop(agg) (egnn.py:236)
    ve = rank1_add(gather_rows(QK, col, 2 * H + 128, H), rn, w["v_r"])
    vec_add_(dalpha, rowdot2(dve, ve))
    scale_rows(dve, alpha)
    scatter_add_rows(dve, col, dQK, col0=2 * H + 128)
    grads["v_r"] = colsum(dve, rn)
    vec_add_(drn, rowdot(dve, w["v_r"]))
    # segment softmax over the destination row, then the logits q . (k + rn k_r) + pair bias
    dlogit = softmax_seg_bwd(alpha, dalpha, row, N)
    dq = rank1_add(gather_rows(QK, col, H, H), rn, w["k_r"])          # kk
    scale_rows(dq, dlogit)
    scatter_add_rows(dq, row, dQK, col0=0)
    dkk = scale_rows(gather_rows(QK, row, 0, H), dlogit)
    scatter_add_rows(dkk, col, dQK, col0=H)
    grads["k_r"] = colsum(dkk, rn)
    vec_add_(drn, rowdot(dkk, w["k_r"]))
    # pair bias on the unique interface pairs
    P = dP0.shape[0]
    dpb_dense = torch.zeros(P, 1, dtype=torch.float32, device=dev)
    scatter_add_rows(dlogit.view(-1, 1), sv["pair"], dpb_dense)
    dpbu = gather_rows(dpb_dense, sv["u_pair"])                      # [U,1]
    grads["pt_c"] = colsum(dpbu)
    dpbu = dpbu.view(-1)
    grads["pt2v"] = colsum(sv["Tzp"] if "Tzp" in sv else act_fwd(sv["Zp"], ACT_RELU), dpbu)
    dZp = outer_act_bwd(sv["Zp"], dpbu, w["pt2v"], ACT_RELU)
    dz = _linear_bwd(grads, "pt1_w", "pt1_b", _pt1_padded(w, "pt1_w_t", H), sv["zcat"], dZp)
    if grads["pt1_w"].shape[1] != H + 64:
        grads["pt1_w"] = grads["pt1_w"][:, :H + 64].contiguous()
    scatter_add_rows(dz, sv["u_pair"], dP0, col0=0, width=H)
    U = dz.shape[0]
    ident_u = torch.arange(U, dtype=torch.int32, device=dev)
    dt = gather_rows(dz, ident_u, H, 32)
    a32, b32 = gather_rows(sv["pc32"], sv["u_pi"]), gather_rows(sv["pc32"], sv["u_ci"])
    dpc32 = torch.zeros(N, 32, dtype=torch.float32, device=dev)
    scatter_add_rows(vec_mul(dt, b32), sv["u_pi"], dpc32)
    scatter_add_rows(vec_mul(dt, a32), sv["u_ci"], dpc32)
    scatter_add_rows(dpc32[Nc:], torch.arange(Nc, N, dtype=torch.int32, device=dev), dQK, col0=2 * H)
    scatter_add_rows(dpc32[:Nc], torch.arange(0, Nc, dtype=torch.int32, device=dev), dQK, col0=2 * H + 32)
    radial_bwd(x, row, col, geo["node_cplx"], sv["nrm"], drn, dx)
    vec_add_(dh2, _linear_bwd(grads, "qk_w", "qk_b", w["qk_w_t"], sv["h2"], dQK))
    dhc, dhp = dh2[:Nc], dh2[Nc:]                                    # contiguous row slices, updated in place below
    # transitions h + linear_2(relu(linear_1(h)))
    dTc = _linear_bwd(grads, "tc2_w", "tc2_b", w["tc2_w_t"], sv["Ttc"], dhc)
    vec_add_(dhc, _linear_bwd(grads, "tc1_w", "tc1_b", w["tc1_w_t"], sv["hc1"], act_bwd(sv["Ttc"], dTc, ACT_RELU)))
    dTp = _linear_bwd(grads, "tp2_w", "tp2_b", w["tp2_w_t"], sv["Ttp"], dhp)
    vec_add_(dhp, _linear_bwd(grads, "tp1_w", "tp1_b", w["tp1_w_t"], sv["hp1"], act_bwd(sv["Ttp"], dTp, ACT_RELU)))
    # compound-side row attention (keys / values from the updated protein side)
    dOc = _linear_bwd(grads, "o_c_w", "o_c_b", w["o_c_w_t"], sv["Oc"], _drop(drop, dhc, layer, "catt"))
    dCAc = torch.zeros_like(sv["CAc"])
    dCAp2 = torch.zeros_like(sv["CAp2"])
    dPB_c = row_attention_bwd(geo, 0, (sv["CAc"], 2 * HD), (sv["CAc"], 3 * HD), (sv["CAp2"], 0), (sv["CAp2"], HD), sv["PB_c"], dOc,
                              (dCAc, 2 * HD), (dCAc, 3 * HD), (dCAp2, 0), (dCAp2, HD))
    vec_add_(dhp, _linear_bwd(grads, "ca_p2_w", None, w["ca_p2_w_t"], sv["hp1"], dCAp2))
    # protein-side row attention
    dOp = _linear_bwd(grads, "o_p_w", "o_p_b", w["o_p_w_t"], sv["Op"], _drop(drop, dhp, layer, "patt", Nc))
    dCAp = torch.zeros_like(sv["CAp"])
    dPB_p = row_attention_bwd(geo, 1, (sv["CAp"], 0), (sv["CAp"], HD), (sv["CAc"], 0), (sv["CAc"], HD), sv["PB_p"], dOp,
                              (dCAp, 0), (dCAp, HD), (dCAc, 0), (dCAc, HD))
    h_in = sv["h_in"]
    vec_add_(dhc, _linear_bwd(grads, "ca_c_w", "ca_c_b", w["ca_c_w_t"], h_in[:Nc], dCAc))
    vec_add_(dhp, _linear_bwd(grads, "ca_p_w", "ca_p_b", w["ca_p_w_t"], h_in[Nc:], dCAp))
    return dh2, dx, grads, dPB_p, dPB_c


def stack_backward_v1(weights, tape, top, geo, edges, consts, dH_out, dX_out, on_group=None):
    """Reverse pass of the LAST refinement iteration of the v1 stack (att_model.py:227-245: earlier iterations run under no_grad):
    linear_out <- out layer <- [LAS <- MC_Att_L <- MC_E_GCL] x L <- linear_in, plus pair_embed0 and the gated pair biases of every
    row-attention block.  Specification: tests/emulate_backward.py::forward_backward_v1.

    weights: {prefix: dict of packed-arena tensors (+ `_t` transposes)} for "" (top level), "gclI.", "attI.", "out.".
    tape: per layer (saved_gcl, saved_att, saved_las); top: Hin, pc, outer, P0, raw_full, h_last, out_saved (internal node order).
    edges: ctx_row/ctx_col, int_row/int_col, las_a/las_b (int32).  consts: cmax, lcl, las_step, xl (LAS reference coordinates).
    dH_out [N,H], dX_out [N,3]: gradients of the outputs in internal order (dX_out already masked to the moving nodes).
    on_group(prefix, {slot: gradient}): called as soon as a layer's weight gradients are complete (out layer first, the top-level
    slots last) -- the hook the overlapped gradient all-reduce hangs on.
    Returns (grads keyed by full slot name, dHin)."""
    L = len(tape)
    Nc = geo["Nc"]
    grads = {}
    drop = consts.get("drop")

    def take(pre, g):
        for k, v in g.items():
            grads[pre + k] = v
        if on_group is not None:        # this group's weight gradients are final: train.py starts their all-reduce behind the reverse pass
            on_group(pre, g)
    g0 = {}
    dh = _drop(drop, _linear_bwd(g0, "out_w", "out_b", weights[""]["out_w_t"], top.get("h_last_d", top["h_last"]), dH_out), -1, "stack_out")
    dh, dx, g = gcl_backward(weights["out."], top["out_saved"], edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], consts["cmax"], dh, dX_out,
                             drop, L)
    take("out.", g)
    P0 = top["P0"]
    dP0 = torch.zeros_like(P0)
    dPB = torch.zeros(P0.shape[0], 2 * L, 4, dtype=torch.float32, device=P0.device)
    for l in reversed(range(L)):
        s_gcl, s_att, s_las = tape[l]
        dx = las_bwd(s_las["x"], consts["xl"], edges["las_a"], edges["las_b"], s_las["acc"], consts["las_step"], consts["lcl"], dx)
        dh, dx, g, dPB_p, dPB_c = att_backward(weights[f"att{l}."], s_att, geo, edges["int_row"], edges["int_col"], consts["cmax"], dh, dx, dP0,
                                               drop, l)
        take(f"att{l}.", g)
        dPB[:, 2 * l].copy_(dPB_p)
        dPB[:, 2 * l + 1].copy_(dPB_c)
        dh, dx, g = gcl_backward(weights[f"gcl{l}."], s_gcl, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], consts["cmax"], dh, dx, drop, l)
        take(f"gcl{l}.", g)
    dHin = _linear_bwd(g0, "in_w", "in_b", weights[""]["in_w_t"], top["Hin"], _drop(drop, dh, -1, "stack_in"))
    draw = pair_bias_gate_bwd(top["raw_full"], dPB)
    vec_add_(dP0, _linear_bwd(g0, "pb_w", "pb_b", weights[""]["pb_w_t"], P0, draw))
    douter = _linear_bwd(g0, "il_o_w", "il_o_b", weights[""]["il_o_w_t"], top["outer"], dP0)
    dpc = pair_outer_bwd(douter, top["pc"], geo)
    vec_add_(dHin[:Nc], _linear_bwd(g0, "il_c_w", "il_c_b", weights[""]["il_c_w_t"], top["Hin"][:Nc], dpc[:Nc]))
    vec_add_(dHin[Nc:], _linear_bwd(g0, "il_p_w", "il_p_b", weights[""]["il_p_w_t"], top["Hin"][Nc:], dpc[Nc:]))
    take("", g0)
    return grads, dHin


# ------------------------------------------------------------------------------------------------------------------------
# Training-mode forward of the last refinement iteration (v1 layout): the same arithmetic as the inference launch sequence
# (csrc/forward.cu), un-fused where the reverse pass needs an intermediate, every saved tensor in the form gcl_backward /
# att_backward / las_bwd consume.  fp32; GEMMs through fb_gemm (SIMT parity kernel), row attention through the inference kernel.
# GPU parity: tests/test_gpu_train_forward.py; the orchestration is also validated on the CPU against the specification's forward
# (tests/test_backward_orchestration.py).
# ------------------------------------------------------------------------------------------------------------------------
def linear(A, W, bias=None, act=ACT_NONE, res=None, drop=None, A16=None):
    """drop(act(A W^T + bias)) + res  (fp32, fb_gemm); drop = (Drop, layer, site name, row0) or None; A16 = bf16 twin of A if its
    producer wrote one"""
    _chk(A), _chk(W, W.dtype if W.dtype == torch.bfloat16 else torch.float32)
    if bias is not None:
        _chk(bias)
    if res is not None:
        _chk(res)
    return _gemm_call(A, W, bias, act, res, A.shape[0], W.shape[0], A.shape[1], drop, A16)


FUSED_FORWARD = True     # training forward: fused edge kernels (False / CPU stand-in tests: the composition of the primitive wrappers)


def edge_pre_train(Pn, row, col, rn, w_rad, b1, act=ACT_SILU):
    """Z1[e] = Pn[row[e], :H] + Pn[col[e], H:] + rn[e] w_rad + b1, A1 = act(Z1) -> (Z1, A1, bf16 twin of A1 or None)   (egnn.py:75-81)"""
    E, H = row.numel(), Pn.shape[1] // 2
    if not (FUSED_FORWARD and Pn.is_cuda and H % 4 == 0):
        Z1 = gather_rows(Pn, row, 0, H)
        gather_add_rows(Pn, col, Z1, col0=H)
        rank1_add(Z1, rn, w_rad)
        rank1_add(Z1, _ones(E, Pn.device), b1)
        return Z1, act_fwd(Z1, act), None
    _chk(Pn), _chk(rn), _chk(w_rad), _chk(b1), _chk(row, torch.int32), _chk(col, torch.int32)
    Z1 = torch.empty(E, H, dtype=torch.float32, device=Pn.device)
    A1 = torch.empty_like(Z1)
    A16 = torch.empty(E, H, dtype=torch.bfloat16, device=Pn.device) if PRECISION == "bf16" else None
    _lib.check(_lib.lib().fb_edge_pre_train(Pn.data_ptr(), row.data_ptr(), col.data_ptr(), E, H, rn.data_ptr(), w_rad.data_ptr(),
                                            b1.data_ptr(), Z1.data_ptr(), A1.data_ptr(), A16.data_ptr() if A16 is not None else None,
                                            act, _st(Pn)), "fb_edge_pre_train")
    return Z1, A1, A16


def act_drop(Z, act, drop, layer, name):
    """drop(act(Z)) -> (fp32, bf16 twin or None): activation, the nn.Dropout behind it and the conversion for the next GEMM in one pass"""
    if not (FUSED_FORWARD and Z.is_cuda and Z.shape[1] % 4 == 0):
        return _drop(drop, act_fwd(Z, act), layer, name), None
    _chk(Z)
    M, N = Z.shape
    Y = torch.empty_like(Z)
    Y16 = torch.empty(M, N, dtype=torch.bfloat16, device=Z.device) if PRECISION == "bf16" else None
    on = drop is not None and drop.p > 0
    _lib.check(_lib.lib().fb_act_drop(Z.data_ptr(), M, N, act, drop.p if on else 0.0, drop.seed if on else 0, drop.site(layer, name) if on else 0,
                                      0, drop.colonly if on else 0, Y.data_ptr(), Y16.data_ptr() if Y16 is not None else None, _st(Z)),
               "fb_act_drop")
    return Y, Y16


def radial_fwd(x, row, col, node_cplx, B):
    """-> d [E,3], d2 [E], rn [E], nrm [B]"""
    _chk(x), _chk(row, torch.int32), _chk(col, torch.int32), _chk(node_cplx, torch.int32)
    E, dev = row.numel(), x.device
    S = torch.zeros(B, dtype=torch.float32, device=dev)
    d, d2, rn, nrm = (torch.empty(E, 3, dtype=torch.float32, device=dev), torch.empty(E, dtype=torch.float32, device=dev),
                      torch.empty(E, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.float32, device=dev))
    _lib.check(_lib.lib().fb_radial_fwd(x.data_ptr(), row.data_ptr(), col.data_ptr(), E, node_cplx.data_ptr(), B, S.data_ptr(), d.data_ptr(),
                                        d2.data_ptr(), rn.data_ptr(), nrm.data_ptr(), _st(x)), "fb_radial_fwd")
    return d, d2, rn, nrm


def coord_apply(x, ssum, cnt, cmax):
    """-> (unclamped step, x + clamp(step))"""
    _chk(x), _chk(ssum)
    step, x_new = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().fb_coord_apply(x.data_ptr(), ssum.data_ptr(), _chk(cnt).data_ptr() if cnt is not None else None, x.shape[0],
                                         float(cmax), step.data_ptr(), x_new.data_ptr(), _st(x)), "fb_coord_apply")
    return step, x_new


def softmax_seg_fwd(logit, rowptr, n_rows):
    _chk(logit), _chk(rowptr, torch.int32)
    alpha = torch.empty_like(logit)
    _lib.check(_lib.lib().fb_softmax_seg_fwd(logit.data_ptr(), rowptr.data_ptr(), n_rows, alpha.data_ptr(), _st(logit)), "fb_softmax_seg_fwd")
    return alpha


def las_acc(x, xref, a_idx, b_idx, step_size):
    _chk(x), _chk(xref), _chk(a_idx, torch.int32), _chk(b_idx, torch.int32)
    acc = torch.zeros_like(x)
    _lib.check(_lib.lib().fb_las_acc(x.data_ptr(), xref.data_ptr(), a_idx.data_ptr(), b_idx.data_ptr(), a_idx.numel(), float(step_size),
                                     acc.data_ptr(), _st(x)), "fb_las_acc")
    return acc


def pair_outer_fwd(pc, geo, n_pairs):
    _chk(pc)
    N, H = pc.shape
    outer = torch.empty(n_pairs, H, dtype=torch.float32, device=pc.device)
    _lib.check(_lib.lib().fb_pair_outer_fwd(pc.data_ptr(), H, geo["c_off"].data_ptr(), geo["p_off"].data_ptr(), geo["pair_base"].data_ptr(),
                                            geo["node_cplx"].data_ptr(), geo["Nc"], N - geo["Nc"], outer.data_ptr(), _st(pc)),
               "fb_pair_outer_fwd")
    return outer


def pair_bias_gate_fwd(raw, nblk):
    _chk(raw)
    PB = torch.empty(raw.shape[0], nblk, 4, dtype=torch.float32, device=raw.device)
    _lib.check(_lib.lib().fb_pair_bias_gate_fwd(raw.data_ptr(), raw.shape[1], raw.shape[0], nblk, PB.data_ptr(), _st(raw)),
               "fb_pair_bias_gate_fwd")
    return PB


def row_attention_fwd(geo, q_is_prot, Q, G, K, V, PB, n_q_rows):
    """(tensor, first column) operands as in row_attention_bwd -> O [n_q_rows, 128]"""
    Nc = geo["Nc"]

    def ptr(tc, prot_side):
        t, c0 = tc
        _chk(t)
        return t.data_ptr() + 4 * c0 - (4 * Nc * t.shape[1] if prot_side else 0), t.shape[1]
    qs = bool(q_is_prot)
    (q, ldq), (g, ldg), (k, ldk), (v, ldv) = ptr(Q, qs), ptr(G, qs), ptr(K, not qs), ptr(V, not qs)
    _chk(PB)
    O = torch.empty(n_q_rows, HD, dtype=torch.float32, device=PB.device)
    o = O.data_ptr() - (4 * Nc * HD if qs else 0)
    max_q, max_k = (geo["max_p"], geo["max_c"]) if qs else (geo["max_c"], geo["max_p"])
    _lib.check(_lib.lib().fb_row_attention_fwd(geo["c_off"].data_ptr(), geo["p_off"].data_ptr(), geo["pair_base"].data_ptr(), geo["B"],
                                               int(qs), max_q, max_k, q, ldq, g, ldg, k, ldk, v, ldv, PB.data_ptr(), o, HD, _st(PB)),
               "fb_row_attention_fwd")
    return O


def _ones(n, dev):
    return torch.ones(n, dtype=torch.float32, device=dev)


def gcl_forward_train(w, h, x, row, col, node_cplx, B, cmax, drop=None, layer=0):
    """MC_E_GCL forward (egnn.py:68-144; inference twin: forward.cu::run_gcl) keeping what gcl_backward consumes"""
    N, H = h.shape
    E, dev = row.numel(), h.device
    d, d2, rn, nrm = radial_fwd(x, row, col, node_cplx, B)
    Pn = linear(h, w["e1_rc"])
    # the activations the reverse pass multiplies with (A1, M, T3, t1) are kept next to their pre-activations: recomputing them cost four
    # passes over [E, H] per sub-layer (35 us each at B = 16)
    Z1, A1, A1h = edge_pre_train(Pn, row, col, rn, w["e1_rad"], w["e1_b"])
    Z2 = linear(A1, w["e2_w"], w["e2_b"], A16=A1h)
    M, Mh = act_drop(Z2, ACT_SILU, drop, layer, "edge2")              # egnn.py:82
    Z3 = linear(M, w["c1_w"], w["c1_b"], A16=Mh)
    T3 = act_fwd(Z3, ACT_SILU)
    s = rowdot(T3, w["c2_w"])
    ssum = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    scatter_add_rows(scale_rows(d.clone(), s), row, ssum)
    deg = torch.zeros(N, 1, dtype=torch.float32, device=dev)
    scatter_add_rows(_ones(E, dev).view(E, 1), row, deg)
    deg = deg.view(N)
    step, x_new = coord_apply(x, ssum, deg, cmax)
    agg = torch.zeros(N, H, dtype=torch.float32, device=dev)
    scatter_add_rows(M, row, agg)
    cat = torch.empty(N, 2 * H, dtype=torch.float32, device=dev)
    cat[:, :H].copy_(h)
    cat[:, H:].copy_(agg)
    Z4 = linear(cat, w["n1_w"], w["n1_b"])
    t1 = act_fwd(Z4, ACT_SILU)
    h_new = linear(t1, w["n2_w"], w["n2_b"], res=h, drop=(drop, layer, "node2", 0))   # egnn.py:106
    return h_new, x_new, dict(h=h, x=x, rn=rn, nrm=nrm, Z1=Z1, Z2=Z2, Z3=Z3, s=s, deg=deg, step=step, agg=agg, Z4=Z4,
                              A1=A1, M=M, T3=T3, t1=t1)


def interface_indices(row, col, geo):
    """pair row of every interface edge and the lig->prot subset (index bookkeeping with integer tensor ops; no float arithmetic)"""
    Nc = geo["Nc"]
    r, c = row.long(), col.long()
    cplx = geo["node_cplx"].long()
    eb = cplx[r]
    is_c = r < Nc
    ci, pi = torch.where(is_c, r, c), torch.where(is_c, c, r)
    c_off, p_off, pair_base = geo["c_off"].long(), geo["p_off"].long(), geo["pair_base"].long()
    nc1 = c_off[1:] - c_off[:-1]
    pair = pair_base[eb] + (pi - p_off[eb]) * nc1[eb] + (ci - c_off[eb])
    u = is_c.nonzero().squeeze(1)
    i32 = lambda t: t.to(torch.int32).contiguous()
    n_rows = cplx.numel()
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=row.device)
    rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=n_rows), 0)
    return dict(pair=i32(pair), u_pair=i32(pair[u]), u_pi=i32(pi[u]), u_ci=i32(ci[u]), rowptr=i32(rowptr))


def _pt1_cols(H):
    return H + 128 if PRECISION == "bf16" else H + 64


def _pt1_padded(w, name, H):
    """pt1_w [2H, H+64] (or its transpose [H+64, 2H]) zero-padded to H+128 input columns under PRECISION = 'bf16'"""
    t = w[name]
    Kp = _pt1_cols(H)
    if Kp == H + 64:
        return t
    if name.endswith("_t"):
        out = torch.zeros(Kp, t.shape[1], dtype=t.dtype, device=t.device)
        out[:H + 64].copy_(t)
    else:
        out = torch.zeros(t.shape[0], Kp, dtype=t.dtype, device=t.device)
        out[:, :H + 64].copy_(t)
    return out


def att_forward_train(w, h, x, geo, row, col, idx, P0, PB_p, PB_c, cmax, drop=None, layer=0):
    """MC_Att_L forward (egnn.py:186-333, cross_att.py:24-54; inference twin: forward.cu::run_att) keeping what att_backward consumes.
    row / col must be sorted by row (the library's interface graph is); idx = interface_indices(row, col, geo)."""
    N, H = h.shape
    Nc, dev, E = geo["Nc"], h.device, row.numel()
    hc0, hp0 = h[:Nc], h[Nc:]
    CAc = linear(hc0, w["ca_c_w"], w["ca_c_b"])
    CAp = linear(hp0, w["ca_p_w"], w["ca_p_b"])
    Op = row_attention_fwd(geo, 1, (CAp, 0), (CAp, HD), (CAc, 0), (CAc, HD), PB_p, N - Nc)
    hp1 = linear(Op, w["o_p_w"], w["o_p_b"], res=hp0, drop=(drop, layer, "patt", Nc))       # cross_att.py:128
    CAp2 = linear(hp1, w["ca_p2_w"])
    Oc = row_attention_fwd(geo, 0, (CAc, 2 * HD), (CAc, 3 * HD), (CAp2, 0), (CAp2, HD), PB_c, Nc)
    hc1 = linear(Oc, w["o_c_w"], w["o_c_b"], res=hc0, drop=(drop, layer, "catt", 0))
    Ttp = linear(hp1, w["tp1_w"], w["tp1_b"], act=ACT_RELU)
    Ttc = linear(hc1, w["tc1_w"], w["tc1_b"], act=ACT_RELU)
    h2 = torch.empty(N, H, dtype=torch.float32, device=dev)
    h2[:Nc].copy_(linear(Ttc, w["tc2_w"], w["tc2_b"], res=hc1))
    h2[Nc:].copy_(linear(Ttp, w["tp2_w"], w["tp2_b"], res=hp1))
    QK = linear(h2, w["qk_w"], w["qk_b"])
    pc32 = torch.empty(N, 32, dtype=torch.float32, device=dev)
    pc32[Nc:].copy_(gather_rows(QK, torch.arange(Nc, N, dtype=torch.int32, device=dev), 2 * H, 32))
    pc32[:Nc].copy_(gather_rows(QK, torch.arange(0, Nc, dtype=torch.int32, device=dev), 2 * H + 32, 32))
    U = idx["u_pair"].numel()
    # [pair0 | t32 | 0]: H + 64 columns; under PRECISION = "bf16" padded to H + 128 so that the weight-gradient and data-gradient
    # GEMMs of pair_transition.linear_1 tile on tcgen05 (N = K_in must be a multiple of 128)
    Kp = _pt1_cols(H)
    zcat = torch.zeros(U, Kp, dtype=torch.float32, device=dev)
    zcat[:, :H].copy_(gather_rows(P0, idx["u_pair"]))
    zcat[:, H:H + 32].copy_(vec_mul(gather_rows(pc32, idx["u_pi"]), gather_rows(pc32, idx["u_ci"])))
    Zp = linear(zcat, _pt1_padded(w, "pt1_w", H), w["pt1_b"])
    Tzp = act_fwd(Zp, ACT_RELU)
    pbu = rowdot(Tzp, w["pt2v"]).view(U, 1)
    rank1_add(pbu, _ones(U, dev), w["pt_c"])
    pb_dense = torch.zeros(P0.shape[0], 1, dtype=torch.float32, device=dev)
    scatter_add_rows(pbu, idx["u_pair"], pb_dense)                      # unique pair rows: the sum is an assignment
    d, d2, rn, nrm = radial_fwd(x, row, col, geo["node_cplx"], geo["B"])
    logit = rowdot2(gather_rows(QK, row, 0, H), rank1_add(gather_rows(QK, col, H, H), rn, w["k_r"]))
    vec_add_(logit, gather_rows(pb_dense, idx["pair"]).view(E))
    alpha = softmax_seg_fwd(logit, idx["rowptr"], N)
    ve = scale_rows(rank1_add(gather_rows(QK, col, 2 * H + 128, H), rn, w["v_r"]), alpha)
    if drop is not None and drop.p > 0:                                  # h3 = h2 + drop(agg)  (egnn.py:235-237)
        agg = torch.zeros_like(h2)
        scatter_add_rows(ve, row, agg)
        h3 = vec_add_(drop.apply(agg, layer, "agg"), h2)
    else:
        h3 = h2.clone()
        scatter_add_rows(ve, row, h3)
    zc = rank1_add(gather_rows(QK, col, 3 * H + 128, H), rn, w["ac_u"])
    rank1_add(zc, _ones(E, dev), w["ac1_b"])
    Tzc = act_fwd(zc, ACT_SILU)
    se = rowdot(Tzc, w["ac2_w"])
    ssum = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    scatter_add_rows(scale_rows(d.clone(), vec_mul(alpha, se)), row, ssum)
    step, x_new = coord_apply(x, ssum, None, cmax)
    sv = dict(Tzc=Tzc, Tzp=Tzp, h_in=h, x=x, CAc=CAc, CAp=CAp, CAp2=CAp2, PB_p=PB_p, PB_c=PB_c, Op=Op, Oc=Oc, hp1=hp1, hc1=hc1, Ttp=Ttp, Ttc=Ttc, h2=h2, QK=QK,
              pc32=pc32, pair=idx["pair"], u_pair=idx["u_pair"], u_pi=idx["u_pi"], u_ci=idx["u_ci"], zcat=zcat, Zp=Zp, rn=rn, nrm=nrm,
              alpha=alpha, se=se, zc=zc, step=step)
    return h3, x_new, sv


def stack_forward_train_v1(weights, Hin, x_state, moves, geo, edges, consts, n_layers):
    """Last refinement iteration of the v1 stack in training mode (internal node order): returns (X_out [N,3], H_out [N,H], tape, top)
    in the form stack_backward_v1 consumes.  x_state: coordinates after the earlier (no_grad) iterations; moves: bool [N]."""
    N, H = Hin.shape
    Nc, dev = geo["Nc"], Hin.device
    wt = weights[""]
    pc = torch.empty(N, H, dtype=torch.float32, device=dev)
    pc[:Nc].copy_(linear(Hin[:Nc], wt["il_c_w"], wt["il_c_b"]))
    pc[Nc:].copy_(linear(Hin[Nc:], wt["il_p_w"], wt["il_p_b"]))
    outer = pair_outer_fwd(pc, geo, consts["n_pairs"])
    P0 = linear(outer, wt["il_o_w"], wt["il_o_b"])
    raw_full = linear(P0, wt["pb_w"], wt["pb_b"])
    PB = pair_bias_gate_fwd(raw_full, 2 * n_layers)
    drop = consts.get("drop")
    h = linear(Hin, wt["in_w"], wt["in_b"], drop=(drop, -1, "stack_in", 0))         # egnn.py:398
    x = x_state
    idx = interface_indices(edges["int_row"], edges["int_col"], geo)
    tape = []
    for l in range(n_layers):
        h, x, s_gcl = gcl_forward_train(weights[f"gcl{l}."], h, x, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], geo["B"], consts["cmax"],
                                        drop, l)
        h, x, s_att = att_forward_train(weights[f"att{l}."], h, x, geo, edges["int_row"], edges["int_col"], idx, P0,
                                        PB[:, 2 * l].contiguous(), PB[:, 2 * l + 1].contiguous(), consts["cmax"], drop, l)
        acc = las_acc(x, consts["xl"], edges["las_a"], edges["las_b"], consts["las_step"])
        x_in = x
        _, x = coord_apply(x_in, acc, None, consts["lcl"])
        tape.append((s_gcl, s_att, dict(x=x_in, acc=acc)))
    h_last, x, s_out = gcl_forward_train(weights["out."], h, x, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], geo["B"], consts["cmax"],
                                         drop, n_layers)
    h_last_d = _drop(drop, h_last, -1, "stack_out")                                 # egnn.py:461
    H_out = linear(h_last_d, wt["out_w"], wt["out_b"])
    X_out = torch.where(moves[:, None], x, x_state)
    top = dict(Hin=Hin, pc=pc, outer=outer, P0=P0, raw_full=raw_full, h_last=h_last, h_last_d=h_last_d, out_saved=s_out)
    return X_out, H_out, tape, top


# ------------------------------------------------------------------------------------------------------------------------
# FABind+ layout (LayerNorm MLPs, LayerNorm folded through the node-level hoisting, propagated pair embedding): reverse pass.
# Specification: tests/emulate_backward.py::{gcl_plus_bwd, att_plus_bwd, forward_backward_plus}.  Orchestration validated on the
# CPU; the kernels of this section have not run on a GPU yet (gated tests).
# ------------------------------------------------------------------------------------------------------------------------
LN_EPS = 1e-5


def layernorm(x, gamma, beta):
    _chk(x), _chk(gamma), _chk(beta)
    out = torch.empty_like(x)
    _lib.check(_lib.lib().fb_layernorm(x.data_ptr(), x.shape[0], x.shape[1], gamma.data_ptr(), beta.data_ptr(), LN_EPS, out.data_ptr(), _st(x)),
               "fb_layernorm")
    return out


def layernorm_bwd(grads, gname, bname, x, gamma, dy):
    """accumulates dgamma, dbeta; returns dx"""
    _chk(x), _chk(gamma), _chk(dy)
    dx, xhat = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().fb_layernorm_bwd(x.data_ptr(), gamma.data_ptr(), dy.data_ptr(), x.shape[0], x.shape[1], LN_EPS, dx.data_ptr(),
                                           xhat.data_ptr(), _st(x)), "fb_layernorm_bwd")
    grads[gname] = colsum(vec_mul(dy, xhat), None, grads.get(gname))
    grads[bname] = colsum(dy, None, grads.get(bname))
    return dx


def row_stats_bwd(h, w, ds1, ds2, ds3, dh):
    """dh[m,:] += ds1[m] + 2 h[m,:] ds2[m] + w ds3[m]"""
    _chk(h), _chk(dh)
    p = lambda t: _chk(t).data_ptr() if t is not None else None
    _lib.check(_lib.lib().fb_row_stats_bwd(h.data_ptr(), h.shape[1], h.shape[0], h.shape[1], p(w), p(ds1), p(ds2), p(ds3), dh.data_ptr(),
                                           dh.shape[1], _st(h)), "fb_row_stats_bwd")
    return dh


def folded_stats_bwd(A3, rn, a0, a1, D, mu, var_raw, rstd, drstd, dmu, drn, want_da):
    """-> dA1, dA2, dA3 (None without A3), da [2] (None unless want_da); drn accumulated in place"""
    E, dev = rn.numel(), rn.device
    p = lambda t: _chk(t).data_ptr() if t is not None else None
    dA1, dA2 = torch.empty_like(rn), torch.empty_like(rn)
    dA3 = torch.empty_like(rn) if A3 is not None else None
    da = torch.zeros(2, dtype=torch.float32, device=dev) if want_da else None
    _lib.check(_lib.lib().fb_folded_stats_bwd(p(A3), rn.data_ptr(), float(a0), float(a1), float(D), E, mu.data_ptr(), var_raw.data_ptr(),
                                              rstd.data_ptr(), drstd.data_ptr(), dmu.data_ptr(), dA1.data_ptr(), dA2.data_ptr(), p(dA3),
                                              drn.data_ptr(), p(da), _st(rn)), "fb_folded_stats_bwd")
    return dA1, dA2, dA3, da


def _neg(v):
    return vec_mul(v, torch.full_like(v, -1.0))


def _scatter_vec(v, idx, n):
    out = torch.zeros(n, 1, dtype=torch.float32, device=v.device)
    scatter_add_rows(v.view(-1, 1), idx, out)
    return out


def gcl_plus_backward(w, sv, row, col, node_cplx, cmax, dh_new, dx_new):
    """Reverse of one FABind+ MC_E_GCL (P/models/egnn.py:44-115).  sv: h, x, rn, nrm, mu, var_raw, rstd [E], U [E,Dp] (the un-normalised
    folded first Linear), Z2, Z3 [E,H] and Z4, Z5 [N,H] (pre-activations; post-ReLU values serve equally), s, deg, step, agg."""
    h, x, rn, mu, rstd = sv["h"], sv["x"], sv["rn"], sv["mu"], sv["rstd"]
    N, H = h.shape
    E, dev = row.numel(), h.device
    Dp, D = sv["U"].shape[1], 2 * H + 1
    grads = {}
    dx, ds = coord_step_bwd(x, row, col, sv["s"], sv["step"], sv["deg"], cmax, dx_new)
    grads["c2_w"] = colsum(act_fwd(sv["Z3"], ACT_RELU), ds)
    dZ3 = outer_act_bwd(sv["Z3"], ds, w["c2_w"], ACT_RELU)
    M = act_fwd(sv["Z2"], ACT_RELU)
    dM2 = _linear_bwd(grads, "c1_w", "c1_b", w["c1_w_t"], layernorm(M, w["cl_g"], w["cl_b"]), dZ3)
    dM = layernorm_bwd(grads, "cl_g", "cl_b", M, w["cl_g"], dM2)
    dt1 = _linear_bwd(grads, "n2_w", "n2_b", w["n2_w_t"], act_fwd(sv["Z4"], ACT_RELU), act_bwd(sv["Z5"], dh_new, ACT_RELU))
    cat = torch.empty(N, 2 * H, dtype=torch.float32, device=dev)
    cat[:, :H].copy_(h)
    cat[:, H:].copy_(sv["agg"])
    dt0 = _linear_bwd(grads, "n1_w", "n1_b", w["n1_w_t"], layernorm(cat, w["nl_g"], w["nl_b"]), act_bwd(sv["Z4"], dt1, ACT_RELU))
    dcat = layernorm_bwd(grads, "nl_g", "nl_b", cat, w["nl_g"], dt0)
    dh = dh_new.clone()
    gather_add_rows(dcat, torch.arange(N, dtype=torch.int32, device=dev), dh, col0=0)
    gather_add_rows(dcat, row, dM, col0=H)
    Z1 = rank1_add(scale_rows(sv["U"].clone(), rstd), _ones(E, dev), w["e1_c0"])
    dA1 = _linear_bwd(grads, "e2_w", "e2_b", w["e2_w_t"], act_fwd(Z1, ACT_RELU), act_bwd(sv["Z2"], dM, ACT_RELU))
    dZ1 = act_bwd(Z1, dA1, ACT_RELU)
    grads["e1_c0"] = colsum(dZ1)
    drstd = rowdot2(dZ1, sv["U"])
    dU = scale_rows(dZ1, rstd)
    grads["e1_rad"] = colsum(dU, rn)
    grads["e1_g"] = colsum(dU, _neg(mu))
    drn = rowdot(dU, w["e1_rad"])
    dmu = _neg(rowdot(dU, w["e1_g"]))
    dPn = torch.zeros(N, 2 * Dp, dtype=torch.float32, device=dev)
    scatter_add_rows(dU, row, dPn, col0=0)
    scatter_add_rows(dU, col, dPn, col0=Dp)
    vec_add_(dh, _linear_bwd(grads, "e1_rc", None, w["e1_rc_t"], h, dPn))
    dA1s, dA2s, _, _ = folded_stats_bwd(None, rn, 1.0, 1.0, D, mu, sv["var_raw"], rstd, drstd, dmu, drn, False)
    ds1 = vec_add_(_scatter_vec(dA1s, row, N), _scatter_vec(dA1s, col, N)).view(N)
    ds2 = vec_add_(_scatter_vec(dA2s, row, N), _scatter_vec(dA2s, col, N)).view(N)
    row_stats_bwd(h, None, ds1, ds2, None, dh)
    radial_bwd(x, row, col, node_cplx, sv["nrm"], drn, dx)
    return dh, dx, grads


def att_plus_backward(w, sv, geo, row, col, cmax, dh3, dx_new, dpair_out):
    """Reverse of one FABind+ MC_Att_L (P/models/egnn.py:129-300, P/models/cross_att.py:20-45): as att_backward, plus the LayerNorm
    transitions, the LayerNorm-folded coordinate head and the pair transition on EVERY pair row (the embedding is propagated:
    dpair_out comes from the next layer / the loss, dpair_in goes to the previous one).
    sv (beyond att_backward's): raw_full [P,ld], pair_in, Zpre, Zh, Zo [P,H] (pre-LayerNorm sum; hidden and output, post-ReLU), t32, a32,
    b32 [P,32], pi_all, ci_all [P] int32, Tc1/Tc2/Tp1/Tp2 (transition hidden / output, post-ReLU), s3 [N], mu, var_raw, rstd [E], Uc [E,H].
    Returns (dh, dx, grads, dpair_in)."""
    Nc, N = geo["Nc"], sv["h2"].shape[0]
    H = sv["h2"].shape[1]
    dev, E = dh3.device, row.numel()
    x, rn, alpha, se, QK, mu, rstd = sv["x"], sv["rn"], sv["alpha"], sv["se"], sv["QK"], sv["mu"], sv["rstd"]
    grads = {}
    dQK = torch.zeros_like(QK)
    ident = torch.arange(N, dtype=torch.int32, device=dev)
    dx, dw = coord_step_bwd(x, row, col, vec_mul(alpha, se), sv["step"], None, cmax, dx_new)
    dalpha, dse = vec_mul(dw, se), vec_mul(dw, alpha)
    # LayerNorm-folded coordinate head
    tco = rank1_add(scale_rows(sv["Uc"].clone(), rstd), _ones(E, dev), w["ac_c0"])
    grads["ac2_w"] = colsum(act_fwd(tco, ACT_RELU), dse)
    dt = outer_act_bwd(tco, dse, w["ac2_w"], ACT_RELU)
    grads["ac_c0"] = colsum(dt)
    drstd = rowdot2(dt, sv["Uc"])
    dUc = scale_rows(dt, rstd)
    scatter_add_rows(dUc, col, dQK, col0=3 * H + 128)
    grads["ac_u"] = colsum(dUc, rn)
    grads["ac_g"] = colsum(dUc, _neg(mu))
    drn = rowdot(dUc, w["ac_u"])
    dmu = _neg(rowdot(dUc, w["ac_g"]))
    acr = sv["acr"]                                               # (sum v_r, sum v_r^2) as python floats
    A3 = gather_rows(sv["s3"].view(N, 1), col).view(E)
    dA1, dA2, dA3, da = folded_stats_bwd(A3, rn, acr[0], acr[1], H, mu, sv["var_raw"], rstd, drstd, dmu, drn, True)
    grads["ac_r"] = da
    ds1, ds2, ds3 = _scatter_vec(dA1, col, N).view(N), _scatter_vec(dA2, col, N).view(N), _scatter_vec(dA3, col, N).view(N)
    V = gather_rows(QK, ident, 2 * H + 128, H)
    dV = row_stats_bwd(V, w["v_r"], ds1, ds2, ds3, torch.zeros(N, H, dtype=torch.float32, device=dev))
    grads["v_r"] = colsum(V, ds3)
    # aggregation, segment softmax, logits (as in the v1 layout)
    dh2 = dh3.clone()
    dve = gather_rows(dh3, row)
    ve = rank1_add(gather_rows(QK, col, 2 * H + 128, H), rn, w["v_r"])
    vec_add_(dalpha, rowdot2(dve, ve))
    scale_rows(dve, alpha)
    scatter_add_rows(dve, col, dV)
    scatter_add_rows(dV, ident, dQK, col0=2 * H + 128)
    grads["v_r"] = colsum(dve, rn, grads["v_r"])
    vec_add_(drn, rowdot(dve, w["v_r"]))
    dlogit = softmax_seg_bwd(alpha, dalpha, row, N)
    dq = scale_rows(rank1_add(gather_rows(QK, col, H, H), rn, w["k_r"]), dlogit)
    scatter_add_rows(dq, row, dQK, col0=0)
    dkk = scale_rows(gather_rows(QK, row, 0, H), dlogit)
    scatter_add_rows(dkk, col, dQK, col0=H)
    grads["k_r"] = colsum(dkk, rn)
    vec_add_(drn, rowdot(dkk, w["k_r"]))
    radial_bwd(x, row, col, geo["node_cplx"], sv["nrm"], drn, dx)
    # pair transition on every pair row, attn_bias_proj as its row-dot
    P = sv["pair_in"].shape[0]
    dpb = torch.zeros(P, 1, dtype=torch.float32, device=dev)
    scatter_add_rows(dlogit.view(-1, 1), sv["pair"], dpb)
    grads["pt_c"] = colsum(dpb)
    dpb = dpb.view(P)
    grads["wb"] = colsum(sv["Zo"], dpb)
    dpo = rank1_add(dpair_out.clone(), dpb, w["wb"])
    dZh = _linear_bwd(grads, "pt2_w", "pt2_b", w["pt2_w_t"], sv["Zh"], act_bwd(sv["Zo"], dpo, ACT_RELU))
    dZl = _linear_bwd(grads, "pt1_w", "pt1_b", w["pt1_w_t"], layernorm(sv["Zpre"], w["zl_g"], w["zl_b"]), act_bwd(sv["Zh"], dZh, ACT_RELU))
    dZpre = layernorm_bwd(grads, "zl_g", "zl_b", sv["Zpre"], w["zl_g"], dZl)
    dpair_in = dZpre.clone()
    grads["zo_b"] = colsum(dZpre)
    grads["zo_w"] = gemm_wgrad(sv["t32"], dZpre)
    dt32 = gemm_dgrad(dZpre, w["zo_w"])
    scatter_add_rows(vec_mul(dt32, sv["b32"]), sv["pi_all"], dQK, col0=2 * H)
    scatter_add_rows(vec_mul(dt32, sv["a32"]), sv["ci_all"], dQK, col0=2 * H + 32)
    vec_add_(dh2, _linear_bwd(grads, "qk_w", "qk_b", w["qk_w_t"], sv["h2"], dQK))
    dhc, dhp = dh2[:Nc], dh2[Nc:]
    # transitions hs + relu(linear2(relu(linear1(LN(hs)))))
    for t, dhs, hs, T1, T2 in (("tc", dhc, sv["hc1"], sv["Tc1"], sv["Tc2"]), ("tp", dhp, sv["hp1"], sv["Tp1"], sv["Tp2"])):
        dT1 = _linear_bwd(grads, t + "2_w", t + "2_b", w[t + "2_w_t"], T1, act_bwd(T2, dhs, ACT_RELU))
        dt0 = _linear_bwd(grads, t + "1_w", t + "1_b", w[t + "1_w_t"], layernorm(hs, w[t + "l_g"], w[t + "l_b"]), act_bwd(T1, dT1, ACT_RELU))
        vec_add_(dhs, layernorm_bwd(grads, t + "l_g", t + "l_b", hs, w[t + "l_g"], dt0))
    dOc = _linear_bwd(grads, "o_c_w", "o_c_b", w["o_c_w_t"], sv["Oc"], dhc)
    dCAc = torch.zeros_like(sv["CAc"])
    dCAp2 = torch.zeros_like(sv["CAp2"])
    dPB_c = row_attention_bwd(geo, 0, (sv["CAc"], 2 * HD), (sv["CAc"], 3 * HD), (sv["CAp2"], 0), (sv["CAp2"], HD), sv["PB_c"], dOc,
                              (dCAc, 2 * HD), (dCAc, 3 * HD), (dCAp2, 0), (dCAp2, HD))
    vec_add_(dhp, _linear_bwd(grads, "ca_p2_w", None, w["ca_p2_w_t"], sv["hp1"], dCAp2))
    dOp = _linear_bwd(grads, "o_p_w", "o_p_b", w["o_p_w_t"], sv["Op"], dhp)
    dCAp = torch.zeros_like(sv["CAp"])
    dPB_p = row_attention_bwd(geo, 1, (sv["CAp"], 0), (sv["CAp"], HD), (sv["CAc"], 0), (sv["CAc"], HD), sv["PB_p"], dOp,
                              (dCAp, 0), (dCAp, HD), (dCAc, 0), (dCAc, HD))
    h_in = sv["h_in"]
    vec_add_(dhc, _linear_bwd(grads, "ca_c_w", "ca_c_b", w["ca_c_w_t"], h_in[:Nc], dCAc))
    vec_add_(dhp, _linear_bwd(grads, "ca_p_w", "ca_p_b", w["ca_p_w_t"], h_in[Nc:], dCAp))
    # gated pair biases of the layer's two row-attention blocks read the INCOMING pair embedding
    dPB = torch.empty(P, 2, 4, dtype=torch.float32, device=dev)
    dPB[:, 0].copy_(dPB_p)
    dPB[:, 1].copy_(dPB_c)
    vec_add_(dpair_in, _linear_bwd(grads, "pb_w", "pb_b", w["pb_w_t"], sv["pair_in"], pair_bias_gate_bwd(sv["raw_full"], dPB)))
    return dh2, dx, grads, dpair_in


def stack_backward_plus(weights, tape, top, geo, edges, consts, dH_out, dX_out, dP_out):
    """Reverse pass of the last refinement iteration of the FABind+ stack (P/models/att_model.py:166-223): as stack_backward_v1, with
    the pair embedding propagated layer to layer (dP_out = gradient of the returned pair embedding, packed rows) down to pair_embed0.
    Specification: tests/emulate_backward.py::forward_backward_plus.  Returns (grads keyed by full slot name, dHin)."""
    L = len(tape)
    Nc = geo["Nc"]
    grads, g0 = {}, {}

    def take(pre, g):
        for k, v in g.items():
            grads[pre + k] = v
    dh = _linear_bwd(g0, "out_w", "out_b", weights[""]["out_w_t"], top["h_last"], dH_out)
    dh, dx, g = gcl_plus_backward(weights["out."], top["out_saved"], edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], consts["cmax"], dh, dX_out)
    take("out.", g)
    dpair = dP_out.clone()
    for l in reversed(range(L)):
        s_gcl, s_att, s_las = tape[l]
        dx = las_bwd(s_las["x"], consts["xl"], edges["las_a"], edges["las_b"], s_las["acc"], consts["las_step"], consts["lcl"], dx)
        dh, dx, g, dpair = att_plus_backward(weights[f"att{l}."], s_att, geo, edges["int_row"], edges["int_col"], consts["cmax"], dh, dx, dpair)
        take(f"att{l}.", g)
        dh, dx, g = gcl_plus_backward(weights[f"gcl{l}."], s_gcl, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], consts["cmax"], dh, dx)
        take(f"gcl{l}.", g)
    dHin = _linear_bwd(g0, "in_w", "in_b", weights[""]["in_w_t"], top["Hin"], dh)
    douter = _linear_bwd(g0, "il_o_w", "il_o_b", weights[""]["il_o_w_t"], top["outer"], dpair)
    dpc = pair_outer_bwd(douter, top["pc"], geo)
    vec_add_(dHin[:Nc], _linear_bwd(g0, "il_c_w", "il_c_b", weights[""]["il_c_w_t"], top["Hin"][:Nc], dpc[:Nc]))
    vec_add_(dHin[Nc:], _linear_bwd(g0, "il_p_w", "il_p_b", weights[""]["il_p_w_t"], top["Hin"][Nc:], dpc[Nc:]))
    take("", g0)
    return grads, dHin


# ---- FABind+ training-mode forward of the last iteration (kernels not yet run on a GPU; orchestration CPU-validated) ----------
def row_stats(h, w=None):
    """-> s1 = sum_f h, s2 = sum_f h^2, s3 = sum_f h w (None without w)"""
    _chk(h)
    M, dev = h.shape[0], h.device
    s1, s2 = torch.empty(M, dtype=torch.float32, device=dev), torch.empty(M, dtype=torch.float32, device=dev)
    s3 = torch.empty(M, dtype=torch.float32, device=dev) if w is not None else None
    p = lambda t: _chk(t).data_ptr() if t is not None else None
    _lib.check(_lib.lib().fb_row_stats(h.data_ptr(), h.shape[1], M, h.shape[1], p(w), s1.data_ptr(), s2.data_ptr(), p(s3), _st(h)), "fb_row_stats")
    return s1, s2, s3


def folded_stats_fwd(A1, A2, A3, rn, a0, a1, D):
    """-> mu, var_raw, rstd of the folded LayerNorm (see csrc/backward.cu::folded_stats_fwd_kernel)"""
    _chk(A1), _chk(A2), _chk(rn)
    mu, var_raw, rstd = torch.empty_like(rn), torch.empty_like(rn), torch.empty_like(rn)
    _lib.check(_lib.lib().fb_folded_stats_fwd(A1.data_ptr(), A2.data_ptr(), _chk(A3).data_ptr() if A3 is not None else None, rn.data_ptr(),
                                              float(a0), float(a1), float(D), LN_EPS, rn.numel(), mu.data_ptr(), var_raw.data_ptr(),
                                              rstd.data_ptr(), _st(rn)), "fb_folded_stats_fwd")
    return mu, var_raw, rstd


def _gather_vec(v, idx):
    return gather_rows(v.view(-1, 1), idx).view(-1)


def gcl_plus_forward_train(w, h, x, row, col, node_cplx, B, cmax):
    """FABind+ MC_E_GCL forward (P/models/egnn.py:44-115; inference twin: plus.cu / forward.cu) keeping what gcl_plus_backward consumes"""
    N, H = h.shape
    E, dev = row.numel(), h.device
    Dp, D = w["e1_rad"].numel(), 2 * H + 1
    d, d2, rn, nrm = radial_fwd(x, row, col, node_cplx, B)
    s1, s2, _ = row_stats(h)
    Pn = linear(h, w["e1_rc"])
    mu, var_raw, rstd = folded_stats_fwd(vec_add_(_gather_vec(s1, row), _gather_vec(s1, col)), vec_add_(_gather_vec(s2, row), _gather_vec(s2, col)),
                                         None, rn, 1.0, 1.0, D)
    U = gather_rows(Pn, row, 0, Dp)
    gather_add_rows(Pn, col, U, col0=Dp)
    rank1_add(U, rn, w["e1_rad"])
    rank1_add(U, _neg(mu), w["e1_g"])
    Z1 = rank1_add(scale_rows(U.clone(), rstd), _ones(E, dev), w["e1_c0"])
    M = linear(act_fwd(Z1, ACT_RELU), w["e2_w"], w["e2_b"], act=ACT_RELU)
    T3 = linear(layernorm(M, w["cl_g"], w["cl_b"]), w["c1_w"], w["c1_b"], act=ACT_RELU)
    s = rowdot(T3, w["c2_w"])
    ssum = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    scatter_add_rows(scale_rows(d.clone(), s), row, ssum)
    deg = torch.zeros(N, 1, dtype=torch.float32, device=dev)
    scatter_add_rows(_ones(E, dev).view(E, 1), row, deg)
    deg = deg.view(N)
    step, x_new = coord_apply(x, ssum, deg, cmax)
    agg = torch.zeros(N, H, dtype=torch.float32, device=dev)
    scatter_add_rows(M, row, agg)
    cat = torch.empty(N, 2 * H, dtype=torch.float32, device=dev)
    cat[:, :H].copy_(h)
    cat[:, H:].copy_(agg)
    t1 = linear(layernorm(cat, w["nl_g"], w["nl_b"]), w["n1_w"], w["n1_b"], act=ACT_RELU)
    t2 = linear(t1, w["n2_w"], w["n2_b"], act=ACT_RELU)
    h_new = vec_add_(h.clone(), t2)
    return h_new, x_new, dict(h=h, x=x, rn=rn, nrm=nrm, mu=mu, var_raw=var_raw, rstd=rstd, U=U, Z2=M, Z3=T3, Z4=t1, Z5=t2, s=s, deg=deg,
                              step=step, agg=agg)


def pair_row_nodes(geo):
    """protein-side / compound-side node of every pair row (index bookkeeping; small host loop over complexes)"""
    c_off, p_off = geo["c_off"].tolist(), geo["p_off"].tolist()
    pi, ci = [], []
    for b in range(geo["B"]):
        nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
        pi.append(torch.arange(p_off[b], p_off[b + 1], dtype=torch.int32).repeat_interleave(nc1))
        ci.append(torch.arange(c_off[b], c_off[b + 1], dtype=torch.int32).repeat(np1))
    dev = geo["c_off"].device
    return torch.cat(pi).to(dev).contiguous(), torch.cat(ci).to(dev).contiguous()


def att_plus_forward_train(w, pair_in, h, x, geo, row, col, idx, pair_nodes, cmax):
    """FABind+ MC_Att_L forward (P/models/egnn.py:129-300, P/models/cross_att.py:20-45) keeping what att_plus_backward consumes"""
    N, H = h.shape
    Nc, dev, E = geo["Nc"], h.device, row.numel()
    P = pair_in.shape[0]
    pi_all, ci_all = pair_nodes
    raw_full = linear(pair_in, w["pb_w"], w["pb_b"])
    PB = pair_bias_gate_fwd(raw_full, 2)
    PB_p, PB_c = PB[:, 0].contiguous(), PB[:, 1].contiguous()
    hc0, hp0 = h[:Nc], h[Nc:]
    CAc = linear(hc0, w["ca_c_w"], w["ca_c_b"])
    CAp = linear(hp0, w["ca_p_w"], w["ca_p_b"])
    Op = row_attention_fwd(geo, 1, (CAp, 0), (CAp, HD), (CAc, 0), (CAc, HD), PB_p, N - Nc)
    hp1 = linear(Op, w["o_p_w"], w["o_p_b"], res=hp0)
    CAp2 = linear(hp1, w["ca_p2_w"])
    Oc = row_attention_fwd(geo, 0, (CAc, 2 * HD), (CAc, 3 * HD), (CAp2, 0), (CAp2, HD), PB_c, Nc)
    hc1 = linear(Oc, w["o_c_w"], w["o_c_b"], res=hc0)
    tr = {}
    h2 = torch.empty(N, H, dtype=torch.float32, device=dev)
    for t, hs, dst in (("tc", hc1, h2[:Nc]), ("tp", hp1, h2[Nc:])):
        T1 = linear(layernorm(hs, w[t + "l_g"], w[t + "l_b"]), w[t + "1_w"], w[t + "1_b"], act=ACT_RELU)
        T2 = linear(T1, w[t + "2_w"], w[t + "2_b"], act=ACT_RELU)
        dst.copy_(hs)
        vec_add_(dst, T2)
        tr[t] = (T1, T2)
    QK = linear(h2, w["qk_w"], w["qk_b"])
    a32, b32 = gather_rows(QK, pi_all, 2 * H, 32), gather_rows(QK, ci_all, 2 * H + 32, 32)
    t32 = vec_mul(a32, b32)
    Zpre = linear(t32, w["zo_w_t"], w["zo_b"], res=pair_in)
    Zh = linear(layernorm(Zpre, w["zl_g"], w["zl_b"]), w["pt1_w"], w["pt1_b"], act=ACT_RELU)
    Zo = linear(Zh, w["pt2_w"], w["pt2_b"], act=ACT_RELU)
    pb_dense = rowdot(Zo, w["wb"]).view(P, 1)
    rank1_add(pb_dense, _ones(P, dev), w["pt_c"])
    d, d2, rn, nrm = radial_fwd(x, row, col, geo["node_cplx"], geo["B"])
    logit = rowdot2(gather_rows(QK, row, 0, H), rank1_add(gather_rows(QK, col, H, H), rn, w["k_r"]))
    vec_add_(logit, gather_rows(pb_dense, idx["pair"]).view(E))
    alpha = softmax_seg_fwd(logit, idx["rowptr"], N)
    ve = scale_rows(rank1_add(gather_rows(QK, col, 2 * H + 128, H), rn, w["v_r"]), alpha)
    h3 = h2.clone()
    scatter_add_rows(ve, row, h3)
    V = gather_rows(QK, torch.arange(N, dtype=torch.int32, device=dev), 2 * H + 128, H)
    s1, s2, s3 = row_stats(V, w["v_r"])
    acr = tuple(float(v) for v in w["ac_r"].tolist())
    mu, var_raw, rstd = folded_stats_fwd(_gather_vec(s1, col), _gather_vec(s2, col), _gather_vec(s3, col), rn, acr[0], acr[1], H)
    Uc = rank1_add(gather_rows(QK, col, 3 * H + 128, H), rn, w["ac_u"])
    rank1_add(Uc, _neg(mu), w["ac_g"])
    tco = rank1_add(scale_rows(Uc.clone(), rstd), _ones(E, dev), w["ac_c0"])
    se = rowdot(act_fwd(tco, ACT_RELU), w["ac2_w"])
    ssum = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    scatter_add_rows(scale_rows(d.clone(), vec_mul(alpha, se)), row, ssum)
    step, x_new = coord_apply(x, ssum, None, cmax)
    sv = dict(h_in=h, x=x, CAc=CAc, CAp=CAp, CAp2=CAp2, PB_p=PB_p, PB_c=PB_c, raw_full=raw_full, pair_in=pair_in, Zpre=Zpre, Zh=Zh, Zo=Zo, t32=t32,
              a32=a32, b32=b32, pi_all=pi_all, ci_all=ci_all, Tc1=tr["tc"][0], Tc2=tr["tc"][1], Tp1=tr["tp"][0], Tp2=tr["tp"][1], Op=Op, Oc=Oc,
              hp1=hp1, hc1=hc1, h2=h2, QK=QK, pair=idx["pair"], rn=rn, nrm=nrm, alpha=alpha, se=se, s3=s3, mu=mu, var_raw=var_raw, rstd=rstd,
              Uc=Uc, step=step, acr=acr)
    return h3, x_new, Zo, sv


def stack_forward_train_plus(weights, Hin, x_state, moves, geo, edges, consts, n_layers):
    """Last refinement iteration of the FABind+ stack in training mode, eval-mode masks (no dropout): returns
    (X_out, H_out, pair_out [P,H], tape, top) in the form stack_backward_plus consumes."""
    N, H = Hin.shape
    Nc, dev = geo["Nc"], Hin.device
    wt = weights[""]
    pc = torch.empty(N, H, dtype=torch.float32, device=dev)
    pc[:Nc].copy_(linear(Hin[:Nc], wt["il_c_w"], wt["il_c_b"]))
    pc[Nc:].copy_(linear(Hin[Nc:], wt["il_p_w"], wt["il_p_b"]))
    outer = pair_outer_fwd(pc, geo, consts["n_pairs"])
    pair = linear(outer, wt["il_o_w"], wt["il_o_b"])
    h = linear(Hin, wt["in_w"], wt["in_b"])
    x = x_state
    idx = interface_indices(edges["int_row"], edges["int_col"], geo)
    pair_nodes = pair_row_nodes(geo)
    tape = []
    for l in range(n_layers):
        h, x, s_gcl = gcl_plus_forward_train(weights[f"gcl{l}."], h, x, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], geo["B"], consts["cmax"])
        h, x, pair, s_att = att_plus_forward_train(weights[f"att{l}."], pair, h, x, geo, edges["int_row"], edges["int_col"], idx, pair_nodes,
                                                   consts["cmax"])
        acc = las_acc(x, consts["xl"], edges["las_a"], edges["las_b"], consts["las_step"])
        x_in = x
        _, x = coord_apply(x_in, acc, None, consts["lcl"])
        tape.append((s_gcl, s_att, dict(x=x_in, acc=acc)))
    h_last, x, s_out = gcl_plus_forward_train(weights["out."], h, x, edges["ctx_row"], edges["ctx_col"], geo["node_cplx"], geo["B"], consts["cmax"])
    H_out = linear(h_last, wt["out_w"], wt["out_b"])
    X_out = torch.where(moves[:, None], x, x_state)
    top = dict(Hin=Hin, pc=pc, outer=outer, h_last=h_last, out_saved=s_out)
    return X_out, H_out, pair, tape, top
