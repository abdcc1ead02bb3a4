"""Reverse pass of the docking stack on the GPU (training path, fp32) -- first part: the primitives and the MC_E_GCL sub-layer.

The reference trains through torch autograd (main_fabind.py:380-401).  Here the reverse pass is hand-derived for the library's
formulation (hoisted first Linears, CSR segment reductions; specification pinned against autograd and the unmodified
reference: tests/emulate_backward.py) and runs as launches of csrc/backward.cu plus fb_gemm for the data-gradient GEMMs
(`dX = dY W` on a transposed weight).  torch only allocates device memory here; there is no torch arithmetic on the path and
no fallback: CPU tensors raise.

Status: MC_E_GCL (egnn.py:68-144) reverse pass + the LAS step (egnn.py:433-449); MC_Att_L's reverse kernels (row attention,
interfacial attention, pair path) are the next round's work."""
import ctypes as C

import torch

from . import _lib
from .runtime import current_stream_ptr

ACT_NONE, ACT_SILU, ACT_RELU = 0, 1, 2


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
    if out is None:
        out = torch.zeros(N, K, dtype=torch.float32, device=dY.device)
    _lib.check(_lib.lib().fb_gemm_wgrad(dY.data_ptr(), N, X.data_ptr(), K, M, N, K, out.data_ptr(), K, _st(dY)), "fb_gemm_wgrad")
    return out


def gemm_dgrad(dY, Wt):
    """dX = dY W, with Wt = W^T stored [K_in, N_out] contiguous (the 'weight' of the data-gradient GEMM); fp32 SIMT path"""
    _chk(dY), _chk(Wt)
    M, N = dY.shape
    K = Wt.shape[0]
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = dY.data_ptr(), N, N
    g.W, g.bias, g.act = Wt.data_ptr(), None, ACT_NONE
    g.M, g.N, g.bf16_mode, g.force_simt = M, K, 0, 0
    out = torch.empty(M, K, dtype=torch.float32, device=dY.device)
    g.C, g.ldc = out.data_ptr(), K
    _lib.check(_lib.lib().fb_gemm(C.byref(g), _st(dY)), "fb_gemm")
    return out


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


def _linear_bwd(grads, wname, bname, Wt, X, dY, need_dx=True):
    """y = x W^T + b: accumulates dW, db into `grads`, returns dX = dY W"""
    grads[wname] = gemm_wgrad(dY, X, grads.get(wname))
    if bname is not None:
        grads[bname] = colsum(dY, None, grads.get(bname))
    return gemm_dgrad(dY, Wt) if need_dx else None


def gcl_backward(w, saved, row, col, node_cplx, cmax, dh_new, dx_new):
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
    T3 = act_fwd(saved["Z3"], ACT_SILU)
    grads["c2_w"] = colsum(T3, ds)
    dZ3 = outer_act_bwd(saved["Z3"], ds, w["c2_w"], ACT_SILU)        # dZ3[e,f] = ds[e] c2[f] silu'(Z3[e,f])
    M = act_fwd(saved["Z2"], ACT_SILU)
    dM = _linear_bwd(grads, "c1_w", "c1_b", w["c1_w_t"], M, dZ3)
    # node branch: h_new = h + n2(silu(n1([h | agg])))
    t1 = act_fwd(saved["Z4"], ACT_SILU)
    dt1 = _linear_bwd(grads, "n2_w", "n2_b", w["n2_w_t"], t1, dh_new)
    dZ4 = act_bwd(saved["Z4"], dt1, ACT_SILU)
    cat = torch.empty(N, 2 * H, dtype=torch.float32, device=h.device)
    cat[:, :H].copy_(h)
    cat[:, H:].copy_(saved["agg"])
    dcat = _linear_bwd(grads, "n1_w", "n1_b", w["n1_w_t"], cat, dZ4)
    gather_add_rows(dcat, row, dM, col0=H)                    # dM[e] += dagg[row[e]]
    # edge MLP
    dZ2 = act_bwd(saved["Z2"], dM, ACT_SILU)
    A1 = act_fwd(saved["Z1"], ACT_SILU)
    dA1 = _linear_bwd(grads, "e2_w", "e2_b", w["e2_w_t"], A1, dZ2)
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
