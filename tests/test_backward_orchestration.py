"""The GPU-side orchestration of the reverse pass (fabind_b200/backward.py: gcl_backward, att_backward, las_bwd) validated on the
CPU: every kernel wrapper is swapped for its torch definition (the semantic contract the GPU primitive tests hold the kernels
to), and the orchestrated result is compared with the pinned specification (tests/emulate_backward.py) on the saved tensors of a
real forward.  No CUDA involved: what is checked here is the launch sequence, the operand slicing and the accumulation order."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import GOLDEN_DIR, load_golden, rel_err
import emulate_backward as spec


def _actf(z, a):
    return F.silu(z) if a == 1 else F.relu(z) if a == 2 else z


def _actg(z, a):
    if a == 1:
        s = torch.sigmoid(z)
        return s * (1 + z * (1 - s))
    return (z > 0).float() if a == 2 else torch.ones_like(z)


def _install_standins(monkeypatch, bw):
    L = lambda t: t.long()

    def colsum(A, w=None, out=None):
        r = (A * (w[:, None] if w is not None else 1)).sum(0)
        return r if out is None else out + r

    def scatter(src, idx, dst, col0=0, width=None):
        width = src.shape[1] if width is None else width
        dst[:, col0:col0 + width].index_add_(0, L(idx), src[:, :width])
        return dst

    def gather_add(src, idx, dst, col0=0):
        dst += src[L(idx), col0:col0 + dst.shape[1]]
        return dst

    def wgrad(dY, X, out=None):
        r = dY.t() @ X
        return r if out is None else out + r

    def coord(x, row, col, s, step, cnt, cmax, dx_new):
        row, col = L(row), L(col)
        g = dx_new * (step.abs() <= cmax)
        if cnt is not None:
            g = g / cnt.clamp(min=1)[:, None]
        d, de = x[row] - x[col], g[row]
        dd = de * s[:, None]
        return dx_new.clone().index_add_(0, row, dd).index_add_(0, col, -dd), (de * d).sum(1)

    def radial(x, row, col, cplx, nrm, drn, dx):
        row, col, cplx = L(row), L(col), L(cplx)
        d = x[row] - x[col]
        d2, eb = (d * d).sum(1), cplx[row]
        dot = torch.zeros_like(nrm).index_add_(0, eb, drn * d2)
        g = 2 * d * (drn / nrm[eb] - d2 * dot[eb] / nrm[eb] ** 3)[:, None]
        dx.index_add_(0, row, g).index_add_(0, col, -g)
        return dx

    def las(x, xref, a, b, acc, step_size, lcl, dx_new):
        a, b = L(a), L(b)
        d = x[a] - x[b]
        diff = (d * d).sum(1) - ((xref[a] - xref[b]) ** 2).sum(1)
        f = (dx_new * (acc.abs() <= lcl) * step_size)[b]
        dd = 4 * diff[:, None] * f + 8 * (f * d).sum(1)[:, None] * d
        return dx_new.clone().index_add_(0, a, dd).index_add_(0, b, -dd)

    def scale_rows(A, u):
        A *= u[:, None]
        return A

    def rank1(A, u, v):
        A += u[:, None] * v[None, :]
        return A

    def vec_add_(c, a):
        c += a
        return c

    def smax(alpha, dalpha, row, n):
        t = torch.zeros(n).index_add_(0, L(row), alpha * dalpha)
        return alpha * (dalpha - t[L(row)])

    def rowatt(geo, q_is_prot, Q, G, K, V, PB, dO, dQ, dG, dK, dV):
        Nc, dPB = geo["Nc"], torch.zeros_like(PB)
        sl = lambda tc, w=128: tc[0][:, tc[1]:tc[1] + w]
        for b in range(geo["B"]):
            c0, c1, p0, p1 = int(geo["c_off"][b]), int(geo["c_off"][b + 1]), int(geo["p_off"][b]) - Nc, int(geo["p_off"][b + 1]) - Nc
            nc1, np1 = c1 - c0, p1 - p0
            pr = slice(int(geo["pair_base"][b]), int(geo["pair_base"][b + 1]))
            qs, ks = (slice(p0, p1), slice(c0, c1)) if q_is_prot else (slice(c0, c1), slice(p0, p1))
            bias = PB[pr].view(np1, nc1, 4)
            bias = bias if q_is_prot else bias.transpose(0, 1)
            _, s = spec.rowatt_fwd(sl(Q)[qs], sl(G)[qs], sl(K)[ks], sl(V)[ks], bias)
            dq, dg, dk, dv, db = spec.rowatt_bwd(s, dO[qs])
            sl(dQ)[qs], sl(dG)[qs], sl(dK)[ks], sl(dV)[ks] = dq, dg, dk, dv
            dPB[pr] = (db if q_is_prot else db.transpose(0, 1)).reshape(-1, 4)
        return dPB

    reps = dict(act_fwd=_actf, act_bwd=lambda Z, dY, a: dY * _actg(Z, a), outer_act_bwd=lambda Z, u, v, a: u[:, None] * v[None, :] * _actg(Z, a),
                colsum=colsum, rowdot=lambda A, v: A @ v, scatter_add_rows=scatter, gather_add_rows=gather_add, gemm_wgrad=wgrad,
                gemm_dgrad=lambda dY, Wt: dY @ Wt.t(), coord_step_bwd=coord, radial_bwd=radial, las_bwd=las,
                rowdot2=lambda A, Bm: (A * Bm).sum(1), scale_rows=scale_rows, rank1_add=rank1, vec_mul=lambda a, b: a * b, vec_add_=vec_add_,
                softmax_seg_bwd=smax, row_attention_bwd=rowatt)
    for k, v in reps.items():
        monkeypatch.setattr(bw, k, v)


def _weights(W, pre, names):
    w = {}
    for n in names:
        t = W.m(pre + n).clone()
        w[n] = t
        if t.dim() == 2:
            w[n + "_t"] = t.t().contiguous()
    return w


def test_orchestration_matches_specification(monkeypatch):
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))[0])
    gen = torch.Generator().manual_seed(4)
    ex = {}
    spec.forward_backward_v1(sd, cfg, b, torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen), export=ex)
    W, geo, tape, N, B, Nc, cmax = ex["W"], ex["geo"], ex["tape"], ex["N"], ex["B"], ex["Nc"], ex["cmax"]
    H = b.H.shape[1]
    i32 = lambda t: t.to(torch.int32)
    s1, s2, s3 = tape[0]
    dh_up, dx_up = torch.randn(N, H, generator=gen), torch.randn(N, 3, generator=gen)

    # ---- LAS step
    a, bb = ex["las"]
    ref = spec.las_bwd(s3, ex["las"], cfg.geometry_reg_step_size, ex["lcl"], dx_up)
    mine = bw.las_bwd(s3["x"], ex["xl"], i32(a), i32(bb), s3["acc"], cfg.geometry_reg_step_size, ex["lcl"], dx_up)
    assert rel_err(mine, ref) < 1e-5

    # ---- MC_E_GCL
    G = spec.Grads()
    rdh, rdx = spec.gcl_bwd(G, W, "gcl0.", s1, ex["ctx"], geo["cplx"], B, cmax, dh_up, dx_up)
    names = ["e1_rc", "e1_rad", "e2_w", "c1_w", "c2_w", "n1_w", "n2_w"]
    d_, d2, nrm = s1["rs"]
    saved = dict(h=s1["h"], x=s1["x"], rn=s1["rn"], nrm=nrm, Z1=s1["Z1"], Z2=s1["Z2"], Z3=s1["Z3"], s=s1["s"], deg=s1["deg"],
                 step=s1["step"], agg=s1["cat"][:, H:].contiguous(), Z4=s1["Z4"])
    dh, dx, grads = bw.gcl_backward(_weights(W, "gcl0.", names), saved, i32(ex["ctx"][0]), i32(ex["ctx"][1]), i32(geo["cplx"]), cmax,
                                    dh_up, dx_up)
    assert rel_err(dh, rdh) < 1e-5 and rel_err(dx, rdx) < 1e-5
    for k, v in grads.items():
        assert rel_err(v.reshape(-1), G["gcl0." + k].reshape(-1)) < 1e-5, k

    # ---- MC_Att_L
    G = spec.Grads()
    P0, PB = ex["P0"], ex["PB"]
    rdP0, rdPB = torch.zeros_like(P0), torch.zeros_like(PB)
    rdh, rdx = spec.att_bwd(G, W, "att0.", 0, s2, geo, ex["inter"], cmax, dh_up, dx_up, rdP0, rdPB)
    names = ["ac2_w", "ac_u", "ac1_b", "v_r", "k_r", "pt2v", "pt_c", "pt1_w", "qk_w", "tc1_w", "tc2_w", "tp1_w", "tp2_w", "o_c_w", "o_p_w",
             "ca_p2_w", "ca_c_w", "ca_p_w"]
    u = s2["u"]
    geo_dev = dict(Nc=Nc, B=B, c_off=i32(torch.from_numpy(geo["c_off"].astype(np.int64))), p_off=i32(torch.from_numpy(geo["p_off"].astype(np.int64))),
                   pair_base=i32(torch.from_numpy(geo["pair_base"].astype(np.int64))), node_cplx=i32(geo["cplx"]),
                   max_c=int(np.diff(geo["c_off"]).max()), max_p=int(np.diff(geo["p_off"]).max()))
    sv = dict(h_in=s2["h_in"], x=s2["x"], CAc=s2["CAc"], CAp=s2["CAp"], CAp2=s2["CAp2"], PB_p=PB[:, 0, 0].contiguous(), PB_c=PB[:, 0, 1].contiguous(),
              Op=s2["Op"], Oc=s2["Oc"], hp1=s2["hp1"], hc1=s2["hc1"], Ttp=s2["Tp"], Ttc=s2["Tc"], h2=s2["h2"], QK=s2["QK"], pc32=s2["pc32"],
              pair=i32(s2["pair"]), u_pair=i32(s2["pair"][u]), u_pi=i32(s2["pi"][u]), u_ci=i32(s2["ci"][u]), zcat=s2["zcat"], Zp=s2["Zp"],
              rn=s2["rn"], nrm=s2["rs"][2], alpha=s2["alpha"], se=s2["se"], zc=s2["zc"], step=s2["step"])
    dP0 = torch.zeros_like(P0)
    dh, dx, grads, dPB_p, dPB_c = bw.att_backward(_weights(W, "att0.", names), sv, geo_dev, i32(ex["inter"][0]), i32(ex["inter"][1]), cmax,
                                                  dh_up.clone(), dx_up.clone(), dP0)
    assert rel_err(dh, rdh) < 1e-5, rel_err(dh, rdh)
    assert rel_err(dx, rdx) < 1e-5, rel_err(dx, rdx)
    assert rel_err(dP0, rdP0) < 1e-5
    assert rel_err(dPB_p, rdPB[:, 0, 0]) < 1e-5 and rel_err(dPB_c, rdPB[:, 0, 1]) < 1e-5
    assert set("att0." + k for k in grads) == set(G), set("att0." + k for k in grads) ^ set(G)
    for k, v in grads.items():
        assert rel_err(v.reshape(-1), G["att0." + k].reshape(-1)) < 1e-5, k
