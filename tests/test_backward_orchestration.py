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
from emulate_packed import forward_emulated


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


def _weights(W, pre, names=None):
    """every arena slot of a prefix (names=None) or the named ones, + `_t` transposes of the matrices"""
    if names is None:
        names = [n[len(pre):] for n in W.s if n.startswith(pre) and "." not in n[len(pre):] and W.s[n][0] * W.s[n][1] > 0]
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
    ex, cfg, H, dh_up, dx_up = spec_case()
    W, geo, tape, N, B, Nc, cmax = ex["W"], ex["geo"], ex["tape"], ex["N"], ex["B"], ex["Nc"], ex["cmax"]
    i32 = lambda t: t.to(torch.int32)
    s1, s2, s3 = tape[0]

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
    case = att_case(ex, dh_up, dx_up)
    dP0 = torch.zeros_like(ex["P0"])
    dh, dx, grads, dPB_p, dPB_c = bw.att_backward(case["w"], case["sv"], case["geo"], case["row"], case["col"], cmax, dh_up.clone(),
                                                  dx_up.clone(), dP0)
    check_att(case, dh, dx, grads, dPB_p, dPB_c, dP0, 1e-5)


def spec_case(seed=4):
    """saved tensors of the last iteration of a real forward (gradient golden batch) + upstream gradients"""
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))[0])
    gen = torch.Generator().manual_seed(seed)
    ex = {}
    spec.forward_backward_v1(sd, cfg, b, torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen), export=ex)
    H = b.H.shape[1]
    return ex, cfg, H, torch.randn(ex["N"], H, generator=gen), torch.randn(ex["N"], 3, generator=gen)


def check_att(case, dh, dx, grads, dPB_p, dPB_c, dP0, tol):
    G = case["G"]
    assert rel_err(dh, case["dh"]) < tol, rel_err(dh, case["dh"])
    assert rel_err(dx, case["dx"]) < tol, rel_err(dx, case["dx"])
    assert rel_err(dP0, case["dP0"]) < tol
    assert rel_err(dPB_p, case["dPB"][:, 0, 0]) < tol and rel_err(dPB_c, case["dPB"][:, 0, 1]) < tol
    assert set("att0." + k for k in grads) == set(G), set("att0." + k for k in grads) ^ set(G)
    gmax = max(float(t.abs().max()) for t in G.values())
    for k, v in grads.items():
        ref = G["att0." + k].reshape(-1)
        # absolute floor: the gradient of the softmax-shift constant pt_c is exactly zero in real arithmetic (sum of dlogit over a row)
        err = float((v.reshape(-1) - ref).abs().max())
        assert err < tol * float(ref.abs().max()) + 5e-2 * tol * gmax, (k, err, float(ref.abs().max()))


def att_case(ex, dh_up, dx_up):
    """inputs of bw.att_backward for layer 0 of the exported forward + the specification's results"""
    W, geo, tape, N, B, Nc, cmax = ex["W"], ex["geo"], ex["tape"], ex["N"], ex["B"], ex["Nc"], ex["cmax"]
    i32 = lambda t: t.to(torch.int32)
    s2 = tape[0][1]
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
    return dict(w=_weights(W, "att0.", names), sv=sv, geo=geo_dev, row=i32(ex["inter"][0]), col=i32(ex["inter"][1]), G=G, dh=rdh, dx=rdx,
                dP0=rdP0, dPB=rdPB)


def _gcl_saved(s1, H):
    return dict(h=s1["h"], x=s1["x"], rn=s1["rn"], nrm=s1["rs"][2], Z1=s1["Z1"], Z2=s1["Z2"], Z3=s1["Z3"], s=s1["s"], deg=s1["deg"],
                step=s1["step"], agg=s1["cat"][:, H:].contiguous(), Z4=s1["Z4"])


def _att_saved(s2, PB, l):
    i32 = lambda t: t.to(torch.int32)
    u = s2["u"]
    return dict(h_in=s2["h_in"], x=s2["x"], CAc=s2["CAc"], CAp=s2["CAp"], CAp2=s2["CAp2"], PB_p=PB[:, l, 0].contiguous(),
                PB_c=PB[:, l, 1].contiguous(), Op=s2["Op"], Oc=s2["Oc"], hp1=s2["hp1"], hc1=s2["hc1"], Ttp=s2["Tp"], Ttc=s2["Tc"], h2=s2["h2"],
                QK=s2["QK"], pc32=s2["pc32"], pair=i32(s2["pair"]), u_pair=i32(s2["pair"][u]), u_pi=i32(s2["pi"][u]), u_ci=i32(s2["ci"][u]),
                zcat=s2["zcat"], Zp=s2["Zp"], rn=s2["rn"], nrm=s2["rs"][2], alpha=s2["alpha"], se=s2["se"], zc=s2["zc"], step=s2["step"])


GCL_W = ["e1_rc", "e1_rad", "e2_w", "c1_w", "c2_w", "n1_w", "n2_w"]
ATT_W = ["ac2_w", "ac_u", "ac1_b", "v_r", "k_r", "pt2v", "pt_c", "pt1_w", "qk_w", "tc1_w", "tc2_w", "tp1_w", "tp2_w", "o_c_w", "o_p_w", "ca_p2_w",
         "ca_c_w", "ca_p_w"]
TOP_W = ["out_w", "in_w", "pb_w", "il_o_w", "il_c_w", "il_p_w"]


def two_layer_problem():
    from fabind_b200 import EfficientMCAttModel
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_batch
    from oracle import fabind_oracle as orc
    from oracle.det_weights import det_state_dict
    H, L = 32, 2
    m = EfficientMCAttModel(published_args(), H, H, 1, n_layers=L, n_iter=2,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 23)
    return make_batch(n_complexes=3, seed=11, embed=H, n_c_range=(4, 9), n_p_range=(10, 20)), sd, orc.make_cfg(n_layers=L, n_iter=2)


def stack_case(path=None, seed=7, problem=None):
    """everything fabind_b200.backward.stack_backward_v1 consumes, from the specification's forward on a golden batch (or on a
    (batch, state_dict, cfg) problem), and the specification's arena gradient"""
    if problem is None:
        g, r, b, sd, cfg = load_golden(path)
    else:
        b, sd, cfg = problem
    gen = torch.Generator().manual_seed(seed)
    gX, gH = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)
    ex = {}
    _, _, garena, gHin = spec.forward_backward_v1(sd, cfg, b, gX, gH, export=ex)
    W, geo, N, B, Nc, H, L = ex["W"], ex["geo"], ex["N"], ex["B"], ex["Nc"], b.H.shape[1], cfg.n_layers
    i32 = lambda t: t.to(torch.int32)
    geo_dev = dict(Nc=Nc, B=B, c_off=i32(torch.from_numpy(geo["c_off"].astype(np.int64))), p_off=i32(torch.from_numpy(geo["p_off"].astype(np.int64))),
                   pair_base=i32(torch.from_numpy(geo["pair_base"].astype(np.int64))), node_cplx=i32(geo["cplx"]),
                   max_c=int(np.diff(geo["c_off"]).max()), max_p=int(np.diff(geo["p_off"]).max()))
    weights = {"": _weights(W, ""), "out.": _weights(W, "out.")}
    tape = []
    for l in range(L):
        weights[f"gcl{l}."] = _weights(W, f"gcl{l}.")
        weights[f"att{l}."] = _weights(W, f"att{l}.")
        s1, s2, s3 = ex["tape"][l]
        tape.append((_gcl_saved(s1, H), _att_saved(s2, ex["PB"], l), dict(x=s3["x"], acc=s3["acc"])))
    top = dict(Hin=ex["Hin"], pc=ex["pc"], outer=ex["outer"], P0=ex["P0"], raw_full=ex["raw_full"], h_last=ex["h_last"],
               out_saved=_gcl_saved(ex["s_out"], H))
    edges = dict(ctx_row=i32(ex["ctx"][0]), ctx_col=i32(ex["ctx"][1]), int_row=i32(ex["inter"][0]), int_col=i32(ex["inter"][1]),
                 las_a=i32(ex["las"][0]), las_b=i32(ex["las"][1]))
    consts = dict(cmax=ex["cmax"], lcl=ex["lcl"], las_step=cfg.geometry_reg_step_size, xl=ex["xl"])
    permt = ex["permt"]
    dH_out, dX_out = gH[permt].contiguous(), (gX[permt, 0] * ex["moves"][:, None]).contiguous()
    consts["n_pairs"] = int(ex["P0"].shape[0])
    return dict(weights=weights, tape=tape, top=top, geo=geo_dev, edges=edges, consts=consts, dH_out=dH_out, dX_out=dX_out, W=W,
                garena=garena, gHin=gHin[permt], x_state=ex["x_state"], moves=ex["moves"], x_out=ex["x_out"], h_final=ex["h_final"], L=L)


def check_stack(case, grads, dHin, tol):
    W, ref = case["W"], case["garena"]
    gmax = float(ref.abs().max())
    seen = 0
    for name, (r, c, off) in W.s.items():
        if r * c == 0:
            continue
        t = ref[off:off + r * c]
        if name not in grads:
            assert float(t.abs().max()) == 0.0, f"{name}: gradient missing"
            continue
        err = float((grads[name].reshape(-1) - t).abs().max())
        assert err < tol * float(t.abs().max()) + 1e-2 * tol * gmax, (name, err, float(t.abs().max()))
        seen += 1
    assert seen >= 40
    assert rel_err(dHin, case["gHin"]) < tol


def _gate_bwd_standin(raw, dPB):
    P, nblk = dPB.shape[0], dPB.shape[1]
    r5 = raw[:, :8 * nblk].reshape(P, nblk, 2, 4)
    sg = torch.sigmoid(r5[:, :, 1])
    out = torch.zeros_like(raw)
    out[:, :8 * nblk] = torch.stack([dPB * sg, dPB * r5[:, :, 0] * sg * (1 - sg)], 2).reshape(P, -1)
    return out


def _outer_bwd_standin(douter, pc, geo):
    dpc = torch.zeros_like(pc)
    for b in range(geo["B"]):
        c0, c1, p0, p1 = int(geo["c_off"][b]), int(geo["c_off"][b + 1]), int(geo["p_off"][b]), int(geo["p_off"][b + 1])
        t = douter[int(geo["pair_base"][b]):int(geo["pair_base"][b + 1])].view(p1 - p0, c1 - c0, -1)
        dpc[p0:p1] += (t * pc[None, c0:c1]).sum(1)
        dpc[c0:c1] += (t * pc[p0:p1, None]).sum(0)
    return dpc


def test_stack_reverse_pass_matches_specification(monkeypatch):
    """the whole last-iteration reverse pass as fabind_b200.backward orchestrates it == the specification's arena gradient"""
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)

    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    cases = [stack_case(path) for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))]
    cases.append(stack_case(problem=two_layer_problem()))      # two layers: pair-bias blocks and pair_embed0 gradients accumulate
    for case in cases:
        grads, dHin = bw.stack_backward_v1(case["weights"], case["tape"], case["top"], case["geo"], case["edges"], case["consts"],
                                           case["dH_out"], case["dX_out"])
        check_stack(case, grads, dHin, 1e-4)


def _install_forward_standins(monkeypatch, bw):
    L = lambda t: t.long()

    def drop_apply(self, X, layer, name, row0=0):
        from fabind_b200.dropout import keep_mask
        return X * keep_mask(self.seed, self.site(layer, name), X.shape[0], X.shape[1], self.p, colonly=bool(self.colonly), row0=row0)
    monkeypatch.setattr(bw.Drop, "apply", drop_apply)

    def linear(A, W, bias=None, act=0, res=None, drop=None, A16=None):
        y = _actf(F.linear(A, W, bias), act)
        if drop is not None and drop[0] is not None and drop[0].p > 0:
            y = drop[0].apply(y, drop[1], drop[2], drop[3])
        return y if res is None else y + res

    def radial_fwd(x, row, col, cplx, B):
        d = x[L(row)] - x[L(col)]
        d2 = (d * d).sum(1)
        nrm = torch.zeros(B).index_add_(0, L(cplx)[L(row)], d2 * d2).sqrt()
        return d, d2, d2 / nrm[L(cplx)[L(row)]], nrm

    def coord_apply(x, ssum, cnt, cmax):
        step = ssum / cnt.clamp(min=1)[:, None] if cnt is not None else ssum.clone()
        return step, x + step.clamp(-cmax, cmax)

    def softmax_seg_fwd(logit, rowptr, n_rows):
        alpha = torch.empty_like(logit)
        for r in range(n_rows):
            lo, hi = int(rowptr[r]), int(rowptr[r + 1])
            if hi > lo:
                alpha[lo:hi] = torch.softmax(logit[lo:hi], 0)
        return alpha

    def las_acc(x, xref, a, b, step_size):
        d = x[L(a)] - x[L(b)]
        diff = (d * d).sum(1) - ((xref[L(a)] - xref[L(b)]) ** 2).sum(1)
        return torch.zeros_like(x).index_add_(0, L(b), 4 * diff[:, None] * d * step_size)

    def pair_outer_fwd(pc, geo, n_pairs):
        H = pc.shape[1]
        return torch.cat([(pc[int(geo["p_off"][b]):int(geo["p_off"][b + 1]), None, :] * pc[None, int(geo["c_off"][b]):int(geo["c_off"][b + 1]), :]
                           ).reshape(-1, H) for b in range(geo["B"])])

    def gate_fwd(raw, nblk):
        r5 = raw[:, :8 * nblk].reshape(raw.shape[0], nblk, 2, 4)
        return (r5[:, :, 0] * torch.sigmoid(r5[:, :, 1])).contiguous()

    def rowatt_fwd(geo, q_is_prot, Q, G, K, V, PB, n_q_rows):
        Nc, O = geo["Nc"], torch.zeros(n_q_rows, 128)
        sl = lambda tc: tc[0][:, tc[1]:tc[1] + 128]
        for b in range(geo["B"]):
            c0, c1, p0, p1 = int(geo["c_off"][b]), int(geo["c_off"][b + 1]), int(geo["p_off"][b]) - Nc, int(geo["p_off"][b + 1]) - Nc
            bias = PB[int(geo["pair_base"][b]):int(geo["pair_base"][b + 1])].view(p1 - p0, c1 - c0, 4)
            qs, ks = (slice(p0, p1), slice(c0, c1)) if q_is_prot else (slice(c0, c1), slice(p0, p1))
            O[qs] = spec.rowatt_fwd(sl(Q)[qs], sl(G)[qs], sl(K)[ks], sl(V)[ks], bias if q_is_prot else bias.transpose(0, 1))[0]
        return O
    for k, v in dict(linear=linear, radial_fwd=radial_fwd, coord_apply=coord_apply, softmax_seg_fwd=softmax_seg_fwd, las_acc=las_acc,
                     pair_outer_fwd=pair_outer_fwd, pair_bias_gate_fwd=gate_fwd, row_attention_fwd=rowatt_fwd).items():
        monkeypatch.setattr(bw, k, v)


def check_forward(case, X_out, H_out, tape, top, tol):
    """the training-mode forward reproduces the specification's outputs and every tensor the reverse pass consumes"""
    assert rel_err(X_out, case["x_out"]) < tol and rel_err(H_out, case["h_final"]) < tol
    for k in ("pc", "outer", "P0", "raw_full", "h_last"):
        assert rel_err(top[k], case["top"][k]) < tol, k
    # activations the product's forward keeps NEXT TO their pre-activations so that the reverse pass does not recompute them (derived
    # values: A1 = silu(Z1), M = drop(silu(Z2)), T3 = silu(Z3), t1 = silu(Z4), Tzc = silu(zc), Tzp = relu(Zp)); the specification keeps
    # the pre-activations only
    cached = {"A1", "M", "T3", "t1", "Tzc", "Tzp"}
    for mine, ref in zip(tape + [(top["out_saved"],)], case["tape"] + [(case["top"]["out_saved"],)]):
        for sm, sr in zip(mine, ref):
            assert set(sm) - cached == set(sr), (set(sm) - cached) ^ set(sr)
            for k in sr:
                if sr[k].dtype == torch.int32:
                    assert torch.equal(sm[k].cpu(), sr[k]), k
                else:
                    assert rel_err(sm[k], sr[k]) < tol, (k, rel_err(sm[k], sr[k]))


def test_training_forward_and_reverse_close_the_loop(monkeypatch):
    """stack_forward_train_v1 -> stack_backward_v1 as orchestrated for the GPU, on torch stand-ins: outputs, every saved tensor and
    the final arena gradient equal the specification's (which is pinned to the unmodified reference)"""
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    cases = [stack_case(path) for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))] + [stack_case(problem=two_layer_problem())]
    for case in cases:
        X_out, H_out, tape, top = bw.stack_forward_train_v1(case["weights"], case["top"]["Hin"], case["x_state"], case["moves"], case["geo"],
                                                            case["edges"], case["consts"], case["L"])
        check_forward(case, X_out, H_out, tape, top, 1e-5)
        grads, dHin = bw.stack_backward_v1(case["weights"], tape, top, case["geo"], case["edges"], case["consts"], case["dH_out"], case["dX_out"])
        check_stack(case, grads, dHin, 1e-4)


def test_training_step_assembly(monkeypatch):
    """fabind_b200/train.py::training_step_v1 end to end on the CPU: kernel wrappers -> torch stand-ins, the two GPU providers
    (earlier iterations, graph builder) -> the oracle; the parameter gradients it returns are those of the UNMODIFIED reference
    (tests/golden/grad_v1_*.pt), and the outputs are the oracle's."""
    from fabind_b200 import EfficientMCAttModel, backward as bw, train
    from fabind_b200.config import published_args
    from oracle import fabind_oracle as orc
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt"))):
        g, r, b, sd, cfg = load_golden(path)
        H = r["hidden"]
        model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                                    normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        model.load_state_dict(sd, strict=True)
        fa = b.forward_args()
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)

        def prev_coords(m, fa):
            if cfg.n_iter <= 1:
                return fa["X"].clone()
            c = orc.make_cfg(n_layers=cfg.n_layers, n_iter=cfg.n_iter - 1)
            with torch.no_grad():
                return orc.model_forward(sd, c, fa["X"], fa["H"], fa["batch_id"], fa["segment_id"], fa["mask"], fa["is_global"],
                                         fa["compound_edge_index"], fa["LAS_edge_index"], fa["batched_complex_coord_LAS"])[0]

        def edge_lists(m, X_prev, fa):
            ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                            cfg.inter_cutoff / cfg.coordinate_scale)
            return ctx, inter
        X_out, H_out, pgrads, gH_in = train.training_step_v1(model, fa, lambda X, Hh: (rx, rh), prev_coords=prev_coords, edge_lists=edge_lists)
        with torch.no_grad():
            Xo, Ho = orc.model_forward(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index,
                                       b.LAS_edge_index, b.X_LAS)
        assert rel_err(X_out, Xo) < 1e-5 and rel_err(H_out, Ho) < 1e-4
        loss = float((X_out * rx).sum() + (H_out * rh).sum())
        assert abs(loss - g["loss"]) < 1e-4 * abs(g["loss"])
        gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
        n = 0
        for k, ref in g["grads"].items():
            if ref is None:
                continue
            err = float((pgrads[k] - ref).abs().max())
            assert err < 5e-4 * float(ref.abs().max()) + 5e-7 * gmax, (k, err, float(ref.abs().max()))
            n += 1
        assert n >= 80


def test_training_step_with_dropout(monkeypatch):
    """the training step in the reference's train() mode (dropout 0.1 at egnn.py:82,106,236,398,461 / cross_att.py:128): kernel
    wrappers -> torch stand-ins, earlier iterations / graph builder -> the oracle with the same column-only masks; outputs, loss
    and parameter gradients are those of the UNMODIFIED reference in train() mode with its nn.Dropout modules patched to the
    library's mask function (tests/golden/graddrop_v1_*.pt, scripts/make_golden.py::main_grad_dropout)."""
    from fabind_b200 import EfficientMCAttModel, backward as bw, train
    from fabind_b200.config import published_args
    from oracle import fabind_oracle as orc
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, "graddrop_v1_*.pt")))
    assert paths
    for path in paths:
        g, r, b, sd, cfg = load_golden(path)
        H = r["hidden"]
        dropout = (r["dropout_p"], r["dropout_seed"], True)
        model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                                    normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        model.load_state_dict(sd, strict=True)
        fa = b.forward_args()
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)

        def prev_coords(m, fa):
            c = orc.make_cfg(n_layers=cfg.n_layers, n_iter=cfg.n_iter - 1)
            with torch.no_grad():
                return orc.model_forward(sd, c, fa["X"], fa["H"], fa["batch_id"], fa["segment_id"], fa["mask"], fa["is_global"],
                                         fa["compound_edge_index"], fa["LAS_edge_index"], fa["batched_complex_coord_LAS"], dropout=dropout)[0]

        def edge_lists(m, X_prev, fa):
            ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                            cfg.inter_cutoff / cfg.coordinate_scale)
            return ctx, inter
        X_out, H_out, pgrads, gH_in = train.training_step(model, fa, lambda X, Hh: (rx, rh), prev_coords=prev_coords, edge_lists=edge_lists,
                                                          dropout=dropout)
        assert rel_err(X_out, g["X"]) < 1e-5 and rel_err(H_out, g["H"]) < 1e-4
        loss = float((X_out * rx).sum() + (H_out * rh).sum())
        assert abs(loss - g["loss"]) < 1e-4 * abs(g["loss"])
        gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
        n = 0
        for k, ref in g["grads"].items():
            if ref is None:
                continue
            err = float((pgrads[k] - ref).abs().max())
            assert err < 5e-4 * float(ref.abs().max()) + 5e-7 * gmax, (k, err, float(ref.abs().max()))
            n += 1
        assert n >= 80


# ------------------------------------------------------------------------------------------------ FABind+ layout
def _install_plus_standins(monkeypatch, bw):
    def layernorm(x, gamma, beta):
        return F.layer_norm(x, (x.shape[1],), gamma, beta, 1e-5)

    def layernorm_bwd(grads, gname, bname, x, gamma, dy):
        mean = x.mean(1, keepdim=True)
        rstd = torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
        xhat = (x - mean) * rstd
        grads[gname] = bw.colsum(dy * xhat, None, grads.get(gname))
        grads[bname] = bw.colsum(dy, None, grads.get(bname))
        g = dy * gamma
        return rstd * (g - g.mean(1, keepdim=True) - xhat * (g * xhat).mean(1, keepdim=True))

    def row_stats_bwd(h, w, ds1, ds2, ds3, dh):
        dh += ds1[:, None] + 2 * h * ds2[:, None]
        if w is not None and ds3 is not None:
            dh += w[None, :] * ds3[:, None]
        return dh

    def folded_stats_bwd(A3, rn, a0, a1, D, mu, var_raw, rstd, drstd, dmu, drn, want_da):
        dvar = drstd * (-0.5) * rstd ** 3 * (var_raw >= 0)
        dmu = dmu - 2 * mu * dvar
        a3 = A3 if A3 is not None else torch.zeros_like(rn)
        drn += a0 * dmu / D + (2 * a3 + 2 * rn * a1) * dvar / D
        da = torch.stack([(rn * dmu / D).sum(), (rn * rn * dvar / D).sum()]) if want_da else None
        return dmu / D, dvar / D, (2 * rn * dvar / D if A3 is not None else None), da
    for k, v in dict(layernorm=layernorm, layernorm_bwd=layernorm_bwd, row_stats_bwd=row_stats_bwd, folded_stats_bwd=folded_stats_bwd).items():
        monkeypatch.setattr(bw, k, v)


def _gcl_plus_saved(s1):
    return dict(h=s1["h"], x=s1["x"], rn=s1["rn"], nrm=s1["rs"][2], mu=s1["mu"], var_raw=s1["var_raw"], rstd=s1["rstd"], U=s1["U"],
                Z2=s1["M"], Z3=s1["T3"], Z4=s1["t1"], Z5=s1["t2"], s=s1["s"], deg=s1["deg"], step=s1["step"], agg=s1["agg"])


def test_gcl_plus_orchestration_matches_specification(monkeypatch):
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)
    _install_plus_standins(monkeypatch, bw)
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_plus_*.pt")))[0])
    gen = torch.Generator().manual_seed(4)
    H = b.H.shape[1]
    # the packed pair rows are needed for gP's shape: run once without a pair gradient by reading the size from the layout
    from fabind_b200.layout import build_layout
    P = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu").P_total
    ex = {}
    spec.forward_backward_plus(sd, cfg, b, torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen),
                               0.1 * torch.randn(P, H, generator=gen), export=ex)
    W, geo, N, B, cmax = ex["W"], ex["geo"], ex["N"], ex["B"], ex["cmax"]
    i32 = lambda t: t.to(torch.int32)
    dh_up, dx_up = torch.randn(N, H, generator=gen), torch.randn(N, 3, generator=gen)
    for pre, s1 in (("gcl0.", ex["tape"][0][0]), ("out.", ex["s_out"])):
        G = spec.Grads()
        rdh, rdx = spec.gcl_plus_bwd(G, W, pre, s1, ex["ctx"], geo["cplx"], B, cmax, dh_up, dx_up)
        dh, dx, grads = bw.gcl_plus_backward(_weights(W, pre), _gcl_plus_saved(s1), i32(ex["ctx"][0]), i32(ex["ctx"][1]), i32(geo["cplx"]), cmax,
                                             dh_up.clone(), dx_up.clone())
        assert rel_err(dh, rdh) < 1e-5 and rel_err(dx, rdx) < 1e-5, (rel_err(dh, rdh), rel_err(dx, rdx))
        assert set(pre + k for k in grads) == set(G), set(pre + k for k in grads) ^ set(G)
        gmax = max(float(t.abs().max()) for t in G.values())
        for k, v in grads.items():
            ref = G[pre + k].reshape(-1)
            err = float((v.reshape(-1) - ref).abs().max())
            assert err < 1e-5 * float(ref.abs().max()) + 1e-7 * gmax, (k, err)


def _att_plus_saved(s2, W, pre):
    i32 = lambda t: t.to(torch.int32)
    acr = W.m(pre + "ac_r")
    return dict(h_in=s2["h_in"], x=s2["x"], CAc=s2["CAc"], CAp=s2["CAp"], CAp2=s2["CAp2"], PB_p=s2["PBl"][:, 0].contiguous(),
                PB_c=s2["PBl"][:, 1].contiguous(), raw_full=s2["raw_full"], pair_in=s2["pair_in"], Zpre=s2["Zpre"], Zh=s2["Zh"], Zo=s2["pair_out"],
                t32=s2["t32"], a32=s2["a32"].contiguous(), b32=s2["b32"].contiguous(), pi_all=i32(s2["pi_all"]), ci_all=i32(s2["ci_all"]),
                Tc1=s2["tr"]["tc"][2], Tc2=s2["tr"]["tc"][3], Tp1=s2["tr"]["tp"][2], Tp2=s2["tr"]["tp"][3], Op=s2["Op"], Oc=s2["Oc"],
                hp1=s2["hp1"], hc1=s2["hc1"], h2=s2["h2"], QK=s2["QK"], pair=i32(s2["pair"]), rn=s2["rn"], nrm=s2["rs"][2], alpha=s2["alpha"],
                se=s2["se"], s3=s2["s3"], mu=s2["mu"], var_raw=s2["var_raw"], rstd=s2["rstd"], Uc=s2["Uc"], step=s2["step"],
                acr=(float(acr[0]), float(acr[1])))


def test_att_plus_orchestration_matches_specification(monkeypatch):
    from fabind_b200 import backward as bw
    from fabind_b200.layout import build_layout
    _install_standins(monkeypatch, bw)
    _install_plus_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_plus_*.pt")))[0])
    gen = torch.Generator().manual_seed(4)
    H = b.H.shape[1]
    P = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu").P_total
    ex = {}
    spec.forward_backward_plus(sd, cfg, b, torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen),
                               0.1 * torch.randn(P, H, generator=gen), export=ex)
    W, geo, N, B, Nc, cmax = ex["W"], ex["geo"], ex["N"], ex["B"], ex["Nc"], ex["cmax"]
    i32 = lambda t: t.to(torch.int32)
    dh_up, dx_up, dp_up = torch.randn(N, H, generator=gen), torch.randn(N, 3, generator=gen), 0.1 * torch.randn(P, H, generator=gen)
    s2 = ex["tape"][0][1]
    G = spec.Grads()
    rdh, rdx, rdp = spec.att_plus_bwd(G, W, "att0.", s2, geo, ex["inter"], cmax, dh_up, dx_up, dp_up)
    geo_dev = dict(Nc=Nc, B=B, c_off=i32(torch.from_numpy(geo["c_off"].astype(np.int64))), p_off=i32(torch.from_numpy(geo["p_off"].astype(np.int64))),
                   pair_base=i32(torch.from_numpy(geo["pair_base"].astype(np.int64))), node_cplx=i32(geo["cplx"]),
                   max_c=int(np.diff(geo["c_off"]).max()), max_p=int(np.diff(geo["p_off"]).max()))
    dh, dx, grads, dp = bw.att_plus_backward(_weights(W, "att0."), _att_plus_saved(s2, W, "att0."), geo_dev, i32(ex["inter"][0]),
                                             i32(ex["inter"][1]), cmax, dh_up.clone(), dx_up.clone(), dp_up.clone())
    assert rel_err(dh, rdh) < 1e-5 and rel_err(dx, rdx) < 1e-5 and rel_err(dp, rdp) < 1e-5, (rel_err(dh, rdh), rel_err(dx, rdx), rel_err(dp, rdp))
    assert set("att0." + k for k in grads) == set(G), set("att0." + k for k in grads) ^ set(G)
    gmax = max(float(t.abs().max()) for t in G.values())
    for k, v in grads.items():
        ref = G["att0." + k].reshape(-1)
        err = float((v.reshape(-1) - ref).abs().max())
        assert err < 1e-5 * float(ref.abs().max()) + 1e-7 * gmax, (k, err, float(ref.abs().max()))


def plus_stack_case(seed=7):
    """inputs of fabind_b200.backward.stack_backward_plus from the specification's forward on the FABind+ gradient golden"""
    from fabind_b200.layout import build_layout
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_plus_*.pt")))[0])
    gen = torch.Generator().manual_seed(seed)
    H, L = b.H.shape[1], cfg.n_layers
    P = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu").P_total
    gX, gH, gP = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen), 0.1 * torch.randn(P, H, generator=gen)
    ex = {}
    _, _, _, garena, gHin = spec.forward_backward_plus(sd, cfg, b, gX, gH, gP, export=ex)
    W, geo, N, B, Nc = ex["W"], ex["geo"], ex["N"], ex["B"], ex["Nc"]
    i32 = lambda t: t.to(torch.int32)
    geo_dev = dict(Nc=Nc, B=B, c_off=i32(torch.from_numpy(geo["c_off"].astype(np.int64))), p_off=i32(torch.from_numpy(geo["p_off"].astype(np.int64))),
                   pair_base=i32(torch.from_numpy(geo["pair_base"].astype(np.int64))), node_cplx=i32(geo["cplx"]),
                   max_c=int(np.diff(geo["c_off"]).max()), max_p=int(np.diff(geo["p_off"]).max()))
    weights = {"": _weights(W, ""), "out.": _weights(W, "out.")}
    tape = []
    for l in range(L):
        weights[f"gcl{l}."], weights[f"att{l}."] = _weights(W, f"gcl{l}."), _weights(W, f"att{l}.")
        s1, s2, s3 = ex["tape"][l]
        tape.append((_gcl_plus_saved(s1), _att_plus_saved(s2, W, f"att{l}."), dict(x=s3["x"], acc=s3["acc"])))
    top = dict(Hin=ex["Hin"], pc=ex["pc"], outer=ex["outer"], h_last=ex["h_last"], out_saved=_gcl_plus_saved(ex["s_out"]))
    edges = dict(ctx_row=i32(ex["ctx"][0]), ctx_col=i32(ex["ctx"][1]), int_row=i32(ex["inter"][0]), int_col=i32(ex["inter"][1]),
                 las_a=i32(ex["las"][0]), las_b=i32(ex["las"][1]))
    consts = dict(cmax=ex["cmax"], lcl=ex["lcl"], las_step=cfg.geometry_reg_step_size, xl=ex["xl"])
    permt = ex["permt"]
    return dict(weights=weights, tape=tape, top=top, geo=geo_dev, edges=edges, consts=consts, dH_out=gH[permt].contiguous(),
                dX_out=(gX[permt, 0] * ex["moves"][:, None]).contiguous(), dP_out=gP, W=W, garena=garena, gHin=gHin[permt],
                moves=ex["moves"], x_state=ex["x_state"])


def test_plus_stack_reverse_pass_matches_specification(monkeypatch):
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)
    _install_plus_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    case = plus_stack_case()
    grads, dHin = bw.stack_backward_plus(case["weights"], case["tape"], case["top"], case["geo"], case["edges"], case["consts"],
                                         case["dH_out"], case["dX_out"], case["dP_out"])
    check_stack(case, grads, dHin, 1e-4)


def test_plus_training_forward_and_reverse_close_the_loop(monkeypatch):
    """FABind+: stack_forward_train_plus -> stack_backward_plus on torch stand-ins == the specification (outputs, saved tensors, arena gradient)"""
    from fabind_b200 import backward as bw
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    _install_plus_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    monkeypatch.setattr(bw, "row_stats", lambda h, w=None: (h.sum(1), (h * h).sum(1), (h * w).sum(1) if w is not None else None))

    def fstats(A1, A2, A3, rn, a0, a1, D):
        mu = (A1 + rn * a0) / D
        var_raw = (A2 + 2 * rn * (A3 if A3 is not None else 0) + rn * rn * a1) / D - mu * mu
        return mu, var_raw, torch.rsqrt(var_raw.clamp(min=0) + 1e-5)
    monkeypatch.setattr(bw, "folded_stats_fwd", fstats)
    case = plus_stack_case()
    case["consts"]["n_pairs"] = int(case["dP_out"].shape[0])
    ex_moves, x_state = case["moves"], case["x_state"]
    L = len(case["tape"])
    X_out, H_out, pair, tape, top = bw.stack_forward_train_plus(case["weights"], case["top"]["Hin"], x_state, ex_moves, case["geo"], case["edges"],
                                                                case["consts"], L)
    for mine, ref in zip(tape + [(top["out_saved"],)], case["tape"] + [(case["top"]["out_saved"],)]):
        for sm, sr in zip(mine, ref):
            assert set(sm) == set(sr), set(sm) ^ set(sr)
            for k in sr:
                if k == "acr":
                    assert max(abs(a - b) for a, b in zip(sm[k], sr[k])) < 1e-6
                elif sr[k].dtype == torch.int32:
                    assert torch.equal(sm[k], sr[k]), k
                else:
                    assert rel_err(sm[k], sr[k]) < 1e-5, (k, rel_err(sm[k], sr[k]))
    grads, dHin = bw.stack_backward_plus(case["weights"], tape, top, case["geo"], case["edges"], case["consts"], case["dH_out"], case["dX_out"],
                                         case["dP_out"])
    check_stack(case, grads, dHin, 1e-4)


def test_plus_training_step_assembly(monkeypatch):
    """train.training_step on the FABind+ layout, CPU: stand-ins for the kernel wrappers, the FABind+ oracle as the two providers;
    parameter gradients of the UNMODIFIED FABind+ reference (tests/golden/grad_plus_*.pt)"""
    from fabind_b200 import backward as bw, train
    from fabind_b200.config import published_args_plus
    from fabind_b200.plus import EfficientMCAttModel as PlusModel
    from oracle import fabind_oracle as orc, fabind_plus_oracle as porc
    from test_formulation_cpu import _dense_pair
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    _install_plus_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    monkeypatch.setattr(bw, "row_stats", lambda h, w=None: (h.sum(1), (h * h).sum(1), (h * w).sum(1) if w is not None else None))
    monkeypatch.setattr(bw, "folded_stats_fwd", lambda A1, A2, A3, rn, a0, a1, D: (
        (A1 + rn * a0) / D, (A2 + 2 * rn * (A3 if A3 is not None else 0) + rn * rn * a1) / D - ((A1 + rn * a0) / D) ** 2,
        torch.rsqrt(((A2 + 2 * rn * (A3 if A3 is not None else 0) + rn * rn * a1) / D - ((A1 + rn * a0) / D) ** 2).clamp(min=0) + 1e-5)))
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_plus_*.pt"))):
        g, r, b, sd, cfg = load_golden(path)
        H = r["hidden"]
        model = PlusModel(published_args_plus(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                          normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        model.load_state_dict(sd, strict=True)
        fa = b.forward_args()
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)
        dims = [(int(b.n_p[i]) + 1, int(b.n_c[i]) + 1) for i in range(len(b.n_c))]
        rp_holder = {}

        def prev_coords(m, fa):
            if cfg.n_iter <= 1:
                return fa["X"].clone()
            from types import SimpleNamespace
            c = SimpleNamespace(**{**vars(cfg), "n_iter": cfg.n_iter - 1})
            with torch.no_grad():
                return forward_emulated(sd, c, b, flavour=1)[0]

        def edge_lists(m, X_prev, fa):
            ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                            cfg.inter_cutoff / cfg.coordinate_scale)
            return ctx, inter

        def output_grads(X, Hh, pair):
            dense = _dense_pair(pair, dims, H)
            rp = torch.randn(dense.shape, generator=gen) * 0.1
            rp_holder["loss"] = float((X * rx).sum() + (Hh * rh).sum() + (dense * rp).sum())
            return rx, rh, torch.cat([rp[i, :n, :c].reshape(-1, H) for i, (n, c) in enumerate(dims)])
        X_out, H_out, pair, pgrads, gH_in = train.training_step(model, fa, output_grads, prev_coords=prev_coords, edge_lists=edge_lists)
        assert abs(rp_holder["loss"] - g["loss"]) < 1e-4 * abs(g["loss"]), (rp_holder["loss"], g["loss"])
        gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
        n = 0
        for k, ref in g["grads"].items():
            if ref is None:
                continue
            err = float((pgrads[k] - ref).abs().max())
            assert err < 5e-4 * float(ref.abs().max()) + 5e-7 * gmax, (k, err, float(ref.abs().max()))
            n += 1
        assert n >= 80


def test_forward_with_grad_trains_the_drop_in_module(monkeypatch):
    """train.forward_with_grad: an unchanged training loop (loss.backward() on the module's outputs) puts the UNMODIFIED reference's
    gradients on the drop-in module's parameters and on the incoming node features (CPU: stand-ins + oracle providers)"""
    from fabind_b200 import EfficientMCAttModel, backward as bw, train
    from fabind_b200.config import published_args
    from oracle import fabind_oracle as orc
    _install_standins(monkeypatch, bw)
    _install_forward_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", _gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", _outer_bwd_standin)
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))[0])
    H = r["hidden"]
    model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                                normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    model.load_state_dict(sd, strict=True)
    fa = b.forward_args()
    fa["H"] = fa["H"].clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(r["readout_seed"])
    rx, rh = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)

    def prev_coords(m, fa):
        c = orc.make_cfg(n_layers=cfg.n_layers, n_iter=cfg.n_iter - 1)
        with torch.no_grad():
            return orc.model_forward(sd, c, fa["X"], fa["H"].detach(), fa["batch_id"], fa["segment_id"], fa["mask"], fa["is_global"],
                                     fa["compound_edge_index"], fa["LAS_edge_index"], fa["batched_complex_coord_LAS"])[0]

    def edge_lists(m, X_prev, fa):
        ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                        cfg.inter_cutoff / cfg.coordinate_scale)
        return ctx, inter
    X, Hh = train.forward_with_grad(model, fa, prev_coords=prev_coords, edge_lists=edge_lists)
    loss = (X * rx).sum() + (Hh * rh).sum()
    assert abs(float(loss) - g["loss"]) < 1e-4 * abs(g["loss"])
    loss.backward()
    params = dict(model.named_parameters())
    gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
    n = 0
    for k, ref in g["grads"].items():
        if ref is None:
            continue
        err = float((params[k].grad - ref).abs().max())
        assert err < 5e-4 * float(ref.abs().max()) + 5e-7 * gmax, (k, err, float(ref.abs().max()))
        n += 1
    assert n >= 80 and fa["H"].grad is not None and float(fa["H"].grad.abs().max()) > 0


class _MarshalCheckLib:
    """stands in for the loaded library: every call is checked against the REAL ctypes prototype (argument count, convertibility of
    each argument) and returns 0 without touching memory -- catches marshalling mistakes of the Python wrappers without a GPU"""

    def __init__(self, real):
        self._real, self.calls = real, {}

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        argtypes = fn.argtypes

        def call(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments for {len(argtypes)} parameters"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    t.from_param(a)
                except Exception as e:          # noqa: BLE001
                    raise AssertionError(f"{name}: argument {i} ({type(a).__name__}) does not convert to {t.__name__}: {e}")
            self.calls[name] = self.calls.get(name, 0) + 1
            return 0
        return call


def test_kernel_wrappers_marshal_their_arguments(monkeypatch):
    """the REAL wrappers of fabind_b200/backward.py (no stand-ins) over a library double that validates every call against the ctypes
    prototypes: the training-mode forward and the reverse pass of both layouts issue well-formed calls for every entry point they use"""
    import ctypes
    from fabind_b200 import _lib, backward as bw
    v1 = stack_case(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))[0])      # (built with the real library: slot tables)
    pl = plus_stack_case()
    fake = _MarshalCheckLib(_lib.lib())
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(bw, "_chk", lambda t, dtype=torch.float32: (_ for _ in ()).throw(AssertionError(f"dtype {t.dtype} != {dtype}"))
                        if t.dtype != dtype or not t.is_contiguous() else t)
    monkeypatch.setattr(bw, "_st", lambda t: ctypes.c_void_p(0))
    X, Hh, tape, top = bw.stack_forward_train_v1(v1["weights"], v1["top"]["Hin"], v1["x_state"], v1["moves"], v1["geo"], v1["edges"], v1["consts"],
                                                 v1["L"])
    bw.stack_backward_v1(v1["weights"], v1["tape"], v1["top"], v1["geo"], v1["edges"], v1["consts"], v1["dH_out"], v1["dX_out"])
    pl["consts"]["n_pairs"] = int(pl["dP_out"].shape[0])
    for layer in pl["tape"]:
        layer[1]["acr"] = tuple(layer[1]["acr"])
    # the forward double leaves outputs uninitialised; ac_r is read back on the host, keep it finite
    bw.stack_forward_train_plus(pl["weights"], pl["top"]["Hin"], pl["x_state"], pl["moves"], pl["geo"], pl["edges"], pl["consts"], len(pl["tape"]))
    bw.stack_backward_plus(pl["weights"], pl["tape"], pl["top"], pl["geo"], pl["edges"], pl["consts"], pl["dH_out"], pl["dX_out"], pl["dP_out"])
    used = set(fake.calls)
    expected = {"fb_gemm", "fb_act_fwd", "fb_act_bwd", "fb_outer_act_bwd", "fb_colsum", "fb_rowdot", "fb_rowdot2", "fb_rows_update", "fb_vec_op",
                "fb_scatter_add_rows", "fb_gather_add_rows", "fb_gemm_wgrad", "fb_coord_step_bwd", "fb_radial_bwd", "fb_las_bwd",
                "fb_softmax_seg_bwd", "fb_row_attention_bwd", "fb_pair_bias_gate_bwd", "fb_pair_outer_bwd", "fb_radial_fwd", "fb_coord_apply",
                "fb_softmax_seg_fwd", "fb_las_acc", "fb_pair_outer_fwd", "fb_pair_bias_gate_fwd", "fb_row_attention_fwd", "fb_layernorm",
                "fb_layernorm_bwd", "fb_row_stats", "fb_row_stats_bwd", "fb_folded_stats_fwd", "fb_folded_stats_bwd"}
    assert expected <= used, expected - used
