"""Hand-derived backward of the launch sequence (FABind v1 layout, FABind+ layout further down) -- TEST INFRASTRUCTURE,
specification of the training kernels (BASELINE config 5).

`tests/emulate_packed.py` mirrors the forward launch sequence of csrc/forward.cu on the packed weight arena;
this file is its explicit reverse pass: every function below is ONE planned backward launch (or one GEMM pair
`dX = dY W`, `dW = dY^T X`), written without autograd, and it lists what the training-mode forward has to keep.
It is pinned two ways (tests/test_formulation_cpu.py):
  * against autograd through the emulation (gradient w.r.t. every arena slot and w.r.t. the node features), and
  * through the differentiable weight packing against parameter gradients of the UNMODIFIED reference
    (tests/golden/grad_v1_*.pt, grad_plus_*.pt).
Reference semantics that shape it (refine_coord, att_model.py:227-236): only the LAST refinement iteration carries
gradients; edges are rebuilt under no_grad, so no gradient flows through the graph construction; the coordinates
entering the last iteration are constants.

Saved by the training-mode forward of the last iteration, per sub-layer (what the CUDA kernels must store, all
other intermediates are recomputed from these):
  gcl : h_in, x_in, Z1 (pre-activation of edge_mlp.0, [E_ctx,H]), Z2 ([E_ctx,H]), Z3 ([E_ctx,H]), Z4 ([N,H]), the
        unclamped coordinate step [N,3]                         (A1, M, T3, t1, agg are SiLU / segment sums of these)
  att : h_in, x_in, CAc/CAp/CAp2 (stacked projections), attention probabilities are RECOMPUTED (row softmax of
        <=201 keys), hp1, hc1, transition pre-activations, h2, QK (stacked), Zp (pair hidden on unique inter pairs),
        logit max / sum per row (alpha recomputed), zc ([E_int,H]), the unclamped coordinate step
  LAS : x_in and the unclamped step.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from emulate_packed import Arena, _edges, forward_emulated, HD
from fabind_b200.layout import build_layout


class Grads(dict):
    """gradient accumulator keyed by arena slot name"""

    def add(self, name, t):
        self[name] = self[name] + t if name in self else t.clone()


def lin_bwd(G, W, wname, bname, x, dy):
    """y = x W^T (+ b):  dW += dy^T x (one GEMM, reduction over rows), db += sum_rows dy, returns dx = dy W (one GEMM)"""
    G.add(wname, dy.t() @ x)
    if bname is not None:
        G.add(bname, dy.sum(0))
    return dy @ W.m(wname)


def silu_bwd(z, dy):
    s = torch.sigmoid(z)
    return dy * (s * (1 + z * (1 - s)))


def radial_fwd(row, col, x, cplx, B):
    """coord2radial with the per-complex norm (egnn.py:767-787)"""
    d = x[row] - x[col]
    d2 = (d * d).sum(1)
    nrm = torch.zeros(B).index_add_(0, cplx[row], d2 * d2).sqrt()
    return d2 / nrm[cplx[row]], (d, d2, nrm)


def radial_bwd(row, col, cplx, B, N, saved, drn):
    """rn_e = d2_e / sqrt(sum_e' d2_e'^2):  dd2_e = drn_e/nrm - d2_e (sum_e' drn_e' d2_e') / nrm^3 ; d2 = |x_r - x_c|^2"""
    d, d2, nrm = saved
    eb = cplx[row]
    dot = torch.zeros(B).index_add_(0, eb, drn * d2)
    dd2 = drn / nrm[eb] - d2 * dot[eb] / nrm[eb] ** 3
    g = 2 * d * dd2[:, None]
    return torch.zeros(N, 3).index_add_(0, row, g).index_add_(0, col, -g)


# ------------------------------------------------------------------------------------------------- MC_E_GCL
def gcl_fwd(W, pre, h, x, ctx, cplx, B, cmax):
    r, c = ctx
    N, H = h.shape
    rn, rs = radial_fwd(r, c, x, cplx, B)
    Pn = F.linear(h, W.m(pre + "e1_rc"))
    Z1 = Pn[r, :H] + Pn[c, H:] + rn[:, None] * W.m(pre + "e1_rad") + W.m(pre + "e1_b")
    A1 = F.silu(Z1)
    Z2 = F.linear(A1, W.m(pre + "e2_w"), W.m(pre + "e2_b"))
    M = F.silu(Z2)
    Z3 = F.linear(M, W.m(pre + "c1_w"), W.m(pre + "c1_b"))
    T3 = F.silu(Z3)
    s = T3 @ W.m(pre + "c2_w")
    deg = torch.zeros(N).index_add_(0, r, torch.ones(r.numel())).clamp(min=1)
    d = x[r] - x[c]
    step = torch.zeros(N, 3).index_add_(0, r, d * s[:, None]) / deg[:, None]
    x_new = x + step.clamp(-cmax, cmax)
    agg = torch.zeros(N, H).index_add_(0, r, M)
    cat = torch.cat([h, agg], 1)
    Z4 = F.linear(cat, W.m(pre + "n1_w"), W.m(pre + "n1_b"))
    t1 = F.silu(Z4)
    h_new = h + F.linear(t1, W.m(pre + "n2_w"), W.m(pre + "n2_b"))
    return h_new, x_new, dict(h=h, x=x, rn=rn, rs=rs, Z1=Z1, A1=A1, Z2=Z2, M=M, Z3=Z3, T3=T3, s=s, deg=deg, d=d, step=step,
                              cat=cat, Z4=Z4, t1=t1)


def gcl_bwd(G, W, pre, sv, ctx, cplx, B, cmax, dh_new, dx_new):
    r, c = ctx
    N, H = sv["h"].shape
    # coordinate branch: x_new = x + clamp(mean_e d_e s_e)
    dx = dx_new.clone()
    dstep = dx_new * (sv["step"].abs() <= cmax)
    de = dstep[r] / sv["deg"][r][:, None]
    ds = (de * sv["d"]).sum(1)
    dd = de * sv["s"][:, None]
    dx.index_add_(0, r, dd).index_add_(0, c, -dd)
    G.add(pre + "c2_w", sv["T3"].t() @ ds)
    dZ3 = silu_bwd(sv["Z3"], ds[:, None] * W.m(pre + "c2_w"))
    dM = lin_bwd(G, W, pre + "c1_w", pre + "c1_b", sv["M"], dZ3)
    # node branch: h_new = h + n2(silu(n1([h | sum_e M])))
    dh = dh_new.clone()
    dt1 = lin_bwd(G, W, pre + "n2_w", pre + "n2_b", sv["t1"], dh_new)
    dcat = lin_bwd(G, W, pre + "n1_w", pre + "n1_b", sv["cat"], silu_bwd(sv["Z4"], dt1))
    dh += dcat[:, :H]
    dM = dM + dcat[:, H:][r]                              # gather of the aggregate's gradient back to the edges
    # edge MLP
    dA1 = lin_bwd(G, W, pre + "e2_w", pre + "e2_b", sv["A1"], silu_bwd(sv["Z2"], dM))
    dZ1 = silu_bwd(sv["Z1"], dA1)
    G.add(pre + "e1_b", dZ1.sum(0))
    G.add(pre + "e1_rad", (dZ1 * sv["rn"][:, None]).sum(0))
    drn = dZ1 @ W.m(pre + "e1_rad")
    dPn = torch.zeros(N, 2 * H)
    dPn[:, :H].index_add_(0, r, dZ1)                      # segment sums over rows / over columns of the edge list
    dPn[:, H:].index_add_(0, c, dZ1)
    dh += lin_bwd(G, W, pre + "e1_rc", None, sv["h"], dPn)
    dx += radial_bwd(r, c, cplx, B, N, sv["rs"], drn)
    return dh, dx


# ------------------------------------------------------------------------------------------------- LAS step
def las_fwd(x, xl, las, step_size, lcl):
    a, b = las
    d = x[a] - x[b]
    diff = (d * d).sum(1) - ((xl[a] - xl[b]) ** 2).sum(1)
    acc = torch.zeros_like(x).index_add_(0, b, 4 * diff[:, None] * d) * step_size
    return x + acc.clamp(-lcl, lcl), dict(x=x, d=d, diff=diff, acc=acc)


def las_bwd(sv, las, step_size, lcl, dx_new):
    a, b = las
    dforce = (dx_new * (sv["acc"].abs() <= lcl) * step_size)[b]
    ddiff = 4 * (dforce * sv["d"]).sum(1)
    dd = 4 * sv["diff"][:, None] * dforce + ddiff[:, None] * 2 * sv["d"]
    return dx_new.clone().index_add_(0, a, dd).index_add_(0, b, -dd)


# ------------------------------------------------------------------------------------------------- row attention
def rowatt_fwd(q, g, k, v, bias):
    I, J = q.shape[0], k.shape[0]
    qh, kh, vh = q.view(I, 4, 32) / math.sqrt(32), k.view(J, 4, 32), v.view(J, 4, 32)
    a = torch.softmax(torch.einsum("ihd,jhd->hij", qh, kh) + bias.permute(2, 0, 1), -1)
    o = torch.einsum("hij,jhd->ihd", a, vh)
    sg = torch.sigmoid(g).view(I, 4, 32)
    return (o * sg).reshape(I, 128), (qh, kh, vh, a, o, sg)


def rowatt_bwd(sv, dout):
    qh, kh, vh, a, o, sg = sv
    I, J = qh.shape[0], kh.shape[0]
    dov = dout.view(I, 4, 32)
    dg = (dov * o * sg * (1 - sg)).reshape(I, 128)
    do = dov * sg
    da = torch.einsum("ihd,jhd->hij", do, vh)
    dv = torch.einsum("hij,ihd->jhd", a, do).reshape(J, 128)
    dl = a * (da - (da * a).sum(-1, keepdim=True))
    dq = (torch.einsum("hij,jhd->ihd", dl, kh) / math.sqrt(32)).reshape(I, 128)
    dk = torch.einsum("hij,ihd->jhd", dl, qh).reshape(J, 128)
    return dq, dg, dk, dv, dl.permute(1, 2, 0)            # dbias [I, J, 4]


# ------------------------------------------------------------------------------------------------- MC_Att_L
def att_fwd(W, pre, l, h, x, geo, P0, PB, inter, cmax):
    Nc, B, c_off, p_off, pair_base, cplx = geo["Nc"], geo["B"], geo["c_off"], geo["p_off"], geo["pair_base"], geo["cplx"]
    int_r, int_c = inter
    N, H = h.shape
    hc0, hp0 = h[:Nc], h[Nc:]
    CAc = F.linear(hc0, W.m(pre + "ca_c_w"), W.m(pre + "ca_c_b"))
    CAp = F.linear(hp0, W.m(pre + "ca_p_w"), W.m(pre + "ca_p_b"))
    blocks = []
    for b in range(B):
        cs, ps = slice(c_off[b], c_off[b + 1]), slice(p_off[b] - Nc, p_off[b + 1] - Nc)
        blocks.append((cs, ps, c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b], slice(pair_base[b], pair_base[b + 1])))
    Op, svp = [], []
    for cs, ps, nc1, np1, pr in blocks:
        o, sv = rowatt_fwd(CAp[ps, :HD], CAp[ps, HD:], CAc[cs, :HD], CAc[cs, HD:2 * HD], PB[pr, l, 0].view(np1, nc1, 4))
        Op.append(o); svp.append(sv)
    Op = torch.cat(Op)
    hp1 = hp0 + F.linear(Op, W.m(pre + "o_p_w"), W.m(pre + "o_p_b"))
    CAp2 = F.linear(hp1, W.m(pre + "ca_p2_w"))
    Oc, svc = [], []
    for cs, ps, nc1, np1, pr in blocks:
        o, sv = rowatt_fwd(CAc[cs, 2 * HD:3 * HD], CAc[cs, 3 * HD:], CAp2[ps, :HD], CAp2[ps, HD:],
                           PB[pr, l, 1].view(np1, nc1, 4).transpose(0, 1))
        Oc.append(o); svc.append(sv)
    Oc = torch.cat(Oc)
    hc1 = hc0 + F.linear(Oc, W.m(pre + "o_c_w"), W.m(pre + "o_c_b"))
    Tp = F.relu(F.linear(hp1, W.m(pre + "tp1_w"), W.m(pre + "tp1_b")))
    hp2 = hp1 + F.linear(Tp, W.m(pre + "tp2_w"), W.m(pre + "tp2_b"))
    Tc = F.relu(F.linear(hc1, W.m(pre + "tc1_w"), W.m(pre + "tc1_b")))
    hc2 = hc1 + F.linear(Tc, W.m(pre + "tc2_w"), W.m(pre + "tc2_b"))
    h2 = torch.cat([hc2, hp2])
    QK = F.linear(h2, W.m(pre + "qk_w"), W.m(pre + "qk_b"))
    pc32 = torch.empty(N, 32)
    pc32[Nc:] = QK[Nc:, 2 * H:2 * H + 32]
    pc32[:Nc] = QK[:Nc, 2 * H + 32:2 * H + 64]
    eb = cplx[int_r]
    is_c = int_r < Nc
    ci, pi = torch.where(is_c, int_r, int_c), torch.where(is_c, int_c, int_r)
    c_off_t, p_off_t = torch.from_numpy(c_off.astype(np.int64)), torch.from_numpy(p_off.astype(np.int64))
    nc1_t = c_off_t[1:] - c_off_t[:-1]
    pair = torch.from_numpy(pair_base.astype(np.int64))[eb] + (pi - p_off_t[eb]) * nc1_t[eb] + (ci - c_off_t[eb])
    u = is_c.nonzero().squeeze(1)
    zcat = torch.cat([P0[pair[u]], pc32[pi[u]] * pc32[ci[u]], torch.zeros(u.numel(), 32)], 1)
    Zp = F.linear(zcat, W.m(pre + "pt1_w"), W.m(pre + "pt1_b"))
    Rp = F.relu(Zp)
    pbu = Rp @ W.m(pre + "pt2v") + W.m(pre + "pt_c")
    pb_dense = torch.zeros(P0.shape[0])
    pb_dense[pair[u]] = pbu
    rn, rs = radial_fwd(int_r, int_c, x, cplx, B)
    V, VC = QK[:, 2 * H + 128:3 * H + 128], QK[:, 3 * H + 128:]
    q = QK[int_r, :H]
    kk = QK[int_c, H:2 * H] + rn[:, None] * W.m(pre + "k_r")
    logit = (q * kk).sum(1) + pb_dense[pair]
    mx = torch.full((N,), float("-inf")).scatter_reduce(0, int_r, logit, reduce="amax")
    e = (logit - mx[int_r]).exp()
    alpha = e / torch.zeros(N).index_add_(0, int_r, e)[int_r]
    ve = V[int_c] + rn[:, None] * W.m(pre + "v_r")
    h3 = h2 + torch.zeros(N, H).index_add_(0, int_r, alpha[:, None] * ve)
    zc = VC[int_c] + rn[:, None] * W.m(pre + "ac_u") + W.m(pre + "ac1_b")
    sc = F.silu(zc)
    se = sc @ W.m(pre + "ac2_w")
    d = x[int_r] - x[int_c]
    step = torch.zeros(N, 3).index_add_(0, int_r, d * (alpha * se)[:, None])
    x_new = x + step.clamp(-cmax, cmax)
    sv = dict(hc0=hc0, hp0=hp0, blocks=blocks, svp=svp, svc=svc, Op=Op, Oc=Oc, hp1=hp1, hc1=hc1, Tp=Tp, Tc=Tc, h2=h2, QK=QK,
              CAc=CAc, CAp=CAp, CAp2=CAp2, x=x, h_in=h, pc32=pc32, pi=pi, ci=ci, pair=pair, u=u, zcat=zcat, Zp=Zp, Rp=Rp, rn=rn, rs=rs, q=q, kk=kk, alpha=alpha, ve=ve,
              zc=zc, sc=sc, se=se, d=d, step=step)
    return h3, x_new, sv


def att_bwd(G, W, pre, l, sv, geo, inter, cmax, dh3, dx_new, dP0, dPB):
    Nc, B, cplx = geo["Nc"], geo["B"], geo["cplx"]
    int_r, int_c = inter
    N, H = sv["h2"].shape
    rn, alpha, se = sv["rn"], sv["alpha"], sv["se"]
    dQK = torch.zeros_like(sv["QK"])
    # interfacial coordinate update: x_new = x + clamp(sum_e d_e alpha_e se_e)
    dx = dx_new.clone()
    de = (dx_new * (sv["step"].abs() <= cmax))[int_r]
    dw = (de * sv["d"]).sum(1)
    dd = de * (alpha * se)[:, None]
    dx.index_add_(0, int_r, dd).index_add_(0, int_c, -dd)
    dalpha, dse = dw * se, dw * alpha
    G.add(pre + "ac2_w", sv["sc"].t() @ dse)
    dzc = silu_bwd(sv["zc"], dse[:, None] * W.m(pre + "ac2_w"))
    G.add(pre + "ac1_b", dzc.sum(0))
    G.add(pre + "ac_u", (dzc * rn[:, None]).sum(0))
    drn = dzc @ W.m(pre + "ac_u")
    dQK[:, 3 * H + 128:].index_add_(0, int_c, dzc)
    # interfacial aggregation: h3 = h2 + sum_e alpha_e ve_e
    dh2 = dh3.clone()
    dagg = dh3[int_r]
    dalpha = dalpha + (dagg * sv["ve"]).sum(1)
    dve = dagg * alpha[:, None]
    dQK[:, 2 * H + 128:3 * H + 128].index_add_(0, int_c, dve)
    G.add(pre + "v_r", (dve * rn[:, None]).sum(0))
    drn = drn + dve @ W.m(pre + "v_r")
    # segment softmax over the destination row
    dlogit = alpha * (dalpha - torch.zeros(N).index_add_(0, int_r, alpha * dalpha)[int_r])
    dq, dkk = dlogit[:, None] * sv["kk"], dlogit[:, None] * sv["q"]
    dQK[:, :H].index_add_(0, int_r, dq)
    dQK[:, H:2 * H].index_add_(0, int_c, dkk)
    G.add(pre + "k_r", (dkk * rn[:, None]).sum(0))
    drn = drn + dkk @ W.m(pre + "k_r")
    # pair bias on the unique interface pairs (both directions of a pair share one value)
    pair, u = sv["pair"], sv["u"]
    dpbu = torch.zeros(dP0.shape[0]).index_add_(0, pair, dlogit)[pair[u]]
    G.add(pre + "pt_c", dpbu.sum().reshape(1))
    G.add(pre + "pt2v", sv["Rp"].t() @ dpbu)
    dZp = (dpbu[:, None] * W.m(pre + "pt2v")) * (sv["Zp"] > 0)
    dz = lin_bwd(G, W, pre + "pt1_w", pre + "pt1_b", sv["zcat"], dZp)
    dP0.index_add_(0, pair[u], dz[:, :H])
    dt = dz[:, H:H + 32]
    pc32, pi, ci = sv["pc32"], sv["pi"], sv["ci"]
    dpc32 = torch.zeros(N, 32).index_add_(0, pi[u], dt * pc32[ci[u]]).index_add_(0, ci[u], dt * pc32[pi[u]])
    dQK[Nc:, 2 * H:2 * H + 32] += dpc32[Nc:]
    dQK[:Nc, 2 * H + 32:2 * H + 64] += dpc32[:Nc]
    dx += radial_bwd(int_r, int_c, cplx, B, N, sv["rs"], drn)
    dh2 += lin_bwd(G, W, pre + "qk_w", pre + "qk_b", sv["h2"], dQK)
    dhc2, dhp2 = dh2[:Nc], dh2[Nc:]
    # transitions
    dTc = lin_bwd(G, W, pre + "tc2_w", pre + "tc2_b", sv["Tc"], dhc2)
    dhc1 = dhc2 + lin_bwd(G, W, pre + "tc1_w", pre + "tc1_b", sv["hc1"], dTc * (sv["Tc"] > 0))
    dTp = lin_bwd(G, W, pre + "tp2_w", pre + "tp2_b", sv["Tp"], dhp2)
    dhp1 = dhp2 + lin_bwd(G, W, pre + "tp1_w", pre + "tp1_b", sv["hp1"], dTp * (sv["Tp"] > 0))
    # compound-side row attention (keys / values from the UPDATED protein side)
    dOc = lin_bwd(G, W, pre + "o_c_w", pre + "o_c_b", sv["Oc"], dhc1)
    dhc0 = dhc1.clone()
    dCAc = torch.zeros(Nc, 4 * HD)
    dCAp2 = torch.zeros(N - Nc, 2 * HD)
    for (cs, ps, nc1, np1, pr), s in zip(sv["blocks"], sv["svc"]):
        dq_, dg_, dk_, dv_, db_ = rowatt_bwd(s, dOc[cs])
        dCAc[cs, 2 * HD:3 * HD], dCAc[cs, 3 * HD:] = dq_, dg_
        dCAp2[ps, :HD], dCAp2[ps, HD:] = dk_, dv_
        dPB[pr, l, 1] += db_.transpose(0, 1).reshape(-1, 4)
    dhp1 = dhp1 + lin_bwd(G, W, pre + "ca_p2_w", None, sv["hp1"], dCAp2)
    # protein-side row attention
    dOp = lin_bwd(G, W, pre + "o_p_w", pre + "o_p_b", sv["Op"], dhp1)
    dhp0 = dhp1.clone()
    dCAp = torch.zeros(N - Nc, 2 * HD)
    for (cs, ps, nc1, np1, pr), s in zip(sv["blocks"], sv["svp"]):
        dq_, dg_, dk_, dv_, db_ = rowatt_bwd(s, dOp[ps])
        dCAp[ps, :HD], dCAp[ps, HD:] = dq_, dg_
        dCAc[cs, :HD], dCAc[cs, HD:2 * HD] = dk_, dv_
        dPB[pr, l, 0] += db_.reshape(-1, 4)
    dhc0 += lin_bwd(G, W, pre + "ca_c_w", pre + "ca_c_b", sv["hc0"], dCAc)
    dhp0 += lin_bwd(G, W, pre + "ca_p_w", pre + "ca_p_b", sv["hp0"], dCAp)
    return torch.cat([dhc0, dhp0]), dx


# ------------------------------------------------------------------------------------------------- whole step
def forward_backward_v1(sd, cfg, batch, gX, gH, arena=None, export=None):
    """Forward (all iterations) + explicit backward of the last one for loss = <X_out, gX> + <H_out, gH>.
    Returns (X_out, H_out, grad of the flat arena, grad of batch.H).  export: optional dict that receives the internals of the
    last iteration (weights, geometry, edge lists, per-layer saved tensors) for tests of GPU-side orchestration."""
    from types import SimpleNamespace
    H = batch.H.shape[1]
    L = cfg.n_layers
    W = Arena(sd, H, L, 0, False, arena)
    with torch.no_grad():
        lay = build_layout(batch.batch_id, batch.segment_id, batch.is_global, batch.mask, "cpu")
        o, blob = lay.offs, lay.blob.numpy()
        N, B, Nc = lay.N, lay.B, lay.Nc_tot
        perm, inv = blob[o["perm"]:o["perm"] + N], blob[o["inv"]:o["inv"] + N]
        lay_np = dict(node_cplx=blob[o["node_cplx"]:o["node_cplx"] + N], flags=lay.flags.numpy(),
                      c_off=blob[o["c_off"]:o["c_off"] + B + 1], p_off=blob[o["p_off"]:o["p_off"] + B + 1])
        geo = dict(Nc=Nc, B=B, c_off=lay_np["c_off"], p_off=lay_np["p_off"], pair_base=blob[o["pair_base"]:o["pair_base"] + B + 1],
                   cplx=torch.from_numpy(lay_np["node_cplx"].astype(np.int64)))
        permt = torch.from_numpy(perm.astype(np.int64))
        Hin = batch.H[permt]
        xl = batch.X_LAS[permt, 0]
        bonds_int = inv[batch.compound_edge_index.numpy()]
        las = tuple(torch.from_numpy(inv[batch.LAS_edge_index.numpy()].astype(np.int64)))
        moves = torch.from_numpy((lay_np["flags"] & 4) != 0)
        intra, inter_cut = cfg.intra_cutoff / cfg.coordinate_scale, cfg.inter_cutoff / cfg.coordinate_scale
        cmax, lcl = 10.0 / cfg.coordinate_scale, 15.0 / cfg.coordinate_scale
        # iterations 0 .. n_iter-2: forward only (the product path: fb_model_forward as it is)
        if cfg.n_iter > 1:
            Xprev = forward_emulated(sd, SimpleNamespace(**{**vars(cfg), "n_iter": cfg.n_iter - 1}), batch, arena=W.a.detach())[0]
        else:
            Xprev = batch.X
        x_state = Xprev[permt, 0].clone()
        ctx, inter = _edges(x_state, lay_np, intra, inter_cut, bonds_int)
        if inter[0].numel() == 0:
            inter = (torch.tensor([lay.fb_atom, lay.fb_res]), torch.tensor([lay.fb_res, lay.fb_atom]))
        c_off, p_off = geo["c_off"], geo["p_off"]

        # ---- forward of the last iteration, keeping what the reverse pass needs
        pc = torch.empty(N, H)
        pc[:Nc] = F.linear(Hin[:Nc], W.m("il_c_w"), W.m("il_c_b"))
        pc[Nc:] = F.linear(Hin[Nc:], W.m("il_p_w"), W.m("il_p_b"))
        outer = [(pc[p_off[b]:p_off[b + 1], None, :] * pc[None, c_off[b]:c_off[b + 1], :]).reshape(-1, H) for b in range(B)]
        outer = torch.cat(outer)
        P0 = F.linear(outer, W.m("il_o_w"), W.m("il_o_b"))
        raw_full = F.linear(P0, W.m("pb_w"), W.m("pb_b"))
        raw = raw_full[:, :16 * L].reshape(-1, L, 2, 2, 4)
        sig = torch.sigmoid(raw[:, :, :, 1])
        PB = raw[:, :, :, 0] * sig
        h = F.linear(Hin, W.m("in_w"), W.m("in_b"))
        x = x_state.clone()
        tape = []
        for l in range(L):
            h, x, s1 = gcl_fwd(W, f"gcl{l}.", h, x, ctx, geo["cplx"], B, cmax)
            h, x, s2 = att_fwd(W, f"att{l}.", l, h, x, geo, P0, PB, inter, cmax)
            x, s3 = las_fwd(x, xl, las, cfg.geometry_reg_step_size, lcl)
            tape.append((s1, s2, s3))
        h_last, x, s_out = gcl_fwd(W, "out.", h, x, ctx, geo["cplx"], B, cmax)
        h_final = F.linear(h_last, W.m("out_w"), W.m("out_b"))
        x_out = torch.where(moves[:, None], x, x_state)
        X_out = torch.empty_like(batch.X)
        X_out[permt, 0] = x_out
        H_out = torch.empty_like(batch.H)
        H_out[permt] = h_final

        if export is not None:
            export.update(W=W, geo=geo, ctx=ctx, inter=inter, las=las, tape=tape, s_out=s_out, P0=P0, PB=PB, cmax=cmax, lcl=lcl,
                          xl=xl, N=N, B=B, Nc=Nc, Hin=Hin, pc=pc, outer=outer, raw_full=raw_full, h_last=h_last, permt=permt,
                          moves=moves, x_state=x_state, x_out=x_out, h_final=h_final)
        # ---- reverse pass
        G = Grads()
        dx = gX[permt, 0] * moves[:, None]
        dh = lin_bwd(G, W, "out_w", "out_b", h_last, gH[permt])
        dh, dx = gcl_bwd(G, W, "out.", s_out, ctx, geo["cplx"], B, cmax, dh, dx)
        dP0 = torch.zeros_like(P0)
        dPB = torch.zeros_like(PB)
        for l in reversed(range(L)):
            s1, s2, s3 = tape[l]
            dx = las_bwd(s3, las, cfg.geometry_reg_step_size, lcl, dx)
            dh, dx = att_bwd(G, W, f"att{l}.", l, s2, geo, inter, cmax, dh, dx, dP0, dPB)
            dh, dx = gcl_bwd(G, W, f"gcl{l}.", s1, ctx, geo["cplx"], B, cmax, dh, dx)
        dHin = lin_bwd(G, W, "in_w", "in_b", Hin, dh)
        # gated pair biases of every row-attention block, then pair_embed0 = il_o(p (x) c)
        draw = torch.zeros_like(raw)
        draw[:, :, :, 0] = dPB * sig
        draw[:, :, :, 1] = dPB * raw[:, :, :, 0] * sig * (1 - sig)
        draw_full = torch.zeros_like(raw_full)
        draw_full[:, :16 * L] = draw.reshape(-1, 16 * L)
        dP0 += lin_bwd(G, W, "pb_w", "pb_b", P0, draw_full)
        douter = lin_bwd(G, W, "il_o_w", "il_o_b", outer, dP0)
        dpc = torch.zeros(N, H)
        pb = geo["pair_base"]
        for b in range(B):
            np1, nc1 = p_off[b + 1] - p_off[b], c_off[b + 1] - c_off[b]
            t = douter[pb[b]:pb[b + 1]].view(np1, nc1, H)
            dpc[p_off[b]:p_off[b + 1]] += (t * pc[None, c_off[b]:c_off[b + 1]]).sum(1)
            dpc[c_off[b]:c_off[b + 1]] += (t * pc[p_off[b]:p_off[b + 1], None]).sum(0)
        dHin[:Nc] += lin_bwd(G, W, "il_c_w", "il_c_b", Hin[:Nc], dpc[:Nc])
        dHin[Nc:] += lin_bwd(G, W, "il_p_w", "il_p_b", Hin[Nc:], dpc[Nc:])
        garena = torch.zeros_like(W.a)
        for name, g in G.items():
            r, c, off = W.s[name]
            garena[off:off + r * c] = g.reshape(-1)
        gH_in = torch.empty_like(batch.H)
        gH_in[permt] = dHin
    return X_out, H_out, garena, gH_in


# =================================================================================================
# FABind+ layout (LayerNorm MLPs folded through the node-level hoisting, propagated pair embedding)
# =================================================================================================
EPS = 1e-5


def ln_fwd(W, z, gname, bname):
    mean = z.mean(1, keepdim=True)
    rstd = torch.rsqrt(z.var(1, unbiased=False, keepdim=True) + EPS)
    xhat = (z - mean) * rstd
    return xhat * W.m(gname) + W.m(bname), (xhat, rstd)


def ln_bwd(G, W, gname, bname, sv, dy):
    xhat, rstd = sv
    G.add(gname, (dy * xhat).sum(0))
    G.add(bname, dy.sum(0))
    dxh = dy * W.m(gname)
    return rstd * (dxh - dxh.mean(1, keepdim=True) - xhat * (dxh * xhat).mean(1, keepdim=True))


def folded_stats_bwd(drstd, rstd, var_raw, mu, dmu):
    """rstd = rsqrt(clamp(ex2 - mu^2, 0) + eps): returns (dex2, dmu_total)"""
    dvar = drstd * (-0.5) * rstd ** 3 * (var_raw >= 0)
    return dvar, dmu - 2 * mu * dvar


def gcl_plus_fwd(W, pre, h, x, ctx, cplx, B, cmax):
    r, c = ctx
    N, H = h.shape
    Dp = (2 * H + 1 + 63) // 64 * 64
    D = 2 * H + 1
    rn, rs = radial_fwd(r, c, x, cplx, B)
    s1, s2 = h.sum(1), (h * h).sum(1)
    Pn = F.linear(h, W.m(pre + "e1_rc"))
    mu = (s1[r] + s1[c] + rn) / D
    var_raw = (s2[r] + s2[c] + rn * rn) / D - mu * mu
    rstd = torch.rsqrt(var_raw.clamp(min=0) + EPS)
    U = Pn[r, :Dp] + Pn[c, Dp:] + rn[:, None] * W.m(pre + "e1_rad") - mu[:, None] * W.m(pre + "e1_g")
    Z1 = rstd[:, None] * U + W.m(pre + "e1_c0")
    A1 = F.relu(Z1)
    M = F.relu(F.linear(A1, W.m(pre + "e2_w"), W.m(pre + "e2_b")))
    M2, lc = ln_fwd(W, M, pre + "cl_g", pre + "cl_b")
    T3 = F.relu(F.linear(M2, W.m(pre + "c1_w"), W.m(pre + "c1_b")))
    s = T3 @ W.m(pre + "c2_w")
    deg = torch.zeros(N).index_add_(0, r, torch.ones(r.numel())).clamp(min=1)
    d = x[r] - x[c]
    step = torch.zeros(N, 3).index_add_(0, r, d * s[:, None]) / deg[:, None]
    x_new = x + step.clamp(-cmax, cmax)
    agg = torch.zeros(N, H).index_add_(0, r, M)
    t0, ln_n = ln_fwd(W, torch.cat([h, agg], 1), pre + "nl_g", pre + "nl_b")
    t1 = F.relu(F.linear(t0, W.m(pre + "n1_w"), W.m(pre + "n1_b")))
    t2 = F.relu(F.linear(t1, W.m(pre + "n2_w"), W.m(pre + "n2_b")))
    sv = dict(h=h, x=x, agg=agg, rn=rn, rs=rs, mu=mu, var_raw=var_raw, rstd=rstd, U=U, Z1=Z1, A1=A1, M=M, M2=M2, lc=lc, T3=T3, s=s, deg=deg,
              d=d, step=step, t0=t0, ln_n=ln_n, t1=t1, t2=t2, Dp=Dp, D=D)
    return h + t2, x_new, sv


def gcl_plus_bwd(G, W, pre, sv, ctx, cplx, B, cmax, dh_new, dx_new):
    r, c = ctx
    h = sv["h"]
    N, H = h.shape
    Dp, D, rn, mu, rstd = sv["Dp"], sv["D"], sv["rn"], sv["mu"], sv["rstd"]
    dx = dx_new.clone()
    de = (dx_new * (sv["step"].abs() <= cmax))[r] / sv["deg"][r][:, None]
    ds = (de * sv["d"]).sum(1)
    dd = de * sv["s"][:, None]
    dx.index_add_(0, r, dd).index_add_(0, c, -dd)
    G.add(pre + "c2_w", sv["T3"].t() @ ds)
    dM2 = lin_bwd(G, W, pre + "c1_w", pre + "c1_b", sv["M2"], (ds[:, None] * W.m(pre + "c2_w")) * (sv["T3"] > 0))
    dM = ln_bwd(G, W, pre + "cl_g", pre + "cl_b", sv["lc"], dM2)
    dh = dh_new.clone()
    dt1 = lin_bwd(G, W, pre + "n2_w", pre + "n2_b", sv["t1"], dh_new * (sv["t2"] > 0))
    dt0 = lin_bwd(G, W, pre + "n1_w", pre + "n1_b", sv["t0"], dt1 * (sv["t1"] > 0))
    dcat = ln_bwd(G, W, pre + "nl_g", pre + "nl_b", sv["ln_n"], dt0)
    dh += dcat[:, :H]
    dM = dM + dcat[:, H:][r]
    dA1 = lin_bwd(G, W, pre + "e2_w", pre + "e2_b", sv["A1"], dM * (sv["M"] > 0))
    dZ1 = dA1 * (sv["Z1"] > 0)
    # folded LayerNorm of the edge MLP's first Linear: Z1 = rstd_e (U_e) + c0,  U = Pn[r] + Pn[c] + rn rad - mu g
    G.add(pre + "e1_c0", dZ1.sum(0))
    dU = dZ1 * rstd[:, None]
    drstd = (dZ1 * sv["U"]).sum(1)
    G.add(pre + "e1_rad", (dU * rn[:, None]).sum(0))
    G.add(pre + "e1_g", -(dU * mu[:, None]).sum(0))
    drn = dU @ W.m(pre + "e1_rad")
    dmu = -(dU @ W.m(pre + "e1_g"))
    dPn = torch.zeros(N, 2 * Dp)
    dPn[:, :Dp].index_add_(0, r, dU)
    dPn[:, Dp:].index_add_(0, c, dU)
    dh += lin_bwd(G, W, pre + "e1_rc", None, h, dPn)
    dex2, dmu = folded_stats_bwd(drstd, rstd, sv["var_raw"], mu, dmu)
    ds1 = torch.zeros(N).index_add_(0, r, dmu / D).index_add_(0, c, dmu / D)
    ds2 = torch.zeros(N).index_add_(0, r, dex2 / D).index_add_(0, c, dex2 / D)
    drn = drn + dmu / D + 2 * rn * dex2 / D
    dh += ds1[:, None] + 2 * h * ds2[:, None]
    dx += radial_bwd(r, c, cplx, B, N, sv["rs"], drn)
    return dh, dx


def _pair_rows(geo):
    pi, ci = [], []
    for b in range(geo["B"]):
        nc1, np1 = geo["c_off"][b + 1] - geo["c_off"][b], geo["p_off"][b + 1] - geo["p_off"][b]
        pi.append(torch.arange(geo["p_off"][b], geo["p_off"][b + 1]).repeat_interleave(nc1))
        ci.append(torch.arange(geo["c_off"][b], geo["c_off"][b + 1]).repeat(np1))
    return torch.cat(pi), torch.cat(ci)


def att_plus_fwd(W, pre, pair_in, h, x, geo, inter, cmax):
    Nc, B, c_off, p_off, pair_base, cplx = geo["Nc"], geo["B"], geo["c_off"], geo["p_off"], geo["pair_base"], geo["cplx"]
    int_r, int_c = inter
    N, H = h.shape
    hc0, hp0 = h[:Nc], h[Nc:]
    raw_full = F.linear(pair_in, W.m(pre + "pb_w"), W.m(pre + "pb_b"))
    raw = raw_full[:, :16].reshape(-1, 2, 2, 4)
    sig = torch.sigmoid(raw[:, :, 1])
    PBl = raw[:, :, 0] * sig
    CAc = F.linear(hc0, W.m(pre + "ca_c_w"), W.m(pre + "ca_c_b"))
    CAp = F.linear(hp0, W.m(pre + "ca_p_w"), W.m(pre + "ca_p_b"))
    blocks = []
    for b in range(B):
        blocks.append((slice(c_off[b], c_off[b + 1]), slice(p_off[b] - Nc, p_off[b + 1] - Nc), c_off[b + 1] - c_off[b],
                       p_off[b + 1] - p_off[b], slice(pair_base[b], pair_base[b + 1])))
    Op, svp = [], []
    for cs, ps, nc1, np1, pr in blocks:
        o, sv = rowatt_fwd(CAp[ps, :HD], CAp[ps, HD:], CAc[cs, :HD], CAc[cs, HD:2 * HD], PBl[pr, 0].view(np1, nc1, 4))
        Op.append(o); svp.append(sv)
    Op = torch.cat(Op)
    hp1 = hp0 + F.linear(Op, W.m(pre + "o_p_w"), W.m(pre + "o_p_b"))
    CAp2 = F.linear(hp1, W.m(pre + "ca_p2_w"))
    Oc, svc = [], []
    for cs, ps, nc1, np1, pr in blocks:
        o, sv = rowatt_fwd(CAc[cs, 2 * HD:3 * HD], CAc[cs, 3 * HD:], CAp2[ps, :HD], CAp2[ps, HD:],
                           PBl[pr, 1].view(np1, nc1, 4).transpose(0, 1))
        Oc.append(o); svc.append(sv)
    Oc = torch.cat(Oc)
    hc1 = hc0 + F.linear(Oc, W.m(pre + "o_c_w"), W.m(pre + "o_c_b"))
    tr = {}
    for t, hs in (("tc", hc1), ("tp", hp1)):
        t0, lsv = ln_fwd(W, hs, pre + t + "l_g", pre + t + "l_b")
        t1 = F.relu(F.linear(t0, W.m(pre + t + "1_w"), W.m(pre + t + "1_b")))
        t2 = F.relu(F.linear(t1, W.m(pre + t + "2_w"), W.m(pre + t + "2_b")))
        tr[t] = (t0, lsv, t1, t2)
    h2 = torch.cat([hc1 + tr["tc"][3], hp1 + tr["tp"][3]])
    QK = F.linear(h2, W.m(pre + "qk_w"), W.m(pre + "qk_b"))
    pi_all, ci_all = _pair_rows(geo)
    a32, b32 = QK[pi_all, 2 * H:2 * H + 32], QK[ci_all, 2 * H + 32:2 * H + 64]
    t32 = a32 * b32
    Zpre = pair_in + (t32 @ W.m(pre + "zo_w") + W.m(pre + "zo_b"))
    Zl, lz = ln_fwd(W, Zpre, pre + "zl_g", pre + "zl_b")
    Zh = F.relu(F.linear(Zl, W.m(pre + "pt1_w"), W.m(pre + "pt1_b")))
    pair_out = F.relu(F.linear(Zh, W.m(pre + "pt2_w"), W.m(pre + "pt2_b")))
    pb_dense = pair_out @ W.m(pre + "wb") + W.m(pre + "pt_c")
    eb = cplx[int_r]
    is_c = int_r < Nc
    ci, pi = torch.where(is_c, int_r, int_c), torch.where(is_c, int_c, int_r)
    c_off_t, p_off_t = torch.from_numpy(c_off.astype(np.int64)), torch.from_numpy(p_off.astype(np.int64))
    nc1_t = c_off_t[1:] - c_off_t[:-1]
    pair = torch.from_numpy(pair_base.astype(np.int64))[eb] + (pi - p_off_t[eb]) * nc1_t[eb] + (ci - c_off_t[eb])
    rn, rs = radial_fwd(int_r, int_c, x, cplx, B)
    V, VC = QK[:, 2 * H + 128:3 * H + 128], QK[:, 3 * H + 128:]
    q = QK[int_r, :H]
    kk = QK[int_c, H:2 * H] + rn[:, None] * W.m(pre + "k_r")
    logit = (q * kk).sum(1) + pb_dense[pair]
    mx = torch.full((N,), float("-inf")).scatter_reduce(0, int_r, logit, reduce="amax")
    e = (logit - mx[int_r]).exp()
    alpha = e / torch.zeros(N).index_add_(0, int_r, e)[int_r]
    v_r = W.m(pre + "v_r")
    ve = V[int_c] + rn[:, None] * v_r
    h3 = h2 + torch.zeros(N, H).index_add_(0, int_r, alpha[:, None] * ve)
    s1, s2, s3 = V.sum(1), (V * V).sum(1), (V * v_r).sum(1)
    acr = W.m(pre + "ac_r")
    mu = (s1[int_c] + rn * acr[0]) / H
    var_raw = (s2[int_c] + 2 * rn * s3[int_c] + rn * rn * acr[1]) / H - mu * mu
    rstd = torch.rsqrt(var_raw.clamp(min=0) + EPS)
    Uc = VC[int_c] + rn[:, None] * W.m(pre + "ac_u") - mu[:, None] * W.m(pre + "ac_g")
    tco = rstd[:, None] * Uc + W.m(pre + "ac_c0")
    se = F.relu(tco) @ W.m(pre + "ac2_w")
    d = x[int_r] - x[int_c]
    step = torch.zeros(N, 3).index_add_(0, int_r, d * (alpha * se)[:, None])
    sv = dict(pair_in=pair_in, raw_full=raw_full, raw=raw, sig=sig, PBl=PBl, Zpre=Zpre, CAc=CAc, CAp=CAp, CAp2=CAp2, x=x, h_in=h, hc0=hc0,
              hp0=hp0, blocks=blocks, svp=svp, svc=svc, Op=Op,
              Oc=Oc, hp1=hp1, hc1=hc1, tr=tr, h2=h2, QK=QK, pi_all=pi_all, ci_all=ci_all, a32=a32, b32=b32, t32=t32, lz=lz, Zl=Zl,
              Zh=Zh, pair_out=pair_out, pair=pair, rn=rn, rs=rs, q=q, kk=kk, alpha=alpha, ve=ve, V=V, s3=s3, mu=mu, var_raw=var_raw,
              rstd=rstd, Uc=Uc, tco=tco, se=se, d=d, step=step)
    return h3, x + step.clamp(-cmax, cmax), pair_out, sv


def att_plus_bwd(G, W, pre, sv, geo, inter, cmax, dh3, dx_new, dpair_out):
    Nc, B, cplx = geo["Nc"], geo["B"], geo["cplx"]
    int_r, int_c = inter
    N, H = sv["h2"].shape
    rn, alpha, se, mu, rstd, V = sv["rn"], sv["alpha"], sv["se"], sv["mu"], sv["rstd"], sv["V"]
    v_r, acr = W.m(pre + "v_r"), W.m(pre + "ac_r")
    dQK = torch.zeros_like(sv["QK"])
    dx = dx_new.clone()
    de = (dx_new * (sv["step"].abs() <= cmax))[int_r]
    dw = (de * sv["d"]).sum(1)
    dd = de * (alpha * se)[:, None]
    dx.index_add_(0, int_r, dd).index_add_(0, int_c, -dd)
    dalpha, dse = dw * se, dw * alpha
    # LayerNorm-folded coordinate head on v_e = V[c] + rn v_r  (per-node sums s1, s2, s3 of V)
    rt = F.relu(sv["tco"])
    G.add(pre + "ac2_w", rt.t() @ dse)
    dt = (dse[:, None] * W.m(pre + "ac2_w")) * (sv["tco"] > 0)
    G.add(pre + "ac_c0", dt.sum(0))
    dUc = dt * rstd[:, None]
    drstd = (dt * sv["Uc"]).sum(1)
    dQK[:, 3 * H + 128:].index_add_(0, int_c, dUc)
    G.add(pre + "ac_u", (dUc * rn[:, None]).sum(0))
    G.add(pre + "ac_g", -(dUc * mu[:, None]).sum(0))
    drn = dUc @ W.m(pre + "ac_u")
    dmu = -(dUc @ W.m(pre + "ac_g"))
    dex2, dmu = folded_stats_bwd(drstd, rstd, sv["var_raw"], mu, dmu)
    s3c = sv["s3"][int_c]
    ds1 = torch.zeros(N).index_add_(0, int_c, dmu / H)
    ds2 = torch.zeros(N).index_add_(0, int_c, dex2 / H)
    ds3 = torch.zeros(N).index_add_(0, int_c, 2 * rn * dex2 / H)
    drn = drn + acr[0] * dmu / H + (2 * s3c + 2 * rn * acr[1]) * dex2 / H
    G.add(pre + "ac_r", torch.stack([(rn * dmu / H).sum(), (rn * rn * dex2 / H).sum()]))
    dV = ds1[:, None] + 2 * V * ds2[:, None] + v_r[None, :] * ds3[:, None]
    G.add(pre + "v_r", (V * ds3[:, None]).sum(0))
    # aggregation + segment softmax + logits (as in the v1 layout)
    dh2 = dh3.clone()
    dagg = dh3[int_r]
    dalpha = dalpha + (dagg * sv["ve"]).sum(1)
    dve = dagg * alpha[:, None]
    dV.index_add_(0, int_c, dve)
    dQK[:, 2 * H + 128:3 * H + 128] += dV
    G.add(pre + "v_r", (dve * rn[:, None]).sum(0))
    drn = drn + dve @ v_r
    dlogit = alpha * (dalpha - torch.zeros(N).index_add_(0, int_r, alpha * dalpha)[int_r])
    dq, dkk = dlogit[:, None] * sv["kk"], dlogit[:, None] * sv["q"]
    dQK[:, :H].index_add_(0, int_r, dq)
    dQK[:, H:2 * H].index_add_(0, int_c, dkk)
    G.add(pre + "k_r", (dkk * rn[:, None]).sum(0))
    drn = drn + dkk @ W.m(pre + "k_r")
    dx += radial_bwd(int_r, int_c, cplx, B, N, sv["rs"], drn)
    # pair transition on EVERY pair row (the embedding is propagated) with attn_bias_proj as a row-dot
    pair_out = sv["pair_out"]
    dpb = torch.zeros(pair_out.shape[0]).index_add_(0, sv["pair"], dlogit)
    G.add(pre + "wb", pair_out.t() @ dpb)
    G.add(pre + "pt_c", dpb.sum().reshape(1))
    dpo = dpair_out + dpb[:, None] * W.m(pre + "wb")
    dZh = lin_bwd(G, W, pre + "pt2_w", pre + "pt2_b", sv["Zh"], dpo * (pair_out > 0))
    dZl = lin_bwd(G, W, pre + "pt1_w", pre + "pt1_b", sv["Zl"], dZh * (sv["Zh"] > 0))
    dZpre = ln_bwd(G, W, pre + "zl_g", pre + "zl_b", sv["lz"], dZl)
    dpair_in = dZpre.clone()
    G.add(pre + "zo_b", dZpre.sum(0))
    G.add(pre + "zo_w", sv["t32"].t() @ dZpre)
    dt32 = dZpre @ W.m(pre + "zo_w").t()
    dQK[:, 2 * H:2 * H + 32].index_add_(0, sv["pi_all"], dt32 * sv["b32"])
    dQK[:, 2 * H + 32:2 * H + 64].index_add_(0, sv["ci_all"], dt32 * sv["a32"])
    dh2 += lin_bwd(G, W, pre + "qk_w", pre + "qk_b", sv["h2"], dQK)
    # transitions: hs + relu(linear2(relu(linear1(LN(hs)))))
    dside = {}
    for t, dhs in (("tc", dh2[:Nc]), ("tp", dh2[Nc:])):
        t0, lsv, t1, t2 = sv["tr"][t]
        dt1 = lin_bwd(G, W, pre + t + "2_w", pre + t + "2_b", t1, dhs * (t2 > 0))
        dt0 = lin_bwd(G, W, pre + t + "1_w", pre + t + "1_b", t0, dt1 * (t1 > 0))
        dside[t] = dhs + ln_bwd(G, W, pre + t + "l_g", pre + t + "l_b", lsv, dt0)
    dhc1, dhp1 = dside["tc"], dside["tp"]
    dPB = torch.zeros_like(sv["raw"][:, :, 0])
    dOc = lin_bwd(G, W, pre + "o_c_w", pre + "o_c_b", sv["Oc"], dhc1)
    dhc0 = dhc1.clone()
    dCAc = torch.zeros(Nc, 4 * HD)
    dCAp2 = torch.zeros(N - Nc, 2 * HD)
    for (cs, ps, nc1, np1, pr), s in zip(sv["blocks"], sv["svc"]):
        dq_, dg_, dk_, dv_, db_ = rowatt_bwd(s, dOc[cs])
        dCAc[cs, 2 * HD:3 * HD], dCAc[cs, 3 * HD:] = dq_, dg_
        dCAp2[ps, :HD], dCAp2[ps, HD:] = dk_, dv_
        dPB[pr, 1] += db_.transpose(0, 1).reshape(-1, 4)
    dhp1 = dhp1 + lin_bwd(G, W, pre + "ca_p2_w", None, sv["hp1"], dCAp2)
    dOp = lin_bwd(G, W, pre + "o_p_w", pre + "o_p_b", sv["Op"], dhp1)
    dhp0 = dhp1.clone()
    dCAp = torch.zeros(N - Nc, 2 * HD)
    for (cs, ps, nc1, np1, pr), s in zip(sv["blocks"], sv["svp"]):
        dq_, dg_, dk_, dv_, db_ = rowatt_bwd(s, dOp[ps])
        dCAp[ps, :HD], dCAp[ps, HD:] = dq_, dg_
        dCAc[cs, :HD], dCAc[cs, HD:2 * HD] = dk_, dv_
        dPB[pr, 0] += db_.reshape(-1, 4)
    dhc0 += lin_bwd(G, W, pre + "ca_c_w", pre + "ca_c_b", sv["hc0"], dCAc)
    dhp0 += lin_bwd(G, W, pre + "ca_p_w", pre + "ca_p_b", sv["hp0"], dCAp)
    # gated pair biases of this layer's two row-attention blocks: read from the INCOMING pair embedding
    raw, sig = sv["raw"], sv["sig"]
    draw = torch.zeros_like(raw)
    draw[:, :, 0] = dPB * sig
    draw[:, :, 1] = dPB * raw[:, :, 0] * sig * (1 - sig)
    draw_full = torch.zeros_like(sv["raw_full"])
    draw_full[:, :16] = draw.reshape(-1, 16)
    dpair_in += lin_bwd(G, W, pre + "pb_w", pre + "pb_b", sv["pair_in"], draw_full)
    return torch.cat([dhc0, dhp0]), dx, dpair_in


def forward_backward_plus(sd, cfg, batch, gX, gH, gP, arena=None, export=None):
    """FABind+ layout, eval-mode masks (no dropout): loss = <X,gX> + <H,gH> + <pair (packed rows), gP>.
    Returns (X_out, H_out, pair_out, grad of the flat arena, grad of batch.H)."""
    from types import SimpleNamespace
    H = batch.H.shape[1]
    L = cfg.n_layers
    W = Arena(sd, H, L, 1, False, arena)
    with torch.no_grad():
        lay = build_layout(batch.batch_id, batch.segment_id, batch.is_global, batch.mask, "cpu")
        o, blob = lay.offs, lay.blob.numpy()
        N, B, Nc = lay.N, lay.B, lay.Nc_tot
        perm, inv = blob[o["perm"]:o["perm"] + N], blob[o["inv"]:o["inv"] + N]
        lay_np = dict(node_cplx=blob[o["node_cplx"]:o["node_cplx"] + N], flags=lay.flags.numpy(),
                      c_off=blob[o["c_off"]:o["c_off"] + B + 1], p_off=blob[o["p_off"]:o["p_off"] + B + 1])
        geo = dict(Nc=Nc, B=B, c_off=lay_np["c_off"], p_off=lay_np["p_off"], pair_base=blob[o["pair_base"]:o["pair_base"] + B + 1],
                   cplx=torch.from_numpy(lay_np["node_cplx"].astype(np.int64)))
        permt = torch.from_numpy(perm.astype(np.int64))
        Hin = batch.H[permt]
        xl = batch.X_LAS[permt, 0]
        bonds_int = inv[batch.compound_edge_index.numpy()]
        las = tuple(torch.from_numpy(inv[batch.LAS_edge_index.numpy()].astype(np.int64)))
        moves = torch.from_numpy((lay_np["flags"] & 4) != 0)
        intra, inter_cut = cfg.intra_cutoff / cfg.coordinate_scale, cfg.inter_cutoff / cfg.coordinate_scale
        cmax, lcl = 10.0 / cfg.coordinate_scale, 15.0 / cfg.coordinate_scale
        if cfg.n_iter > 1:
            Xprev = forward_emulated(sd, SimpleNamespace(**{**vars(cfg), "n_iter": cfg.n_iter - 1}), batch, flavour=1,
                                     arena=W.a.detach())[0]
        else:
            Xprev = batch.X
        x_state = Xprev[permt, 0].clone()
        ctx, inter = _edges(x_state, lay_np, intra, inter_cut, bonds_int)
        if inter[0].numel() == 0:
            inter = (torch.tensor([lay.fb_atom, lay.fb_res]), torch.tensor([lay.fb_res, lay.fb_atom]))
        c_off, p_off = geo["c_off"], geo["p_off"]
        pc = torch.empty(N, H)
        pc[:Nc] = F.linear(Hin[:Nc], W.m("il_c_w"), W.m("il_c_b"))
        pc[Nc:] = F.linear(Hin[Nc:], W.m("il_p_w"), W.m("il_p_b"))
        outer = torch.cat([(pc[p_off[b]:p_off[b + 1], None, :] * pc[None, c_off[b]:c_off[b + 1], :]).reshape(-1, H)
                           for b in range(B)])
        P0 = F.linear(outer, W.m("il_o_w"), W.m("il_o_b"))
        h = F.linear(Hin, W.m("in_w"), W.m("in_b"))
        x = x_state.clone()
        pair = P0
        tape = []
        for l in range(L):
            h, x, s1 = gcl_plus_fwd(W, f"gcl{l}.", h, x, ctx, geo["cplx"], B, cmax)
            h, x, pair, s2 = att_plus_fwd(W, f"att{l}.", pair, h, x, geo, inter, cmax)
            x, s3 = las_fwd(x, xl, las, cfg.geometry_reg_step_size, lcl)
            tape.append((s1, s2, s3))
        h_last, x, s_out = gcl_plus_fwd(W, "out.", h, x, ctx, geo["cplx"], B, cmax)
        h_final = F.linear(h_last, W.m("out_w"), W.m("out_b"))
        X_out = torch.empty_like(batch.X)
        X_out[permt, 0] = torch.where(moves[:, None], x, x_state)
        H_out = torch.empty_like(batch.H)
        H_out[permt] = h_final

        if export is not None:
            export.update(W=W, geo=geo, ctx=ctx, inter=inter, las=las, tape=tape, s_out=s_out, P0=P0, cmax=cmax, lcl=lcl, xl=xl, N=N, B=B,
                          Nc=Nc, Hin=Hin, pc=pc, outer=outer, h_last=h_last, permt=permt, moves=moves, x_state=x_state)
        G = Grads()
        dx = gX[permt, 0] * moves[:, None]
        dh = lin_bwd(G, W, "out_w", "out_b", h_last, gH[permt])
        dh, dx = gcl_plus_bwd(G, W, "out.", s_out, ctx, geo["cplx"], B, cmax, dh, dx)
        dpair = gP.clone()
        for l in reversed(range(L)):
            s1, s2, s3 = tape[l]
            dx = las_bwd(s3, las, cfg.geometry_reg_step_size, lcl, dx)
            dh, dx, dpair = att_plus_bwd(G, W, f"att{l}.", s2, geo, inter, cmax, dh, dx, dpair)
            dh, dx = gcl_plus_bwd(G, W, f"gcl{l}.", s1, ctx, geo["cplx"], B, cmax, dh, dx)
        dHin = lin_bwd(G, W, "in_w", "in_b", Hin, dh)
        douter = lin_bwd(G, W, "il_o_w", "il_o_b", outer, dpair)
        dpc = torch.zeros(N, H)
        pb = geo["pair_base"]
        for b in range(B):
            np1, nc1 = p_off[b + 1] - p_off[b], c_off[b + 1] - c_off[b]
            t = douter[pb[b]:pb[b + 1]].view(np1, nc1, H)
            dpc[p_off[b]:p_off[b + 1]] += (t * pc[None, c_off[b]:c_off[b + 1]]).sum(1)
            dpc[c_off[b]:c_off[b + 1]] += (t * pc[p_off[b]:p_off[b + 1], None]).sum(0)
        dHin[:Nc] += lin_bwd(G, W, "il_c_w", "il_c_b", Hin[:Nc], dpc[:Nc])
        dHin[Nc:] += lin_bwd(G, W, "il_p_w", "il_p_b", Hin[Nc:], dpc[Nc:])
        garena = torch.zeros_like(W.a)
        for name, g in G.items():
            r, c, off = W.s[name]
            garena[off:off + r * c] = g.reshape(-1)
        gH_in = torch.empty_like(batch.H)
        gH_in[permt] = dHin
    return X_out, H_out, pair, garena, gH_in
