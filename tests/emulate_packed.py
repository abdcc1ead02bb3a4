"""CPU emulation of the launch sequence in fabind_b200/csrc/forward.cu (TEST INFRASTRUCTURE).

It consumes the SAME packed weight arena and internal node order as the CUDA library and mirrors every
kernel with a few lines of torch, so the refactored formulation (per-node first layers, collapsed
pair-bias vector, pair work on inter pairs only, type-sorted node order) and `fabind_b200/weights.py`
can be validated against the oracle without a GPU.  It is not part of the product and is never used as
a fallback.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from fabind_b200.weights import slots, pack_state_dict
from fabind_b200.layout import build_layout

HD = 128


class Arena:
    def __init__(self, sd, H, L, flavour=0, differentiable=False, arena=None):
        # arena: an already packed (possibly requires_grad) flat tensor -- gradients w.r.t. the arena itself
        self.a = pack_state_dict(sd, H, L, flavour, differentiable=differentiable) if arena is None else arena
        self.s = {n: (r, c, o) for n, r, c, o in slots(H, L, flavour)}

    def m(self, name):
        r, c, o = self.s[name]
        t = self.a[o:o + r * c]
        return t.view(r, c) if r > 1 else t.view(c)


def _edges(x, lay_np, intra, inter, bonds_int):
    """internal-order ctx CSR (bond edges first per row, then geometric by ascending internal col) and inter CSR."""
    N = x.shape[0]
    cplx, flags, c_off, p_off = lay_np["node_cplx"], lay_np["flags"], lay_np["c_off"], lay_np["p_off"]
    ctx_r, ctx_c, int_r, int_c = [], [], [], []
    xs = x.detach().numpy().astype(np.float32)
    for r in range(N):
        b = cplx[r]
        for e in range(bonds_int.shape[1]):
            if bonds_int[0, e] == r:
                ctx_r.append(r); ctx_c.append(int(bonds_int[1, e]))
        cands = list(range(c_off[b], c_off[b + 1])) + list(range(p_off[b], p_off[b + 1]))
        for c in cands:
            if c == r:
                continue
            sr, sc, gr, gc = flags[r] & 1, flags[c] & 1, flags[r] & 2, flags[c] & 2
            d = xs[r] - xs[c]
            # sqrt(fma(dz,dz,fma(dy,dy,dx*dx))) evaluated with float64 intermediates rounded like fma
            t = np.float32(np.float64(d[0]) * np.float64(d[0]))
            t = np.float32(np.float64(d[1]) * np.float64(d[1]) + np.float64(t))
            t = np.float32(np.float64(d[2]) * np.float64(d[2]) + np.float64(t))
            dist = np.sqrt(t)
            if not gr and not gc:
                if sr == sc:
                    if sr and dist <= np.float32(intra):
                        ctx_r.append(r); ctx_c.append(c)
                elif dist <= np.float32(inter):
                    int_r.append(r); int_c.append(c)
            elif sr == sc or (gr and gc):
                ctx_r.append(r); ctx_c.append(c)
    return (torch.tensor(ctx_r), torch.tensor(ctx_c)), (torch.tensor(int_r, dtype=torch.long), torch.tensor(int_c, dtype=torch.long))


def _radial(row, col, x, cplx_t, B):
    d2 = ((x[row] - x[col]) ** 2).sum(1)
    nrm = torch.zeros(B).index_add_(0, cplx_t[row], d2 * d2).sqrt()
    return d2 / nrm[cplx_t[row]]


def forward_emulated(sd, cfg, batch, flavour=0, dropout=None, differentiable=False, arena=None):
    """flavour 0: FABind v1 layout -> (X, H, stats); 1: FABind+ layout -> (X, H, stats, pair [P_total, H] packed rows).
    dropout = (p, seed, colonly): FABind+ train-mode masks with the library's counter-based mask function
    (fabind_b200/dropout.py), at the sites and row ids csrc/forward.cu uses."""
    from fabind_b200.dropout import keep_mask, site_id, iter_seed
    H = batch.H.shape[1]
    L = cfg.n_layers
    plus = flavour == 1
    W = Arena(sd, H, L, flavour, differentiable, arena)
    Dp = (2 * H + 1 + 63) // 64 * 64
    EPS = 1e-5
    lay = build_layout(batch.batch_id, batch.segment_id, batch.is_global, batch.mask, "cpu")
    o = lay.offs
    blob = lay.blob.numpy()
    N, B, Nc = lay.N, lay.B, lay.Nc_tot
    perm = blob[o["perm"]:o["perm"] + N]
    inv = blob[o["inv"]:o["inv"] + N]
    lay_np = dict(node_cplx=blob[o["node_cplx"]:o["node_cplx"] + N], flags=lay.flags.numpy(),
                  c_off=blob[o["c_off"]:o["c_off"] + B + 1], p_off=blob[o["p_off"]:o["p_off"] + B + 1])
    pair_base = blob[o["pair_base"]:o["pair_base"] + B + 1]
    cplx_t = torch.from_numpy(lay_np["node_cplx"].astype(np.int64))
    permt = torch.from_numpy(perm.astype(np.int64))
    Hin = batch.H[permt]
    x_state = batch.X[permt, 0].clone()
    xl = batch.X_LAS[permt, 0]
    bonds_int = inv[batch.compound_edge_index.numpy()]
    las_int = torch.from_numpy(inv[batch.LAS_edge_index.numpy()].astype(np.int64))
    moves = torch.from_numpy((lay_np["flags"] & 4) != 0)
    intra, inter = cfg.intra_cutoff / cfg.coordinate_scale, cfg.inter_cutoff / cfg.coordinate_scale
    cmax, lcl = 10.0 / cfg.coordinate_scale, 15.0 / cfg.coordinate_scale
    c_off, p_off = lay_np["c_off"], lay_np["p_off"]

    # pair0 and the gated biases of every row-attention block
    pc = torch.empty(N, H)
    pc[:Nc] = F.linear(Hin[:Nc], W.m("il_c_w"), W.m("il_c_b"))
    pc[Nc:] = F.linear(Hin[Nc:], W.m("il_p_w"), W.m("il_p_b"))
    P0 = []
    for b in range(B):
        pp, cc = pc[p_off[b]:p_off[b + 1]], pc[c_off[b]:c_off[b + 1]]
        P0.append(F.linear((pp[:, None, :] * cc[None, :, :]).reshape(-1, H), W.m("il_o_w"), W.m("il_o_b")))
    P0 = torch.cat(P0)
    if not plus:
        raw = F.linear(P0, W.m("pb_w"), W.m("pb_b"))[:, :16 * L].reshape(-1, L, 2, 2, 4)
        PB = raw[:, :, :, 0] * torch.sigmoid(raw[:, :, :, 1])          # [P, L, blk, head]
    pair_last = None
    state = dict(it=0)

    def drop(t, layer, name, row0=0):
        if dropout is None or dropout[0] <= 0:
            return t
        pdrop, seed, colonly = dropout
        return t * keep_mask(iter_seed(seed, state["it"]), site_id(layer, name), t.shape[0], t.shape[1], pdrop, colonly, row0)

    (ctx_r, ctx_c), _ = _edges(x_state, lay_np, intra, inter, bonds_int)
    stats = []
    h_final = None
    for it in range(cfg.n_iter):
        last = it == cfg.n_iter - 1
        state["it"] = it
        _, (int_r, int_c) = _edges(x_state, lay_np, intra, inter, bonds_int)
        if int_r.numel() == 0:
            int_r, int_c = torch.tensor([lay.fb_atom, lay.fb_res]), torch.tensor([lay.fb_res, lay.fb_atom])
        stats.append(int(int_r.numel()))
        h = F.linear(Hin, W.m("in_w"), W.m("in_b"))
        if plus:
            h = drop(h, -1, "stack_in")
        x = x_state.clone()

        def gcl(pre, h, x, need_h=True):
            rn = _radial(ctx_r, ctx_c, x, cplx_t, B)
            Pn = F.linear(h, W.m(pre + "e1_rc"))
            A1 = F.silu(Pn[ctx_r, :H] + Pn[ctx_c, H:] + rn[:, None] * W.m(pre + "e1_rad") + W.m(pre + "e1_b"))
            M = F.silu(F.linear(A1, W.m(pre + "e2_w"), W.m(pre + "e2_b")))
            s = F.silu(F.linear(M, W.m(pre + "c1_w"), W.m(pre + "c1_b"))) @ W.m(pre + "c2_w")
            deg = torch.zeros(N).index_add_(0, ctx_r, torch.ones(ctx_r.numel())).clamp(min=1)
            dx = torch.zeros(N, 3).index_add_(0, ctx_r, (x[ctx_r] - x[ctx_c]) * s[:, None]) / deg[:, None]
            x_new = x + dx.clamp(-cmax, cmax)
            if need_h:
                agg = torch.zeros(N, H).index_add_(0, ctx_r, M)
                t1 = F.silu(F.linear(torch.cat([h, agg], 1), W.m(pre + "n1_w"), W.m(pre + "n1_b")))
                h = h + F.linear(t1, W.m(pre + "n2_w"), W.m(pre + "n2_b"))
            return h, x_new

        def rowatt(q, g, k, v, bias):  # q [I,128] k,v [J,128] bias [I,J,4]
            qh = q.view(-1, 4, 32) / math.sqrt(32)
            a = torch.einsum("ihd,jhd->hij", qh, k.view(-1, 4, 32)) + bias.permute(2, 0, 1)
            a = torch.softmax(a, -1)
            o = torch.einsum("hij,jhd->ihd", a, v.view(-1, 4, 32))
            return (o * torch.sigmoid(g).view(-1, 4, 32)).reshape(-1, 128)

        def att(pre, l, h, x):
            # (no in-place updates of h: the emulation is also differentiated, see test_refactored_formulation_gradients)
            hc, hp = h[:Nc], h[Nc:]
            CAc = F.linear(hc, W.m(pre + "ca_c_w"), W.m(pre + "ca_c_b"))
            CAp = F.linear(hp, W.m(pre + "ca_p_w"), W.m(pre + "ca_p_b"))
            Op = []
            for b in range(B):
                cs = slice(c_off[b], c_off[b + 1])
                psl = slice(p_off[b] - Nc, p_off[b + 1] - Nc)
                nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
                bias = PB[pair_base[b]:pair_base[b + 1], l, 0].view(np1, nc1, 4)
                Op.append(rowatt(CAp[psl, :HD], CAp[psl, HD:], CAc[cs, :HD], CAc[cs, HD:2 * HD], bias))
            hp = hp + F.linear(torch.cat(Op), W.m(pre + "o_p_w"), W.m(pre + "o_p_b"))
            CAp2 = F.linear(hp, W.m(pre + "ca_p2_w"))
            Oc = []
            for b in range(B):
                cs = slice(c_off[b], c_off[b + 1])
                psl = slice(p_off[b] - Nc, p_off[b + 1] - Nc)
                nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
                bias = PB[pair_base[b]:pair_base[b + 1], l, 1].view(np1, nc1, 4).transpose(0, 1)
                Oc.append(rowatt(CAc[cs, 2 * HD:3 * HD], CAc[cs, 3 * HD:], CAp2[psl, :HD], CAp2[psl, HD:], bias))
            hc = hc + F.linear(torch.cat(Oc), W.m(pre + "o_c_w"), W.m(pre + "o_c_b"))
            hp = hp + F.linear(F.relu(F.linear(hp, W.m(pre + "tp1_w"), W.m(pre + "tp1_b"))), W.m(pre + "tp2_w"), W.m(pre + "tp2_b"))
            hc = hc + F.linear(F.relu(F.linear(hc, W.m(pre + "tc1_w"), W.m(pre + "tc1_b"))), W.m(pre + "tc2_w"), W.m(pre + "tc2_b"))
            h = torch.cat([hc, hp])
            QK = F.linear(h, W.m(pre + "qk_w"), W.m(pre + "qk_b"))     # q | k | linear_p32 | linear_c32 | pad
            pc32 = torch.empty(N, 32)
            pc32[Nc:] = QK[Nc:, 2 * H:2 * H + 32]
            pc32[:Nc] = QK[:Nc, 2 * H + 32:2 * H + 64]
            # pair index of every inter edge
            eb = cplx_t[int_r]
            is_c = int_r < Nc
            ci = torch.where(is_c, int_r, int_c)
            pi = torch.where(is_c, int_c, int_r)
            c_off_t, p_off_t = torch.from_numpy(c_off.astype(np.int64)), torch.from_numpy(p_off.astype(np.int64))
            nc1_t = c_off_t[1:] - c_off_t[:-1]
            pair = torch.from_numpy(pair_base.astype(np.int64))[eb] + (pi - p_off_t[eb]) * nc1_t[eb] + (ci - c_off_t[eb])
            u = is_c.nonzero().squeeze(1)
            zcat = torch.cat([P0[pair[u]], pc32[pi[u]] * pc32[ci[u]], torch.zeros(u.numel(), 32)], 1)   # [pair0 | t | 0]
            pbu = F.relu(F.linear(zcat, W.m(pre + "pt1_w"), W.m(pre + "pt1_b"))) @ W.m(pre + "pt2v") + W.m(pre + "pt_c")
            pb_dense = torch.zeros(P0.shape[0])
            pb_dense[pair[u]] = pbu
            rn = _radial(int_r, int_c, x, cplx_t, B)
            V, VC = QK[:, 2 * H + 128:3 * H + 128], QK[:, 3 * H + 128:]      # stacked GEMM: ... || v | vc
            logit = (QK[int_r, :H] * (QK[int_c, H:2 * H] + rn[:, None] * W.m(pre + "k_r"))).sum(1) + pb_dense[pair]
            mx = torch.full((N,), float("-inf")).scatter_reduce(0, int_r, logit, reduce="amax")
            e = (logit - mx[int_r]).exp()
            alpha = e / torch.zeros(N).index_add_(0, int_r, e)[int_r]
            ve = V[int_c] + rn[:, None] * W.m(pre + "v_r")
            h = h + torch.zeros(N, H).index_add_(0, int_r, alpha[:, None] * ve)
            se = F.silu(VC[int_c] + rn[:, None] * W.m(pre + "ac_u") + W.m(pre + "ac1_b")) @ W.m(pre + "ac2_w")
            dx = torch.zeros(N, 3).index_add_(0, int_r, (x[int_r] - x[int_c]) * (alpha * se)[:, None])
            return h, x + dx.clamp(-cmax, cmax)

        def ln(z, gname, bname):
            return F.layer_norm(z, (z.shape[-1],), W.m(gname), W.m(bname), EPS)

        def row_stats(t, w=None):
            return t.sum(1), (t * t).sum(1), (t * w).sum(1) if w is not None else None

        def gcl_plus(pre, h, x, need_h=True, layer=0):
            rn = _radial(ctx_r, ctx_c, x, cplx_t, B)
            s1, s2, _ = row_stats(h)
            Pn = F.linear(h, W.m(pre + "e1_rc"))          # [N, 2*Dp]
            D = 2 * H + 1
            mu = (s1[ctx_r] + s1[ctx_c] + rn) / D
            var = ((s2[ctx_r] + s2[ctx_c] + rn * rn) / D - mu * mu).clamp(min=0)
            rstd = torch.rsqrt(var + EPS)
            A1 = F.relu(rstd[:, None] * (Pn[ctx_r, :Dp] + Pn[ctx_c, Dp:] + rn[:, None] * W.m(pre + "e1_rad")
                                         - mu[:, None] * W.m(pre + "e1_g")) + W.m(pre + "e1_c0"))
            A1 = drop(A1, layer, "edge1")
            M = drop(F.relu(F.linear(A1, W.m(pre + "e2_w"), W.m(pre + "e2_b"))), layer, "edge2")
            M2 = ln(M, pre + "cl_g", pre + "cl_b")
            s = drop(F.relu(F.linear(M2, W.m(pre + "c1_w"), W.m(pre + "c1_b"))), layer, "gcoord") @ W.m(pre + "c2_w")
            deg = torch.zeros(N).index_add_(0, ctx_r, torch.ones(ctx_r.numel())).clamp(min=1)
            dx = torch.zeros(N, 3).index_add_(0, ctx_r, (x[ctx_r] - x[ctx_c]) * s[:, None]) / deg[:, None]
            x_new = x + dx.clamp(-cmax, cmax)
            if need_h:
                agg = torch.zeros(N, H).index_add_(0, ctx_r, M)
                t0 = ln(torch.cat([h, agg], 1), pre + "nl_g", pre + "nl_b")
                t1 = drop(F.relu(F.linear(t0, W.m(pre + "n1_w"), W.m(pre + "n1_b"))), layer, "node1")
                h = h + drop(F.relu(F.linear(t1, W.m(pre + "n2_w"), W.m(pre + "n2_b"))), layer, "node2")
            return h, x_new

        def att_plus(pre, pair_in, h, x, layer=0):
            hc, hp = h[:Nc], h[Nc:]
            raw = F.linear(pair_in, W.m(pre + "pb_w"), W.m(pre + "pb_b"))[:, :16].reshape(-1, 2, 2, 4)
            PBl = raw[:, :, 0] * torch.sigmoid(raw[:, :, 1])          # [P, blk, head]
            CAc = F.linear(hc, W.m(pre + "ca_c_w"), W.m(pre + "ca_c_b"))
            CAp = F.linear(hp, W.m(pre + "ca_p_w"), W.m(pre + "ca_p_b"))
            Op = []
            for b in range(B):
                cs = slice(c_off[b], c_off[b + 1])
                psl = slice(p_off[b] - Nc, p_off[b + 1] - Nc)
                nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
                bias = PBl[pair_base[b]:pair_base[b + 1], 0].view(np1, nc1, 4)
                Op.append(rowatt(CAp[psl, :HD], CAp[psl, HD:], CAc[cs, :HD], CAc[cs, HD:2 * HD], bias))
            hp = hp + drop(F.linear(torch.cat(Op), W.m(pre + "o_p_w"), W.m(pre + "o_p_b")), layer, "patt", Nc)
            CAp2 = F.linear(hp, W.m(pre + "ca_p2_w"))
            Oc = []
            for b in range(B):
                cs = slice(c_off[b], c_off[b + 1])
                psl = slice(p_off[b] - Nc, p_off[b + 1] - Nc)
                nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
                bias = PBl[pair_base[b]:pair_base[b + 1], 1].view(np1, nc1, 4).transpose(0, 1)
                Oc.append(rowatt(CAc[cs, 2 * HD:3 * HD], CAc[cs, 3 * HD:], CAp2[psl, :HD], CAp2[psl, HD:], bias))
            hc = hc + drop(F.linear(torch.cat(Oc), W.m(pre + "o_c_w"), W.m(pre + "o_c_b")), layer, "catt")
            sides = {}
            for t, hs, r0 in (("tc", hc, 0), ("tp", hp, Nc)):
                t0 = ln(hs, pre + t + "l_g", pre + t + "l_b")
                t1 = drop(F.relu(F.linear(t0, W.m(pre + t + "1_w"), W.m(pre + t + "1_b"))), layer, "ctr1" if t == "tc" else "ptr1", r0)
                sides[t] = hs + drop(F.relu(F.linear(t1, W.m(pre + t + "2_w"), W.m(pre + t + "2_b"))), layer, "ctr2" if t == "tc" else "ptr2", r0)
            h = torch.cat([sides["tc"], sides["tp"]])
            QK = F.linear(h, W.m(pre + "qk_w"), W.m(pre + "qk_b"))
            # pair <- MLPwithLastAct(pair + inter32(p, c)) on every pair row
            pi_all, ci_all = [], []
            for b in range(B):
                nc1, np1 = c_off[b + 1] - c_off[b], p_off[b + 1] - p_off[b]
                pi_all.append(torch.arange(p_off[b], p_off[b + 1]).repeat_interleave(nc1))
                ci_all.append(torch.arange(c_off[b], c_off[b + 1]).repeat(np1))
            pi_all, ci_all = torch.cat(pi_all), torch.cat(ci_all)
            t32 = QK[pi_all, 2 * H:2 * H + 32] * QK[ci_all, 2 * H + 32:2 * H + 64]
            Zl = ln(pair_in + (t32 @ W.m(pre + "zo_w") + W.m(pre + "zo_b")), pre + "zl_g", pre + "zl_b")
            Zh = drop(F.relu(F.linear(Zl, W.m(pre + "pt1_w"), W.m(pre + "pt1_b"))), layer, "pair1")
            pair_out = drop(F.relu(F.linear(Zh, W.m(pre + "pt2_w"), W.m(pre + "pt2_b"))), layer, "pair2")
            pb_dense = pair_out @ W.m(pre + "wb") + W.m(pre + "pt_c")
            eb = cplx_t[int_r]
            is_c = int_r < Nc
            ci = torch.where(is_c, int_r, int_c)
            pi = torch.where(is_c, int_c, int_r)
            c_off_t, p_off_t = torch.from_numpy(c_off.astype(np.int64)), torch.from_numpy(p_off.astype(np.int64))
            nc1_t = c_off_t[1:] - c_off_t[:-1]
            pair = torch.from_numpy(pair_base.astype(np.int64))[eb] + (pi - p_off_t[eb]) * nc1_t[eb] + (ci - c_off_t[eb])
            rn = _radial(int_r, int_c, x, cplx_t, B)
            V, VC = QK[:, 2 * H + 128:3 * H + 128], QK[:, 3 * H + 128:]
            logit = (QK[int_r, :H] * (QK[int_c, H:2 * H] + rn[:, None] * W.m(pre + "k_r"))).sum(1) + pb_dense[pair]
            mx = torch.full((N,), float("-inf")).scatter_reduce(0, int_r, logit, reduce="amax")
            e = (logit - mx[int_r]).exp()
            alpha = e / torch.zeros(N).index_add_(0, int_r, e)[int_r]
            v_r = W.m(pre + "v_r")
            ve = V[int_c] + rn[:, None] * v_r
            h = h + drop(torch.zeros(N, H).index_add_(0, int_r, alpha[:, None] * ve), layer, "agg")
            s1, s2, s3 = row_stats(V, v_r)
            acr = W.m(pre + "ac_r")
            mu = (s1[int_c] + rn * acr[0]) / H
            ex2 = (s2[int_c] + 2 * rn * s3[int_c] + rn * rn * acr[1]) / H
            rstd = torch.rsqrt((ex2 - mu * mu).clamp(min=0) + EPS)
            t = rstd[:, None] * (VC[int_c] + rn[:, None] * W.m(pre + "ac_u") - mu[:, None] * W.m(pre + "ac_g")) + W.m(pre + "ac_c0")
            se = drop(F.relu(t), layer, "acoord") @ W.m(pre + "ac2_w")
            dx = torch.zeros(N, 3).index_add_(0, int_r, (x[int_r] - x[int_c]) * (alpha * se)[:, None])
            return h, x + dx.clamp(-cmax, cmax), pair_out

        pair_cur = P0
        for l in range(L):
            if plus:
                h, x = gcl_plus(f"gcl{l}.", h, x, layer=l)
                h, x, pair_cur = att_plus(f"att{l}.", pair_cur, h, x, layer=l)
            else:
                h, x = gcl(f"gcl{l}.", h, x)
                h, x = att(f"att{l}.", l, h, x)
            a, bb = las_int
            cur = ((x[a] - x[bb]) ** 2).sum(1)
            ref = ((xl[a] - xl[bb]) ** 2).sum(1)
            force = 2 * (cur - ref)[:, None] * (2 * (x[a] - x[bb]))
            x = x + (torch.zeros(N, 3).index_add_(0, bb, force) * cfg.geometry_reg_step_size).clamp(-lcl, lcl)
        h, x = gcl_plus("out.", h, x, need_h=last, layer=L) if plus else gcl("out.", h, x, need_h=last)
        pair_last = pair_cur
        if last:
            if plus:
                h = drop(h, -1, "stack_out")
            h_final = F.linear(h, W.m("out_w"), W.m("out_b"))
        x_state = torch.where(moves[:, None], x, x_state)
        if not last:
            x_state = x_state.detach()    # refine_coord: earlier iterations run under no_grad in the reference (att_model.py:227-236)
    X_out = torch.empty_like(batch.X)
    X_out[permt, 0] = x_state
    H_out = torch.empty_like(batch.H)
    H_out[permt] = h_final
    if plus:
        return X_out, H_out, stats, pair_last
    return X_out, H_out, stats
