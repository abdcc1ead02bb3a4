"""CPU oracle for the FABind iterative docking stack (v1 weight layout).

TEST INFRASTRUCTURE.  This file is the checker, never the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
The shipped path (`fabind_b200`) never imports anything under `oracle/` and raises if its CUDA
library is missing.

It is a from-scratch, functional restatement (plain torch fp32 on CPU, no nn.Module, no
torch_scatter / torch_geometric) of the reference algorithm *as written* -- including the work the
CUDA path proves dead (full pair transition, per-edge first-layer GEMMs) -- so that it doubles as
the CPU baseline "port".  Every function cites the reference lines it follows.  All paths are
relative to /root/reference/FABind/fabind/.

Parity pinning: the reference ships no tests, golden vectors or checkpoints,
so this oracle is pinned against the reference ITSELF: `scripts/make_golden.py` runs the unmodified
reference modules (imported through `oracle/ref_shims.py`) on seeded inputs/weights in the dev
container and commits inputs+outputs under `tests/golden/`; `tests/test_oracle_golden.py` checks this
file against those vectors on any machine, and `tests/test_oracle_vs_reference.py` checks it
tensor-for-tensor against the live reference whenever /root/reference is present.

Weights are passed as a flat ``state_dict`` with the reference's own key names
(`gnn.gcl_0.edge_mlp.0.weight`, ...), so a reference checkpoint drops in unchanged.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------
def _lin(sd, prefix, x):
    w = sd[prefix + ".weight"]
    b = sd.get(prefix + ".bias")
    return F.linear(x, w, b)


def segment_sum(data, seg, n):
    """models/egnn.py:790-803 (unsorted_segment_sum): zero-filled result, scatter-add along dim 0."""
    out = data.new_zeros((n,) + tuple(data.shape[1:]))
    out.index_add_(0, seg, data)
    return out


def segment_mean(data, seg, n):
    """models/egnn.py:806-821: sum / count.clamp(min=1)."""
    s = segment_sum(data, seg, n)
    c = segment_sum(torch.ones_like(data), seg, n)
    return s / c.clamp(min=1)


def segment_softmax(val, seg, n):
    """torch_scatter.scatter_softmax as called at models/egnn.py:221 (max-shifted, per segment)."""
    mx = val.new_full((n,), float("-inf")).scatter_reduce(0, seg, val, reduce="amax", include_self=True)
    e = (val - mx[seg]).exp()
    s = val.new_zeros((n,)).index_add_(0, seg, e)
    return e / s[seg]


def complex_layout(batch_id, segment_id):
    """Per-complex node offsets and compound/protein block sizes.

    The reference builds dense [B, max_n, ...] blocks with torch_geometric.to_dense_batch
    (models/egnn.py:260-265, models/att_model.py:199-204); with the dataloader's node order
    (utils/utils.py:328-335) the compound side (glb_c + atoms) and protein side (glb_p + residues)
    of every complex are contiguous runs, so a block is a slice."""
    B = int(batch_id.max()) + 1
    counts = torch.bincount(batch_id, minlength=B)
    offs = torch.cumsum(counts, 0) - counts
    seg = segment_id.to(torch.bool)
    ncp = torch.bincount(batch_id[~seg], minlength=B)  # compound-side nodes (incl. glb_c)
    return B, offs.tolist(), counts.tolist(), ncp.tolist()


# ------------------------------------------------------------------------------------------------
# graph construction  (models/att_model.py:38-128)
# ------------------------------------------------------------------------------------------------
def build_edges(X, batch_id, segment_id, is_global, intra_cutoff, inter_cutoff):
    """ComplexGraph.construct_edges + _radial_edges.

    Candidate pairs are all ordered (row, col) with row != col in the same complex, enumerated
    row-major (torch.nonzero over the [N, max_n] same-complex mask, att_model.py:55-62).  Returned
    lists keep that order, which the reference inherits from boolean-mask indexing:
      ctx   = [prot-prot non-global within intra_cutoff ; global-normal same segment ; global-global]
      inter = compound<->protein non-global within inter_cutoff (both directions)
    Distance test is `torch.norm(xi - xj) <= cutoff` on fp32 (att_model.py:123-126).
    If no inter edge survives, the first candidate pair is used in both directions (:85-86)."""
    N = batch_id.shape[0]
    B, offs, counts, _ = complex_layout(batch_id, segment_id)
    rows, cols = [], []
    for b in range(B):
        idx = torch.arange(offs[b], offs[b] + counts[b])
        r = idx.repeat_interleave(counts[b])
        c = idx.repeat(counts[b])
        keep = r != c
        rows.append(r[keep])
        cols.append(c[keep])
    row = torch.cat(rows)
    col = torch.cat(cols)
    seg = segment_id.to(torch.long)
    rg, cg = is_global[row], is_global[col]
    not_glb = ~(rg | cg)
    rs, cs = seg[row], seg[col]
    xyz = X[:, 0]

    def within(sel, cutoff):
        r, c = row[sel], col[sel]
        d = torch.norm(xyz[r] - xyz[c], dim=-1)
        k = d <= cutoff
        return torch.stack([r[k], c[k]])

    pp = within((rs == cs) & (rs == 1) & not_glb, intra_cutoff)
    sel_int = (rs != cs) & not_glb
    inter = within(sel_int, inter_cutoff)
    if inter.shape[1] == 0:
        r0, c0 = row[sel_int][0], col[sel_int][0]
        inter = torch.stack([torch.stack([r0, c0]), torch.stack([c0, r0])])
    sel = (rs == cs) & ~not_glb
    glb_normal = torch.stack([row[sel], col[sel]])
    sel = rg & cg
    glb_glb = torch.stack([row[sel], col[sel]])
    ctx = torch.cat([pp, glb_normal, glb_glb], dim=1)
    # reduced tuple (att_model.py:87-89): per compound->protein edge its complex id and node offset
    fwd = inter[0] < inter[1]
    red_b = batch_id[inter[0][fwd]]
    red_off = torch.tensor(offs, dtype=torch.long)[red_b]
    return ctx, inter, (red_b, red_off)


# ------------------------------------------------------------------------------------------------
# geometry helpers
# ------------------------------------------------------------------------------------------------
def radial_per_sample(edges, x, batch_id):
    """models/egnn.py:767-787 (coord2radial, norm_type='per_sample', n_channel=1).

    radial_e = |x_row - x_col|^2, divided by sqrt(sum over the edges of the same complex of
    radial^2)."""
    row, col = edges
    diff = x[row] - x[col]                              # [E, 1, 3]
    radial = torch.bmm(diff, diff.transpose(-1, -2))    # [E, 1, 1]
    eb = batch_id[row]
    B = int(eb.max()) + 1 if eb.numel() else 0
    nrm = segment_sum(radial ** 2, eb, B).sqrt()
    return radial / nrm[eb], diff


# ------------------------------------------------------------------------------------------------
# MC_E_GCL  (models/egnn.py:20-144)
# ------------------------------------------------------------------------------------------------
def _nodrop(x, layer, name):
    return x


def make_drop(dropout, it):
    """Training-mode dropout of the v1 stack (models/egnn.py:82,106,236,398,461, models/cross_att.py:128) with the deterministic
    COLUMN-ONLY masks of fabind_b200/dropout.py (the form the goldens of scripts/make_golden.py::main_grad_dropout pin against the
    unmodified reference): dropout = (p, seed, colonly=True) or None; `it` = refinement iteration."""
    if dropout is None or dropout[0] <= 0:
        return _nodrop
    p, seed, colonly = dropout
    if not colonly:
        raise NotImplementedError("the oracle restates column-only masks (row x column masks depend on the library's internal row order)")
    from fabind_b200.dropout import keep_mask, site_id, iter_seed
    return lambda x, layer, name: x * keep_mask(iter_seed(seed, it), site_id(layer, name), 1, x.shape[-1], p, colonly=True)[0]


def gcl_forward(sd, pre, h, edges, x, batch_id, clamp, drop=_nodrop, layer=0):
    row, col = edges
    n = h.shape[0]
    radial, diff = radial_per_sample(edges, x, batch_id)
    # edge_model (egnn.py:68-87): cat[h_row, h_col, radial] -> Linear -> SiLU -> Linear -> SiLU
    m = torch.cat([h[row], h[col], radial.reshape(radial.shape[0], -1)], dim=1)
    m = F.silu(_lin(sd, pre + "edge_mlp.0", m))
    m = drop(F.silu(_lin(sd, pre + "edge_mlp.2", m)), layer, "edge2")      # egnn.py:82
    # coord_model (egnn.py:111-128): mean-aggregated, clamped update
    s = F.linear(F.silu(_lin(sd, pre + "coord_mlp.0", m)), sd[pre + "coord_mlp.2.weight"])
    trans = diff * s.unsqueeze(-1)
    x = x + segment_mean(trans, row, n).clamp(-clamp, clamp)
    # node_model (egnn.py:89-109): sum-aggregate, cat[h, agg] -> Linear -> SiLU -> Linear, residual
    agg = segment_sum(m, row, n)
    out = _lin(sd, pre + "node_mlp.2", F.silu(_lin(sd, pre + "node_mlp.0", torch.cat([h, agg], 1))))
    return h + drop(out, layer, "node2"), x                                # egnn.py:106


# ------------------------------------------------------------------------------------------------
# cross attention on the dense per-complex blocks (models/cross_att.py, models/model_utils.py)
# ------------------------------------------------------------------------------------------------
def interaction(sd, pre, p, c):
    """models/model_utils.py:200-223 with rm_layernorm: Linear_out(Linear_p(p)_i * Linear_c(c)_j)."""
    pp = _lin(sd, pre + "linear_p", p)
    cc = _lin(sd, pre + "linear_c", c)
    return _lin(sd, pre + "linear_out", pp[:, None, :] * cc[None, :, :])


def row_attention(sd, pre, xi, xj, pair, heads=4, dh=32, drop=_nodrop, layer=0, site="patt"):
    """RowAttentionBlock (models/cross_att.py:118-134) + gated multi-head Attention
    (models/model_utils.py:96-159, _attention :21-38) for ONE complex (no padding, so the 1e9
    mask bias of cross_att.py:124 never applies to a real entry).
    xi [I, C] queries, xj [J, C] keys/values, pair [I, J, C]."""
    bias = _lin(sd, pre + "linear", pair) * torch.sigmoid(_lin(sd, pre + "linear_g", pair))  # [I,J,h]
    q = F.linear(xi, sd[pre + "mha.linear_q.weight"]).view(-1, heads, dh) / math.sqrt(dh)
    k = F.linear(xj, sd[pre + "mha.linear_k.weight"]).view(-1, heads, dh)
    v = F.linear(xj, sd[pre + "mha.linear_v.weight"]).view(-1, heads, dh)
    a = torch.einsum("ihd,jhd->hij", q, k) + bias.permute(2, 0, 1)
    a = torch.softmax(a, dim=-1)
    o = torch.einsum("hij,jhd->ihd", a, v)
    g = torch.sigmoid(_lin(sd, pre + "mha.linear_g", xi)).view(-1, heads, dh)
    o = (o * g).reshape(-1, heads * dh)
    return xi + drop(_lin(sd, pre + "mha.linear_o", o), layer, site)      # cross_att.py:128


def transition(sd, pre, x):
    """models/model_utils.py:171-175 with rm_layernorm."""
    return _lin(sd, pre + "linear_2", _lin(sd, pre + "linear_1", x).relu())


def cross_attention(sd, pre, p, c, pair, drop=_nodrop, layer=0):
    """CrossAttentionModule.forward (models/cross_att.py:24-54), one complex."""
    p = row_attention(sd, pre + "p_attention_block.", p, c, pair, drop=drop, layer=layer, site="patt")
    c = row_attention(sd, pre + "c_attention_block.", c, p, pair.transpose(0, 1), drop=drop, layer=layer, site="catt")   # uses the NEW p
    p = p + transition(sd, pre + "p_transition.", p)
    c = c + transition(sd, pre + "c_transition.", c)
    pair = pair + interaction(sd, pre + "inter_layer.", p, c)
    pair = transition(sd, pre + "pair_transition.", pair)
    return p, c, pair


# ------------------------------------------------------------------------------------------------
# MC_Att_L  (models/egnn.py:147-333)
# ------------------------------------------------------------------------------------------------
def att_forward(sd, pre, h, inter, x, batch_id, segment_id, pair0, clamp, layout, drop=_nodrop, layer=0):
    B, offs, counts, ncp = layout
    row, col = inter
    n = h.shape[0]
    # --- trio_encoder (egnn.py:254-305): cross attention block per complex, re-flatten, pair gather
    new_h = torch.empty_like(h)
    pair_new = []
    for b in range(B):
        o, nc1, nn = offs[b], ncp[b], counts[b]
        c = h[o:o + nc1]
        p = h[o + nc1:o + nn]
        p, c, pr = cross_attention(sd, pre + "cross_attn_module.", p, c, pair0[b], drop, layer)
        new_h[o:o + nc1] = c
        new_h[o + nc1:o + nn] = p
        pair_new.append(pr)
    h = new_h
    # pair rows at the inter edges, in the reference's per-sample [lig->prot ; prot->lig] order
    # (egnn.py:287-304).  Edge lists are row-sorted and the compound side precedes the protein side
    # inside a complex, so that order coincides with the order of `inter` itself.
    eb = batch_id[row]
    off_t = torch.tensor(offs, dtype=torch.long)
    nc1_t = torch.tensor(ncp, dtype=torch.long)
    lr, lc = row - off_t[eb], col - off_t[eb]
    fwd = row < col                                    # compound -> protein
    pi = torch.where(fwd, lc - nc1_t[eb], lr - nc1_t[eb])   # protein index inside the block
    ci = torch.where(fwd, lr, lc)                           # compound index inside the block
    np1_t = torch.tensor(counts, dtype=torch.long) - nc1_t
    base = torch.cumsum(np1_t * nc1_t, 0) - np1_t * nc1_t         # first pair row of each complex
    pair_flat = torch.cat([pr.reshape(-1, pr.shape[-1]) for pr in pair_new], dim=0)
    pair_off = pair_flat[base[eb] + pi * nc1_t[eb] + ci]
    # --- interfacial attention (egnn.py:186-252)
    radial, diff = radial_per_sample(inter, x, batch_id)
    q = _lin(sd, pre + "linear_q", h[row])
    kv = _lin(sd, pre + "linear_kv", torch.cat([radial.reshape(-1, 1), h[col]], dim=1))
    k, v = kv[..., 0::2], kv[..., 1::2]
    alpha = (q * k).sum(1) + _lin(sd, pre + "attn_bias_proj", pair_off).squeeze(-1)
    alpha = segment_softmax(alpha, row, n)
    aw = alpha.unsqueeze(-1)
    h = h + drop(segment_sum(aw * v, row, n), layer, "agg")                # egnn.py:235-237
    cv = aw * F.linear(F.silu(_lin(sd, pre + "coord_mlp.0", v)), sd[pre + "coord_mlp.2.weight"])
    x = x + segment_sum(diff * cv.unsqueeze(-1), row, n).clamp(-clamp, clamp)
    return h, x, alpha, pair_new


# ------------------------------------------------------------------------------------------------
# LAS constrained step  (models/egnn.py:433-449)
# ------------------------------------------------------------------------------------------------
def las_step(x, x_ref, las, step, clamp):
    xs, rs = x.squeeze(1), x_ref.squeeze(1)
    a, b = las
    cur = ((xs[a] - xs[b]) ** 2).sum(1)
    ref = ((rs[a] - rs[b]) ** 2).sum(1)
    force = 2 * (cur - ref)[:, None] * (2 * (xs[a] - xs[b]))
    delta = segment_sum(force, b, xs.shape[0])
    return (xs + (delta * step).clamp(min=-clamp, max=clamp)).unsqueeze(1)


# ------------------------------------------------------------------------------------------------
# MCAttEGNN.forward  (models/egnn.py:392-466)
# ------------------------------------------------------------------------------------------------
def egnn_forward(sd, pre, cfg, h, x, ctx, inter, las, x_ref, batch_id, segment_id, pair0, layout,
                 trace=None, drop=_nodrop):
    clamp = 10.0 / cfg.coordinate_scale                       # normalize_coord(10), egnn.py:378
    h = drop(_lin(sd, pre + "linear_in", h), -1, "stack_in")  # egnn.py:397-398
    x = x.clone()
    atts = []
    for i in range(cfg.n_layers):
        h, x = gcl_forward(sd, f"{pre}gcl_{i}.", h, ctx, x, batch_id, clamp, drop, i)
        if trace is not None:
            trace.append((f"gcl_{i}", h.clone(), x.clone()))
        h, x, att, _ = att_forward(sd, f"{pre}att_{i}.", h, inter, x, batch_id, segment_id, pair0,
                                   clamp, layout, drop, i)
        atts.append(att)
        if trace is not None:
            trace.append((f"att_{i}", h.clone(), x.clone()))
        x = las_step(x, x_ref, las, cfg.geometry_reg_step_size, 15.0 / cfg.coordinate_scale)
        if trace is not None:
            trace.append((f"las_{i}", h.clone(), x.clone()))
    h, x = gcl_forward(sd, pre + "out_layer.", h, ctx, x, batch_id, clamp, drop, cfg.n_layers)
    h = _lin(sd, pre + "linear_out", drop(h, -1, "stack_out"))   # egnn.py:461-462
    return h, x, atts


# ------------------------------------------------------------------------------------------------
# EfficientMCAttModel.forward  (models/att_model.py:170-246), eval mode, refine='refine_coord'
# ------------------------------------------------------------------------------------------------
def make_cfg(n_layers=4, n_iter=8, coordinate_scale=5.0, intra_cutoff=8.0, inter_cutoff=10.0,
             geometry_reg_step_size=0.001):
    return SimpleNamespace(n_layers=n_layers, n_iter=n_iter, coordinate_scale=coordinate_scale,
                           intra_cutoff=intra_cutoff, inter_cutoff=inter_cutoff,
                           geometry_reg_step_size=geometry_reg_step_size)


def initial_pair(sd, H, layout):
    """att_model.py:198-206: pair_embed0 = InteractionModule(p, c) per complex (hidden = H)."""
    B, offs, counts, ncp = layout
    out = []
    for b in range(B):
        o, nc1, nn = offs[b], ncp[b], counts[b]
        out.append(interaction(sd, "inter_layer.", H[o + nc1:o + nn], H[o:o + nc1]))
    return out


def model_forward(sd, cfg, X, H, batch_id, segment_id, mask, is_global, compound_edge_index,
                  LAS_edge_index, X_LAS, trace=None, return_edges=False, grad_last_iter_only=False, dropout=None):
    """Returns (X, H) like the reference; X is updated on a copy (the reference mutates its
    argument in place, att_model.py:236,245 -- callers that want that effect copy back).
    grad_last_iter_only: autograd semantics of refine='refine_coord' (att_model.py:227-245): every iteration but the last runs
    under no_grad (graph construction always does); used by the gradient goldens that pin this oracle for the training path.
    dropout = (p, seed, True): train() mode of the reference with column-only masks (see make_drop)."""
    X = X.clone()
    layout = complex_layout(batch_id, segment_id)
    pair0 = initial_pair(sd, H, layout)
    intra = cfg.intra_cutoff / cfg.coordinate_scale           # att_model.py:34-35
    inter_c = cfg.inter_cutoff / cfg.coordinate_scale
    edges_seen = []
    H_out = None
    for r in range(cfg.n_iter):
        with torch.no_grad():
            ctx, inter, _ = build_edges(X, batch_id, segment_id, is_global, intra, inter_c)
        ctx = torch.cat([compound_edge_index, ctx], dim=1)   # att_model.py:231
        if return_edges:
            edges_seen.append((ctx, inter))
        tr = [] if trace is not None else None
        with torch.set_grad_enabled(torch.is_grad_enabled() and (not grad_last_iter_only or r == cfg.n_iter - 1)):
            h_new, Z, atts = egnn_forward(sd, "gnn.", cfg, H, X, ctx, inter, LAS_edge_index, X_LAS,
                                          batch_id, segment_id, pair0, layout, trace=tr, drop=make_drop(dropout, r))
        if trace is not None:
            trace.append((r, tr, atts))
        X[mask] = Z[mask]
        H_out = h_new                                         # only the last iteration's H is kept
    if return_edges:
        return X, H_out, edges_seen
    return X, H_out
