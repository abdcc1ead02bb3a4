"""CPU oracle for the FABind+ ("plus") weight layout of the iterative docking stack.

TEST INFRASTRUCTURE (same rules as oracle/fabind_oracle.py): the checker, never the product.

Functional restatement (torch fp32, CPU) of the FABind+ variant of the path, published configuration
(`FABind_plus/README.md:125-141`: --use-ln-mlp --mlp-hidden-scale 1 --mean-layers 5 --n-iter 8 --rm-layernorm
--add-attn-pair-bias --explicit-pair-embed --add-cross-attn-layer, eval mode so --dropout 0.1 is inactive;
argparse defaults `utils/parsing.py:169-195`: mha_heads 4, rel_dis_pair_bias 'no', inter_additional_mlp off,
only_last_LAS off).  All paths below are relative to /root/reference/FABind_plus/fabind/.

Deltas against v1 (everything else is shared with oracle/fabind_oracle.py):
  * every Sequential MLP becomes LayerNorm -> Linear -> ReLU -> Linear (-> ReLU)  (models/model_utils.py:10-74)
  * the pair embedding is PROPAGATED layer to layer inside one `gnn` call and returned (models/egnn.py:380-392,
    431-433; models/att_model.py:209-223); every refinement iteration restarts from pair_embed0
  * pair_transition has no residual: pair <- mask * MLPwithLastAct(pair + inter32(p, c))  (models/cross_att.py:43-45)

Pinned like the v1 oracle: `scripts/make_golden.py` runs the unmodified FABind+ modules through
`oracle/ref_shims.py` and commits `tests/golden/plus_*.pt`; `tests/test_oracle_golden.py` checks this file against
them anywhere, `tests/test_oracle_vs_reference.py` against the live reference in the dev container.
"""
import torch
import torch.nn.functional as F

from .fabind_oracle import (_lin, segment_sum, segment_mean, segment_softmax, complex_layout, build_edges,
                            radial_per_sample, interaction, row_attention, las_step, make_cfg, initial_pair)

LN_EPS = 1e-5   # torch.nn.LayerNorm default (models/model_utils.py:15,37,60)


def _ln(sd, pre, z):
    return F.layer_norm(z, (z.shape[-1],), sd[pre + "layernorm.weight"], sd[pre + "layernorm.bias"], LN_EPS)


def mlp_last_act(sd, pre, z):
    """MLPwithLastAct (models/model_utils.py:32-53), eval: relu(linear2(relu(linear1(LN(z)))))."""
    return _lin(sd, pre + "linear2", _lin(sd, pre + "linear1", _ln(sd, pre, z)).relu()).relu()


def mlp_wo_bias(sd, pre, z):
    """MLPwoBias (models/model_utils.py:55-74), eval: linear2(relu(linear1(LN(z)))), linear2 has no bias."""
    return F.linear(_lin(sd, pre + "linear1", _ln(sd, pre, z)).relu(), sd[pre + "linear2.weight"])


def gcl_forward(sd, pre, h, edges, x, batch_id, clamp):
    """MC_E_GCL.forward (models/egnn.py:44-115)."""
    row, col = edges
    n = h.shape[0]
    radial, diff = radial_per_sample(edges, x, batch_id)
    m = mlp_last_act(sd, pre + "edge_mlp.", torch.cat([h[row], h[col], radial.reshape(radial.shape[0], -1)], dim=1))
    s = mlp_wo_bias(sd, pre + "coord_mlp.", m)
    x = x + segment_mean(diff * s.unsqueeze(-1), row, n).clamp(-clamp, clamp)
    agg = segment_sum(m, row, n)
    return h + mlp_last_act(sd, pre + "node_mlp.", torch.cat([h, agg], 1)), x


def cross_attention(sd, pre, p, c, pair):
    """CrossAttentionModule.forward (models/cross_att.py:20-45), one complex (no padding: mask == 1)."""
    p = row_attention(sd, pre + "p_attention_block.", p, c, pair)
    c = row_attention(sd, pre + "c_attention_block.", c, p, pair.transpose(0, 1))   # uses the NEW p
    p = p + mlp_last_act(sd, pre + "p_transition.", p)
    c = c + mlp_last_act(sd, pre + "c_transition.", c)
    pair = pair + interaction(sd, pre + "inter_layer.", p, c)
    return p, c, mlp_last_act(sd, pre + "pair_transition.", pair)


def att_forward(sd, pre, h, inter, x, batch_id, segment_id, pair, clamp, layout):
    """MC_Att_L.forward (models/egnn.py:269-300); `pair` is the list of per-complex [Np', Nc', H] blocks."""
    B, offs, counts, ncp = layout
    row, col = inter
    n = h.shape[0]
    new_h = torch.empty_like(h)
    pair_new = []
    for b in range(B):
        o, nc1, nn = offs[b], ncp[b], counts[b]
        p, c, pr = cross_attention(sd, pre + "cross_attn_module.", h[o + nc1:o + nn], h[o:o + nc1], pair[b])
        new_h[o:o + nc1] = c
        new_h[o + nc1:o + nn] = p
        pair_new.append(pr)
    h = new_h
    eb = batch_id[row]
    off_t = torch.tensor(offs, dtype=torch.long)
    nc1_t = torch.tensor(ncp, dtype=torch.long)
    lr, lc = row - off_t[eb], col - off_t[eb]
    fwd = row < col
    pi = torch.where(fwd, lc - nc1_t[eb], lr - nc1_t[eb])
    ci = torch.where(fwd, lr, lc)
    np1_t = torch.tensor(counts, dtype=torch.long) - nc1_t
    base = torch.cumsum(np1_t * nc1_t, 0) - np1_t * nc1_t
    pair_flat = torch.cat([pr.reshape(-1, pr.shape[-1]) for pr in pair_new], dim=0)
    pair_off = pair_flat[base[eb] + pi * nc1_t[eb] + ci]
    radial, diff = radial_per_sample(inter, x, batch_id)
    q = _lin(sd, pre + "linear_q", h[row])
    kv = _lin(sd, pre + "linear_kv", torch.cat([radial.reshape(-1, 1), h[col]], dim=1))
    k, v = kv[..., 0::2], kv[..., 1::2]
    alpha = (q * k).sum(1) + _lin(sd, pre + "attn_bias_proj", pair_off).squeeze(-1)
    alpha = segment_softmax(alpha, row, n)
    aw = alpha.unsqueeze(-1)
    h = h + segment_sum(aw * v, row, n)
    cv = aw * mlp_wo_bias(sd, pre + "coord_mlp.", v)
    x = x + segment_sum(diff * cv.unsqueeze(-1), row, n).clamp(-clamp, clamp)
    return h, x, alpha, pair_new


def egnn_forward(sd, pre, cfg, h, x, ctx, inter, las, x_ref, batch_id, segment_id, pair0, layout, trace=None):
    """MCAttEGNN.forward (models/egnn.py:359-433)."""
    clamp = 10.0 / cfg.coordinate_scale
    h = _lin(sd, pre + "linear_in", h)
    x = x.clone()
    pair = pair0
    atts = []
    for i in range(cfg.n_layers):
        h, x = gcl_forward(sd, f"{pre}gcl_{i}.", h, ctx, x, batch_id, clamp)
        if trace is not None:
            trace.append((f"gcl_{i}", h.clone(), x.clone()))
        h, x, att, pair = att_forward(sd, f"{pre}att_{i}.", h, inter, x, batch_id, segment_id, pair, clamp, layout)
        atts.append(att)
        if trace is not None:
            trace.append((f"att_{i}", h.clone(), x.clone()))
        x = las_step(x, x_ref, las, cfg.geometry_reg_step_size, 15.0 / cfg.coordinate_scale)
    h, x = gcl_forward(sd, pre + "out_layer.", h, ctx, x, batch_id, clamp)
    h = _lin(sd, pre + "linear_out", h)
    return h, x, atts, pair


def dense_pair(pair_blocks):
    """[B, max Np', max Nc', H] zero-padded, as the reference returns it (to_dense_batch layout)."""
    B = len(pair_blocks)
    mp = max(p.shape[0] for p in pair_blocks)
    mc = max(p.shape[1] for p in pair_blocks)
    out = pair_blocks[0].new_zeros((B, mp, mc, pair_blocks[0].shape[-1]))
    for b, p in enumerate(pair_blocks):
        out[b, :p.shape[0], :p.shape[1]] = p
    return out


def model_forward(sd, cfg, X, H, batch_id, segment_id, mask, is_global, compound_edge_index, LAS_edge_index, X_LAS,
                  trace=None, return_edges=False, grad_last_iter_only=False):
    """EfficientMCAttModel.forward (models/att_model.py:166-223), eval mode -> (X, H, pair_embed_batched)."""
    X = X.clone()
    layout = complex_layout(batch_id, segment_id)
    pair0 = initial_pair(sd, H, layout)
    intra = cfg.intra_cutoff / cfg.coordinate_scale
    inter_c = cfg.inter_cutoff / cfg.coordinate_scale
    edges_seen = []
    H_out, pair = None, None
    for r in range(cfg.n_iter):
        with torch.no_grad():
            ctx, inter, _ = build_edges(X, batch_id, segment_id, is_global, intra, inter_c)
        ctx = torch.cat([compound_edge_index, ctx], dim=1)
        if return_edges:
            edges_seen.append((ctx, inter))
        tr = [] if trace is not None else None
        # refine='refine_coord' (att_model.py:196-218): every iteration but the last under no_grad
        with torch.set_grad_enabled(torch.is_grad_enabled() and (not grad_last_iter_only or r == cfg.n_iter - 1)):
            h_new, Z, atts, pair = egnn_forward(sd, "gnn.", cfg, H, X, ctx, inter, LAS_edge_index, X_LAS, batch_id,
                                                segment_id, pair0, layout, trace=tr)
        if trace is not None:
            trace.append((r, tr, atts))
        X[mask] = Z[mask]
        H_out = h_new
    out = (X, H_out, dense_pair(pair))
    return out + (edges_seen,) if return_edges else out
