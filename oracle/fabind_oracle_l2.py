"""CPU oracle for the L2 wrapper around the docking stack (TEST INFRASTRUCTURE, see fabind_oracle.py).

Restates `IaBNet_mean_and_pocket_prediction_cls_coords_dependent.forward` (eval mode, stage 2) and
`.inference` of FABind/fabind/models/model.py, citing lines, on top of `fabind_oracle.model_forward`
for the two EfficientMCAttModel instances (pocket stage: hidden 128, 1 layer, 1 iteration, whole
protein; docking stage: hidden 512, mean_layers x n_iter).  Weights: the reference's state_dict.
Pinned against the unmodified reference by tests/test_oracle_vs_reference.py and tests/golden/l2_*.pt.
"""
import torch
import torch.nn.functional as F

from . import fabind_oracle as orc


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _assemble(glb_c, comp, comp_batch, glb_p, prot, prot_batch, B):
    """model.py:104-115 / 205-215: per complex [glb_c | compound rows | glb_p | protein rows]."""
    parts = []
    for i in range(B):
        parts += [glb_c, comp[comp_batch == i], glb_p, prot[prot_batch == i]]
    return torch.cat(parts, dim=0)


def keep_node(protein_xyz, radius, center):
    """utils/utils.py:147-158 (get_keepNode_tensor, no noise)."""
    dis = torch.sqrt(torch.sum((protein_xyz - center.unsqueeze(0)) ** 2, dim=-1))
    return dis < radius


def gumbel_softmax_no_random(logits, tau, hard):
    """utils/utils.py:687-699."""
    y_soft = (logits / tau).softmax(-1)
    if hard:
        index = y_soft.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
        return y_hard - y_soft + y_soft
    return y_soft


def pocket_stage(sd, args, data):
    """model.py:98-144: input linears, global nodes, shrink, pocket_pred_model, enlarge, per-residue logit."""
    scale = args.coordinate_scale
    cb, pbw = data['compound'].batch, data['protein_whole'].batch
    wp = data['complex_whole_protein']
    B = int(wp.batch.max()) + 1
    comp = F.linear(data['compound'].node_feats, sd['compound_linear_whole_protein.weight'], sd['compound_linear_whole_protein.bias'])
    prot = F.linear(data['protein_whole'].node_feats, sd['protein_linear_whole_protein.weight'], sd['protein_linear_whole_protein.bias'])
    x = _assemble(sd['glb_c'], comp, cb, sd['glb_p'], prot, pbw, B)
    x = F.linear(x, sd['embedding_shrink.weight'], sd['embedding_shrink.bias'])
    cfg = orc.make_cfg(n_layers=args.pocket_pred_layers, n_iter=args.pocket_pred_n_iter, coordinate_scale=scale,
                       intra_cutoff=args.intra_cutoff, inter_cutoff=args.inter_cutoff,
                       geometry_reg_step_size=args.geometry_reg_step_size)
    X = (wp.node_coords / scale).unsqueeze(-2)
    XL = (wp.node_coords_LAS / scale).unsqueeze(-2)
    _, Hout = orc.model_forward(_sub(sd, 'pocket_pred_model.'), cfg, X, x, wp.batch, wp.segment, wp.mask, wp.is_global,
                                data['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index,
                                data['complex_whole_protein', 'LAS', 'complex_whole_protein'].edge_index, XL)
    out = F.linear(Hout, sd['embedding_enlarge.weight'], sd['embedding_enlarge.bias'])
    seg = wp.segment.to(torch.bool)
    comp_out = out[(~seg) & (~wp.is_global)]
    prot_out = out[seg & (~wp.is_global)]
    # protein_to_pocket = Transition_diff_out_dim (model.py:11-24): LayerNorm, Linear(H,4H), ReLU, Linear(4H,1)
    z = F.layer_norm(prot_out, (prot_out.shape[-1],), sd['protein_to_pocket.layernorm.weight'], sd['protein_to_pocket.layernorm.bias'])
    z = F.linear(F.linear(z, sd['protein_to_pocket.linear1.weight'], sd['protein_to_pocket.linear1.bias']).relu(),
                 sd['protein_to_pocket.linear2.weight'], sd['protein_to_pocket.linear2.bias']).squeeze(-1)
    return B, comp_out, prot_out, z      # z: per-residue pocket logit, flat in protein order


def _dense(flat, batch, B, fill=0.0):
    counts = torch.bincount(batch, minlength=B)
    Lmax = int(counts.max())
    out = flat.new_full((B, Lmax) + tuple(flat.shape[1:]), fill)
    mask = torch.zeros((B, Lmax), dtype=torch.bool)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(flat.shape[0]) - starts[batch]
    out[batch, pos] = flat
    mask[batch, pos] = True
    return out, mask


def _docking_inputs(sd, args, data, B, comp_out, prot_out, centers):
    """model.py:173-300 (stage 2) == :439-560 (inference): crop by predicted centre, re-assemble the complex graph."""
    cb, pbw = data['compound'].batch, data['protein_whole'].batch
    feats, coords, coords_las, seg, msk, glb, bat, c2c, las, pocket_xyz, pocket_bat, dis_map = [], [], [], [], [], [], [], [], [], [], [], []
    less5 = 0
    n_nodes = 0
    for i in range(B):
        prot_i = data.node_xyz_whole[pbw == i]
        keep = keep_node(prot_i, args.pocket_radius, centers[i])
        if keep.sum() < 5:
            keep[:100] = True
            less5 += 1
        pemb = prot_out[pbw == i][keep]
        cemb = comp_out[cb == i]
        feats += [sd['glb_c'], cemb, sd['glb_p'], pemb]
        pc = prot_i[keep]
        lig = data['compound'].node_coords[cb == i]
        z = torch.zeros((1, 3))
        coords += [z, lig - lig.mean(dim=0).reshape(1, 3) + pc.mean(dim=0).reshape(1, 3), z, pc]
        coords_las += [z, data['compound'].rdkit_coords[cb == i], z, torch.zeros_like(pc)]
        n_p, n_c = pemb.shape[0], cemb.shape[0]
        s = torch.zeros(n_p + n_c + 2, dtype=torch.bool); s[n_c + 1:] = True
        m = torch.zeros(n_p + n_c + 2, dtype=torch.bool); m[:n_c + 2] = True
        g = torch.zeros(n_p + n_c + 2, dtype=torch.bool); g[0] = True; g[n_c + 1] = True
        seg.append(s); msk.append(m); glb.append(g)
        c2c.append(data['compound_atom_edge_list'].x[data['compound_atom_edge_list'].batch == i].t() + n_nodes)
        las.append(data['LAS_edge_list'].x[data['LAS_edge_list'].batch == i].t() + n_nodes)
        bat.append(torch.full((n_p + n_c + 2,), i, dtype=torch.long))
        pocket_bat.append(torch.full((n_p,), i, dtype=torch.long))
        pocket_xyz.append(pc)
        dm = torch.cdist(pc, lig.to(torch.float32)).flatten()
        dis_map.append(torch.clamp(dm, max=10.0))
        n_nodes += n_p + n_c + 2
    return dict(H=torch.cat(feats), X=torch.cat(coords).float(), XL=torch.cat(coords_las).float(), seg=torch.cat(seg),
                mask=torch.cat(msk), glb=torch.cat(glb), batch=torch.cat(bat), c2c=torch.cat(c2c, 1).long(),
                las=torch.cat(las, 1).long(), pocket_xyz=torch.cat(pocket_xyz), pocket_batch=torch.cat(pocket_bat),
                dis_map=torch.cat(dis_map), less5=less5)


def _dock(sd, args, di):
    scale = args.coordinate_scale
    cfg = orc.make_cfg(n_layers=args.mean_layers, n_iter=args.n_iter, coordinate_scale=scale,
                       intra_cutoff=args.intra_cutoff, inter_cutoff=args.inter_cutoff,
                       geometry_reg_step_size=args.geometry_reg_step_size)
    return orc.model_forward(_sub(sd, 'complex_model.'), cfg, (di['X'] / scale).unsqueeze(-2), di['H'], di['batch'],
                             di['seg'], di['mask'], di['glb'], di['c2c'], di['las'], (di['XL'] / scale).unsqueeze(-2))


def _stage1_inputs(sd, args, data, B, comp_out, prot_out):
    """model.py:302-320: dataloader pocket crop (`data['pocket'].keepNode`) and the pre-built `data['complex']` graph."""
    cb, pkb = data['compound'].batch, data['pocket'].batch
    pemb = prot_out[data['pocket'].keepNode]
    feats = []
    for i in range(B):
        feats += [sd['glb_c'], comp_out[cb == i], sd['glb_p'], pemb[pkb == i]]
    cx = data['complex']
    return dict(H=torch.cat(feats), X=cx.node_coords, XL=cx.node_coords_LAS, seg=cx.segment.to(torch.bool), mask=cx.mask,
                glb=cx.is_global, batch=cx.batch, c2c=data['complex', 'c2c', 'complex'].edge_index,
                las=data['complex', 'LAS', 'complex'].edge_index, pocket_xyz=data.node_xyz, pocket_batch=pkb,
                dis_map=data.dis_map, less5=0)


def forward_stage2(sd, args, data):
    return forward_eval(sd, args, data, 2)


def forward_eval(sd, args, data, stage):
    """model.py:82-369 with model.eval(), train=False: final_stage = stage (:170-171)."""
    scale = args.coordinate_scale
    B, comp_out, prot_out, logit = pocket_stage(sd, args, data)
    pbw = data['protein_whole'].batch
    cls_dense, pmask = _dense(logit, pbw, B)
    cls_dense = cls_dense * pmask
    pocket_cls, _ = _dense(data.pocket_idx, pbw, B, fill=0)
    coords_dense, _ = _dense(data.node_xyz_whole, pbw, B)
    # soft pocket centre (model.py:146-158)
    p_true = cls_dense.sigmoid().unsqueeze(-1)
    prob = torch.clamp(torch.cat([1.0 - p_true, p_true], dim=-1), min=1e-6, max=1 - 1e-6)
    one_hot = gumbel_softmax_no_random(torch.log(prob), args.gs_tau, args.gs_hard)
    w = (one_hot[:, :, 1] * pmask).unsqueeze(-1)
    centers = (w * coords_dense).sum(dim=1) / w.sum(dim=1)
    di = (_docking_inputs(sd, args, data, B, comp_out, prot_out, centers) if stage == 2
          else _stage1_inputs(sd, args, data, B, comp_out, prot_out))
    X, H = _dock(sd, args, di)
    # head (model.py:336-367)
    seg, glb = di['seg'], di['glb']
    pocket_out = H[seg & ~glb]
    compound_out = H[~seg & ~glb]
    lig_xyz = X[~seg & ~glb].squeeze(-2)
    cb = data['compound'].batch
    pocket_xyz_n = di['pocket_xyz'] / scale
    ln = lambda t: F.layer_norm(t, (t.shape[-1],), sd['layernorm.weight'], sd['layernorm.bias'])
    y_pred, y_coords = [], []
    for i in range(B):
        po, co = ln(pocket_out[di['pocket_batch'] == i]), ln(compound_out[cb == i])
        z = po[:, None, :] * co[None, :, :]
        bmap = F.linear(F.linear(z, sd['distmap_mlp.0.weight'], sd['distmap_mlp.0.bias']).relu(),
                        sd['distmap_mlp.2.weight'], sd['distmap_mlp.2.bias']).squeeze(-1)
        y_pred.append(bmap.reshape(-1).sigmoid() * 10)
        dist = torch.cdist(pocket_xyz_n[di['pocket_batch'] == i], lig_xyz[cb == i])
        y_coords.append(torch.clamp(dist.reshape(-1) * scale, 0, 10))
    return (lig_xyz * scale, cb, torch.cat(y_pred), torch.cat(y_coords), cls_dense, pocket_cls, pmask, coords_dense,
            centers, di['dis_map'], di['less5'])


def inference(sd, args, data):
    """model.py:371-580."""
    scale = args.coordinate_scale
    B, comp_out, prot_out, logit = pocket_stage(sd, args, data)
    pbw = data['protein_whole'].batch
    centers = torch.zeros((B, 3))
    for i in range(B):
        li = logit[pbw == i]
        xyz = data.node_xyz_whole[pbw == i]
        sel = li.sigmoid().round().int() == 1
        if sel.sum() != 0:
            centers[i] = xyz[sel].mean(dim=0)
        else:
            p_true = li.sigmoid().unsqueeze(-1)
            one_hot = gumbel_softmax_no_random(torch.log(torch.cat([1.0 - p_true, p_true], dim=-1)), args.gs_tau, args.gs_hard)
            w = one_hot[:, 1].unsqueeze(-1)
            centers[i] = (w * xyz).sum(dim=0) / w.sum(dim=0)
    di = _docking_inputs(sd, args, data, B, comp_out, prot_out, centers)
    X, _ = _dock(sd, args, di)
    return X[~di['seg'] & ~di['glb']].squeeze(-2) * scale, data['compound'].batch
