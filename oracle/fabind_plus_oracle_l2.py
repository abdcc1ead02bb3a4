"""CPU oracle for the FABind+ L2 wrapper `FABindPlus` (TEST INFRASTRUCTURE, see fabind_oracle.py).

Restates `FABindPlus.forward(data, stage=2, train=False)` (eval mode) and `.inference(data)` of
FABind_plus/fabind/models/model.py on top of `fabind_plus_oracle.model_forward` for the two EfficientMCAttModel
instances.  Configuration = the published one (FABind_plus/README.md:125-141): --use-for-radius-pred ligand,
--pocket-radius-buffer 5, --min-pocket-radius 20, --dis-map-thres 15, no clustering, no confidence head.
Deltas against the v1 wrapper (oracle/fabind_oracle_l2.py):
  * pocket radius head  relu(MLP(sum of ligand atom embeddings))  -> per-complex crop radius  (model.py:110-122,223-231)
  * `protein_to_pocket` / `distmap_mlp` are `MLP`s (LayerNorm, Linear, ReLU, Linear; models/model_utils.py:10-30)
  * the cropped pocket is re-centred on its own mean; that mean is returned as `pocket_center_bias` and subtracted
    from `data.coords` in place (model.py:255-258)
  * the distance head reads the propagated pair embedding `pair[:, 1:, 1:]` (model.py:379-384), range `dis_map_thres`
Pinned by tests/golden/l2plus_*.pt (generated from the unmodified reference) and, in the dev container, against the
live reference.
"""
import torch
import torch.nn.functional as F

from . import fabind_oracle as orc
from . import fabind_plus_oracle as orcp
from .fabind_oracle_l2 import _sub, _assemble, keep_node, gumbel_softmax_no_random, _dense


def mlp(sd, pre, z):
    """MLP (models/model_utils.py:10-30), eval: linear2(relu(linear1(LN(z))))."""
    z = F.layer_norm(z, (z.shape[-1],), sd[pre + "layernorm.weight"], sd[pre + "layernorm.bias"], orcp.LN_EPS)
    return F.linear(F.linear(z, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"]).relu(),
                    sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])


def _cfg(args, n_layers, n_iter):
    return orc.make_cfg(n_layers=n_layers, n_iter=n_iter, coordinate_scale=args.coordinate_scale, intra_cutoff=args.intra_cutoff,
                        inter_cutoff=args.inter_cutoff, geometry_reg_step_size=args.geometry_reg_step_size)


def pocket_stage(sd, args, data):
    """model.py:72-139: input linears, global nodes, shrink, pocket_pred_model, enlarge, radius head, per-residue logit."""
    if args.use_for_radius_pred != "ligand":
        raise NotImplementedError("published configuration: --use-for-radius-pred ligand")
    scale = args.coordinate_scale
    cb, pbw = data['compound'].batch, data['protein_whole'].batch
    wp = data['complex_whole_protein']
    B = int(wp.batch.max()) + 1
    comp = F.linear(data['compound'].node_feats, sd['compound_linear_whole_protein.weight'], sd['compound_linear_whole_protein.bias'])
    prot = F.linear(data['protein_whole'].node_feats, sd['protein_linear_whole_protein.weight'], sd['protein_linear_whole_protein.bias'])
    x = _assemble(sd['glb_c'], comp, cb, sd['glb_p'], prot, pbw, B)
    x = F.linear(x, sd['embedding_shrink.weight'], sd['embedding_shrink.bias'])
    X = (wp.node_coords / scale).unsqueeze(-2)
    XL = (wp.node_coords_LAS / scale).unsqueeze(-2)
    _, Hout, _ = orcp.model_forward(_sub(sd, 'pocket_pred_model.'), _cfg(args, args.pocket_pred_layers, args.pocket_pred_n_iter),
                                    X, x, wp.batch, wp.segment, wp.mask, wp.is_global,
                                    data['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index,
                                    data['complex_whole_protein', 'LAS', 'complex_whole_protein'].edge_index, XL)
    out = F.linear(Hout, sd['embedding_enlarge.weight'], sd['embedding_enlarge.bias'])
    seg = wp.segment.to(torch.bool)
    comp_out = out[(~seg) & (~wp.is_global)]
    prot_out = out[seg & (~wp.is_global)]
    # pocket radius head on the per-complex SUM of ligand atom embeddings (model.py:110-114)
    comp_sum = _dense(comp_out, cb, B)[0].sum(dim=1)        # to_dense_batch(...).sum(dim=1), as written
    radius_pred = mlp(sd, 'pocket_radius_head.', comp_sum).relu()             # [B, 1]
    logit = mlp(sd, 'protein_to_pocket.', prot_out).squeeze(-1)                # flat, protein order
    return B, comp_out, prot_out, logit, radius_pred


def crop_radius(args, radius_pred_i):
    """model.py:223-231 (python-float arithmetic on `.item()` of an fp32 tensor op)"""
    if args.pocket_radius_buffer <= 2.0:
        r = (radius_pred_i * args.pocket_radius_buffer).item()
    else:
        r = (radius_pred_i + args.pocket_radius_buffer).item()
    if r < args.min_pocket_radius:
        r = args.min_pocket_radius
    if args.force_fix_radius:
        r = args.pocket_radius
    return r


def soft_centers(args, logit, pbw, B, data):
    """model.py:130-145 (eval: gumbel_softmax_no_random)"""
    cls_dense, pmask = _dense(logit, pbw, B)
    cls_dense = cls_dense * pmask
    coords_dense, _ = _dense(data.node_xyz_whole, pbw, B)
    p_true = cls_dense.sigmoid().unsqueeze(-1)
    prob = torch.clamp(torch.cat([1.0 - p_true, p_true], dim=-1), min=1e-6, max=1 - 1e-6)
    one_hot = gumbel_softmax_no_random(torch.log(prob), args.gs_tau, args.gs_hard)
    w = (one_hot[:, :, 1] * pmask).unsqueeze(-1)
    return cls_dense, pmask, coords_dense, (w * coords_dense).sum(dim=1) / w.sum(dim=1)


def docking_inputs(sd, args, data, B, comp_out, prot_out, centers, radius_pred, shift_data_coords):
    """model.py:202-327 (stage 2) == :505-611 (inference, which leaves data.coords alone)."""
    cb, pbw = data['compound'].batch, data['protein_whole'].batch
    feats, coords, coords_las, seg, msk, glb, bat, c2c, las, pocket_xyz, pocket_bat, dis_map = [], [], [], [], [], [], [], [], [], [], [], []
    bias = torch.zeros((B, 3))
    less5, n_nodes = 0, 0
    for i in range(B):
        prot_i = data.node_xyz_whole[pbw == i]
        keep = keep_node(prot_i, crop_radius(args, radius_pred[i]), centers[i])
        if keep.sum() < 5:
            keep[:100] = True
            less5 += 1
        pemb, cemb = prot_out[pbw == i][keep], comp_out[cb == i]
        feats += [sd['glb_c'], cemb, sd['glb_p'], pemb]
        pc = prot_i[keep]
        center = pc.mean(dim=0).reshape(1, 3)
        pc = pc - center
        if shift_data_coords:
            data.coords[cb == i] = data.coords[cb == i] - center
        bias[i] = center.squeeze()
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
        dm = torch.cdist(pc, lig.to(torch.float32) - center).flatten()
        dis_map.append(torch.clamp(dm, max=args.dis_map_thres))
        n_nodes += n_p + n_c + 2
    return dict(H=torch.cat(feats), X=torch.cat(coords).float(), XL=torch.cat(coords_las).float(), seg=torch.cat(seg),
                mask=torch.cat(msk), glb=torch.cat(glb), batch=torch.cat(bat), c2c=torch.cat(c2c, 1).long(),
                las=torch.cat(las, 1).long(), pocket_xyz=torch.cat(pocket_xyz), pocket_batch=torch.cat(pocket_bat),
                dis_map=torch.cat(dis_map), less5=less5, bias=bias)


def _dock(sd, args, di):
    scale = args.coordinate_scale
    return orcp.model_forward(_sub(sd, 'complex_model.'), _cfg(args, args.mean_layers, args.n_iter), (di['X'] / scale).unsqueeze(-2),
                              di['H'], di['batch'], di['seg'], di['mask'], di['glb'], di['c2c'], di['las'],
                              (di['XL'] / scale).unsqueeze(-2))


def stage1_inputs(sd, args, data, B, comp_out, prot_out):
    """model.py:169-197 (stage == 1, teacher forcing with the dataloader's pocket): ligand re-centred on its own mean, pocket
    moved by `data.pocket_residue_center`; mutates data['complex'].node_coords and data.coords like the reference."""
    cb, pkb, cxb = data['compound'].batch, data['pocket'].batch, data['complex'].batch
    pemb = prot_out[data['pocket'].keepNode]
    feats = []
    for i in range(B):
        n_c = int((cb == i).sum())
        tc = data['complex'].node_coords[cxb == i]
        tc[1:n_c + 1] = tc[1:n_c + 1] - tc[1:n_c + 1].mean(dim=0)
        tc[n_c + 2:] = tc[n_c + 2:] - data.pocket_residue_center[i].unsqueeze(0)
        data['complex'].node_coords[cxb == i] = tc
        data.coords[cb == i] = data.coords[cb == i] - data.pocket_residue_center[i].unsqueeze(0)
        feats += [sd['glb_c'], comp_out[cb == i], sd['glb_p'], pemb[pkb == i]]
    cx = data['complex']
    return dict(H=torch.cat(feats), X=cx.node_coords, XL=cx.node_coords_LAS, seg=cx.segment.to(torch.bool), mask=cx.mask,
                glb=cx.is_global, batch=cx.batch, c2c=data['complex', 'c2c', 'complex'].edge_index,
                las=data['complex', 'LAS', 'complex'].edge_index, pocket_xyz=data.node_xyz, pocket_batch=pkb,
                dis_map=data.dis_map, less5=0, bias=torch.zeros((B, 3)))


def forward_stage2(sd, args, data):
    return forward_eval(sd, args, data, 2)


def forward_eval(sd, args, data, stage):
    """model.py:63-401 with model.eval(), train=False -> the reference's 13-tuple (mutates data.coords).  stage 2: predicted
    pocket; stage 1: the dataloader's pocket (the periodic test evaluation of main_fabind.py:179)."""
    if getattr(args, "only_last_LAS", False) or getattr(args, "use_clustering", False):
        raise NotImplementedError
    scale = args.coordinate_scale
    B, comp_out, prot_out, logit, radius_pred = pocket_stage(sd, args, data)
    pbw, cb = data['protein_whole'].batch, data['compound'].batch
    cls_dense, pmask, coords_dense, centers = soft_centers(args, logit, pbw, B, data)
    pocket_cls, _ = _dense(data.pocket_idx, pbw, B, fill=0)
    if stage == 1:
        di = stage1_inputs(sd, args, data, B, comp_out, prot_out)
    else:
        di = docking_inputs(sd, args, data, B, comp_out, prot_out, centers, radius_pred, shift_data_coords=True)
    X, H, pair = _dock(sd, args, di)
    seg, glb = di['seg'], di['glb']
    lig_xyz = X[~seg & ~glb].squeeze(-2)
    pocket_xyz_n = di['pocket_xyz'] / scale
    y_pred, y_coords = [], []
    for i in range(B):
        n_p, n_c = int((di['pocket_batch'] == i).sum()), int((cb == i).sum())
        z = pair[i, 1:n_p + 1, 1:n_c + 1]                                      # model.py:379
        y_pred.append(mlp(sd, 'distmap_mlp.', z).reshape(-1).sigmoid() * args.dis_map_thres)
        dist = torch.cdist(pocket_xyz_n[di['pocket_batch'] == i], lig_xyz[cb == i])
        y_coords.append(torch.clamp(dist.reshape(-1) * scale, 0, args.dis_map_thres))
    return (lig_xyz * scale, cb, torch.cat(y_pred), torch.cat(y_coords), cls_dense, pocket_cls, pmask, coords_dense, centers,
            di['dis_map'], di['less5'], radius_pred, di['bias'])


def inference(sd, args, data):
    """model.py:403-697: the same soft centre as forward; coordinates moved back by pocket_center_bias."""
    scale = args.coordinate_scale
    B, comp_out, prot_out, logit, radius_pred = pocket_stage(sd, args, data)
    pbw, cb = data['protein_whole'].batch, data['compound'].batch
    _, _, _, centers = soft_centers(args, logit, pbw, B, data)
    di = docking_inputs(sd, args, data, B, comp_out, prot_out, centers, radius_pred, shift_data_coords=False)
    X, _, _ = _dock(sd, args, di)
    return X[~di['seg'] & ~di['glb']].squeeze(-2) * scale + di['bias'][cb], cb
