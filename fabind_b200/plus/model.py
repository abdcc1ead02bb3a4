"""Drop-in for FABind_plus/fabind/models/model.py: `FABindPlus`, the L2 wrapper around the FABind+ docking stack.

Same constructor signature, parameter names/shapes (reference checkpoints load with strict=True) and the return tuples of
`forward(data, stage=2)` in eval mode (model.py:63-401: the 13-tuple, plus the in-place shift of `data.coords`) and
`inference(data)` (model.py:403-697).  Published configuration (README.md:125-141): --use-for-radius-pred ligand, no clustering,
`stage=1` (the dataloader's pocket) and `stage=2` (predicted pocket) in eval mode, sampling mode under train() + no_grad,
optional DBSCAN clustering and confidence head; `train=True` raises.
All arithmetic runs in libfabind_b200 kernels; the host side does index bookkeeping only.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..model import _assemble, _gemm, _i32, _layernorm, _select, complex_layout_np
from ..runtime import current_stream_ptr
from .att_model import EfficientMCAttModel
from .dbscan import dbscan_labels
from .model_utils import MLP, MLP4Confidence


HEAD_SITES = dict(pocket_radius_head=3000, protein_to_pocket=3001, distmap_mlp=3002)   # dropout sites of the wrapper's MLP heads


def _mlp_scalar(x, mlp, rows=None, bf16=False, drop=None):
    """MLP with out_channels=1 (LayerNorm -> Linear+ReLU -> Linear(.,1)): fb_layernorm_rows + tcgen05/FFMA GEMM whose epilogue
    carries the final Linear(.,1) as a row-dot.  Returns (dot partials [tiles, M], tiles, M); the caller adds linear2.bias."""
    l = _lib.lib()
    dev = x.device
    M = int(rows.numel()) if rows is not None else x.shape[0]
    D = x.shape[1]
    z = torch.empty((max(M, 1), D), dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    _lib.check(l.fb_layernorm_rows(x.data_ptr(), rows.data_ptr() if rows is not None else None, M, D, mlp.layernorm.weight.data_ptr(),
                                   mlp.layernorm.bias.data_ptr(), float(mlp.layernorm.eps), z.data_ptr(), int(bf16),
                                   current_stream_ptr(dev)), "fb_layernorm_rows")
    W1 = mlp.linear1.weight.to(torch.bfloat16) if bf16 else mlp.linear1.weight
    dot, tiles = _gemm(z[:M], W1, mlp.linear1.bias, act=2, dotv=mlp.linear2.weight[0].contiguous(), bf16=bf16, drop=drop)
    return dot, tiles, M


def _dot_finish(dot, tiles, M, bias):
    l = _lib.lib()
    out = torch.empty(M, dtype=torch.float32, device=dot.device)
    _lib.check(l.fb_dot_finish(dot.data_ptr(), tiles, dot.shape[1], M, bias.data_ptr(), out.data_ptr(), current_stream_ptr(dot.device)),
               "fb_dot_finish")
    return out


class FABindPlus(nn.Module):
    def __init__(self, args, embedding_channels=128, pocket_pred_embedding_channels=128):
        super().__init__()
        if getattr(args, "use_for_radius_pred", "ligand") != "ligand":
            raise NotImplementedError("fabind_b200 implements the published FABind+ configuration: --use-for-radius-pred ligand")
        if getattr(args, "only_last_LAS", False):
            raise NotImplementedError("fabind_b200: only_last_LAS is not built (off in the published configuration)")
        self.args = args
        self.coordinate_scale = args.coordinate_scale
        self.normalize_coord = lambda x: x / self.coordinate_scale
        self.unnormalize_coord = lambda x: x * self.coordinate_scale
        self.glb_c = nn.Parameter(torch.ones(1, embedding_channels))
        self.glb_p = nn.Parameter(torch.ones(1, embedding_channels))
        self.protein_linear_whole_protein = nn.Linear(1280, embedding_channels)
        self.compound_linear_whole_protein = nn.Linear(56, embedding_channels)
        self.embedding_shrink = nn.Linear(embedding_channels, pocket_pred_embedding_channels)
        self.embedding_enlarge = nn.Linear(pocket_pred_embedding_channels, embedding_channels)
        n_channel = 1
        self.pocket_pred_model = EfficientMCAttModel(
            args, pocket_pred_embedding_channels, pocket_pred_embedding_channels, n_channel, n_edge_feats=0,
            n_layers=args.pocket_pred_layers, n_iter=args.pocket_pred_n_iter, inter_cutoff=args.inter_cutoff,
            intra_cutoff=args.intra_cutoff, normalize_coord=self.normalize_coord, unnormalize_coord=self.unnormalize_coord)
        self.pocket_radius_head = MLP(args, embedding_channels=embedding_channels, n=args.mlp_hidden_scale, out_channels=1)
        self.protein_to_pocket = MLP(args, embedding_channels=embedding_channels, n=args.mlp_hidden_scale, out_channels=1)
        self.complex_model = EfficientMCAttModel(
            args, embedding_channels, embedding_channels, n_channel, n_edge_feats=0, n_layers=args.mean_layers, n_iter=args.n_iter,
            inter_cutoff=args.inter_cutoff, intra_cutoff=args.intra_cutoff, normalize_coord=self.normalize_coord,
            unnormalize_coord=self.unnormalize_coord)
        self.distmap_mlp = MLP(args, embedding_channels=embedding_channels, n=args.mlp_hidden_scale, out_channels=1)
        for lin in (self.protein_linear_whole_protein, self.compound_linear_whole_protein, self.embedding_shrink,
                    self.embedding_enlarge):
            nn.init.xavier_uniform_(lin.weight, gain=0.001)
        # confidence head of the sampling-based model (model.py:51-56,393-398): MLP4Confidence on the per-complex sum of node features
        self.confidence_training = getattr(args, "confidence_training", False)
        if self.confidence_training:
            if args.stack_mlp:
                self.ranking_mlp_pre = MLP4Confidence(args, embedding_channels=embedding_channels, n=args.confidence_mlp_hidden_scale,
                                                      out_channels=embedding_channels)
            self.ranking_score_mlp = MLP4Confidence(args, embedding_channels=embedding_channels, n=args.confidence_mlp_hidden_scale,
                                                    out_channels=1)
        if getattr(args, "use_clustering", False):
            from sklearn.cluster import DBSCAN       # host-side, exactly as the reference (model.py:57-61)
            self.dbscan_module = DBSCAN(eps=args.dbscan_eps, min_samples=args.dbscan_min_samples)
        self.precision = "fp32"     # also forwarded to the two stacks
        # train() under no_grad = the reference's sampling mode (P/test_sampling_fabind.py:118-124): dropout active in both stacks
        # and the three MLP heads (confidence modules stay in eval mode there); masks keyed by `dropout_seed` (None: drawn per call)
        self.dropout_seed = None
        self.dropout_colonly = False

    def _sampling_setup(self):
        """returns (p, seed, colonly) or None; seeds the two stacks"""
        if not self.training:
            return None
        if torch.is_grad_enabled():
            raise NotImplementedError("fabind_b200: backward kernels are not built; train() mode is served as the reference's "
                                      "dropout SAMPLING mode only - wrap the call in torch.no_grad()")
        seed = self.dropout_seed if self.dropout_seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        self.complex_model.dropout_seed = seed
        self.pocket_pred_model.dropout_seed = (seed + 0x51ED27) & 0xFFFFFFFF
        for m in (self.complex_model, self.pocket_pred_model):
            m.dropout_colonly = self.dropout_colonly
        return (float(self.args.dropout), seed, self.dropout_colonly)

    def _head_drop(self, name):
        d = getattr(self, "_drop", None)
        return None if d is None else (d[0], d[1], HEAD_SITES[name], d[2])

    def _lin(self, x, lin, act=0):
        return _gemm(x.contiguous(), lin.weight, lin.bias, act)

    # ---------------------------------------------------------------------------------------------------
    def _pocket_stage(self, data, gumbel=False):
        """model.py:72-139"""
        l = _lib.lib()
        dev = self.glb_c.device
        H = self.glb_c.shape[1]
        wp = data['complex_whole_protein']
        cb = data['compound'].batch.cpu().numpy()
        pbw = data['protein_whole'].batch.cpu().numpy()
        B = int(wp.batch[-1]) + 1
        nA, nL = np.bincount(cb, minlength=B), np.bincount(pbw, minlength=B)
        comp_off = np.concatenate([[0], np.cumsum(nA)]).astype(np.int32)
        prot_off = np.concatenate([[0], np.cumsum(nL)]).astype(np.int32)
        comp = self._lin(data['compound'].node_feats.to(dev, torch.float32), self.compound_linear_whole_protein)
        prot = self._lin(data['protein_whole'].node_feats.to(dev, torch.float32), self.protein_linear_whole_protein)
        kind, idx = complex_layout_np(nA, nL)
        Nw = len(kind)
        x = _assemble(Nw, H, kind, idx, [self.glb_c, comp, self.glb_p, prot], 1.0, dev)
        x = self._lin(x, self.embedding_shrink)
        rows = np.arange(Nw)
        X = _select(wp.node_coords.to(dev, torch.float32), rows, 1.0 / self.coordinate_scale).unsqueeze(-2)
        XL = _select(wp.node_coords_LAS.to(dev, torch.float32), rows, 1.0 / self.coordinate_scale).unsqueeze(-2)
        self.pocket_pred_model.precision = self.precision
        self.pocket_pred_model.return_pair = False
        _, Hout, _ = self.pocket_pred_model(
            X, x, batch_id=wp.batch, segment_id=wp.segment, mask=wp.mask, is_global=wp.is_global,
            compound_edge_index=data['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index.to(dev),
            LAS_edge_index=data['complex_whole_protein', 'LAS', 'complex_whole_protein'].edge_index.to(dev),
            batched_complex_coord_LAS=XL, LAS_mask=None)
        out = self._lin(Hout, self.embedding_enlarge)
        seg = wp.segment.cpu().numpy().astype(bool)
        glb = wp.is_global.cpu().numpy().astype(bool)
        comp_out = _select(out, np.nonzero(~seg & ~glb)[0])
        prot_out = _select(out, np.nonzero(seg & ~glb)[0])
        st = current_stream_ptr(dev)
        # pocket radius head on the per-complex sum of the ligand atom embeddings (model.py:110-114); relu applied in fb_pocket_mask_r
        co = _i32(comp_off, dev)
        comp_sum = torch.empty((B, H), dtype=torch.float32, device=dev)
        _lib.check(l.fb_segment_sum_rows(comp_out.data_ptr(), H, co.data_ptr(), B, comp_sum.data_ptr(), st), "fb_segment_sum_rows")
        dot, tiles, _ = _mlp_scalar(comp_sum, self.pocket_radius_head, drop=self._head_drop("pocket_radius_head"))
        radius_raw = _dot_finish(dot, tiles, B, self.pocket_radius_head.linear2.bias)
        # per-residue pocket logit (model.py:124-129)
        dot, tiles, M = _mlp_scalar(prot_out, self.protein_to_pocket, drop=self._head_drop("protein_to_pocket"))
        logit = _dot_finish(dot, tiles, M, self.protein_to_pocket.linear2.bias)
        xyz = data.node_xyz_whole.to(dev, torch.float32).contiguous()
        po = _i32(prot_off, dev)
        centers = torch.empty((B, 3), dtype=torch.float32, device=dev)
        if gumbel:
            # train()-mode forward: F.gumbel_softmax (model.py:136-137); the noise comes from torch's generator exactly as there
            # (gumbels = -empty.exponential_().log()), or from `self.gumbel_noise` ([n_res, 2], tests)
            noise = getattr(self, "gumbel_noise", None)
            if noise is None:
                noise = -torch.empty((logit.shape[0], 2), dtype=torch.float32, device=dev).exponential_().log()
            noise = noise.to(dev, torch.float32).contiguous()
            _lib.check(l.fb_pocket_center_gumbel(logit.data_ptr(), noise.data_ptr(), xyz.data_ptr(), po.data_ptr(), B,
                                                 float(self.args.gs_tau), int(bool(self.args.gs_hard)), centers.data_ptr(), st),
                       "fb_pocket_center_gumbel")
        else:
            _lib.check(l.fb_pocket_center(logit.data_ptr(), xyz.data_ptr(), po.data_ptr(), B, float(self.args.gs_tau),
                                          int(bool(self.args.gs_hard)), 0, centers.data_ptr(), st), "fb_pocket_center")
        return dict(B=B, dev=dev, H=H, cb=cb, pbw=pbw, nA=nA, nL=nL, comp_off=comp_off, prot_off=prot_off, comp_out=comp_out,
                    prot_out=prot_out, logit=logit, radius_raw=radius_raw, xyz_whole=xyz, prot_off_dev=po, co=co, centers=centers)

    def _cluster_centers(self, s):
        """model.py:147-167 (--use-clustering): host-side DBSCAN over the predicted pocket residues and python `random` draws,
        exactly as the reference does it (it moves the coordinates to numpy there too); the chosen centres go back to the device."""
        if not getattr(self.args, "use_clustering", False):
            return
        import random
        B, nL, prot_off = s["B"], s["nL"], s["prot_off"]
        Lmax = int(nL.max())
        logit = s["logit"].cpu()
        xyz = s["xyz_whole"].cpu()
        centers = s["centers"].cpu().clone()
        out = torch.zeros_like(centers)
        for i in range(B):
            dense_logit = torch.zeros(Lmax)
            dense_xyz = torch.zeros(Lmax, 3)
            dense_logit[:nL[i]] = logit[prot_off[i]:prot_off[i + 1]]
            dense_xyz[:nL[i]] = xyz[prot_off[i]:prot_off[i + 1]]
            prob = dense_logit.sigmoid()                       # padded positions: sigmoid(0) = 0.5, as in the reference
            pos = dense_xyz[prob > 0.5]
            if pos.shape[0] < 50:
                top = torch.argsort(prob)[-50:]
                sel = torch.full(prob.shape, False)
                sel[top] = True
                pos = dense_xyz[sel]
            pos = pos.numpy()
            # same labels as self.dbscan_module.fit(pos).labels_ (tests/test_dbscan.py) without sklearn's per-call overhead
            labels = dbscan_labels(pos, self.args.dbscan_eps, self.args.dbscan_min_samples)
            cid = random.randint(0, labels.max())
            if random.random() < self.args.choose_cluster_prob:
                out[i] = torch.tensor(pos[labels == cid].mean(axis=0))
            else:
                out[i] = centers[i]
        s["centers"] = out.to(s["dev"])

    def _mlp4conf(self, x, mlp, act_last=0):
        z = _layernorm(x, mlp.layernorm) if self.args.confidence_use_ln_mlp else x
        return _gemm(_gemm(z.contiguous(), mlp.linear1.weight, mlp.linear1.bias, act=2), mlp.linear2.weight, mlp.linear2.bias, act=act_last)

    def _confidence(self, s, Ho):
        """model.py:393-398: MLP4Confidence heads on scatter_add(complex_out, complex_batch); evaluated without dropout (the
        sampling scripts keep `ranking*` modules in eval mode, P/test_sampling_fabind.py:120-123)."""
        l = _lib.lib()
        dev, B, H = s["dev"], s["B"], s["H"]
        node_off = _i32(np.concatenate([[0], np.cumsum(s["nA"] + s["nP"] + 2)]), dev)
        hidden = torch.empty((B, H), dtype=torch.float32, device=dev)
        _lib.check(l.fb_segment_sum_rows(Ho.contiguous().data_ptr(), H, node_off.data_ptr(), B, hidden.data_ptr(),
                                         current_stream_ptr(dev)), "fb_segment_sum_rows")
        if self.args.stack_mlp:
            hidden = self._mlp4conf(hidden, self.ranking_mlp_pre, act_last=2)
        return self._mlp4conf(hidden, self.ranking_score_mlp).squeeze(-1)

    def _dock(self, s, data, want_pair):
        """model.py:202-341 / 505-626: crop by the predicted centre and radius, re-centre, re-assemble, run the docking stack."""
        l = _lib.lib()
        dev, B, H = s["dev"], s["B"], s["H"]
        a = self.args
        scale = self.coordinate_scale
        xyz, po = s["xyz_whole"], s["prot_off_dev"]
        st = current_stream_ptr(dev)
        keep = torch.empty(xyz.shape[0], dtype=torch.uint8, device=dev)
        less5 = torch.empty(B, dtype=torch.int32, device=dev)
        radius_pred = torch.empty((B, 1), dtype=torch.float32, device=dev)
        _lib.check(l.fb_pocket_mask_r(xyz.data_ptr(), po.data_ptr(), B, s["centers"].data_ptr(), s["radius_raw"].data_ptr(),
                                      float(a.pocket_radius_buffer), float(a.min_pocket_radius),
                                      float(a.pocket_radius) if a.force_fix_radius else -1.0, keep.data_ptr(), less5.data_ptr(),
                                      radius_pred.data_ptr(), st), "fb_pocket_mask_r")
        keep_h = keep.cpu().numpy().astype(bool)         # the one host read of this stage: sizes of the cropped graphs
        s["less5"], s["radius_pred"] = int(less5.sum().item()), radius_pred
        kept = np.nonzero(keep_h)[0]
        nP = np.add.reduceat(keep_h.astype(np.int64), s["prot_off"][:-1]) if B else np.zeros(0, np.int64)
        pocket_off = np.concatenate([[0], np.cumsum(nP)]).astype(np.int32)
        nA, comp_off, co = s["nA"], s["comp_off"], s["co"]
        pk = _i32(pocket_off, dev)
        pocket_raw = _select(xyz, kept)
        pocket_xyz = torch.empty_like(pocket_raw)                                   # re-centred on its own mean (model.py:255-256)
        bias = torch.empty((B, 3), dtype=torch.float32, device=dev)
        _lib.check(l.fb_center_rows3(pocket_raw.data_ptr(), pk.data_ptr(), B, pocket_xyz.data_ptr(), bias.data_ptr(), st), "fb_center_rows3")
        lig = data['compound'].node_coords.to(dev, torch.float32).contiguous()
        lig_init = torch.empty_like(lig)
        _lib.check(l.fb_ligand_place(lig.data_ptr(), co.data_ptr(), pocket_xyz.data_ptr(), pk.data_ptr(), B, lig_init.data_ptr(), st),
                   "fb_ligand_place")
        kind, idx_f = complex_layout_np(nA, nP, kept)             # features: residue rows come from the whole-protein table
        _, idx_x = complex_layout_np(nA, nP)                      # coordinates: from the cropped pocket table
        seg, msk, glb = kind >= 2, kind <= 2, (kind == 0) | (kind == 2)
        bat = np.repeat(np.arange(B, dtype=np.int64), nA + nP + 2)
        Ncx = len(kind)
        Hc = _assemble(Ncx, H, kind, idx_f, [self.glb_c, s["comp_out"], self.glb_p, s["prot_out"]], 1.0, dev)
        X = _assemble(Ncx, 3, kind, idx_x, [None, lig_init, None, pocket_xyz], 1.0 / scale, dev).unsqueeze(-2)
        XL = _assemble(Ncx, 3, kind, idx_x, [None, data['compound'].rdkit_coords.to(dev, torch.float32), None, None], 1.0 / scale,
                       dev).unsqueeze(-2)
        node_off = np.concatenate([[0], np.cumsum(nA + nP + 2)])
        ael, lel = data['compound_atom_edge_list'], data['LAS_edge_list']
        c2c = (ael.x.cpu().numpy() + node_off[ael.batch.cpu().numpy()][:, None]).T
        las = (lel.x.cpu().numpy() + node_off[lel.batch.cpu().numpy()][:, None]).T
        self.complex_model.precision = self.precision
        self.complex_model.return_pair = want_pair
        Xo, Ho, pair = self.complex_model(
            X.contiguous(), Hc, batch_id=torch.from_numpy(bat), segment_id=torch.from_numpy(seg), mask=torch.from_numpy(msk),
            is_global=torch.from_numpy(glb), compound_edge_index=torch.from_numpy(np.ascontiguousarray(c2c)).to(dev),
            LAS_edge_index=torch.from_numpy(np.ascontiguousarray(las)).to(dev), batched_complex_coord_LAS=XL.contiguous(),
            LAS_mask=None)
        s.update(nP=nP, pocket_off=pocket_off, pocket_xyz=pocket_xyz, seg=seg, glb=glb, pk=pk, lig=lig, bias=bias)
        return Xo, Ho, pair

    def _dock_stage1(self, s, data, want_pair):
        """model.py:169-197 (stage == 1): the dataloader's pocket (`data['pocket'].keepNode`) and pre-built complex graph; the ligand
        is re-centred on its own mean, the pocket moved by `data.pocket_residue_center`; `data['complex'].node_coords` and
        `data.coords` are updated like the reference does in place."""
        l = _lib.lib()
        dev, B, H, scale = s["dev"], s["B"], s["H"], self.coordinate_scale
        st = current_stream_ptr(dev)
        cx = data['complex']
        kept = np.nonzero(data['pocket'].keepNode.cpu().numpy().astype(bool))[0]
        nP = np.bincount(data['pocket'].batch.cpu().numpy(), minlength=B)
        pocket_off = np.concatenate([[0], np.cumsum(nP)]).astype(np.int32)
        nA, co = s["nA"], s["co"]
        pk = _i32(pocket_off, dev)
        kind, idx_f = complex_layout_np(nA, nP, kept)
        _, idx_x = complex_layout_np(nA, nP)
        Ncx = len(kind)
        Hc = _assemble(Ncx, H, kind, idx_f, [self.glb_c, s["comp_out"], self.glb_p, s["prot_out"]], 1.0, dev)
        cxc = cx.node_coords.to(dev, torch.float32).contiguous()
        lig_raw = _select(cxc, np.nonzero(kind == 1)[0])
        pocket_raw = _select(cxc, np.nonzero(kind == 3)[0])
        lig_c = torch.empty_like(lig_raw)
        lig_mean = torch.empty((B, 3), dtype=torch.float32, device=dev)
        _lib.check(l.fb_center_rows3(lig_raw.data_ptr(), co.data_ptr(), B, lig_c.data_ptr(), lig_mean.data_ptr(), st), "fb_center_rows3")
        prc = data.pocket_residue_center.to(dev, torch.float32).contiguous()
        pocket_s = torch.empty_like(pocket_raw)
        _lib.check(l.fb_shift_rows3(pocket_raw.data_ptr(), pk.data_ptr(), B, pocket_raw.shape[0], prc.data_ptr(), -1.0,
                                    pocket_s.data_ptr(), st), "fb_shift_rows3")
        Xun = _assemble(Ncx, 3, kind, idx_x, [None, lig_c, None, pocket_s], 1.0, dev)
        cx.node_coords = Xun.to(cx.node_coords.device)                                # in-place effect of model.py:179-182
        if getattr(data, "coords", None) is not None:                                 # model.py:184
            gt = data.coords.to(dev, torch.float32).contiguous()
            gt_out = torch.empty_like(gt)
            _lib.check(l.fb_shift_rows3(gt.data_ptr(), co.data_ptr(), B, gt.shape[0], prc.data_ptr(), -1.0, gt_out.data_ptr(), st),
                       "fb_shift_rows3")
            data.coords = gt_out.to(data.coords.device)
        X = _select(Xun, np.arange(Ncx), 1.0 / scale).unsqueeze(-2)
        XL = _select(cx.node_coords_LAS.to(dev, torch.float32), np.arange(Ncx), 1.0 / scale).unsqueeze(-2)
        seg, glb = kind >= 2, (kind == 0) | (kind == 2)
        self.complex_model.precision = self.precision
        self.complex_model.return_pair = want_pair
        Xo, Ho, pair = self.complex_model(
            X.contiguous(), Hc, batch_id=cx.batch, segment_id=cx.segment, mask=cx.mask, is_global=cx.is_global,
            compound_edge_index=data['complex', 'c2c', 'complex'].edge_index.to(dev),
            LAS_edge_index=data['complex', 'LAS', 'complex'].edge_index.to(dev), batched_complex_coord_LAS=XL.contiguous(), LAS_mask=None)
        bias = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        # pocket_radius_pred = relu(head) is still returned in stage 1 (model.py:114,399): the crop kernel delivers it (its mask is unused)
        xyz, a = s["xyz_whole"], self.args
        keep = torch.empty(xyz.shape[0], dtype=torch.uint8, device=dev)
        less5 = torch.empty(B, dtype=torch.int32, device=dev)
        radius_pred = torch.empty((B, 1), dtype=torch.float32, device=dev)
        _lib.check(l.fb_pocket_mask_r(xyz.data_ptr(), s["prot_off_dev"].data_ptr(), B, s["centers"].data_ptr(), s["radius_raw"].data_ptr(),
                                      float(a.pocket_radius_buffer), float(a.min_pocket_radius), -1.0, keep.data_ptr(), less5.data_ptr(),
                                      radius_pred.data_ptr(), st), "fb_pocket_mask_r")
        s.update(nP=nP, pocket_off=pocket_off, pocket_xyz=data.node_xyz.to(dev, torch.float32).contiguous(), seg=seg, glb=glb, pk=pk,
                 lig=lig_raw, bias=bias, less5=0, radius_pred=radius_pred)
        return Xo, Ho, pair

    # ---------------------------------------------------------------------------------------------------
    def forward(self, data, stage=2, train=False):
        """eval semantics of model.py:63-401 with stage=2, train=False (the predicted pocket is used for docking)."""
        if train:
            raise NotImplementedError("fabind_b200: the training path (teacher forcing + backward) is not built; pass train=False")
        self._drop = self._sampling_setup()
        if stage not in (1, 2):
            raise ValueError("stage must be 1 or 2")
        l = _lib.lib()
        a = self.args
        with torch.no_grad():
            s = self._pocket_stage(data, gumbel=self.pocket_pred_model.training)
            dev, B, H, scale = s["dev"], s["B"], s["H"], self.coordinate_scale
            self._cluster_centers(s)
            if stage == 1:
                Xo, Ho, pair = self._dock_stage1(s, data, want_pair=not self.confidence_training)
            else:
                Xo, Ho, pair = self._dock(s, data, want_pair=not self.confidence_training)
            if self.confidence_training:
                return self._forward_confidence(s, data, Xo, Ho, shift_coords=stage == 2)
            st = current_stream_ptr(dev)
            seg, glb = s["seg"], s["glb"]
            c_rows = np.nonzero(~seg & ~glb)[0]
            lig_n = _select(Xo.view(-1, 3), c_rows)
            pocket_n = _select(s["pocket_xyz"], np.arange(s["pocket_xyz"].shape[0]), 1.0 / scale)
            nP, nA = s["nP"], s["nA"]
            nQ = nP * nA
            pair_off = np.concatenate([[0], np.cumsum(nQ)]).astype(np.int32)
            Q = int(pair_off[-1])
            qo = _i32(pair_off, dev)
            # distance head on pair[:, 1:, 1:] (model.py:379-387): row list into the dense [B, max_p, max_c, H] block
            _, max_p, max_c, _ = pair.shape
            rows = np.concatenate([((b * max_p + 1 + np.arange(nP[b]))[:, None] * max_c + 1 + np.arange(nA[b])[None, :]).reshape(-1)
                                   for b in range(B)]) if Q else np.zeros(0, np.int64)
            bf = self.precision == "bf16"
            dot, tiles, _ = _mlp_scalar(pair.view(-1, H), self.distmap_mlp, rows=_i32(rows, dev), bf16=bf, drop=self._head_drop("distmap_mlp"))
            y_pred = torch.empty(Q, dtype=torch.float32, device=dev)
            y_coords = torch.empty(Q, dtype=torch.float32, device=dev)
            _lib.check(l.fb_head_finish_cap(dot.data_ptr(), tiles, dot.shape[1], self.distmap_mlp.linear2.bias.data_ptr(),
                                            pocket_n.data_ptr(), lig_n.data_ptr(), s["pk"].data_ptr(), s["co"].data_ptr(), qo.data_ptr(),
                                            B, Q, float(scale), float(a.dis_map_thres), y_pred.data_ptr(), y_coords.data_ptr(), st),
                       "fb_head_finish_cap")
            if stage == 1:
                compound_coords_out = _select(Xo.view(-1, 3), c_rows, scale)
                cls_dense, pmask, kind, idx, Lmax = self._dense_cls(s)
                coords_dense = _assemble(B * Lmax, 3, kind, idx, [None, s["xyz_whole"]], 1.0, dev).view(B, Lmax, 3)
                pocket_cls = torch.zeros((B, Lmax), dtype=data.pocket_idx.dtype, device=dev)
                pocket_cls[pmask] = data.pocket_idx.to(dev)
                return (compound_coords_out, data['compound'].batch, y_pred, y_coords, cls_dense, pocket_cls, pmask, coords_dense,
                        s["centers"], data.dis_map.to(dev), 0, s["radius_pred"], s["bias"])
            # label-side bookkeeping the reference does inside forward: dis_map against the shifted ligand, data.coords -= centre
            lig_shift = torch.empty_like(s["lig"])
            _lib.check(l.fb_shift_rows3(s["lig"].data_ptr(), s["co"].data_ptr(), B, s["lig"].shape[0], s["bias"].data_ptr(), -1.0,
                                        lig_shift.data_ptr(), st), "fb_shift_rows3")
            dis_map = torch.empty(Q, dtype=torch.float32, device=dev)
            _lib.check(l.fb_pair_dist(s["pocket_xyz"].data_ptr(), lig_shift.data_ptr(), s["pk"].data_ptr(), s["co"].data_ptr(),
                                      qo.data_ptr(), B, Q, float(a.dis_map_thres), dis_map.data_ptr(), st), "fb_pair_dist")
            if getattr(data, "coords", None) is not None:
                gt = data.coords.to(dev, torch.float32).contiguous()
                gt_out = torch.empty_like(gt)
                _lib.check(l.fb_shift_rows3(gt.data_ptr(), s["co"].data_ptr(), B, gt.shape[0], s["bias"].data_ptr(), -1.0,
                                            gt_out.data_ptr(), st), "fb_shift_rows3")
                data.coords = gt_out.to(data.coords.device)
            compound_coords_out = _select(Xo.view(-1, 3), c_rows, scale)
            nL, prot_off = s["nL"], s["prot_off"]
            Lmax = int(nL.max())
            kind = np.zeros(B * Lmax, np.uint8); idx = np.zeros(B * Lmax, np.int32)
            mask_h = np.zeros((B, Lmax), bool)
            for b in range(B):
                kind[b * Lmax:b * Lmax + nL[b]] = 1
                idx[b * Lmax:b * Lmax + nL[b]] = np.arange(prot_off[b], prot_off[b + 1])
                mask_h[b, :nL[b]] = True
            cls_dense = _assemble(B * Lmax, 1, kind, idx, [None, s["logit"].view(-1, 1)], 1.0, dev).view(B, Lmax)
            coords_dense = _assemble(B * Lmax, 3, kind, idx, [None, s["xyz_whole"]], 1.0, dev).view(B, Lmax, 3)
            pocket_cls = torch.zeros((B, Lmax), dtype=data.pocket_idx.dtype, device=dev)
            pmask = torch.from_numpy(mask_h).to(dev)
            pocket_cls[pmask] = data.pocket_idx.to(dev)
            return (compound_coords_out, data['compound'].batch, y_pred, y_coords, cls_dense, pocket_cls, pmask, coords_dense,
                    s["centers"], dis_map, s["less5"], s["radius_pred"], s["bias"])

    def _dense_cls(self, s):
        B, nL, prot_off, dev = s["B"], s["nL"], s["prot_off"], s["dev"]
        Lmax = int(nL.max())
        kind = np.zeros(B * Lmax, np.uint8); idx = np.zeros(B * Lmax, np.int32)
        mask_h = np.zeros((B, Lmax), bool)
        for b in range(B):
            kind[b * Lmax:b * Lmax + nL[b]] = 1
            idx[b * Lmax:b * Lmax + nL[b]] = np.arange(prot_off[b], prot_off[b + 1])
            mask_h[b, :nL[b]] = True
        cls_dense = _assemble(B * Lmax, 1, kind, idx, [None, s["logit"].view(-1, 1)], 1.0, dev).view(B, Lmax)
        return cls_dense, torch.from_numpy(mask_h).to(dev), kind, idx, Lmax

    def _forward_confidence(self, s, data, Xo, Ho, shift_coords=True):
        """return tuple of model.py:399 (confidence_training)"""
        l = _lib.lib()
        dev, B = s["dev"], s["B"]
        c_rows = np.nonzero(~s["seg"] & ~s["glb"])[0]
        if shift_coords and getattr(data, "coords", None) is not None:        # data.coords -= pocket centre (model.py:257)
            gt = data.coords.to(dev, torch.float32).contiguous()
            gt_out = torch.empty_like(gt)
            _lib.check(l.fb_shift_rows3(gt.data_ptr(), s["co"].data_ptr(), B, gt.shape[0], s["bias"].data_ptr(), -1.0, gt_out.data_ptr(),
                                        current_stream_ptr(dev)), "fb_shift_rows3")
            data.coords = gt_out.to(data.coords.device)
        cls_dense, pmask, _, _, _ = self._dense_cls(s)
        return (_select(Xo.view(-1, 3), c_rows, self.coordinate_scale), data['compound'].batch, cls_dense, pmask, s["less5"],
                self._confidence(s, Ho), s["bias"])

    def sample(self, data_fn, n_samples, seed=0):
        """Sampling-based FABind+ (P/test_sampling_fabind.py:126-131, P/inference_sampling_fabind.py:168-185): `n_samples` passes of
        `inference` in train() mode (dropout masks keyed by seed + k), returning the list of (coords, batch[, confidence]).
        `data_fn()` must return a fresh batch per pass (inference mutates nothing, the forward path shifts data.coords)."""
        flags = {m: m.training for m in self.modules()}   # per-module: sub-modules the caller keeps in eval() stay there
        self.train()
        for name, sub in self.named_modules():
            if name.startswith("confidence") or name.startswith("ranking"):
                sub.eval()
        outs = []
        try:
            with torch.no_grad():
                for k in range(n_samples):
                    self.dropout_seed = (int(seed) + 7919 * k) & 0x7FFFFFFF
                    outs.append(self.inference(data_fn()))
        finally:
            self.dropout_seed = None
            for m, f in flags.items():
                m.training = f
        return outs

    def inference(self, data):
        self._drop = self._sampling_setup()
        l = _lib.lib()
        with torch.no_grad():
            s = self._pocket_stage(data)
            self._cluster_centers(s)
            Xo, Ho, _ = self._dock(s, data, want_pair=False)
            dev = s["dev"]
            c_rows = np.nonzero(~s["seg"] & ~s["glb"])[0]
            pred = _select(Xo.view(-1, 3), c_rows, self.coordinate_scale)
            out = torch.empty_like(pred)      # move back to whole-protein coordinates (model.py:684)
            _lib.check(l.fb_shift_rows3(pred.data_ptr(), s["co"].data_ptr(), s["B"], pred.shape[0], s["bias"].data_ptr(), 1.0,
                                        out.data_ptr(), current_stream_ptr(dev)), "fb_shift_rows3")
            if self.confidence_training:
                return out, data['compound'].batch, self._confidence(s, Ho)
            return out, data['compound'].batch


def get_model(args, logger):
    """model.py:700-703"""
    logger.log_message("FABind plus")
    return FABindPlus(args, args.hidden_size, args.pocket_pred_hidden_size)
