"""Drop-in for FABind/fabind/models/model.py: the L2 wrapper around the docking stack.

`IaBNet_mean_and_pocket_prediction_cls_coords_dependent` keeps the reference's constructor signature, parameter
names/shapes (reference checkpoints load with strict=True) and the return tuples of `forward(data, stage=2)`
(eval semantics, model.py:82-369) and `inference(data)` (model.py:371-580).  All arithmetic runs in
libfabind_b200 kernels; the host side only does index bookkeeping (which residue/atom goes to which row),
which the reference does with per-sample python loops and torch.cat.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .att_model import EfficientMCAttModel
from .runtime import current_stream_ptr


class Transition_diff_out_dim(nn.Module):
    """model.py:11-24 (weights only; evaluated by fb_layernorm + fb_gemm with a fused Linear(4H,1) row-dot)."""

    def __init__(self, embedding_channels=256, out_channels=256, n=4):
        super().__init__()
        self.layernorm = nn.LayerNorm(embedding_channels)
        self.linear1 = nn.Linear(embedding_channels, n * embedding_channels)
        self.linear2 = nn.Linear(n * embedding_channels, out_channels)
        nn.init.xavier_uniform_(self.linear1.weight, gain=0.001)
        nn.init.xavier_uniform_(self.linear2.weight, gain=0.001)


# ---------------------------------------------------------------------------------------------------------
# thin op wrappers (device tensors in, device tensors out; everything on the current stream)
# ---------------------------------------------------------------------------------------------------------
def _p(t):
    return t.data_ptr() if t is not None else None


def _i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def _gemm(A, W, bias=None, act=0, dotv=None, bf16=False, drop=None):
    """act(A W^T + bias)  (returns [M,N]);  with dotv: returns (dot partials [tiles, M], tiles) and stores nothing.
    drop = (p, seed, site, colonly): dropout after the activation (FABind+ sampling mode)."""
    l = _lib.lib()
    dev = A.device
    M, K = A.shape
    N = W.shape[0]
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K
    g.W, g.bias, g.act = W.data_ptr(), _p(bias), act
    g.M, g.N, g.bf16_mode, g.force_simt = M, N, int(bf16), 0
    if drop is not None and drop[0] > 0:
        g.drop_p, g.drop_seed, g.drop_site, g.drop_row0, g.drop_colonly = float(drop[0]), int(drop[1]) & 0xFFFFFFFF, int(drop[2]), 0, int(bool(drop[3]))
    st = current_stream_ptr(dev)
    if dotv is not None:
        tiles = l.fb_gemm_dot_tiles(M, N, K, int(bf16), 0)
        dot = torch.empty((tiles, max(M, 1)), dtype=torch.float32, device=dev)
        g.dotv, g.dot_out, g.dot_stride = dotv.data_ptr(), dot.data_ptr(), max(M, 1)
        _lib.check(l.fb_gemm(C.byref(g), st), "fb_gemm")
        return dot, tiles
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    g.C, g.ldc = out.data_ptr(), N
    _lib.check(l.fb_gemm(C.byref(g), st), "fb_gemm")
    return out


def _assemble(M, D, kind, idx, srcs, scale, dev):
    l = _lib.lib()
    out = torch.empty((M, D), dtype=torch.float32, device=dev)
    k = torch.from_numpy(np.ascontiguousarray(kind, dtype=np.uint8)).to(dev)
    ix = _i32(idx, dev)
    srcs = [s.contiguous() if s is not None else None for s in srcs] + [None] * (4 - len(srcs))
    _lib.check(l.fb_assemble_rows(out.data_ptr(), M, D, k.data_ptr(), ix.data_ptr(), _p(srcs[0]), _p(srcs[1]), _p(srcs[2]),
                                  _p(srcs[3]), float(scale), current_stream_ptr(dev)), "fb_assemble_rows")
    for t in [k, ix] + [s for s in srcs if s is not None]:
        t.record_stream(torch.cuda.current_stream(dev))
    return out


def complex_layout_np(nA, nX, src_x=None):
    """Row plan of the reference's per-complex [glb_c | atoms | glb_p | residues] assembly (model.py:104-115,205-253) for all
    complexes at once: kind (0 glb_c, 1 atom, 2 glb_p, 3 residue), the source row of every atom / residue row (atoms and residues
    are consumed in order; `src_x` gives the residue source rows, default 0..sum(nX)-1), and the node attributes derived from it."""
    nA, nX = np.asarray(nA, dtype=np.int64), np.asarray(nX, dtype=np.int64)
    B = len(nA)
    lens = np.stack([np.ones(B, np.int64), nA, np.ones(B, np.int64), nX], 1).reshape(-1)
    kind = np.repeat(np.tile(np.arange(4, dtype=np.uint8), B), lens)
    idx = np.zeros(kind.shape[0], dtype=np.int32)
    idx[kind == 1] = np.arange(int(nA.sum()), dtype=np.int32)
    idx[kind == 3] = np.arange(int(nX.sum()), dtype=np.int32) if src_x is None else np.asarray(src_x, dtype=np.int32)
    return kind, idx


def _select(src, idx, scale=1.0):
    """rows src[idx] (the reference's boolean-mask selections), optionally scaled"""
    return _assemble(len(idx), src.shape[1], np.ones(len(idx), np.uint8), idx, [None, src], scale, src.device)


def _layernorm(x, ln):
    l = _lib.lib()
    out = torch.empty_like(x)
    _lib.check(l.fb_layernorm(x.data_ptr(), x.shape[0], x.shape[1], ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps),
                              out.data_ptr(), current_stream_ptr(x.device)), "fb_layernorm")
    return out


class IaBNet_mean_and_pocket_prediction_cls_coords_dependent(nn.Module):
    def __init__(self, args, embedding_channels=128, pocket_pred_embedding_channels=128):
        super().__init__()
        self.layernorm = nn.LayerNorm(embedding_channels)
        self.args = args
        self.coordinate_scale = args.coordinate_scale
        self.normalize_coord = lambda x: x / self.coordinate_scale
        self.unnormalize_coord = lambda x: x * self.coordinate_scale
        self.stage_prob = args.stage_prob
        n_channel = 1
        self.complex_model = EfficientMCAttModel(
            args, embedding_channels, embedding_channels, n_channel, n_edge_feats=0, n_layers=args.mean_layers,
            n_iter=args.n_iter, inter_cutoff=args.inter_cutoff, intra_cutoff=args.intra_cutoff,
            normalize_coord=self.normalize_coord, unnormalize_coord=self.unnormalize_coord)
        self.pocket_pred_model = EfficientMCAttModel(
            args, pocket_pred_embedding_channels, pocket_pred_embedding_channels, n_channel, n_edge_feats=0,
            n_layers=args.pocket_pred_layers, n_iter=args.pocket_pred_n_iter, inter_cutoff=args.inter_cutoff,
            intra_cutoff=args.intra_cutoff, normalize_coord=self.normalize_coord, unnormalize_coord=self.unnormalize_coord)
        self.protein_to_pocket = Transition_diff_out_dim(embedding_channels=embedding_channels, n=4, out_channels=1)
        self.glb_c = nn.Parameter(torch.ones(1, embedding_channels))
        self.glb_p = nn.Parameter(torch.ones(1, embedding_channels))
        protein_hidden = 1280 if args.use_esm2_feat else 15
        if args.esm2_concat_raw:
            protein_hidden = 1295
        self.protein_linear_whole_protein = nn.Linear(protein_hidden, embedding_channels)
        self.compound_linear_whole_protein = nn.Linear(56, embedding_channels)
        self.embedding_shrink = nn.Linear(embedding_channels, pocket_pred_embedding_channels)
        self.embedding_enlarge = nn.Linear(pocket_pred_embedding_channels, embedding_channels)
        self.distmap_mlp = nn.Sequential(nn.Linear(embedding_channels, embedding_channels), nn.ReLU(),
                                         nn.Linear(embedding_channels, 1))
        for lin in (self.protein_linear_whole_protein, self.compound_linear_whole_protein, self.embedding_shrink,
                    self.embedding_enlarge, self.distmap_mlp[0], self.distmap_mlp[2]):
            nn.init.xavier_uniform_(lin.weight, gain=0.001)
        self.precision = "fp32"     # also forwarded to the two stacks

    # ---------------------------------------------------------------------------------------------------
    def _lin(self, x, lin, act=0):
        return _gemm(x.contiguous(), lin.weight, lin.bias, act)

    def _pocket_stage(self, data):
        """model.py:98-144.  Returns bookkeeping + per-residue pocket logits (flat, protein order)."""
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
        _, Hout = self.pocket_pred_model(
            X, x, batch_id=wp.batch, segment_id=wp.segment, mask=wp.mask, is_global=wp.is_global,
            compound_edge_index=data['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index.to(dev),
            LAS_edge_index=data['complex_whole_protein', 'LAS', 'complex_whole_protein'].edge_index.to(dev),
            batched_complex_coord_LAS=XL, LAS_mask=None)
        out = self._lin(Hout, self.embedding_enlarge)
        seg = wp.segment.cpu().numpy().astype(bool)
        glb = wp.is_global.cpu().numpy().astype(bool)
        comp_out = _select(out, np.nonzero(~seg & ~glb)[0])
        prot_out = _select(out, np.nonzero(seg & ~glb)[0])
        # protein_to_pocket: LayerNorm -> Linear(H,4H) ReLU -> Linear(4H,1) fused as a row-dot epilogue
        z = _layernorm(prot_out, self.protein_to_pocket.layernorm)
        dot, tiles = _gemm(z, self.protein_to_pocket.linear1.weight, self.protein_to_pocket.linear1.bias, act=2,
                           dotv=self.protein_to_pocket.linear2.weight[0].contiguous())
        logit = torch.empty(prot_out.shape[0], dtype=torch.float32, device=dev)
        _lib.check(l.fb_dot_finish(dot.data_ptr(), tiles, dot.shape[1], prot_out.shape[0],
                                   self.protein_to_pocket.linear2.bias.data_ptr(), logit.data_ptr(), current_stream_ptr(dev)),
                   "fb_dot_finish")
        return dict(B=B, dev=dev, H=H, cb=cb, pbw=pbw, nA=nA, nL=nL, comp_off=comp_off, prot_off=prot_off,
                    comp_out=comp_out, prot_out=prot_out, logit=logit)

    def _centers(self, s, data, mode):
        l = _lib.lib()
        dev = s["dev"]
        xyz = data.node_xyz_whole.to(dev, torch.float32).contiguous()
        centers = torch.empty((s["B"], 3), dtype=torch.float32, device=dev)
        po = _i32(s["prot_off"], dev)
        _lib.check(l.fb_pocket_center(s["logit"].data_ptr(), xyz.data_ptr(), po.data_ptr(), s["B"], float(self.args.gs_tau),
                                      int(bool(self.args.gs_hard)), mode, centers.data_ptr(), current_stream_ptr(dev)),
                   "fb_pocket_center")
        s["xyz_whole"], s["prot_off_dev"] = xyz, po
        return centers

    def _dock(self, s, data, centers):
        """model.py:173-334 (stage 2) / :439-574 (inference): crop by the predicted centre, re-assemble, run the docking stack."""
        l = _lib.lib()
        dev, B, H = s["dev"], s["B"], s["H"]
        scale = self.coordinate_scale
        xyz, po = s["xyz_whole"], s["prot_off_dev"]
        keep = torch.empty(xyz.shape[0], dtype=torch.uint8, device=dev)
        less5 = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(l.fb_pocket_mask(xyz.data_ptr(), po.data_ptr(), B, centers.data_ptr(), float(self.args.pocket_radius),
                                    keep.data_ptr(), less5.data_ptr(), current_stream_ptr(dev)), "fb_pocket_mask")
        keep_h = keep.cpu().numpy().astype(bool)         # the one host read of this stage: sizes of the cropped graphs
        s["keep"], s["less5"] = keep_h, int(less5.sum().item())
        kept = np.nonzero(keep_h)[0]
        nP = np.add.reduceat(keep_h.astype(np.int64), s["prot_off"][:-1]) if B else np.zeros(0, np.int64)
        pocket_off = np.concatenate([[0], np.cumsum(nP)]).astype(np.int32)
        nA, comp_off = s["nA"], s["comp_off"]
        pocket_xyz = _select(xyz, kept)                                             # [Pk,3], original units
        lig = data['compound'].node_coords.to(dev, torch.float32).contiguous()
        lig_init = torch.empty_like(lig)
        co, pk = _i32(comp_off, dev), _i32(pocket_off, dev)
        _lib.check(l.fb_ligand_place(lig.data_ptr(), co.data_ptr(), pocket_xyz.data_ptr(), pk.data_ptr(), B, lig_init.data_ptr(),
                                     current_stream_ptr(dev)), "fb_ligand_place")
        kind, idx_f = complex_layout_np(nA, nP, kept)             # features: residue rows come from the whole-protein table
        _, idx_x = complex_layout_np(nA, nP)                      # coordinates: from the cropped pocket table
        seg, msk, glb = kind >= 2, kind <= 2, (kind == 0) | (kind == 2)
        bat = np.repeat(np.arange(B, dtype=np.int64), nA + nP + 2)
        Ncx = len(kind)
        Hc = _assemble(Ncx, H, kind, idx_f, [self.glb_c, s["comp_out"], self.glb_p, s["prot_out"]], 1.0, dev)
        X = _assemble(Ncx, 3, kind, idx_x, [None, lig_init, None, pocket_xyz], 1.0 / scale, dev).unsqueeze(-2)
        if self.args.compound_coords_init_mode in ('redocking', 'redocking_no_rotate'):
            las_src = data['compound'].node_coords.to(dev, torch.float32)
        else:
            las_src = data['compound'].rdkit_coords.to(dev, torch.float32)
        XL = _assemble(Ncx, 3, kind, idx_x, [None, las_src, None, None], 1.0 / scale, dev).unsqueeze(-2)
        node_off = np.concatenate([[0], np.cumsum(nA + nP + 2)])
        ael, lel = data['compound_atom_edge_list'], data['LAS_edge_list']
        c2c = (ael.x.cpu().numpy() + node_off[ael.batch.cpu().numpy()][:, None]).T
        las = (lel.x.cpu().numpy() + node_off[lel.batch.cpu().numpy()][:, None]).T
        self.complex_model.precision = self.precision
        Xo, Ho = self.complex_model(
            X.contiguous(), Hc, batch_id=torch.from_numpy(bat), segment_id=torch.from_numpy(seg), mask=torch.from_numpy(msk),
            is_global=torch.from_numpy(glb), compound_edge_index=torch.from_numpy(np.ascontiguousarray(c2c)).to(dev),
            LAS_edge_index=torch.from_numpy(np.ascontiguousarray(las)).to(dev), batched_complex_coord_LAS=XL.contiguous(),
            LAS_mask=None)
        s.update(nP=nP, pocket_off=pocket_off, pocket_xyz=pocket_xyz, seg=seg, glb=glb, co=co, pk=pk, lig=lig)
        return Xo, Ho

    def _dock_stage1(self, s, data):
        """model.py:302-334 (final_stage == 1): the dataloader's ground-truth pocket crop and pre-built complex graph."""
        dev, B, H, scale = s["dev"], s["B"], s["H"], self.coordinate_scale
        cx = data['complex']
        kept = np.nonzero(data['pocket'].keepNode.cpu().numpy().astype(bool))[0]
        pb = data['pocket'].batch.cpu().numpy()
        nP = np.bincount(pb, minlength=B)
        pocket_off = np.concatenate([[0], np.cumsum(nP)]).astype(np.int32)
        nA, comp_off = s["nA"], s["comp_off"]
        kind, idx_f = complex_layout_np(nA, nP, kept)
        Ncx = len(kind)
        Hc = _assemble(Ncx, H, kind, idx_f, [self.glb_c, s["comp_out"], self.glb_p, s["prot_out"]], 1.0, dev)
        rows = np.arange(Ncx)
        X = _select(cx.node_coords.to(dev, torch.float32), rows, 1.0 / scale).unsqueeze(-2)
        XL = _select(cx.node_coords_LAS.to(dev, torch.float32), rows, 1.0 / scale).unsqueeze(-2)
        seg = cx.segment.cpu().numpy().astype(bool)
        glb = cx.is_global.cpu().numpy().astype(bool)
        self.complex_model.precision = self.precision
        Xo, Ho = self.complex_model(
            X.contiguous(), Hc, batch_id=cx.batch, segment_id=cx.segment, mask=cx.mask, is_global=cx.is_global,
            compound_edge_index=data['complex', 'c2c', 'complex'].edge_index.to(dev),
            LAS_edge_index=data['complex', 'LAS', 'complex'].edge_index.to(dev), batched_complex_coord_LAS=XL.contiguous(),
            LAS_mask=None)
        s.update(nP=nP, pocket_off=pocket_off, pocket_xyz=data.node_xyz.to(dev, torch.float32).contiguous(), seg=seg, glb=glb,
                 co=_i32(comp_off, dev), pk=_i32(pocket_off, dev), lig=data['compound'].node_coords.to(dev, torch.float32).contiguous(),
                 less5=0)
        return Xo, Ho

    # ---------------------------------------------------------------------------------------------------
    def forward(self, data, stage=1, train=False):
        """eval semantics of model.py:82-369: `final_stage = stage` (:170-171); stage 1 = dataloader pocket, 2 = predicted."""
        if self.training:
            raise NotImplementedError("fabind_b200: the training path is not built yet; call .eval()")
        if stage not in (1, 2):
            raise ValueError("stage must be 1 or 2")
        l = _lib.lib()
        with torch.no_grad():
            s = self._pocket_stage(data)
            dev, B, H, scale = s["dev"], s["B"], s["H"], self.coordinate_scale
            centers = self._centers(s, data, mode=0)
            if stage == 2:
                Xo, Ho = self._dock(s, data, centers)
            else:
                Xo, Ho = self._dock_stage1(s, data)
            seg, glb = s["seg"], s["glb"]
            p_rows, c_rows = np.nonzero(seg & ~glb)[0], np.nonzero(~seg & ~glb)[0]
            pocket_out = _layernorm(_select(Ho, p_rows), self.layernorm)
            comp_out = _layernorm(_select(Ho, c_rows), self.layernorm)
            lig_n = _select(Xo.view(-1, 3), c_rows)                      # normalised predicted ligand coordinates
            pocket_n = _select(s["pocket_xyz"], np.arange(s["pocket_xyz"].shape[0]), 1.0 / scale)
            nQ = s["nP"] * s["nA"]
            pair_off = np.concatenate([[0], np.cumsum(nQ)]).astype(np.int32)
            Q = int(pair_off[-1])
            qo = _i32(pair_off, dev)
            bf = self.precision == "bf16"
            Z = torch.empty((max(Q, 1), H), dtype=torch.bfloat16 if bf else torch.float32, device=dev)
            st = current_stream_ptr(dev)
            _lib.check(l.fb_head_outer(pocket_out.data_ptr(), comp_out.data_ptr(), s["pk"].data_ptr(), s["co"].data_ptr(),
                                       qo.data_ptr(), B, Q, H, Z.data_ptr(), int(bf), st), "fb_head_outer")
            W0 = self.distmap_mlp[0].weight
            W0 = W0.to(torch.bfloat16) if bf else W0
            dot, tiles = _gemm(Z[:Q], W0, self.distmap_mlp[0].bias, act=2, dotv=self.distmap_mlp[2].weight[0].contiguous(), bf16=bf)
            y_pred = torch.empty(Q, dtype=torch.float32, device=dev)
            y_coords = torch.empty(Q, dtype=torch.float32, device=dev)
            _lib.check(l.fb_head_finish(dot.data_ptr(), tiles, dot.shape[1], self.distmap_mlp[2].bias.data_ptr(),
                                        pocket_n.data_ptr(), lig_n.data_ptr(), s["pk"].data_ptr(), s["co"].data_ptr(),
                                        qo.data_ptr(), B, Q, float(scale), y_pred.data_ptr(), y_coords.data_ptr(), st),
                       "fb_head_finish")
            dis_map = torch.empty(Q, dtype=torch.float32, device=dev)
            if stage == 1:
                dis_map = data.dis_map.to(dev)
            else:
                _lib.check(l.fb_pair_dist(s["pocket_xyz"].data_ptr(), s["lig"].data_ptr(), s["pk"].data_ptr(), s["co"].data_ptr(),
                                          qo.data_ptr(), B, Q, 10.0, dis_map.data_ptr(), st), "fb_pair_dist")
            compound_coords_out = _select(Xo.view(-1, 3), c_rows, scale)
            # dense [B, Lmax] views of the per-residue quantities (to_dense_batch, model.py:138-144)
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
                    centers, dis_map, s["less5"])

    def inference(self, data):
        if self.training:
            raise NotImplementedError("fabind_b200: the training path is not built yet; call .eval()")
        with torch.no_grad():
            s = self._pocket_stage(data)
            centers = self._centers(s, data, mode=1)
            Xo, _ = self._dock(s, data, centers)
            c_rows = np.nonzero(~s["seg"] & ~s["glb"])[0]
            return _select(Xo.view(-1, 3), c_rows, self.coordinate_scale), data['compound'].batch


def get_model(args, logger, device):
    """model.py:582-586"""
    if args.mode == 5:
        logger.log_message("FABind")
        return IaBNet_mean_and_pocket_prediction_cls_coords_dependent(args, args.hidden_size, args.pocket_pred_hidden_size)
