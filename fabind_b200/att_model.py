"""Drop-in replacements for FABind/fabind/models/att_model.py: `ComplexGraph` and `EfficientMCAttModel`.

Same constructor and forward signatures, same state_dict keys; the forward pass is one call into
libfabind_b200 (hand-written sm_100a kernels).  eval(): inference semantics (refine='refine_coord', att_model.py:227-245);
train() under no_grad: the same with every nn.Dropout active; train() with autograd: the training step of fabind_b200/train.py
(earlier iterations without gradient, the last one differentiated by the reverse-pass kernels).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .egnn import MCAttEGNN, _check_args
from .model_utils import InteractionModule
from .runtime import PackedWeights, model_forward, current_stream_ptr


class ComplexGraph(nn.Module):
    """att_model.py:29-120.  `construct_edges` returns the reference's edge lists (same order, int64)."""

    def __init__(self, args, inter_cutoff=10, intra_cutoff=8, normalize_coord=None, unnormalize_coord=None):
        super().__init__()
        self.args = args
        self.inter_cutoff = normalize_coord(inter_cutoff)
        self.intra_cutoff = normalize_coord(intra_cutoff)

    @torch.no_grad()
    def construct_edges(self, X, batch_id, segment_ids, is_global):
        l = _lib.lib()
        dev = X.device
        if dev.type != "cuda":
            raise RuntimeError("fabind_b200 runs on CUDA tensors only (no CPU fallback)")
        N = X.shape[0]
        x = X[:, 0].to(torch.float32).contiguous()
        bid = batch_id.to(torch.int32).contiguous()
        B = int(batch_id[-1]) + 1
        counts = torch.bincount(batch_id, minlength=B)
        off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        flags = (segment_ids.to(torch.bool).to(torch.uint8) | (is_global.to(torch.uint8) << 1)).contiguous()
        ws = torch.zeros(4 * N + 4 * (N + 1) + 8, dtype=torch.int32, device=dev)
        st = current_stream_ptr(dev)
        _lib.check(l.fb_edges_ref_count(N, bid.data_ptr(), off.data_ptr(), flags.data_ptr(), x.data_ptr(),
                                        float(self.intra_cutoff), float(self.inter_cutoff), ws.data_ptr(), st),
                   "fb_edges_ref_count")
        rp = ws[4 * N:4 * N + 4 * (N + 1)].view(4, N + 1)
        host = torch.cat([rp[:, N], ws[4 * N + 4 * (N + 1):4 * N + 4 * (N + 1) + 1]]).cpu().numpy()
        counts_h = (C.c_int32 * 5)(*[int(v) for v in host])
        e_ctx = int(host[0] + host[1] + host[2])
        e_int = 2 if host[4] else int(host[3])
        counts_h[3] = e_int
        ctx = torch.empty((2, e_ctx), dtype=torch.int64, device=dev)
        inter = torch.empty((2, e_int), dtype=torch.int64, device=dev)
        _lib.check(l.fb_edges_ref_fill(N, bid.data_ptr(), off.data_ptr(), flags.data_ptr(), x.data_ptr(),
                                       float(self.intra_cutoff), float(self.inter_cutoff), ws.data_ptr(), counts_h,
                                       ctx.data_ptr(), inter.data_ptr(), st), "fb_edges_ref_fill")
        if self.args.add_attn_pair_bias:
            fwd = inter[0] < inter[1]
            red_b = batch_id[inter[0][fwd]]
            red_off = off[:-1].to(torch.int64)[red_b]
            return ctx, inter, (red_b, red_off)
        return ctx, inter, None

    def forward(self, X, batch_id, segment_id, is_global):
        return self.construct_edges(X, batch_id, segment_id, is_global)


class EfficientMCAttModel(nn.Module):
    """att_model.py:131-246."""

    def __init__(self, args, embed_size, hidden_size, n_channel, n_edge_feats=0, n_layers=5, dropout=0.1, n_iter=5,
                 dense=False, inter_cutoff=10, intra_cutoff=8, normalize_coord=None, unnormalize_coord=None):
        super().__init__()
        _check_args(args)
        if getattr(args, "ablation_no_attention", False) or getattr(args, "ablation_no_attention_with_cross_attn", False):
            raise NotImplementedError("ablation variants are out of scope (not used by any published configuration)")
        if getattr(args, "refine", "refine_coord") != "refine_coord":
            raise NotImplementedError("only refine='refine_coord' (the published mode) is built")
        if embed_size != hidden_size:
            raise NotImplementedError("embed_size must equal hidden_size (true for both FABind stages)")
        self.n_iter = n_iter
        self.args = args
        self.random_n_iter = args.random_n_iter
        self.hidden_size, self.n_layers = hidden_size, n_layers
        self.gnn = MCAttEGNN(args, embed_size, hidden_size, hidden_size, n_channel, n_edge_feats, n_layers=n_layers,
                             residual=True, dropout=dropout, dense=dense, normalize_coord=normalize_coord,
                             unnormalize_coord=unnormalize_coord, geometry_reg_step_size=args.geometry_reg_step_size)
        self.extract_edges = ComplexGraph(args, inter_cutoff=inter_cutoff, intra_cutoff=intra_cutoff,
                                          normalize_coord=normalize_coord, unnormalize_coord=unnormalize_coord)
        self.inter_layer = InteractionModule(hidden_size, hidden_size, hidden_size, rm_layernorm=args.rm_layernorm)
        self._cfg = dict(hidden=hidden_size, n_layers=n_layers, n_iter=n_iter,
                         intra_cutoff=float(normalize_coord(intra_cutoff)), inter_cutoff=float(normalize_coord(inter_cutoff)),
                         coord_clamp=float(normalize_coord(10)), las_clamp=float(normalize_coord(15)),
                         las_step=float(args.geometry_reg_step_size))
        self._packed = PackedWeights()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._packed.invalidate())
        # "fp32": FFMA GEMMs (parity mode);  "fp32_tc": fp32 activations, GEMMs on tcgen05 as six bf16 products per term (the
        # tensor-core parity mode, <= 1e-4 rel vs the reference);  "bf16x3": three products;  "bf16": tcgen05 GEMMs with bf16
        # operands / fp32 accumulation (coordinates, softmax statistics and the residual stream stay fp32)
        self.precision = os.environ.get("FABIND_B200_PRECISION", "fp32")
        # attention core of the RowAttentionBlocks in bf16 mode: "simt" (default, the faster kernel at PDBbind block sizes) or "tcgen05"
        # (csrc/xatt_tc.cu: TMA-fed tiles, scores and outputs in TMEM)
        self.attention = "simt"
        # train() mode: every nn.Dropout of the reference's stack is active (egnn.py:82,106,236,398,461, cross_att.py:128), in all
        # refinement iterations.  Masks are applied in-kernel, keyed by `dropout_seed` (None: drawn from torch's global generator per
        # call, like nn.Dropout; see fabind_b200/dropout.py); `dropout_colonly` = column-only masks (tests: pins mask placement)
        self.dropout_p = float(dropout)
        self.dropout_seed = None
        self.dropout_colonly = False
        self.last_stats = None
        self.debug_trace = False   # tests: record h/x after every sub-layer of the last iteration

    def _apply(self, fn, *a, **k):
        self._packed.invalidate()
        return super()._apply(fn, *a, **k)

    def layout_cutoff(self):
        """the normalised intra cutoff a dataloader-side layout counts the residue-residue edges with
        (fabind_b200.dataloader.layout_hint / prepare_batch)"""
        return self._cfg["intra_cutoff"]

    def invalidate_packed_weights(self):
        """call after editing parameters through `.data` (EMA swaps, manual re-initialisation): such writes change neither
        `_version` nor `data_ptr()`, which is what the packed-arena cache keys on"""
        self._packed.invalidate()

    def forward(self, X, H, batch_id, segment_id, mask, is_global, compound_edge_index, LAS_edge_index,
                batched_complex_coord_LAS, LAS_mask=None):
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        dropout, n_iter = None, None
        if self.training:
            if self.dropout_p > 0:
                seed = self.dropout_seed if self.dropout_seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
                dropout = (self.dropout_p, seed, self.dropout_colonly)
            if torch.is_grad_enabled():
                # training step (att_model.py:210-245): iterations 0..n-2 under no_grad, the last one differentiated -- one autograd
                # node whose backward runs the reverse-pass kernels (fabind_b200/train.py)
                from . import train
                return train.forward_with_grad(self, dict(X=X, H=H, batch_id=batch_id, segment_id=segment_id, mask=mask,
                                                          is_global=is_global, compound_edge_index=compound_edge_index,
                                                          LAS_edge_index=LAS_edge_index,
                                                          batched_complex_coord_LAS=batched_complex_coord_LAS, LAS_mask=LAS_mask),
                                               dropout=dropout)
            if self.random_n_iter:        # att_model.py:210-211: iter_i = random.randint(1, n_iter) in training mode
                import random
                n_iter = random.randint(1, self.n_iter)
        with torch.no_grad():
            H_out, stats, e_ctx, tr = model_forward(self, self._packed, X, H, batch_id, segment_id, mask, is_global,
                                                    compound_edge_index, LAS_edge_index, batched_complex_coord_LAS,
                                                    self._cfg, self.precision, trace=self.debug_trace, dropout=dropout, n_iter=n_iter)
        self.last_stats = dict(inter_edges_per_iter=stats, ctx_edges=e_ctx, trace=tr)
        return X, H_out
