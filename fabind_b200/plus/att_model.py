"""Drop-in replacement for FABind_plus/fabind/models/att_model.py: `ComplexGraph` (shared with v1: the edge rules are
identical, att_model.py:29-126) and `EfficientMCAttModel` whose forward returns `(X, H, pair_embed_batched)`
(att_model.py:166-223).  eval(): inference semantics (refine='refine_coord').  train() under torch.no_grad(): the reference's
dropout SAMPLING mode (P/test_sampling_fabind.py:118-124) with in-kernel masks at every nn.Dropout site and the `random_n_iter`
draw of att_model.py:199-202; train() with autograd enabled routes to the training step (`train.forward_with_grad`, FABind+ reverse
pass; it carries no dropout masks and therefore refuses dropout_p > 0)."""
import os

import torch
import torch.nn as nn

from .. import _lib
from ..att_model import ComplexGraph  # noqa: F401
from ..runtime import PackedWeights, model_forward
from .egnn import MCAttEGNN, _check_args_plus
from .model_utils import InteractionModule


class EfficientMCAttModel(nn.Module):
    """att_model.py:131-223"""

    def __init__(self, args, embed_size, hidden_size, n_channel, n_edge_feats=0, n_layers=5, dropout=0.1, n_iter=5,
                 dense=False, inter_cutoff=10, intra_cutoff=8, normalize_coord=None, unnormalize_coord=None):
        super().__init__()
        _check_args_plus(args)
        if getattr(args, "ablation_no_attention", False) or getattr(args, "ablation_no_attention_with_cross_attn", False):
            raise NotImplementedError("ablation variants are out of scope (not used by any published configuration)")
        if getattr(args, "refine", "refine_coord") != "refine_coord":
            raise NotImplementedError("only refine='refine_coord' (the published mode) is built")
        if embed_size != hidden_size:
            raise NotImplementedError("embed_size must equal hidden_size (true for both FABind+ stages)")
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
                         las_step=float(args.geometry_reg_step_size), flavour=_lib.FLAVOUR_PLUS)
        self._packed = PackedWeights()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._packed.invalidate())
        self.precision = os.environ.get("FABIND_B200_PRECISION", "fp32")
        self.return_pair = True    # False: skip the dense [B, max_p, max_c, H] fp32 copy of the pair embedding (returns None)
        # train() mode = the reference's SAMPLING mode (P/test_sampling_fabind.py:118-124 runs the model in train() mode
        # under no_grad so that every nn.Dropout is active): masks are applied in-kernel, keyed by `dropout_seed`
        # (None: drawn from torch's global generator per call, like nn.Dropout) -- see fabind_b200/dropout.py
        self.dropout_p = float(getattr(args, "dropout", 0.0))
        self.dropout_seed = None
        self.dropout_colonly = False   # tests: column-only masks (row-order invariant) to pin mask placement
        self.last_stats = None
        self.debug_trace = False

    def _apply(self, fn, *a, **k):
        self._packed.invalidate()
        return super()._apply(fn, *a, **k)

    def layout_cutoff(self):
        """the normalised intra cutoff a dataloader-side layout counts the residue-residue edges with
        (fabind_b200.dataloader.layout_hint / prepare_batch)"""
        return self._cfg["intra_cutoff"]

    def invalidate_packed_weights(self):
        """call after editing parameters through `.data` (see fabind_b200.runtime.PackedWeights)"""
        self._packed.invalidate()

    def forward(self, X, H, batch_id, segment_id, mask, is_global, compound_edge_index, LAS_edge_index,
                batched_complex_coord_LAS, LAS_mask=None):
        dropout, n_iter = None, None
        if self.training:
            if torch.is_grad_enabled():
                # FABind+ training step: the reverse pass of this layout (fabind_b200/train.py) differentiates the eval-mode
                # arithmetic; its dropout masks are not threaded through the reverse kernels yet, so refuse rather than train a
                # different objective silently
                if self.dropout_p > 0:
                    raise NotImplementedError("fabind_b200.plus: train() with autograd needs dropout_p = 0 (the FABind+ reverse pass "
                                              "carries no dropout masks yet); sampling mode = train() under torch.no_grad()")
                from .. import train
                return train.forward_with_grad(self, dict(X=X, H=H, batch_id=batch_id, segment_id=segment_id, mask=mask,
                                                          is_global=is_global, compound_edge_index=compound_edge_index,
                                                          LAS_edge_index=LAS_edge_index,
                                                          batched_complex_coord_LAS=batched_complex_coord_LAS, LAS_mask=LAS_mask))
            seed = self.dropout_seed if self.dropout_seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
            dropout = (self.dropout_p, seed, self.dropout_colonly)
            if self.random_n_iter:        # att_model.py:199-202: iter_i = random.randint(1, n_iter) in training mode
                import random
                n_iter = random.randint(1, self.n_iter)
        if self.precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        with torch.no_grad():
            H_out, stats, e_ctx, tr, pair = model_forward(self, self._packed, X, H, batch_id, segment_id, mask, is_global,
                                                          compound_edge_index, LAS_edge_index, batched_complex_coord_LAS,
                                                          self._cfg, self.precision, trace=self.debug_trace,
                                                          want_pair=self.return_pair, dropout=dropout, n_iter=n_iter)
        self.last_stats = dict(inter_edges_per_iter=stats, ctx_edges=e_ctx, trace=tr)
        return X, H_out, pair
