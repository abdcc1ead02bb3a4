"""Weights of the equivariant message-passing layers (FABind/fabind/models/egnn.py:20-466).

Same class names, constructor signatures and parameter names as the reference so that
``load_state_dict(strict=True)`` of a reference checkpoint works.  The layers run inside the fused
CUDA stack driven by ``att_model.EfficientMCAttModel``.
"""
import torch
import torch.nn as nn

from . import _lib
from .cross_att import CrossAttentionModule
from .model_utils import InteractionModule


def _geom(args, coord_clamp, step=0.001):
    scale = float(getattr(args, "coordinate_scale", 5.0))
    return dict(intra_cutoff=float(args.intra_cutoff) / scale, inter_cutoff=float(args.inter_cutoff) / scale,
                coord_clamp=float(coord_clamp), las_clamp=15.0 / scale, las_step=float(step))


def _eval_only(mod):
    if mod.training:
        raise NotImplementedError("fabind_b200: the training path (dropout + backward kernels) is not built yet; call .eval()")


def _check_args(args):
    need = dict(rm_F_norm=False, norm_type="per_sample", rm_layernorm=True, add_attn_pair_bias=True,
                explicit_pair_embed=True, add_cross_attn_layer=True, keep_trig_attn=False, opm=False,
                fix_pocket=False, rm_LAS_constrained_optim=False)
    bad = {k: getattr(args, k, None) for k, v in need.items() if getattr(args, k, v) != v}
    if bad:
        raise NotImplementedError(
            f"fabind_b200 implements the published FABind configuration; unsupported flags: {bad}")


class MC_E_GCL(nn.Module):
    """egnn.py:20-66.  edge_mlp: Linear(2H+1,H) SiLU Linear(H,H) SiLU; node_mlp: Linear(2H,H) SiLU Linear(H,H);
    coord_mlp: Linear(H,H) SiLU Linear(H,1,no bias, xavier gain 0.001)."""

    def __init__(self, args, input_nf, output_nf, hidden_nf, n_channel, edges_in_d=0, act_fn=nn.SiLU(), residual=True,
                 attention=False, normalize=False, coords_agg='mean', tanh=False, dropout=0.1, coord_change_maximum=10):
        super().__init__()
        if n_channel != 1 or edges_in_d != 0 or attention or normalize or tanh or coords_agg != 'mean' or not residual \
                or not isinstance(act_fn, nn.SiLU) or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MC_E_GCL supports the configuration FABind instantiates")
        self.args, self.residual, self.coords_agg = args, residual, coords_agg
        self.dropout = nn.Dropout(dropout)
        self.edge_mlp = nn.Sequential(nn.Linear(input_nf * 2 + n_channel ** 2 + edges_in_d, hidden_nf), act_fn,
                                      nn.Linear(hidden_nf, hidden_nf), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(hidden_nf + input_nf, hidden_nf), act_fn, nn.Linear(hidden_nf, output_nf))
        layer = nn.Linear(hidden_nf, n_channel, bias=False)
        torch.nn.init.xavier_uniform_(layer.weight, gain=0.001)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_nf, hidden_nf), act_fn, layer)
        self.coord_change_maximum = coord_change_maximum
        self.precision = "fp32"

    def forward(self, h, edge_index, coord, edge_attr=None, node_attr=None, batch_id=None):
        """egnn.py:130-144 on a caller-supplied edge list -> (h, coord)."""
        from .substack import egnn_forward
        _eval_only(self)
        if edge_attr is not None or node_attr is not None or batch_id is None:
            raise NotImplementedError("fabind_b200: MC_E_GCL.forward needs batch_id and takes no edge/node attributes")
        hidden = h.shape[1]
        ho, xo, _ = egnn_forward(self, self.args, "gnn.gcl_0.", hidden, 1, _lib.STEP_GCL, h, coord, edge_index, None, None, None,
                                 batch_id, None, None, _geom(self.args, self.coord_change_maximum), bf16=self.precision == "bf16")
        return ho, xo


class MC_Att_L(nn.Module):
    """egnn.py:147-183."""

    def __init__(self, args, input_nf, output_nf, hidden_nf, n_channel, edges_in_d=0, act_fn=nn.SiLU(), dropout=0.1,
                 coord_change_maximum=10, opm=False, normalize_coord=None):
        super().__init__()
        _check_args(args)
        if n_channel != 1 or edges_in_d != 0 or opm or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MC_Att_L supports the configuration FABind instantiates")
        self.args, self.hidden_nf = args, hidden_nf
        self.dropout = nn.Dropout(dropout)
        self.linear_q = nn.Linear(input_nf, hidden_nf)
        self.linear_kv = nn.Linear(input_nf + n_channel ** 2 + edges_in_d, hidden_nf * 2)
        layer = nn.Linear(hidden_nf, n_channel, bias=False)
        torch.nn.init.xavier_uniform_(layer.weight, gain=0.001)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_nf, hidden_nf), act_fn, layer)
        self.coord_change_maximum = coord_change_maximum
        self.cross_attn_module = CrossAttentionModule(node_hidden_dim=input_nf, pair_hidden_dim=input_nf,
                                                      rm_layernorm=args.rm_layernorm, keep_trig_attn=args.keep_trig_attn,
                                                      dist_hidden_dim=input_nf, normalize_coord=normalize_coord)
        # constructed (and checkpointed) by the reference but never used when add_cross_attn_layer is on
        # (egnn.py:180-182 vs 266-284); kept so that state_dict keys match
        self.inter_layer = InteractionModule(input_nf, output_nf, hidden_nf, opm=opm, rm_layernorm=args.rm_layernorm)
        self.attn_bias_proj = nn.Linear(hidden_nf, 1)
        self.precision = "fp32"

    def forward(self, h, edge_index, coord, edge_attr=None, segment_id=None, batch_id=None, reduced_tuple=None,
                pair_embed_batched=None, pair_mask=None, LAS_mask=None, p_p_dist_embed=None, c_c_dist_embed=None):
        """egnn.py:308-333 on a caller-supplied (symmetric) inter edge list -> (h, coord, attention weights)."""
        from .substack import egnn_forward
        _eval_only(self)
        if edge_attr is not None or segment_id is None or batch_id is None or pair_embed_batched is None:
            raise NotImplementedError("fabind_b200: MC_Att_L.forward needs segment_id, batch_id and pair_embed_batched")
        ho, xo, atts = egnn_forward(self, self.args, "gnn.att_0.", self.hidden_nf, 1, _lib.STEP_ATT, h, coord, None, edge_index, None,
                                    None, batch_id, segment_id, pair_embed_batched, _geom(self.args, self.coord_change_maximum),
                                    bf16=self.precision == "bf16", want_att=True)
        return ho, xo, atts[0]


class MCAttEGNN(nn.Module):
    """egnn.py:336-390."""

    def __init__(self, args, in_node_nf, hidden_nf, out_node_nf, n_channel, in_edge_nf=0, act_fn=nn.SiLU(), n_layers=4,
                 residual=True, dropout=0.1, dense=False, normalize_coord=None, unnormalize_coord=None,
                 geometry_reg_step_size=0.001):
        super().__init__()
        _check_args(args)
        if dense or in_edge_nf != 0 or n_channel != 1 or in_node_nf != hidden_nf or out_node_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MCAttEGNN needs embed == hidden == out size, dense=False")
        self.args = args
        self.geometry_reg_step_size = geometry_reg_step_size
        self.geom_reg_steps = 1
        self.hidden_nf, self.n_layers, self.dense = hidden_nf, n_layers, dense
        self.normalize_coord, self.unnormalize_coord = normalize_coord, unnormalize_coord
        self.dropout = nn.Dropout(dropout)
        self.linear_in = nn.Linear(in_node_nf, hidden_nf)
        self.linear_out = nn.Linear(hidden_nf, out_node_nf)
        cmax = normalize_coord(10)
        for i in range(n_layers):
            self.add_module(f'gcl_{i}', MC_E_GCL(args, hidden_nf, hidden_nf, hidden_nf, n_channel, edges_in_d=in_edge_nf,
                                                 act_fn=act_fn, residual=residual, dropout=dropout, coord_change_maximum=cmax))
            self.add_module(f'att_{i}', MC_Att_L(args, hidden_nf, hidden_nf, hidden_nf, n_channel, edges_in_d=0, act_fn=act_fn,
                                                 dropout=dropout, coord_change_maximum=cmax, opm=args.opm,
                                                 normalize_coord=normalize_coord))
        self.out_layer = MC_E_GCL(args, hidden_nf, hidden_nf, hidden_nf, n_channel, edges_in_d=in_edge_nf, act_fn=act_fn,
                                  residual=residual, coord_change_maximum=cmax)

    def forward(self, h, x, ctx_edges, att_edges, LAS_edge_list, batched_complex_coord_LAS, segment_id=None, batch_id=None,
                reduced_tuple=None, pair_embed_batched=None, pair_mask=None, LAS_mask=None, p_p_dist_embed=None,
                c_c_dist_embed=None, mask=None, ctx_edge_attr=None, att_edge_attr=None, return_attention=False):
        """egnn.py:392-466 on caller-supplied graphs -> (h, x[, atts])."""
        from .substack import egnn_forward
        _eval_only(self)
        if ctx_edge_attr is not None or att_edge_attr is not None or segment_id is None or batch_id is None:
            raise NotImplementedError("fabind_b200: MCAttEGNN.forward needs segment_id/batch_id and takes no edge attributes")
        steps = (_lib.STEP_LINEAR_IN | _lib.STEP_GCL | _lib.STEP_ATT | _lib.STEP_LAS | _lib.STEP_OUT_LAYER | _lib.STEP_LINEAR_OUT)
        ho, xo, atts = egnn_forward(self, self.args, "gnn.", self.hidden_nf, self.n_layers, steps, h, x, ctx_edges, att_edges,
                                    LAS_edge_list, batched_complex_coord_LAS, batch_id, segment_id, pair_embed_batched,
                                    _geom(self.args, self.normalize_coord(10), self.geometry_reg_step_size),
                                    bf16=getattr(self, "precision", "fp32") == "bf16", want_att=return_attention)
        return (ho, xo, atts) if return_attention else (ho, xo)
