"""Weights of the FABind+ equivariant message-passing layers (FABind_plus/fabind/models/egnn.py:20-433).

Same class names, constructor signatures and parameter names as the reference so that `load_state_dict(strict=True)`
of a FABind+ checkpoint works.  The layers run inside the fused CUDA stack driven by `att_model.EfficientMCAttModel`;
called on their own they run the same kernels on caller-supplied graphs through `fb_egnn_forward` (eval semantics)."""
import torch
import torch.nn as nn

from .. import _lib
from ..egnn import _check_args, _geom, _eval_only
from ..model_utils import _standalone
from .cross_att import CrossAttentionModule
from .model_utils import InteractionModule, MLPwithLastAct, MLPwoBias


def _check_args_plus(args):
    _check_args(args)
    need = dict(use_ln_mlp=True, mlp_hidden_scale=1, mha_heads=4, rel_dis_pair_bias="no", inter_additional_mlp=False,
                only_last_LAS=False)
    bad = {k: getattr(args, k, None) for k, v in need.items() if getattr(args, k, v) != v}
    if bad:
        raise NotImplementedError(f"fabind_b200 implements the published FABind+ configuration; unsupported flags: {bad}")


class MC_E_GCL(nn.Module):
    """egnn.py:20-42: edge_mlp / node_mlp = MLPwithLastAct, coord_mlp = MLPwoBias (linear2 xavier gain 0.001)."""

    def __init__(self, args, input_nf, output_nf, hidden_nf, n_channel, edges_in_d=0, act_fn=nn.SiLU(), residual=True,
                 attention=False, normalize=False, coords_agg='mean', tanh=False, dropout=0.1, coord_change_maximum=10):
        super().__init__()
        if n_channel != 1 or edges_in_d != 0 or attention or normalize or tanh or coords_agg != 'mean' or not residual \
                or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MC_E_GCL supports the configuration FABind+ instantiates")
        self.args, self.residual, self.coords_agg = args, residual, coords_agg
        self.attention, self.normalize, self.tanh, self.epsilon = attention, normalize, tanh, 1e-8
        n = args.mlp_hidden_scale
        self.edge_mlp = MLPwithLastAct(args, embedding_channels=input_nf * 2 + n_channel ** 2 + edges_in_d, n=n, out_channels=hidden_nf)
        self.node_mlp = MLPwithLastAct(args, embedding_channels=hidden_nf + input_nf, n=n, out_channels=output_nf)
        self.coord_mlp = MLPwoBias(args, embedding_channels=hidden_nf, n=n, out_channels=n_channel)
        torch.nn.init.xavier_uniform_(self.coord_mlp.linear2.weight, gain=0.001)
        self.coord_change_maximum = coord_change_maximum
        self.precision = "fp32"

    def forward(self, h, edge_index, coord, edge_attr=None, node_attr=None, batch_id=None):
        """egnn.py:100-115 on a caller-supplied edge list -> (h, coord)."""
        from ..substack import egnn_forward
        _eval_only(self)
        if edge_attr is not None or node_attr is not None or batch_id is None:
            raise NotImplementedError("fabind_b200: MC_E_GCL.forward needs batch_id and takes no edge/node attributes")
        ho, xo, _, _ = egnn_forward(self, self.args, "gnn.gcl_0.", h.shape[1], 1, _lib.STEP_GCL, h, coord, edge_index, None, None, None,
                                    batch_id, None, None, _geom(self.args, self.coord_change_maximum), bf16=self.precision == "bf16",
                                    flavour=_lib.FLAVOUR_PLUS)
        return ho, xo


class MC_Att_L(nn.Module):
    """egnn.py:118-150"""

    def __init__(self, args, input_nf, output_nf, hidden_nf, n_channel, edges_in_d=0, act_fn=nn.SiLU(), dropout=0.1,
                 coord_change_maximum=10, opm=False, normalize_coord=None):
        super().__init__()
        _check_args_plus(args)
        if n_channel != 1 or edges_in_d != 0 or opm or input_nf != hidden_nf or output_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MC_Att_L supports the configuration FABind+ instantiates")
        self.args, self.hidden_nf = args, hidden_nf
        self.dropout = nn.Dropout(args.dropout)
        self.linear_q = nn.Linear(input_nf, hidden_nf)
        self.linear_kv = nn.Linear(input_nf + n_channel ** 2 + edges_in_d, hidden_nf * 2)
        self.coord_mlp = MLPwoBias(args, embedding_channels=hidden_nf, n=args.mlp_hidden_scale, out_channels=n_channel)
        torch.nn.init.xavier_uniform_(self.coord_mlp.linear2.weight, gain=0.001)
        self.coord_change_maximum = coord_change_maximum
        self.cross_attn_module = CrossAttentionModule(args, node_hidden_dim=input_nf, pair_hidden_dim=input_nf,
                                                      rm_layernorm=args.rm_layernorm, keep_trig_attn=args.keep_trig_attn,
                                                      dist_hidden_dim=input_nf, normalize_coord=normalize_coord)
        # constructed (and checkpointed) by the reference but unused when add_cross_attn_layer is on (egnn.py:147-149)
        self.inter_layer = InteractionModule(input_nf, output_nf, hidden_nf, opm=opm, rm_layernorm=args.rm_layernorm)
        self.attn_bias_proj = nn.Linear(hidden_nf, 1)
        self.precision = "fp32"

    def forward(self, h, edge_index, coord, edge_attr=None, segment_id=None, batch_id=None, reduced_tuple=None,
                pair_embed_batched=None, pair_mask=None, LAS_mask=None, p_p_dist_embed=None, c_c_dist_embed=None):
        """egnn.py:280-300 on a caller-supplied (symmetric) inter edge list -> (h, coord, attention weights, pair_embed_batched)."""
        from ..substack import egnn_forward
        _eval_only(self)
        if edge_attr is not None or segment_id is None or batch_id is None or pair_embed_batched is None:
            raise NotImplementedError("fabind_b200: MC_Att_L.forward needs segment_id, batch_id and pair_embed_batched")
        ho, xo, atts, pair = egnn_forward(self, self.args, "gnn.att_0.", self.hidden_nf, 1, _lib.STEP_ATT, h, coord, None, edge_index, None,
                                          None, batch_id, segment_id, pair_embed_batched, _geom(self.args, self.coord_change_maximum),
                                          bf16=self.precision == "bf16", want_att=True, flavour=_lib.FLAVOUR_PLUS)
        return ho, xo, atts[0], pair


class MCAttEGNN(nn.Module):
    """egnn.py:303-357"""

    def __init__(self, args, in_node_nf, hidden_nf, out_node_nf, n_channel, in_edge_nf=0, act_fn=nn.SiLU(), n_layers=4,
                 residual=True, dropout=0.1, dense=False, normalize_coord=None, unnormalize_coord=None,
                 geometry_reg_step_size=0.001):
        super().__init__()
        _check_args_plus(args)
        if dense or in_edge_nf != 0 or n_channel != 1 or in_node_nf != hidden_nf or out_node_nf != hidden_nf:
            raise NotImplementedError("fabind_b200: MCAttEGNN needs embed == hidden == out size, dense=False")
        self.args = args
        self.geometry_reg_step_size = geometry_reg_step_size
        self.geom_reg_steps = 1
        self.hidden_nf, self.n_layers, self.dense = hidden_nf, n_layers, dense
        self.normalize_coord, self.unnormalize_coord = normalize_coord, unnormalize_coord
        self.dropout = nn.Dropout(args.dropout)
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
        """egnn.py:359-433 on caller-supplied graphs -> (h, x[, atts], pair_embed_batched)."""
        from ..substack import egnn_forward
        _eval_only(self)
        if ctx_edge_attr is not None or att_edge_attr is not None or segment_id is None or batch_id is None:
            raise NotImplementedError("fabind_b200: MCAttEGNN.forward needs segment_id/batch_id and takes no edge attributes")
        steps = (_lib.STEP_LINEAR_IN | _lib.STEP_GCL | _lib.STEP_ATT | _lib.STEP_LAS | _lib.STEP_OUT_LAYER | _lib.STEP_LINEAR_OUT)
        ho, xo, atts, pair = egnn_forward(self, self.args, "gnn.", self.hidden_nf, self.n_layers, steps, h, x, ctx_edges, att_edges,
                                          LAS_edge_list, batched_complex_coord_LAS, batch_id, segment_id, pair_embed_batched,
                                          _geom(self.args, self.normalize_coord(10), self.geometry_reg_step_size),
                                          bf16=getattr(self, "precision", "fp32") == "bf16", want_att=return_attention,
                                          flavour=_lib.FLAVOUR_PLUS)
        return (ho, xo, atts, pair) if return_attention else (ho, xo, pair)
