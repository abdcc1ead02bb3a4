"""Weights of the FABind+ protein<->ligand cross attention (FABind_plus/fabind/models/cross_att.py:7-92)."""
import torch.nn as nn

from .model_utils import Attention, InteractionModule, MLPwithLastAct, _standalone


class RowAttentionBlock(nn.Module):
    """cross_att.py:51-92"""

    def __init__(self, args, node_hidden_dim, pair_hidden_dim, attention_hidden_dim=32, no_heads=4, dropout=0.1,
                 rm_layernorm=False, mha_permu=False):
        super().__init__()
        if not rm_layernorm or attention_hidden_dim != 32 or no_heads != 4:
            raise NotImplementedError("fabind_b200: RowAttentionBlock is built for 4 heads x 32 channels, --rm-layernorm")
        self.args, self.mha_permu = args, mha_permu   # mha_permu only changes a permute order inside the reference
        self.no_heads, self.attention_hidden_dim = no_heads, attention_hidden_dim
        self.pair_hidden_dim, self.node_hidden_dim, self.rm_layernorm = pair_hidden_dim, node_hidden_dim, rm_layernorm
        self.linear = nn.Linear(pair_hidden_dim, no_heads)
        self.linear_g = nn.Linear(pair_hidden_dim, no_heads)
        self.dropout = nn.Dropout(args.dropout)
        self.mha = Attention(node_hidden_dim, node_hidden_dim, node_hidden_dim, attention_hidden_dim, no_heads)

    def forward(self, *a, **k):
        _standalone("RowAttentionBlock")


class CrossAttentionModule(nn.Module):
    """cross_att.py:7-45"""

    def __init__(self, args, node_hidden_dim, pair_hidden_dim, rm_layernorm=False, keep_trig_attn=False, dist_hidden_dim=32,
                 normalize_coord=None):
        super().__init__()
        if keep_trig_attn:
            raise NotImplementedError("keep_trig_attn is off in the published configuration and not built")
        if int(getattr(args, "mha_heads", 4)) != 4:
            raise NotImplementedError("fabind_b200: --mha-heads 4 only (the published value)")
        self.args, self.pair_hidden_dim, self.keep_trig_attn = args, pair_hidden_dim, keep_trig_attn
        self.p_attention_block = RowAttentionBlock(args, node_hidden_dim, pair_hidden_dim, no_heads=args.mha_heads, rm_layernorm=rm_layernorm, mha_permu=True)
        self.c_attention_block = RowAttentionBlock(args, node_hidden_dim, pair_hidden_dim, no_heads=args.mha_heads, rm_layernorm=rm_layernorm, mha_permu=False)
        n = args.mlp_hidden_scale
        self.p_transition = MLPwithLastAct(args, embedding_channels=node_hidden_dim, n=n, out_channels=node_hidden_dim)
        self.c_transition = MLPwithLastAct(args, embedding_channels=node_hidden_dim, n=n, out_channels=node_hidden_dim)
        self.pair_transition = MLPwithLastAct(args, embedding_channels=pair_hidden_dim, n=n, out_channels=pair_hidden_dim)
        self.inter_layer = InteractionModule(node_hidden_dim, pair_hidden_dim, 32, opm=False, rm_layernorm=rm_layernorm)

    def forward(self, *a, **k):
        _standalone("CrossAttentionModule")
