"""Weights of the protein<->ligand cross attention (FABind/fabind/models/cross_att.py:7-54,95-134)."""
import torch.nn as nn

from .model_utils import Attention, Transition, InteractionModule, _standalone


class RowAttentionBlock(nn.Module):
    def __init__(self, node_hidden_dim, pair_hidden_dim, attention_hidden_dim=32, no_heads=4, dropout=0.1,
                 rm_layernorm=False):
        super().__init__()
        if not rm_layernorm or attention_hidden_dim != 32 or no_heads != 4:
            raise NotImplementedError("fabind_b200: RowAttentionBlock is built for 4 heads x 32 channels, --rm-layernorm")
        self.no_heads, self.attention_hidden_dim = no_heads, attention_hidden_dim
        self.pair_hidden_dim, self.node_hidden_dim, self.rm_layernorm = pair_hidden_dim, node_hidden_dim, rm_layernorm
        self.linear = nn.Linear(pair_hidden_dim, no_heads)
        self.linear_g = nn.Linear(pair_hidden_dim, no_heads)
        self.dropout = nn.Dropout(dropout)
        self.mha = Attention(node_hidden_dim, node_hidden_dim, node_hidden_dim, attention_hidden_dim, no_heads)

    def forward(self, *a, **k):
        _standalone("RowAttentionBlock")


class CrossAttentionModule(nn.Module):
    def __init__(self, node_hidden_dim, pair_hidden_dim, rm_layernorm=False, keep_trig_attn=False, dist_hidden_dim=32,
                 normalize_coord=None):
        super().__init__()
        if keep_trig_attn:
            raise NotImplementedError("keep_trig_attn is off in the published configuration and not built")
        self.pair_hidden_dim, self.keep_trig_attn = pair_hidden_dim, keep_trig_attn
        self.p_attention_block = RowAttentionBlock(node_hidden_dim, pair_hidden_dim, rm_layernorm=rm_layernorm)
        self.c_attention_block = RowAttentionBlock(node_hidden_dim, pair_hidden_dim, rm_layernorm=rm_layernorm)
        self.p_transition = Transition(node_hidden_dim, 2, rm_layernorm=rm_layernorm)
        self.c_transition = Transition(node_hidden_dim, 2, rm_layernorm=rm_layernorm)
        self.pair_transition = Transition(pair_hidden_dim, 2, rm_layernorm=rm_layernorm)
        self.inter_layer = InteractionModule(node_hidden_dim, pair_hidden_dim, 32, opm=False, rm_layernorm=rm_layernorm)

    def forward(self, *a, **k):
        _standalone("CrossAttentionModule")
