"""Parameter containers with the reference's names and shapes (FABind/fabind/models/model_utils.py).

These modules own the weights (identical ``state_dict`` keys, so reference checkpoints load with
``strict=True``); the arithmetic lives in libfabind_b200 and is driven from
``att_model.EfficientMCAttModel.forward``.  A sub-module called on its own raises instead of silently
running a PyTorch path.
"""
import torch.nn as nn


def _standalone(name):
    raise NotImplementedError(
        f"{name} is executed inside the fused CUDA stack (EfficientMCAttModel.forward); "
        "fabind_b200 has no per-module PyTorch path")


class Attention(nn.Module):
    """Gated multi-head attention weights (model_utils.py:41-94): q/k/v without bias, gate and output with."""

    def __init__(self, c_q, c_k, c_v, c_hidden, no_heads, gating=True):
        super().__init__()
        self.c_q, self.c_k, self.c_v, self.c_hidden, self.no_heads, self.gating = c_q, c_k, c_v, c_hidden, no_heads, gating
        self.linear_q = nn.Linear(c_q, c_hidden * no_heads, bias=False)
        self.linear_k = nn.Linear(c_k, c_hidden * no_heads, bias=False)
        self.linear_v = nn.Linear(c_v, c_hidden * no_heads, bias=False)
        self.linear_o = nn.Linear(c_hidden * no_heads, c_q)
        self.linear_g = nn.Linear(c_q, c_hidden * no_heads) if gating else None

    def forward(self, *a, **k):
        _standalone("Attention")


class Transition(nn.Module):
    """model_utils.py:162-175 (rm_layernorm only)."""

    def __init__(self, hidden_dim=128, n=4, rm_layernorm=False):
        super().__init__()
        if not rm_layernorm:
            raise NotImplementedError("fabind_b200 implements the published configuration (--rm-layernorm)")
        self.rm_layernorm = rm_layernorm
        self.linear_1 = nn.Linear(hidden_dim, n * hidden_dim)
        self.linear_2 = nn.Linear(n * hidden_dim, hidden_dim)

    def forward(self, *a, **k):
        _standalone("Transition")


class InteractionModule(nn.Module):
    """model_utils.py:177-223 (opm=False, rm_layernorm only)."""

    def __init__(self, node_hidden_dim, pair_hidden_dim, hidden_dim, opm=False, rm_layernorm=False):
        super().__init__()
        if opm or not rm_layernorm:
            raise NotImplementedError("fabind_b200 implements the published configuration (opm off, --rm-layernorm)")
        self.hidden_dim, self.pair_hidden_dim, self.node_hidden_dim = hidden_dim, pair_hidden_dim, node_hidden_dim
        self.opm, self.rm_layernorm = opm, rm_layernorm
        self.linear_p = nn.Linear(node_hidden_dim, hidden_dim)
        self.linear_c = nn.Linear(node_hidden_dim, hidden_dim)
        self.linear_out = nn.Linear(hidden_dim, pair_hidden_dim)

    def forward(self, *a, **k):
        _standalone("InteractionModule")
