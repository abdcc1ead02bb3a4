"""Parameter containers of FABind_plus/fabind/models/model_utils.py:10-74 (MLP / MLPwithLastAct / MLPwoBias with
`--use-ln-mlp`) plus the attention / interaction containers shared with v1.  The arithmetic lives in libfabind_b200;
a container called on its own raises."""
import torch.nn as nn

from ..model_utils import Attention, InteractionModule, _standalone  # noqa: F401  (same shapes and keys as v1)


def _check_mlp_args(args):
    if not getattr(args, "use_ln_mlp", False):
        raise NotImplementedError("fabind_b200 (FABind+ layout) implements the published configuration: --use-ln-mlp")
    if int(getattr(args, "mlp_hidden_scale", 1)) != 1:
        raise NotImplementedError("fabind_b200 (FABind+ layout): --mlp-hidden-scale 1 only (the published value)")


class _LnMlp(nn.Module):
    """layernorm -> linear1 -> ReLU -> linear2; dropout modules own no parameters and are kept for parity of
    `named_modules` only (model_utils.py:16-17,38-40,61-62)."""
    _last_bias = True
    _dropouts = ("dropout",)

    def __init__(self, args, embedding_channels=256, out_channels=256, n=4):
        super().__init__()
        _check_mlp_args(args)
        self.args = args
        self.layernorm = nn.LayerNorm(embedding_channels)
        if args.dropout > 0:
            for d in self._dropouts:
                setattr(self, d, nn.Dropout(args.dropout))
        self.linear1 = nn.Linear(embedding_channels, int(n * embedding_channels))
        self.linear2 = nn.Linear(int(n * embedding_channels), out_channels, bias=self._last_bias)

    def forward(self, *a, **k):
        _standalone(type(self).__name__)


class MLP(_LnMlp):
    """model_utils.py:10-30"""


class MLPwithLastAct(_LnMlp):
    """model_utils.py:32-53 (ReLU after linear2 as well)"""
    _dropouts = ("dropout1", "dropout2")


class MLPwoBias(_LnMlp):
    """model_utils.py:55-74 (linear2 without bias)"""
    _last_bias = False


class MLP4Confidence(nn.Module):
    """model_utils.py:76-97: confidence / ranking head MLP (optional LayerNorm, own dropout rate; evaluated in eval mode by the
    sampling scripts, P/test_sampling_fabind.py:120-123)."""

    def __init__(self, args, embedding_channels=256, out_channels=256, n=4):
        super().__init__()
        self.args = args
        if args.confidence_use_ln_mlp:
            self.layernorm = nn.LayerNorm(embedding_channels)
        if args.confidence_dropout > 0:
            self.dropout = nn.Dropout(args.confidence_dropout)
        self.linear1 = nn.Linear(embedding_channels, n * embedding_channels)
        self.linear2 = nn.Linear(n * embedding_channels, out_channels)

    def forward(self, *a, **k):
        _standalone("MLP4Confidence")
