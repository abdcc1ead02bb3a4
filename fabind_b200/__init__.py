"""fabind_b200: B200-native (sm_100a) implementation of FABind's iterative docking stack.

Drop-in for the reference's `models.att_model` / `models.egnn` / `models.cross_att` / `models.model_utils`
on the hot path (same class names, signatures and state_dict keys); the arithmetic is hand-written
CUDA behind the C ABI in include/fabind_b200.h.
"""
from .att_model import EfficientMCAttModel, ComplexGraph  # noqa: F401
from .egnn import MC_E_GCL, MC_Att_L, MCAttEGNN  # noqa: F401
from .cross_att import CrossAttentionModule, RowAttentionBlock  # noqa: F401
from .model_utils import Attention, Transition, InteractionModule  # noqa: F401
