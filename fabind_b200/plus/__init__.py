"""Drop-in replacements for FABind_plus/fabind/models/{model_utils,cross_att,egnn,att_model}.py (FABind+ weight layout:
LayerNorm MLPs, pair embedding propagated layer to layer).  Same class names, constructor signatures and state_dict
keys as the reference; the forward pass is `fb_model_forward` with `flavour = FB_FLAVOUR_PLUS`."""
from .att_model import ComplexGraph, EfficientMCAttModel  # noqa: F401
from .egnn import MC_E_GCL, MC_Att_L, MCAttEGNN  # noqa: F401
from .model import FABindPlus, get_model  # noqa: F401
