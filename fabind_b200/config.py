"""The published FABind (v1) flags that the docking stack reads at construction/forward time
(reference: the training command rebuilt in FABind/fabind/test_fabind.py:182 and the argparse defaults
of FABind/fabind/main_fabind.py:34-192)."""
from types import SimpleNamespace


def published_args(**over):
    a = SimpleNamespace(
        rm_F_norm=False, norm_type="per_sample", rm_layernorm=True, add_attn_pair_bias=True,
        explicit_pair_embed=True, add_cross_attn_layer=True, keep_trig_attn=False, opm=False,
        fix_pocket=False, rm_LAS_constrained_optim=False, random_n_iter=True, refine="refine_coord",
        ablation_no_attention=False, ablation_no_attention_with_cross_attn=False,
        geometry_reg_step_size=0.001, coordinate_scale=5.0, inter_cutoff=10, intra_cutoff=8,
    )
    for k, v in over.items():
        setattr(a, k, v)
    return a
