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
        # L2 wrapper (models/model.py): published command + argparse defaults (main_fabind.py:43-191)
        mean_layers=4, n_iter=8, pocket_pred_layers=1, pocket_pred_n_iter=1, hidden_size=512,
        pocket_pred_hidden_size=128, stage_prob=0.25, use_esm2_feat=True, esm2_concat_raw=False, gs_tau=1.0,
        gs_hard=False, pocket_radius=20.0, local_eval=False, train_pred_pocket_noise=0.0,
        compound_coords_init_mode="pocket_center_rdkit", center_dist_threshold=4.0,
    )
    for k, v in over.items():
        setattr(a, k, v)
    return a


def published_args_plus(**over):
    """FABind+ published training/evaluation flags (FABind_plus/README.md:125-141) on top of the shared ones;
    argparse defaults from FABind_plus/fabind/utils/parsing.py:169-195."""
    a = published_args(mean_layers=5, use_ln_mlp=True, mlp_hidden_scale=1, dropout=0.1, mha_heads=4,
                       rel_dis_pair_bias="no", inter_additional_mlp=False, only_last_LAS=False,
                       # L2 wrapper FABindPlus (models/model.py): README.md:125-141 + parsing.py:106,158-160,196-199
                       use_for_radius_pred="ligand", pocket_radius_buffer=5.0, min_pocket_radius=20.0, force_fix_radius=False,
                       dis_map_thres=15.0, use_clustering=False, confidence_training=False, stack_mlp=False, geom_reg_steps=1,
                       # sampling-based model (README.md:196-211, parsing.py): confidence head + DBSCAN pocket clustering
                       confidence_dropout=0.2, confidence_use_ln_mlp=True, confidence_mlp_hidden_scale=1, dbscan_eps=9.0,
                       dbscan_min_samples=2, choose_cluster_prob=0.5, infer_dropout=True)
    for k, v in over.items():
        setattr(a, k, v)
    return a
