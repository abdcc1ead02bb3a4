"""ctypes binding of libfabind_b200.so (C ABI declared in include/fabind_b200.h).

The product path has NO fallback: if the shared library is missing or does not export the ABI this
module raises, and every op raises on a negative return code.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfabind_b200.so")
ABI_VERSION = 6

ERRORS = {-1: "bad argument", -2: "workspace too small", -3: "CUDA launch error", -4: "unsupported configuration"}


class ModelParams(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("B", C.c_int32), ("Nc_tot", C.c_int32), ("P_total", C.c_int32),
        ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_iter", C.c_int32), ("n_bond", C.c_int32),
        ("n_las", C.c_int32), ("E_ctx", C.c_int32), ("cap_int", C.c_int32), ("bf16_mode", C.c_int32),
        ("max_c", C.c_int32), ("max_p", C.c_int32), ("fb_atom", C.c_int32), ("fb_res", C.c_int32),
        ("intra_cutoff", C.c_float), ("inter_cutoff", C.c_float), ("coord_clamp", C.c_float),
        ("las_clamp", C.c_float), ("las_step", C.c_float),
        ("X_in", C.c_void_p), ("H_in", C.c_void_p), ("X_las", C.c_void_p), ("bonds", C.c_void_p), ("las", C.c_void_p),
        ("perm", C.c_void_p), ("inv", C.c_void_p), ("node_cplx", C.c_void_p), ("node_flags", C.c_void_p),
        ("c_off", C.c_void_p), ("p_off", C.c_void_p), ("pair_base", C.c_void_p),
        ("w32", C.c_void_p), ("w16", C.c_void_p),
        ("ws_graph", C.c_void_p), ("ws_graph_bytes", C.c_size_t),
        ("ws_main", C.c_void_p), ("ws_main_bytes", C.c_size_t),
        ("X_out", C.c_void_p), ("H_out", C.c_void_p), ("stats", C.c_void_p),
        ("trace_h", C.c_void_p), ("trace_x", C.c_void_p),
        ("flavour", C.c_int32), ("pair_out", C.c_void_p),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint32), ("dropout_colonly", C.c_int32), ("attn_tc", C.c_int32),
        ("n_mv", C.c_int32), ("E_ctx_mv", C.c_int32), ("layout_flag", C.c_void_p),
    ]


class EgnnExtra(C.Structure):
    _fields_ = [
        ("steps", C.c_int32), ("E_int", C.c_int32),
        ("ctx_rowptr", C.c_void_p), ("ctx_row", C.c_void_p), ("ctx_col", C.c_void_p),
        ("int_rowptr", C.c_void_p), ("int_row", C.c_void_p), ("int_col", C.c_void_p), ("int_pair", C.c_void_p),
        ("pair0", C.c_void_p), ("att_out", C.c_void_p),
    ]


FLAVOUR_V1, FLAVOUR_PLUS = 0, 1
STEP_LINEAR_IN, STEP_GCL, STEP_ATT, STEP_LAS, STEP_OUT_LAYER, STEP_LINEAR_OUT = 1, 2, 4, 8, 16, 32


class GemmParams(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_int32), ("K1", C.c_int32),
        ("A2", C.c_void_p), ("lda2", C.c_int32), ("K2", C.c_int32),
        ("W", C.c_void_p), ("bias", C.c_void_p), ("act", C.c_int32),
        ("res", C.c_void_p), ("ldres", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int32),
        ("Cb", C.c_void_p), ("ldcb", C.c_int32),
        ("dotv", C.c_void_p), ("dot_out", C.c_void_p), ("dot_stride", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("m_dev", C.c_void_p),
        ("bf16_mode", C.c_int32), ("force_simt", C.c_int32),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_site", C.c_uint32), ("drop_row0", C.c_int32),
        ("drop_colonly", C.c_int32),
        ("split_ws", C.c_void_p), ("split_ws_bytes", C.c_size_t), ("W_f32", C.c_void_p), ("n_split", C.c_int32), ("ldw", C.c_int32),
    ]


PREC_FP32, PREC_BF16, PREC_SPLIT3, PREC_SPLIT6 = 0, 1, 2, 3
# precision names of the drop-in modules -> FB_PREC_*: "fp32_tc" = fp32 activations, GEMMs on tcgen05 as six bf16 products per
# term (fp32-grade accuracy); "bf16x3" = three products (terms down to 2^-9)
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_SPLIT3, "fp32_tc": PREC_SPLIT6}

EXPORTS = {
    "fb_abi_version": (C.c_int32, []),
    "fb_source_hash": (C.c_int64, []),
    "fb_split_rows": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_launch_count": (C.c_int64, []),
    "fb_prof_enable": (C.c_int32, [C.c_int32]),
    "fb_prof_read": (C.c_int32, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "fb_prof_flops": (C.c_int32, [C.POINTER(C.c_double), C.c_int32]),
    "fb_weight_slot_count": (C.c_int32, [C.c_int32, C.c_int32]),
    "fb_weight_slot_info": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fb_weight_arena_elems": (C.c_int64, [C.c_int32, C.c_int32]),
    "fb_weight_slot_count_f": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "fb_weight_slot_info_f": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32,
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fb_weight_arena_elems_f": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "fb_graph_workspace_bytes": (C.c_int64, [C.POINTER(ModelParams)]),
    "fb_model_workspace_bytes": (C.c_int64, [C.POINTER(ModelParams)]),
    "fb_graph_static": (C.c_int32, [C.POINTER(ModelParams), C.c_void_p]),
    "fb_graph_ctx_count_ptr": (C.c_void_p, [C.POINTER(ModelParams)]),
    "fb_model_forward": (C.c_int32, [C.POINTER(ModelParams), C.c_void_p]),
    "fb_egnn_forward": (C.c_int32, [C.POINTER(ModelParams), C.POINTER(EgnnExtra), C.c_void_p]),
    "fb_edges_ref_count": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p]),
    "fb_edges_ref_fill": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_float, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "fb_row_attention_tc": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_gemm": (C.c_int32, [C.POINTER(GemmParams), C.c_void_p]),
    "fb_gemm_pair": (C.c_int32, [C.POINTER(GemmParams), C.POINTER(GemmParams), C.c_void_p]),
    "fb_graph_counts_ptr": (C.c_void_p, [C.POINTER(ModelParams)]),
    "fb_derive_weights": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "fb_gemm_multi": (C.c_int32, [C.POINTER(GemmParams), C.c_int32, C.c_int32, C.c_void_p]),
    "fb_gemm_dot_tiles": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "fb_gemm_set_debug": (C.c_int32, [C.c_void_p]),
    "fb_assemble_rows": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "fb_layernorm": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "fb_pocket_center": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "fb_pocket_mask": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_ligand_place": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_head_outer": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_head_finish": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_pair_dist": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "fb_dot_finish": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_layernorm_rows": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                      C.c_int32, C.c_void_p]),
    "fb_segment_sum_rows": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_pocket_mask_r": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_center_rows3": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_shift_rows3": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "fb_pocket_center_gumbel": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_void_p,
                                            C.c_void_p]),
    "fb_post_optimize": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_act_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "fb_edge_pre_train": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_act_drop": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "fb_act_bwd_drop": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_outer_act_bwd2": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "fb_act_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "fb_outer_act_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "fb_colsum": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_rowdot": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_scatter_add_rows": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_gather_add_rows": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_gemm_wgrad": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int32, C.c_void_p]),
    "fb_coord_step_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_radial_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "fb_las_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_float, C.c_float,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_rowdot2": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_rows_update": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_vec_op": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "fb_transpose_bf16": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_transpose_slots_bf16": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_dropout_apply": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_uint32, C.c_uint32,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "fb_softmax_seg_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_pair_bias_gate_bwd": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_pair_outer_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_row_attention_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_radial_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_coord_apply": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_softmax_seg_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_las_acc": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "fb_pair_outer_fwd": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_pair_bias_gate_fwd": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "fb_row_attention_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_layernorm_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_row_stats": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_row_stats_bwd": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fb_folded_stats_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_folded_stats_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fb_head_finish_cap": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m fabind_b200.build` (nvcc, sm_100a). "
            "fabind_b200 has no CPU or PyTorch fallback.")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(l, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if l.fb_abi_version() != ABI_VERSION:
        raise RuntimeError("libfabind_b200.so ABI version mismatch; rebuild (python -m fabind_b200.build)")
    from .build import source_hash
    want = source_hash()
    if want is not None and l.fb_source_hash() != want:
        raise RuntimeError("libfabind_b200.so was built from different sources than the ones next to it (stale library): "
                           "rebuild with `python -m fabind_b200.build`")
    _lib = l
    return l


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"fabind_b200: {what} failed: {ERRORS.get(rc, rc)}")
