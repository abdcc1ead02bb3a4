/* fabind_b200 -- C ABI of the B200 (sm_100a) implementation of FABind's iterative docking stack.
 *
 * The reference (QizhiPei/FABind) has no FFI/plugin interface: the path sits behind torch.nn.Module
 * classes.  Each entry point below names the reference method(s) it replaces; the Python classes in
 * fabind_b200/ (same names, constructor/forward signatures and state_dict keys as the reference)
 * bind them through ctypes.  INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only (no torch types); every device pointer is owned by the
 * caller (PyTorch's caching allocator in practice); nothing here allocates device memory, synchronises
 * the device or throws; work is enqueued on `stream` (a cudaStream_t passed as void*); the return value
 * is 0 on success or a negative FB_ERR_* code.  Functions marked [host] touch no GPU state and work
 * without a device.
 */
#ifndef FABIND_B200_H
#define FABIND_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 6

/* One batch of complexes + model dimensions.  Node layout in caller order is the reference
 * dataloader's [glb_c | atoms | glb_p | residues] per complex (utils/utils.py:328-335). */
typedef struct fb_model_params {
  /* ---- dimensions ---- */
  int32_t N;          /* nodes in the batch */
  int32_t B;          /* complexes */
  int32_t Nc_tot;     /* compound-side nodes (glb_c + atoms) over the batch */
  int32_t P_total;    /* sum over complexes of (n_p+1)*(n_c+1) dense pair rows */
  int32_t hidden;     /* hidden_size == embed_size (512 docking stack, 128 pocket stage) */
  int32_t n_layers;   /* MCAttEGNN depth (mean_layers) */
  int32_t n_iter;     /* refinement iterations */
  int32_t n_bond;     /* columns of compound_edge_index */
  int32_t n_las;      /* columns of LAS_edge_index */
  int32_t E_ctx;      /* context edges (bonds + geometric); read back after fb_graph_static */
  int32_t cap_int;    /* capacity of the interface edge list: 2 * sum n_c*n_p (worst case) */
  int32_t bf16_mode;  /* precision mode, FB_PREC_*: 0 fp32 (FFMA GEMMs); 1 bf16 operands, fp32 accumulate (tcgen05);
                       * 2 / 3 (ABI 4) fp32 activations with the GEMMs on tcgen05 as 3 / 6 bf16 products per term
                       * (x = x0 + x1 + x2 in bf16 planes, fp32 accumulation in TMEM): 3 = fp32-grade accuracy ("fp32_tc") */
  int32_t max_c, max_p;    /* largest compound-side / protein-side node count of any complex (launch bounds) */
  int32_t fb_atom, fb_res; /* internal ids of the first ligand atom / first residue of complex 0 (att_model.py:85-86) */
  /* ---- geometry constants, already divided by coordinate_scale ---- */
  float intra_cutoff, inter_cutoff;   /* att_model.py:34-35 */
  float coord_clamp;                  /* normalize_coord(10), egnn.py:378 */
  float las_clamp, las_step;          /* normalize_coord(15), geometry_reg_step_size, egnn.py:447-448 */
  /* ---- inputs, caller node order (device) ---- */
  const float* X_in;        /* [N,3] */
  const float* H_in;        /* [N,hidden] */
  const float* X_las;       /* [N,3] batched_complex_coord_LAS */
  const int64_t* bonds;     /* [2,n_bond] compound_edge_index */
  const int64_t* las;       /* [2,n_las]  LAS_edge_index */
  /* ---- layout, host-built by the binding (device, int32 unless noted) ---- */
  const int32_t* perm;      /* [N] internal -> caller */
  const int32_t* inv;       /* [N] caller -> internal */
  const int32_t* node_cplx; /* [N] complex of each internal node */
  const uint8_t* node_flags;/* [N] bit0 protein side, bit1 global, bit2 updated between iterations (mask) */
  const int32_t* c_off;     /* [B+1] */
  const int32_t* p_off;     /* [B+1] */
  const int32_t* pair_base; /* [B+1] */
  /* ---- weights: flat arenas laid out by fb_weight_slot_* ---- */
  const float* w32;
  const void* w16;          /* FB_PREC_BF16: bf16 copy of the same arena; FB_PREC_SPLIT3/6: the split arena, 3 x the elements:
                             * slot (rows, cols, off) holds [rows, 3*cols] bf16 at element 3*off = fb_split_rows of the fp32 slot */
  /* ---- scratch ---- */
  void* ws_graph; size_t ws_graph_bytes;   /* >= fb_graph_workspace_bytes */
  void* ws_main;  size_t ws_main_bytes;    /* >= fb_model_workspace_bytes */
  /* ---- outputs, caller node order (device) ---- */
  float* X_out;             /* [N,3]  (may alias X_in: the reference updates X in place, att_model.py:236) */
  float* H_out;             /* [N,hidden] */
  int32_t* stats;           /* optional [n_iter]: interface edges per iteration */
  /* optional debug taps (INTERNAL node order): h and x after every gcl_i / att_i of the LAST iteration,
   * slot 2*i = gcl_i, 2*i+1 = att_i (before the LAS step); [2*n_layers, N, hidden] and [2*n_layers, N, 3] */
  float* trace_h; float* trace_x;
  /* ---- ABI 2: weight layout of the stack ---- */
  int32_t flavour;          /* FB_FLAVOUR_V1 (FABind, models/egnn.py) or FB_FLAVOUR_PLUS (FABind+: LayerNorm MLPs,
                             * pair embedding propagated layer to layer, FABind_plus/fabind/models/{egnn,cross_att,model_utils}.py) */
  float* pair_out;          /* FB_FLAVOUR_PLUS, optional: pair embedding after the last layer of the last iteration in the
                             * reference's dense layout [B, max_p, max_c, hidden] (P/models/att_model.py:223); zero-filled by the caller */
  /* ---- ABI 3: dropout (FABind+ sampling mode runs the model in train() mode, P/test_sampling_fabind.py:118-124) ----
   * dropout_p > 0 applies a mask at every nn.Dropout site of the FABind+ stack (P/models/model_utils.py:26,49-50,70;
   * P/models/egnn.py:204,365,428; P/models/cross_att.py:82), keep = hash(seed + iteration, site, row, col) >= p * 2^32, kept
   * values scaled by 1/(1-p).  dropout_colonly = 1 drops whole feature columns (row ignored): test mode that pins the
   * placement of every mask against the unmodified reference.  FB_FLAVOUR_PLUS only. */
  float dropout_p; uint32_t dropout_seed; int32_t dropout_colonly;
  /* ABI 4: attention core of the RowAttentionBlocks in bf16 mode: 0 = SIMT kernel (layers.cu::row_attention_kernel, the faster one at
   * PDBbind block sizes: 22 / 20 us per launch at B = 16), 1 = tcgen05 kernel (xatt_tc.cu: TMA-fed Q / K tiles, S and O in TMEM; 37 / 62 us,
   * profiles/r2e_xatt_*) whenever the per-complex key lists fit its tiles (<= 256 keys) */
  int32_t attn_tc;
  /* ABI 4: dropout is also served for FB_FLAVOUR_V1 -- the training-mode forward of the reference (models/egnn.py:82,106,236,
   * 398,461; models/cross_att.py:128), same sites / hash as the FABind+ stack where the layouts share them. */
  /* ABI 5: moving rows.  Non-final iterations discard H and keep only the coordinates of the masked nodes (node_flags bit2;
   * att_model.py:232-236, model.py:261-262: ligand atoms + both global nodes), so their out_layer MC_E_GCL is evaluated on the
   * context edges INTO those rows only.  n_mv = number of masked nodes (host-known; 0 switches the subset off); fb_graph_static
   * compacts their rows, and E_ctx_mv = the number of such edges, read back next to E_ctx (fb_graph_counts_ptr), sizes the list. */
  int32_t n_mv;
  int32_t E_ctx_mv;
  /* ABI 6: edge counts supplied by the HOST (a dataloader-side layout, fabind_b200/dataloader.py: the context graph depends only on
   * the bond list and the protein coordinates, which the collate step knows; reference utils/utils.py:202-442 + att_model.py:38-116).
   * layout_flag != NULL switches the read-back of fb_graph_counts_ptr off: E_ctx / E_ctx_mv are then EXPECTATIONS set before
   * fb_graph_static, which compares them with the device-side counts in a one-block kernel; on a mismatch it sets *layout_flag = 1
   * (device int32, never cleared by the library) and clamps the row pointers to the expected sizes, fb_model_forward zero-fills the
   * edge lists before writing them and bounds every write by the expected sizes -- a wrong hint gives flagged garbage, never an
   * out-of-bounds access.  NULL = round-1 protocol (host reads the counts after fb_graph_static). */
  int32_t* layout_flag;
} fb_model_params;
#define FB_FLAVOUR_V1 0
#define FB_FLAVOUR_PLUS 1
#define FB_PREC_FP32 0
#define FB_PREC_BF16 1
#define FB_PREC_SPLIT3 2
#define FB_PREC_SPLIT6 3

/* [host] library identification.  fb_source_hash: 63-bit digest of include/fabind_b200.h + every file under csrc/ at build
 * time (fabind_b200/build.py passes it as -DFB_SOURCE_HASH); the binding recomputes it from the sources next to it and refuses
 * a library built from different ones -- a stale .so with shifted arguments must not load. */
int32_t fb_abi_version(void);
int64_t fb_source_hash(void);

/* Instrumentation for bench.py.  fb_launch_count: kernels launched by this library since load.
 * fb_prof_enable(1): every stage of fb_model_forward is bracketed by CUDA events on its stream, tagged
 * with a category (0 edge GEMMs, 1 node GEMMs, 2 pair-path GEMM, 3 pair0 GEMMs, 4 edge elementwise /
 * segment reduce, 5 attention kernels, 6 graph + geometry).  fb_prof_read (after a stream sync) returns
 * the summed milliseconds and span count per category and clears the record. */
int64_t fb_launch_count(void);
int32_t fb_prof_enable(int32_t on);
int32_t fb_prof_read(double* ms, int64_t* spans, int32_t n_cat);
/* FLOPs (2 M N K, as launched) of the GEMMs enqueued per category since the last call while profiling was on; M = the row
 * capacity for problems whose row count lives on the device (the pair-path GEMM) */
int32_t fb_prof_flops(double* flops, int32_t n_cat);

/* [host] weight arena layout for (hidden, n_layers): slot i has a name ("gcl0.e2_w", "att1.qk_w", ...),
 * a [rows, cols] shape and an element offset into the arena.  fabind_b200/weights.py maps every slot to
 * the reference state_dict keys it is derived from. */
int32_t fb_weight_slot_count(int32_t hidden, int32_t n_layers);
int32_t fb_weight_slot_info(int32_t hidden, int32_t n_layers, int32_t i, char* name, int32_t name_cap,
                            int64_t* rows, int64_t* cols, int64_t* offset);
int64_t fb_weight_arena_elems(int32_t hidden, int32_t n_layers);
/* same for a given flavour (the three functions above describe FB_FLAVOUR_V1) */
int32_t fb_weight_slot_count_f(int32_t hidden, int32_t n_layers, int32_t flavour);
int32_t fb_weight_slot_info_f(int32_t hidden, int32_t n_layers, int32_t flavour, int32_t i, char* name, int32_t name_cap,
                              int64_t* rows, int64_t* cols, int64_t* offset);
int64_t fb_weight_arena_elems_f(int32_t hidden, int32_t n_layers, int32_t flavour);
/* ABI 5: DERIVED slots (names "att<l>.f_*").  Consecutive Linear maps of the cross-attention block (models/cross_att.py:24-54:
 * linear_o -> linear_k/linear_v of the other side -> transition.linear_1; node_mlp.2 of the preceding MC_E_GCL -> the block's first
 * projections; transition.linear_2 -> linear_q / linear_kv) are exact compositions W_b(W_a x + b_a) + b_b = (W_b W_a) x + (W_b b_a +
 * b_b); the library keeps the pre-multiplied matrices in the arena so that those dependent launches become independent problems of
 * one launch (fb_gemm_multi).  fb_derive_weights fills them from the base slots of the fp32 arena `w32` (device memory, on `stream`,
 * fp64 accumulation); call it after every (re)pack and BEFORE converting the arena to its bf16 / split copies.  A packer leaves
 * these slots untouched (fabind_b200/weights.py lists them only on request).  FB_FLAVOUR_PLUS has none (no-op). */
int32_t fb_derive_weights(float* w32, int32_t hidden, int32_t n_layers, int32_t flavour, void* stream);

/* [host] scratch sizes */
/* ABI 5: device pointer to int32[2] = {E_ctx, E_ctx_mv}, valid after fb_graph_static and a stream sync (one read for both) */
const int32_t* fb_graph_counts_ptr(const fb_model_params* p);
int64_t fb_graph_workspace_bytes(const fb_model_params* p);
int64_t fb_model_workspace_bytes(const fb_model_params* p);

/* Static part of the graph: converts bond/LAS lists to the internal order, builds the LAS CSR and counts
 * the context edges.  After it (and a stream sync) the int at fb_graph_ctx_count_ptr(p) holds E_ctx.
 * Replaces the context half of ComplexGraph.construct_edges (att_model.py:38-116). */
int32_t fb_graph_static(const fb_model_params* p, void* stream);
const int32_t* fb_graph_ctx_count_ptr(const fb_model_params* p);

/* Full EfficientMCAttModel.forward (att_model.py:170-246; eval mode, refine='refine_coord'):
 * pair_embed0, n_iter x {inter-edge rebuild, MCAttEGNN.forward (egnn.py:392-466)}, X[mask]=Z[mask]. */
int32_t fb_model_forward(const fb_model_params* p, void* stream);

/* One MCAttEGNN pass -- or one of its sub-layers -- on caller-supplied graphs: the entry behind the stand-alone
 * `MCAttEGNN.forward` (egnn.py:392-466), `MC_E_GCL.forward` (:130-144) and `MC_Att_L.forward` (:308-333).
 * Graph arrays are device int32 in the INTERNAL node numbering, sorted by row (CSR); `pair0` is the dense pair
 * embedding re-packed to [P_total, hidden] rows (prot_local * nc1 + comp_local inside each complex). */
#define FB_STEP_LINEAR_IN  1
#define FB_STEP_GCL        2   /* gcl_i of every layer */
#define FB_STEP_ATT        4   /* att_i of every layer */
#define FB_STEP_LAS        8   /* LAS step after every att_i */
#define FB_STEP_OUT_LAYER  16
#define FB_STEP_LINEAR_OUT 32
typedef struct fb_egnn_extra {
  int32_t steps;                 /* FB_STEP_* mask */
  int32_t E_int;                 /* interface edges supplied */
  const int32_t* ctx_rowptr;     /* [N+1] */
  const int32_t* ctx_row; const int32_t* ctx_col;           /* [E_ctx] */
  const int32_t* int_rowptr;     /* [N+1] */
  const int32_t* int_row; const int32_t* int_col; const int32_t* int_pair;   /* [E_int] */
  const float* pair0;            /* [P_total, hidden] fp32 or NULL when FB_STEP_ATT is off */
  float* att_out;                /* optional [n_layers, E_int]: attention weights in the supplied edge order */
} fb_egnn_extra;
int32_t fb_egnn_forward(const fb_model_params* p, const fb_egnn_extra* e, void* stream);

/* ComplexGraph.construct_edges in the caller's node order, reference edge order (API parity).
 * Two phases around one host read of counts[0..4] = {E_pp, E_glb_normal, E_glb_glb, E_inter, fallback}. */
int32_t fb_edges_ref_count(int32_t N, const int32_t* cplx, const int32_t* off, const uint8_t* flags,
                           const float* x, float intra, float inter, int32_t* ws /*[4N + 4(N+1) + 8]*/,
                           void* stream);
int32_t fb_edges_ref_fill(int32_t N, const int32_t* cplx, const int32_t* off, const uint8_t* flags,
                          const float* x, float intra, float inter, int32_t* ws, const int32_t* counts_host,
                          int64_t* ctx_out /*[2,E_ctx]*/, int64_t* inter_out /*[2,E_int]*/, void* stream);

/* Attention core of a RowAttentionBlock (cross_att.py:118-134; model_utils.py:21-38,96-133) on tcgen05, the kernel the stack uses in
 * bf16 mode: O[q, h*32+d] = sigmoid(G) softmax_j(q_h . k_jh / sqrt(32) + PB[pair(q, j), h]) v_jh per complex, 4 heads x 32 channels.
 * QG / KV: bf16 outputs of the stacked projections of the query side / key side (q_rows / k_rows rows, side-local row index =
 * internal node id minus Nc_tot on the protein side); qcol / gcol / kcol / vcol: first column of Q, the gate, K, V (multiples of 64);
 * PB [pair rows, 4] gated pair bias (pair row = pair_base[b] + prot_local * nc1 + comp_local); O [N, ldo] bf16, rows = internal node id.
 * Keys per complex <= 256 (returns FB_ERR_UNSUPPORTED otherwise: the stack then runs its SIMT attention kernel). */
int32_t fb_row_attention_tc(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t Nc_tot,
                            int32_t q_is_prot, int32_t max_q, int32_t max_k, const void* QG, int32_t ldqg, int32_t qcol, int32_t gcol,
                            int32_t q_rows, const void* KV, int32_t ldkv, int32_t kcol, int32_t vcol, int32_t k_rows, const float* PB,
                            void* O, int32_t ldo, void* stream);

/* Generic fused linear layer used by the stack (unit-test surface):
 * C = act([A|A2] W^T + bias) (+res).  bf16_mode selects operand type of A/A2/W/Cb. */
typedef struct fb_gemm_params {
  const void* A; int32_t lda; int32_t K1;
  const void* A2; int32_t lda2; int32_t K2;
  const void* W;
  const float* bias; int32_t act;
  const float* res; int32_t ldres;
  float* C; int32_t ldc;
  void* Cb; int32_t ldcb;
  const float* dotv; float* dot_out; int32_t dot_stride;
  int32_t M, N;
  const int32_t* m_dev;
  int32_t bf16_mode;
  int32_t force_simt;   /* 1: never take the tcgen05 path */
  /* ABI 3: dropout after the activation, before residual / row-dot / stores (0 = off); row index = row + drop_row0 */
  float drop_p; uint32_t drop_seed; uint32_t drop_site; int32_t drop_row0; int32_t drop_colonly;
  /* ABI 4: split-precision modes (bf16_mode = FB_PREC_SPLIT3 / 6): A, A2 fp32; W = fb_split_rows of the fp32 weight ([N, 3K] bf16);
   * split_ws = scratch of at least M * 3K * 2 bytes for the split A operand; W_f32 = the fp32 weight, used when the shape does not
   * tile on tcgen05 (optional: without it such shapes return FB_ERR_UNSUPPORTED).  All outputs are fp32 (Cb too, under n_split). */
  void* split_ws; size_t split_ws_bytes; const float* W_f32;
  int32_t n_split;      /* >0: columns < n_split go to C, the rest to Cb at column n - n_split (multiple of 128) */
  /* ABI 6: row stride of W in elements (0 = K1 + K2, a contiguous [N, K] weight).  A strided W -- a K-range of a wider matrix: the
   * row-split weight gradients of the training step, dW = sum over row chunks of dY_c^T X_c -- is served by fb_gemm_multi's tcgen05
   * kernel only (bf16 mode, multiple of 8); every other path returns FB_ERR_UNSUPPORTED rather than ignore it. */
  int32_t ldw;
} fb_gemm_params;
/* fp32 rows -> three bf16 planes, dst[m, p*K + k] (p = 0..2) with src = p0 + p1 + p2 to 2^-27: the operand format of the split modes */
int32_t fb_split_rows(const float* src, int32_t ld, int32_t M, int32_t K, void* dst, void* stream);
int32_t fb_gemm(const fb_gemm_params* g, void* stream);
/* two layers over disjoint row ranges of ONE activation buffer (g1->A = g0->A + r*lda rows, r >= g0->M, same K):
 * one grouped tcgen05 launch when both qualify, otherwise the two launches in order.  This is how the stack runs
 * the compound-side / protein-side linears of a stage (cross_att.py:24-54, model_utils.py:171-175). */
int32_t fb_gemm_pair(const fb_gemm_params* g0, const fb_gemm_params* g1, void* stream);
/* ABI 5: n (1..4) INDEPENDENT layers (no problem reads what another one writes; same precision mode) -- one multi-problem tcgen05
 * launch when the mode is bf16 and every problem tiles (N % 128 == 0, no row-dot / dropout / device-side row count), otherwise the n
 * launches in order.  prefetch_w != 0: the caller guarantees that the weights were not written by the kernels immediately before
 * this launch in the stream (their first slabs are requested before the dependency wait).  This is how the stack runs the folded
 * per-side projection groups of a layer (cross_att.py:24-54 with o_p -> k/v -> transition collapsed into pre-multiplied weights). */
int32_t fb_gemm_multi(const fb_gemm_params* g, int32_t n, int32_t prefetch_w, void* stream);
int32_t fb_gemm_dot_tiles(int32_t M, int32_t N, int32_t K, int32_t bf16_mode, int32_t force_simt);
/* development probe: when non-null, sampled CTAs of the tcgen05 GEMMs write globaltimer stamps; the buffer must hold at least
 * 8192 int64 (slots up to 2048 + 19 * 16 + 15 are written) */
int32_t fb_gemm_set_debug(int64_t* dbg);

/* ---- L2 wrapper ops (reference: models/model.py, IaBNet_mean_and_pocket_prediction_cls_coords_dependent) ---- */
/* out[i,:] = scale * src_{kind[i]}[idx[i],:] (a null source gives zeros): the [glb_c | atoms | glb_p | residues]
 * feature / coordinate assembly of model.py:104-115,205-253 and every boolean-mask row selection */
int32_t fb_assemble_rows(float* out, int32_t M, int32_t D, const uint8_t* kind, const int32_t* idx, const float* s0,
                         const float* s1, const float* s2, const float* s3, float scale, void* stream);
/* torch.nn.LayerNorm (model.py:22,352-353) */
int32_t fb_layernorm(const float* x, int32_t M, int32_t D, const float* gamma, const float* beta, float eps, float* out,
                     void* stream);
/* predicted pocket centre: mode 0 = model.forward eval path (model.py:146-158), mode 1 = model.inference (:423-437) */
int32_t fb_pocket_center(const float* logit, const float* xyz, const int32_t* prot_off, int32_t B, float tau, int32_t hard,
                         int32_t mode, float* centers, void* stream);
/* get_keepNode_tensor (utils/utils.py:147-158) + the "<5 residues -> first 100" rule (model.py:199-202) */
int32_t fb_pocket_mask(const float* xyz, const int32_t* prot_off, int32_t B, const float* centers, float radius, uint8_t* keep,
                       int32_t* less5, void* stream);
/* ligand start pose: lig - mean(lig) + mean(pocket) per complex (model.py:227) */
int32_t fb_ligand_place(const float* lig, const int32_t* comp_off, const float* pocket, const int32_t* pocket_off, int32_t B,
                        float* out, void* stream);
/* distance head (model.py:349-365): outer product operand, then (after fb_gemm with the row-dot epilogue) sigmoid*10 and cdist */
int32_t fb_head_outer(const float* pocket_ln, const float* comp_ln, const int32_t* pocket_off, const int32_t* comp_off,
                      const int32_t* pair_off, int32_t B, int32_t n_pairs, int32_t H, void* Z, int32_t bf16_mode, void* stream);
int32_t fb_head_finish(const float* dot, int32_t tiles, int32_t stride, const float* b2, const float* pocket_xyz,
                       const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off, const int32_t* pair_off,
                       int32_t B, int32_t n_pairs, float scale, float* y_pred, float* y_coords, void* stream);
int32_t fb_pair_dist(const float* pocket_xyz, const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off,
                     const int32_t* pair_off, int32_t B, int32_t n_pairs, float cap, float* out, void* stream);
int32_t fb_dot_finish(const float* dot, int32_t tiles, int32_t stride, int32_t M, const float* bias, float* out, void* stream);

/* ---- FABind+ wrapper ops (reference: FABind_plus/fabind/models/model.py, FABindPlus) ---- */
/* LayerNorm of the rows x[rows[i], :] (rows == NULL: identity) to fp32 or bf16: the A operand of the MLP distance head on
 * pair[:, 1:, 1:] (model.py:379-384) */
int32_t fb_layernorm_rows(const float* x, const int32_t* rows, int32_t M, int32_t D, const float* gamma, const float* beta, float eps,
                          void* out, int32_t out_bf16, void* stream);
/* out[b,:] = sum of rows off[b]..off[b+1]: ligand-atom sum in front of pocket_radius_head (model.py:110-114) */
int32_t fb_segment_sum_rows(const float* src, int32_t D, const int32_t* off, int32_t B, float* out, void* stream);
/* per-complex crop radius from the radius head (relu, buffer rule, floor, optional fixed radius; model.py:223-231), then
 * get_keepNode_tensor + the "<5 -> first 100" rule; radius_pred[b] = relu(radius_raw[b]) */
int32_t fb_pocket_mask_r(const float* xyz, const int32_t* prot_off, int32_t B, const float* centers, const float* radius_raw,
                         float buffer, float min_radius, float fixed_radius /* < 0: not forced */, uint8_t* keep, int32_t* less5,
                         float* radius_pred, void* stream);
/* pocket re-centred on its own mean + that mean = pocket_center_bias (model.py:255-258) */
int32_t fb_center_rows3(const float* xyz, const int32_t* off, int32_t B, float* centered, float* mean, void* stream);
/* out = xyz + sign * shift[segment]: data.coords -= centre (model.py:257), prediction + pocket_center_bias (model.py:684) */
int32_t fb_shift_rows3(const float* xyz, const int32_t* off, int32_t B, int32_t n_rows, const float* shift, float sign, float* out,
                       void* stream);
/* soft pocket centre with gumbel noise: F.gumbel_softmax(log_prob, tau, hard) of the train()-mode forward (model.py:136-137);
 * noise = [n_res, 2] samples of -log(Exp(1)) supplied by the caller (torch's generator) */
int32_t fb_pocket_center_gumbel(const float* logit, const float* noise, const float* xyz, const int32_t* prot_off, int32_t B, float tau,
                                int32_t hard, float* centers, void* stream);
/* fb_head_finish with the value range as a parameter (--dis-map-thres, model.py:385-390) */
int32_t fb_head_finish_cap(const float* dot, int32_t tiles, int32_t stride, const float* b2, const float* pocket_xyz,
                           const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off, const int32_t* pair_off,
                           int32_t B, int32_t n_pairs, float scale, float cap, float* y_pred, float* y_coords, void* stream);

/* ---- ligand post-optimisation (reference: utils/post_optim_utils.py:36-64 `post_optimize_compound_coords`, called per ligand
 * on the CPU by fabind_inference.py:285-316) ----
 * B ligands in one launch (one CTA each, all `epochs` Adam steps inside the kernel).  Atoms of ligand b are rows
 * atom_off[b]..atom_off[b+1] of ref_coords / pred_coords / out_coords ([n,3] fp32).  las_edges = NULL: rigid mode (all pairs
 * constrained, post_optim_utils.py:31); otherwise [2, n_las_total] ligand-LOCAL atom ids, edges of ligand b at
 * las_off[b]..las_off[b+1] (LAS mask + 1.22 A excluded volume, post_optim_utils.py:25-29).  out_loss[b] = loss before the last
 * step, out_rmsd[b] = RMSD of the result to ref_coords (the function's 2nd and 3rd return values). */
int32_t fb_post_optimize(const float* ref_coords, const float* pred_coords, const int32_t* atom_off, int32_t B, int32_t max_atoms,
                         const int32_t* las_edges, const int32_t* las_off, int32_t n_las_total, int32_t epochs, float lr,
                         float* out_coords, float* out_loss, float* out_rmsd, void* stream);

/* ---- reverse-pass primitives of the training path (fp32; BASELINE config 5).  The reference trains through torch autograd
 * (main_fabind.py:380-401: loss.backward() over models/egnn.py, cross_att.py, model_utils.py); these are the launches of the
 * hand-derived reverse pass of this library's formulation (specified and pinned in tests/emulate_backward.py), orchestrated
 * by fabind_b200/backward.py.  Data-gradient GEMMs dX = dY W are fb_gemm on a transposed weight.  `act`: 0 none, 1 SiLU,
 * 2 ReLU.  Scatter directions use fp32 atomics into caller-initialised buffers. ---- */
/* Y = act(Z): re-materialise an activation from the saved pre-activation */
int32_t fb_act_fwd(const float* Z, float* Y, int64_t n, int32_t act, void* stream);
/* ABI 6: fused forms for the training-mode forward (same reference lines as fb_act_fwd / fb_gather_add_rows / fb_rows_update /
 * fb_dropout_apply, models/egnn.py:75-82).  fb_edge_pre_train: Z1[e,:] = Pn[row[e], 0:H] + Pn[col[e], H:2H] + rn[e] w_rad + b1 (the first
 * edge-MLP Linear hoisted per node, pre-activation kept for the reverse pass), A1 = act(Z1) in fp32 and optionally bf16 (A16 may be
 * NULL).  fb_act_drop: Y = drop(act(Z)) in fp32 and optionally bf16, the library's counter-based mask (p = 0: no dropout). */
int32_t fb_edge_pre_train(const float* Pn, const int32_t* row, const int32_t* col, int32_t E, int32_t H, const float* rn,
                          const float* w_rad, const float* b1, float* Z1, float* A1, void* A16, int32_t act, void* stream);
int32_t fb_act_drop(const float* Z, int32_t M, int32_t N, int32_t act, float p, uint32_t seed, uint32_t site, int32_t row0,
                    int32_t colonly, float* Y, void* Y16, void* stream);
/* ABI 6: their reverse-pass twins (autograd's backward of the same reference lines): dZ = drop(dY) * act'(Z) and dZ[m,n] = u[m] v[n]
 * act'(Z[m,n]), each also written as bf16 (dZ16 may be NULL) -- the operand of the data-gradient GEMM that follows. */
int32_t fb_act_bwd_drop(const float* Z, const float* dY, int32_t M, int32_t N, int32_t act, float p, uint32_t seed, uint32_t site,
                        int32_t row0, int32_t colonly, float* dZ, void* dZ16, void* stream);
int32_t fb_outer_act_bwd2(const float* Z, const float* u, const float* v, float* dZ, void* dZ16, int32_t M, int32_t N, int32_t act,
                          void* stream);
/* dZ = dY * act'(Z)  (in place allowed: dZ == dY) */
int32_t fb_act_bwd(const float* Z, const float* dY, float* dZ, int64_t n, int32_t act, void* stream);
/* dZ[m,n] = u[m] v[n] act'(Z[m,n]): reverse of a Linear(H,1) head behind an activation (coord_mlp, egnn.py:54-60) */
int32_t fb_outer_act_bwd(const float* Z, const float* u, const float* v, float* dZ, int32_t M, int32_t N, int32_t act, void* stream);
/* out[n] += sum_m w[m] A[m,n] (w == NULL: column sums): bias gradients and rank-1 (radial) column gradients */
int32_t fb_colsum(const float* A, int32_t lda, int32_t M, int32_t N, const float* w, float* out, void* stream);
/* out[m] = sum_n A[m,n] v[n] */
int32_t fb_rowdot(const float* A, int32_t lda, int32_t M, int32_t N, const float* v, float* out, void* stream);
/* dst[idx[e], :D] += src[e, :D]: reverse of the per-edge gathers of node rows (egnn.py:78 `h[row], h[col]`) */
int32_t fb_scatter_add_rows(const float* src, int32_t lds, const int32_t* idx, int32_t E, int32_t D, float* dst, int32_t ldd,
                            void* stream);
/* dst[e, :D] += src[idx[e], :D]: reverse of unsorted_segment_sum (egnn.py:790-805) */
int32_t fb_gather_add_rows(const float* src, int32_t lds, const int32_t* idx, int32_t E, int32_t D, float* dst, int32_t ldd,
                           void* stream);
/* dW[n,k] += sum_m dY[m,n] X[m,k]: weight gradient of y = x W^T */
int32_t fb_gemm_wgrad(const float* dY, int32_t ldy, const float* X, int32_t ldx, int32_t M, int32_t N, int32_t K, float* dW,
                      int32_t ldw, void* stream);
/* reverse of x_new = x + clamp(sum_e (x[row]-x[col]) s_e / cnt, +-cmax) (egnn.py:85-98 with cnt = max(degree,1); egnn.py:228-233
 * with cnt == NULL): ds[e], dx += ... ; dx holds dx_new on entry, step = the unclamped forward step */
int32_t fb_coord_step_bwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const float* s, const float* step,
                          const float* cnt, float cmax, const float* dx_new, float* dx, float* ds, void* stream);
/* reverse of coord2radial with norm_type per_sample (egnn.py:767-787): nrm[b] = the forward's per-complex norm, drn[e] the
 * gradient of the normalised radial, dot_zeroed[B] scratch (zero on entry), dx accumulated */
int32_t fb_radial_bwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const int32_t* node_cplx, const float* nrm,
                      const float* drn, float* dot_zeroed, float* dx, void* stream);
/* reverse of the LAS constrained step (egnn.py:433-449): acc = the unclamped forward step, dx holds dx_new on entry */
int32_t fb_las_bwd(const float* x, const float* xref, const int32_t* a_idx, const int32_t* b_idx, int32_t E, const float* acc,
                   float step_size, float lcl, const float* dx_new, float* dx, void* stream);
/* -- second group: MC_Att_L reverse (row attention, interfacial attention, pair path); fabind_b200/backward.py::att_backward and
 *    stack_backward_v1 orchestrate them (GPU parity: tests/test_gpu_train_reverse_att.py) -- */
/* out[m] = sum_n A[m,n] B[m,n] */
int32_t fb_rowdot2(const float* A, int32_t lda, const float* B, int32_t ldb, int32_t M, int32_t N, float* out, void* stream);
/* mode 0: A[m,:] *= u[m];  mode 1: A[m,n] += u[m] v[n]  (the rank-1 radial terms of linear_kv, egnn.py:203-205) */
int32_t fb_rows_update(float* A, int32_t lda, int32_t M, int32_t N, const float* u, const float* v, int32_t mode, void* stream);
/* op 0: c = a*b, 1: c = a+b, 2: c += a*b */
int32_t fb_vec_op(const float* a, const float* b, float* c, int64_t n, int32_t op, void* stream);
/* dst[n, m] = bf16(src[m, n]) (m < M; zero for M <= m < Mp), dst row stride Mp: the K-major operands of the weight-gradient GEMM
 * dW = dY^T X on tcgen05 (reduction over the rows) */
int32_t fb_transpose_bf16(const float* src, int32_t ld, int32_t M, int32_t N, void* dst, int32_t Mp, void* stream);
/* ABI 6: the same for a whole weight arena in ONE launch (training step: the transposed bf16 twin of every matrix slot, once per optimizer
 * step; reference: the `weight.t()` views autograd takes in `loss.backward()`, main_fabind.py:380-401).  desc[4 i .. 4 i + 3] = {source
 * element offset, rows, cols, destination element offset} (device int64), tile_begin[0 .. n] (device int32) = running count of 32 x 32
 * tiles, n_tiles = tile_begin[n]; slot i is written as [cols, rows] bf16 at dst + destination offset. */
int32_t fb_transpose_slots_bf16(const float* arena, const int64_t* desc, const int32_t* tile_begin, int32_t n, int32_t n_tiles, void* dst,
                                void* stream);
/* nn.Dropout of the training step (egnn.py:82,106,236,398,461; cross_att.py:128) as a stand-alone op, forward and reverse alike:
 * dst[m,n] = keep(seed, site, row0 + m, n) ? src[m,n] / (1-p) : 0 with the library's counter-based mask (fb_model_params.dropout_*);
 * dst may alias src */
int32_t fb_dropout_apply(const float* src, float* dst, int32_t ld, int32_t M, int32_t N, float p, uint32_t seed, uint32_t site,
                         int32_t row0, int32_t colonly, void* stream);
/* reverse of scatter_softmax over the destination row (egnn.py:221): dlogit = alpha (dalpha - segsum(alpha dalpha)); t_zeroed[N] scratch */
int32_t fb_softmax_seg_bwd(const float* alpha, const float* dalpha, const int32_t* row, int32_t E, float* t_zeroed, float* dlogit,
                           void* stream);
/* reverse of the gated pair bias linear(pair)*sigmoid(linear_g(pair)) (model_utils.py:96-133): raw [P,ld] = per block 4 values + 4 gates */
int32_t fb_pair_bias_gate_bwd(const float* raw, int32_t ld, int64_t P, int32_t nblk, const float* dPB, float* draw, void* stream);
/* reverse of the InteractionModule outer product p_i * c_j per complex (model_utils.py:216-220): dpc rows of both sides */
int32_t fb_pair_outer_bwd(const float* dO, const float* pc, int32_t H, const int32_t* c_off, const int32_t* p_off,
                          const int32_t* pair_base, const int32_t* node_cplx, int32_t p_begin, int32_t n_p_rows, float* dpc,
                          void* stream);
/* reverse of the RowAttentionBlock core (cross_att.py:118-134, model_utils.py:21-38), probabilities recomputed; PB / dPB = [P,4];
 * dK / dV zeroed by the caller (they are accumulated with atomics: query tiles of one complex and head run in different CTAs; up to
 * 2048 keys) */
int32_t fb_row_attention_bwd(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t q_is_prot,
                             int32_t max_q, int32_t max_k, const float* Q, int32_t ldq, const float* G, int32_t ldg, const float* K, int32_t ldk,
                             const float* V, int32_t ldv, const float* PB, const float* dO, int32_t ldo, float* dQ, int32_t lddq,
                             float* dG, int32_t lddg, float* dK, int32_t lddk, float* dV, int32_t lddv, float* dPB, void* stream);
/* -- training-mode forward pieces: the sub-steps whose intermediates the reverse pass consumes (the fused inference kernels do not
 *    keep them).  fabind_b200/backward.py::stack_forward_train_v1 runs the last refinement iteration over these + fb_gemm. -- */
/* coord2radial with norm_type per_sample (egnn.py:767-787): d [E,3], d2 [E], rn [E] = d2 / nrm[complex], nrm [B]; S_zeroed [B] scratch */
int32_t fb_radial_fwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const int32_t* node_cplx, int32_t B,
                      float* S_zeroed, float* d, float* d2, float* rn, float* nrm, void* stream);
/* step = sum / max(cnt,1) (cnt == NULL: sum); x_new = x + clamp(step, +-cmax) (egnn.py:85-98, 228-233; LAS step :446-449) */
int32_t fb_coord_apply(const float* x, const float* sum, const float* cnt, int32_t N, float cmax, float* step, float* x_new, void* stream);
/* scatter_softmax over destination rows in CSR order (egnn.py:221) */
int32_t fb_softmax_seg_fwd(const float* logit, const int32_t* rowptr, int32_t n_rows, float* alpha, void* stream);
/* unclamped LAS step (egnn.py:433-445): acc[j] += step * 4 (|x_i-x_j|^2 - |ref_i-ref_j|^2)(x_i-x_j); acc_zeroed [N,3] */
int32_t fb_las_acc(const float* x, const float* xref, const int32_t* a_idx, const int32_t* b_idx, int32_t E, float step_size,
                   float* acc_zeroed, void* stream);
/* InteractionModule outer-product operand p_i * c_j per complex (model_utils.py:216-220), fp32 rows in pair order */
int32_t fb_pair_outer_fwd(const float* pc, int32_t H, const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base,
                          const int32_t* node_cplx, int32_t p_begin, int32_t n_p_rows, float* outer, void* stream);
/* gated pair bias linear(pair)*sigmoid(linear_g(pair)) (model_utils.py:96-133): raw [P,ld] -> PB [P,nblk,4] */
int32_t fb_pair_bias_gate_fwd(const float* raw, int32_t ld, int64_t P, int32_t nblk, float* PB, void* stream);
/* the inference row-attention kernel (layers.cu::row_attention_kernel, fp32 output) on caller-supplied projections; PB = [P,4] */
int32_t fb_row_attention_fwd(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t q_is_prot,
                             int32_t max_q, int32_t max_k, const float* Q, int32_t ldq, const float* G, int32_t ldg, const float* K,
                             int32_t ldk, const float* V, int32_t ldv, const float* PB, float* O, int32_t ldo, void* stream);
/* -- FABind+ layout: LayerNorm MLPs (FABind_plus/fabind/models/model_utils.py:10-74) and the LayerNorm folded through the node-level
 *    hoisting of the per-edge first Linear.  GPU parity tests gated (FB_EXPERIMENTAL) until run on a B200. -- */
/* torch.nn.LayerNorm reverse with recomputed statistics: dx, and xhat for the gamma gradient (dgamma = colsum(dy*xhat), dbeta = colsum(dy)) */
int32_t fb_layernorm_bwd(const float* x, const float* gamma, const float* dy, int32_t M, int32_t D, float eps, float* dx, float* xhat,
                         void* stream);
/* s1 = sum_f h, s2 = sum_f h^2, s3 = sum_f h w (w, s3 optional): the per-node statistics of the folded LayerNorm */
int32_t fb_row_stats(const float* h, int32_t ld, int32_t M, int32_t D, const float* w, float* s1, float* s2, float* s3, void* stream);
/* dh[m,:] += ds1[m] + 2 h[m,:] ds2[m] + w ds3[m] */
int32_t fb_row_stats_bwd(const float* h, int32_t ld, int32_t M, int32_t D, const float* w, const float* ds1, const float* ds2,
                         const float* ds3, float* dh, int32_t lddh, void* stream);
/* per-edge folded LayerNorm statistics: mu = (A1 + rn a0)/D, ex2 = (A2 + 2 rn A3 + rn^2 a1)/D, var_raw = ex2 - mu^2,
 * rstd = rsqrt(max(var_raw,0) + eps); A3 may be NULL */
int32_t fb_folded_stats_fwd(const float* A1, const float* A2, const float* A3, const float* rn, float a0, float a1, float D, float eps,
                            int32_t E, float* mu, float* var_raw, float* rstd, void* stream);
/* its reverse: dA1, dA2, dA3 (NULL with A3), drn accumulated, da[2] (gradients of a0, a1; NULL when they are constants) accumulated */
int32_t fb_folded_stats_bwd(const float* A3, const float* rn, float a0, float a1, float D, int32_t E, const float* mu, const float* var_raw,
                            const float* rstd, const float* drstd, const float* dmu_in, float* dA1, float* dA2, float* dA3, float* drn,
                            float* da, void* stream);

#ifdef __cplusplus
}
#endif
#endif
