#pragma once
#include "graph.h"

namespace fb {

int permute_in(const GraphDev& g, const float* H_in, const float* X_in, const float* XL_in, int D, float* h32,
               void* hT, bool bf16_mode, float* x, float* xl, cudaStream_t st);
int permute_x(const GraphDev& g, const float* X_in, float* x, cudaStream_t st);
int permute_out_x(const GraphDev& g, const float* x, float* X_out, cudaStream_t st);
int convert_copy(const float* src, size_t n, float* d32, void* dT, bool bf16_mode, cudaStream_t st);
int permute_out_h(const GraphDev& g, const float* h, int D, float* H_out, cudaStream_t st);
int masked_update_x(const GraphDev& g, float* x_state, const float* z, float* x_out_caller, cudaStream_t st);
int radial(const GraphDev& g, const int* rowptr, const int* erow, const int* ecol, const float* x, float* rad,
           float* norm, cudaStream_t st);
int gcl_edge_pre(int E, int H, const int* erow, const int* ecol, const int* node_cplx, const void* P,
                 const float* rad, const float* norm, const float* w_rad, const float* b1, void* A1, bool bf16_mode,
                 cudaStream_t st, const int* emap = nullptr /* rad is indexed by emap[e] (compact edge subsets) */);
int gcl_node(int N, int H, const int* rowptr, const int* ecol, const void* M, const float* dot, int dot_tiles,
             int dot_stride, const float* x, float cmax, void* agg, float* x_out, bool bf16_mode, cudaStream_t st,
             const int* rmap = nullptr /* row j of rowptr is node rmap[j] (compact row subsets; agg must be null) */,
             const GraphDev* heads = nullptr /* layout of the batch: the high-degree first-of-side rows (global nodes) get CTAs of
                                               their own (rows must be node ids in the library's order) */);
int pair_outer(const GraphDev& g, int P_total, int H, const float* pc, void* A0, bool bf16_mode, cudaStream_t st);
int pair_bias_gate(int P_total, int L, const float* raw, int ld_raw, float* PB, cudaStream_t st);
int row_attention(const GraphDev& g, int q_is_prot, int max_q, int max_k, const float* Q, int ldq, const float* G, int ldg,
                  const float* K, int ldk, const float* V, int ldv, const float* PB, void* O, int ldo, bool bf16_mode,
                  cudaStream_t st);
int pair_gather(const GraphDev& g, int cap_u, int H, const void* P0, const float* pc32, int ld32, void* Zg, void* T64,
                bool bf16_mode, cudaStream_t st);
int pair_bias_finish(const GraphDev& g, int cap_u, const float* dot, int tiles, int stride, const float* cst, float* pb_dense,
                     cudaStream_t st);
int inter_attention(const GraphDev& g, int cap_int, int H, const float* QK, int ldqk, const float* Kt, int ldk, const void* V, const void* VC, int ldv, const float* k_r,
                    const float* v_r, const float* ac_u, const float* ac_b, const float* ac_w2, const float* rad,
                    const float* norm, const float* pb_dense, const float* x, float cmax, float* h, void* hT,
                    float* x_out, float* att, float* logit_ws, float* sdot_ws, bool bf16_mode, cudaStream_t st,
                    const float* ac_g = nullptr, const float* ac_r = nullptr, const float* vstat = nullptr, float eps = 1e-5f,
                    DropCfg drop_coord = DropCfg(), DropCfg drop_agg = DropCfg(),
                    // v1: pair bias from the pair GEMM's row-dot partials instead of pb_dense (see inter_logit_kernel)
                    const float* pb_dot = nullptr, int pb_tiles = 0, int pb_stride = 0, const float* pb_cst = nullptr);
// ---- FABind+ layout (plus.cu) ----
int row_stats(const void* x, int ld, int M, int H, const float* w, float* out, bool typed_bf16, cudaStream_t st);
int ln_rows(const void* x1, bool x1_typed, int ld1, int H1, const void* x2, int ld2, int H2, int M, const float* gamma,
            const float* beta, float eps, void* out, int ldo, bool bf16_mode, cudaStream_t st);
int gcl_edge_pre_plus(int E, int H, int Dp, const int* erow, const int* ecol, const int* node_cplx, const void* P,
                      const float* hstat, const float* rad, const float* norm, const float* w_rad, const float* gsum,
                      const float* c0, float eps, void* A1, bool bf16_mode, cudaStream_t st, DropCfg drop = DropCfg(),
                      const int* emap = nullptr /* rad is indexed by emap[e] (compact edge subsets) */);
int dropout_rows(float* x, void* xT, int M, int H, bool bf16_mode, DropCfg drop, cudaStream_t st);
int pair_zin_plus(const GraphDev& g, int P_total, int H, const void* pair, const float* pc32, int ld32, const float* Wo,
                  const float* bo, const float* gamma, const float* beta, float eps, void* Zl, bool bf16_mode, cudaStream_t st);
int pair_bias_all(int P_total, const float* dot, int tiles, int stride, const float* cst, float* pb_dense, cudaStream_t st);
int pair_unpack(const GraphDev& g, int P_total, int H, int max_p, int max_c, const void* pair, float* out, bool bf16_mode, cudaStream_t st);
int las_step(const GraphDev& g, const float* x, const float* xref, float step, float cl, float* x_out, cudaStream_t st);
// ---- cross-attention core on tcgen05 (xatt_tc.cu): bf16 projections in, gated attention output out ----
bool row_attention_tc_supported(int max_q, int max_k);
int row_attention_tc(const GraphDev& g, int q_is_prot, int max_q, int max_k, const void* QG, int ldqg, int qcol, int gcol, int q_rows,
                     const void* KV, int ldkv, int kcol, int vcol, int k_rows, const float* PB, void* O, int ldo, cudaStream_t st);

}  // namespace fb
