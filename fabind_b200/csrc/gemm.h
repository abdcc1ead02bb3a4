// Internal GEMM interface:  C[M,N] = epilogue( [A | A2][M,K1+K2] * W[N,K1+K2]^T )
// W is the torch.nn.Linear weight layout (row-major [out, in]), i.e. both operands are K-major.
#pragma once
#include "common.cuh"

namespace fb {

struct GemmArgs {
  // A operand(s): row-major, element type = activation type of the precision mode
  const void* A = nullptr;  int lda = 0;  int K1 = 0;
  const void* A2 = nullptr; int lda2 = 0; int K2 = 0;   // optional second block, concatenated on K
  const void* W = nullptr;                               // [N, K1+K2], same element type as A
  int ldw = 0;                                           // row stride of W (0 = K1 + K2); != K only on the multi-problem kernel (gemm_tc5.cu)
  const float* bias = nullptr;                           // [N] or null
  int act = FB_ACT_NONE;
  const float* res = nullptr; int ldres = 0;             // fp32 residual added after the activation
  float* C = nullptr; int ldc = 0;                       // fp32 output (optional)
  void* Cb = nullptr; int ldcb = 0;                      // activation-typed output (optional)
  // row-dot epilogue: dot_out[nt * dot_stride + m] = sum_{n in N-tile nt} dotv[n] * value(m, n)
  const float* dotv = nullptr; float* dot_out = nullptr; int dot_stride = 0;
  int M = 0, N = 0;
  int n_split = 0;                                       // >0: columns < n_split go to C only, the rest to Cb only (at column n - n_split)
  const int* m_dev = nullptr;                            // optional device-side row count (<= M)
  DropCfg drop;                                          // dropout after the activation, before residual / row-dot / stores
  // split-precision modes (GEMM_SPLIT3 / GEMM_SPLIT6): A / A2 are fp32, W is the split weight [N, 3K] bf16 (three bf16 planes
  // w0 | w1 | w2 with w0 + w1 + w2 == w, made by split_rows); split_ws = scratch for the split A operand, M x 3K bf16
  void* split_ws = nullptr; size_t split_ws_bytes = 0;
  const void* W_f32 = nullptr;                           // the same weight in fp32 [N, K]: shapes tcgen05 cannot take run on the FFMA kernel
  int nprod = 0;                                         // set by the dispatcher: products per k-block (tc3 / tc4 kernels)
  bool exact_act = false;                                // full-precision SiLU in the tcgen05 epilogues (split modes)
  bool w_static = false;                                 // W was not written by the kernels immediately before this launch in the stream:
                                                         // the multi-producer kernel (gemm_tc5.cu) may request it before griddepcontrol.wait
};

// precision modes of gemm_launch.  SPLITn: fp32 operands as sums of bf16 planes (x = x0 + x1 + x2, 8 mantissa bits each), the
// product as n bf16 tensor-core products accumulated in fp32 (TMEM): n = 3 keeps the terms down to 2^-9 (x0 w0, x0 w1, x1 w0),
// n = 6 the terms down to 2^-18 (+ x1 w1, x0 w2, x2 w0) -- fp32-grade accuracy on tcgen05.
enum { GEMM_FP32 = 0, GEMM_BF16 = 1, GEMM_SPLIT3 = 2, GEMM_SPLIT6 = 3 };
inline int gemm_mode_products(int mode) { return mode == GEMM_SPLIT3 ? 3 : mode == GEMM_SPLIT6 ? 6 : 0; }

// number of N tiles (= number of row-dot partials per row) the kernel chosen for `bf16` will use
int gemm_dot_tiles(int M, int N, int K, int mode);

// returns FB_OK or an error code; never synchronises
int gemm_launch(const GemmArgs& g, int mode, cudaStream_t st);

// two independent problems over disjoint row ranges of one activation buffer (g1.A = g0.A + r * lda, same K):
// one grouped launch when the tcgen05 path can take it, otherwise two launches
int gemm_launch_pair(const GemmArgs& g0, const GemmArgs& g1, int mode, cudaStream_t st);

// up to four INDEPENDENT problems (no problem reads what another one writes) as one multi-problem tcgen05 launch (gemm_tc5.cu) when
// the mode is bf16 and every problem qualifies, otherwise one after the other in the given order.  prefetch_w: the weights were NOT
// written by the kernels immediately before this launch in the stream (their first slabs are requested before griddepcontrol.wait)
int gemm_launch_multi(const GemmArgs* g, int np, int mode, bool prefetch_w, cudaStream_t st);

// fp32 rows -> three bf16 planes: dst[m, p * K + k] (p = 0..2), K = K1 + K2 ([A | A2] concatenated); K1, K2 multiples of 8
int split_rows(const float* A, int lda, int K1, const float* A2, int lda2, int K2, int M, void* dst, cudaStream_t st);

// implemented per backend
int gemm_simt_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st);
int gemm_simt_dot_tiles(int N);
bool gemm_tc_supported(const GemmArgs& g);
bool gemm_tc_shape_ok(int N, int K);
int gemm_tc_launch(const GemmArgs& g, cudaStream_t st);
int gemm_tc_dot_tiles(int N);

}  // namespace fb
