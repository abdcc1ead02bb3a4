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
};

// number of N tiles (= number of row-dot partials per row) the kernel chosen for `bf16` will use
int gemm_dot_tiles(int M, int N, int K, bool bf16_mode);

// returns FB_OK or an error code; never synchronises
int gemm_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st);

// two independent problems over disjoint row ranges of one activation buffer (g1.A = g0.A + r * lda, same K):
// one grouped launch when the tcgen05 path can take it, otherwise two launches
int gemm_launch_pair(const GemmArgs& g0, const GemmArgs& g1, bool bf16_mode, cudaStream_t st);

// implemented per backend
int gemm_simt_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st);
int gemm_simt_dot_tiles(int N);
bool gemm_tc_supported(const GemmArgs& g);
bool gemm_tc_shape_ok(int N, int K);
int gemm_tc_launch(const GemmArgs& g, cudaStream_t st);
int gemm_tc_dot_tiles(int N);

}  // namespace fb
