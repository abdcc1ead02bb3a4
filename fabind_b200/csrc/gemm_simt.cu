// SIMT fp32-accumulate GEMM (FFMA): the fp32 parity path, and the small-shape path of bf16 mode.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile, operands transposed into shared memory.
#include "gemm.h"

namespace fb {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256, PAD = 4;

template <typename T>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g) {
  pdl_entry();
  int M = g.M;
  if (g.m_dev) M = min(M, *g.m_dev);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;
  const int K = g.K1 + g.K2, N = g.N;
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row within tile, k offset
  const T* A = reinterpret_cast<const T*>(g.A);
  const T* A2 = reinterpret_cast<const T*>(g.A2);
  const T* W = reinterpret_cast<const T*>(g.W);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int k = k0 + lk;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    const int am = m0 + lr;
    if (am < M && k < K) {
      if (k < g.K1) av = ld4(A + (size_t)am * g.lda + k);
      else av = ld4(A2 + (size_t)am * g.lda2 + (k - g.K1));
    }
    const int wn = n0 + lr;
    if (wn < N && k < K) wv = ld4(W + (size_t)wn * K + k);
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk + 0][lr] = wv.x; Bs[lk + 1][lr] = wv.y; Bs[lk + 2][lr] = wv.z; Bs[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* Cb = reinterpret_cast<T*>(g.Cb);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    float dsum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float v = acc[i][j];
        if (g.bias) v += g.bias[n];
        v = apply_act_rt(v, g.act);
        if (g.drop.p > 0.f) v = drop_apply(v, g.drop, m, n);
        if (g.res) v += g.res[(size_t)m * g.ldres + n];
        if (g.n_split > 0) {
          if (n < g.n_split) { if (g.C) g.C[(size_t)m * g.ldc + n] = v; }
          else if (Cb) Cb[(size_t)m * g.ldcb + (n - g.n_split)] = from_f<T>(v);
        } else {
          if (g.C) g.C[(size_t)m * g.ldc + n] = v;
          if (Cb) Cb[(size_t)m * g.ldcb + n] = from_f<T>(v);
        }
        if (g.dotv) dsum = fmaf(g.dotv[n], v, dsum);
      }
    }
    if (g.dotv) {
      // the 16 threads that share `ty` are 16 consecutive lanes: reduce inside the half-warp
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      if (tx == 0 && m < M) g.dot_out[(size_t)blockIdx.x * g.dot_stride + m] = dsum;
    }
  }
}
}  // namespace

int gemm_simt_dot_tiles(int N) { return (N + BN - 1) / BN; }

int gemm_simt_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return FB_OK;
  const int K = g.K1 + g.K2;
  if (K <= 0 || (g.K1 & 3) || (g.K2 & 3) || (g.lda & 3) || (g.A2 && (g.lda2 & 3))) return FB_ERR_BAD_ARG;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  if (bf16_mode) fb_launch(gemm_simt_kernel<bf16>, dim3(grid), dim3(NT), 0, st, g);
  else fb_launch(gemm_simt_kernel<float>, dim3(grid), dim3(NT), 0, st, g);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
