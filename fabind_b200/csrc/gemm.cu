// GEMM dispatch.  bf16 mode: tcgen05 (bf16 operands, TMA-fed, TMEM accumulators) when the shape qualifies, otherwise the SIMT
// kernel.  tcgen05 kernels: v4 (CTA pairs, long problems), v3 (persistent, double-buffered TMEM, TMA-store epilogue), v2 (same main
// loop, register/transposing epilogue; takes what v3 declines).  Split modes: fp32 operands on the SAME tcgen05 kernels as sums of
// bf16 planes (split_rows below + the table-driven k-block walk of the v3 / v4 producers), 3 or 6 products per k-block.
// Diagnostic knobs (FB_TC_V, FB_NO_GROUP) exist only in -DFB_DIAG builds.
#include <cstdlib>

#include "gemm.h"

namespace fb {

bool gemm_tc2_shape_ok(int N);
int gemm_tc2_dot_tiles(int M, int N);
int gemm_tc2_launch(const GemmArgs& g, cudaStream_t st);
int gemm_tc2_launch_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t st);
int gemm_tc3_launch(const GemmArgs& g, cudaStream_t st);
int gemm_tc4_launch(const GemmArgs& g, cudaStream_t st);
int gemm_tc3_launch_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t st);
int gemm_tc5_launch(const GemmArgs* g, int np, bool prefetch_w, cudaStream_t st);

static int tc_version() {
#ifdef FB_DIAG
  static int v = [] { const char* e = getenv("FB_TC_V"); return e ? atoi(e) : 3; }();
  return v;
#else
  return 3;
#endif
}

// ---- split-precision operand: x (fp32) -> x0 | x1 | x2 (bf16 planes, x0 + x1 + x2 == x up to 2^-27 |x|) ----------------------
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ A, int lda, int K1, const float* __restrict__ A2,
                                                         int lda2, int K2, int M, bf16* __restrict__ S) {
  pdl_entry();
  const int K = K1 + K2, K8 = K >> 3;
  const long long total = (long long)M * K8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / K8), k = (int)(i - (long long)m * K8) * 8;
    const float* src = k < K1 ? A + (size_t)m * lda + k : A2 + (size_t)m * lda2 + (k - K1);
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    __align__(16) bf16 p0[8], p1[8], p2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bf16 h0 = __float2bfloat16_rn(x[j]);
      const float r1 = x[j] - __bfloat162float(h0);          // exact
      const bf16 h1 = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(h1);            // exact
      p0[j] = h0; p1[j] = h1; p2[j] = __float2bfloat16_rn(r2);
    }
    bf16* d = S + (size_t)m * 3 * K + k;
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(p0);
    *reinterpret_cast<uint4*>(d + K) = *reinterpret_cast<const uint4*>(p1);
    *reinterpret_cast<uint4*>(d + 2 * K) = *reinterpret_cast<const uint4*>(p2);
  }
}

int split_rows(const float* A, int lda, int K1, const float* A2, int lda2, int K2, int M, void* dst, cudaStream_t st) {
  if (M <= 0) return FB_OK;
  if ((K1 % 8) || (K2 % 8) || (lda % 4) || ((uintptr_t)A & 15) || ((uintptr_t)dst & 15)) return FB_ERR_BAD_ARG;
  if (K2 > 0 && ((lda2 % 4) || ((uintptr_t)A2 & 15))) return FB_ERR_BAD_ARG;
  const long long total = (long long)M * ((K1 + K2) >> 3);
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  fb_launch(split_rows_kernel, dim3(grid), dim3(256), 0, st, A, lda, K1, A2, lda2, K2, M, (bf16*)dst);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int gemm_dot_tiles(int M, int N, int K, int mode) {
  if (mode != GEMM_FP32 && gemm_tc_shape_ok(N, K)) {
    if (tc_version() >= 2 && gemm_tc2_shape_ok(N)) return gemm_tc2_dot_tiles(M, N);
    return gemm_tc_dot_tiles(N);
  }
  return gemm_simt_dot_tiles(N);
}

// the operand view the tcgen05 kernels see in a split mode: A = three bf16 planes [M, 3K]
static bool split_view_ok(const GemmArgs& g) {
  const int K = g.K1 + g.K2;
  if (!g.split_ws || g.split_ws_bytes < (size_t)g.M * 3 * K * sizeof(bf16)) return false;
  if ((g.lda % 4) || ((uintptr_t)g.A & 15) || (g.A2 && ((g.lda2 % 4) || ((uintptr_t)g.A2 & 15)))) return false;
  GemmArgs v = g;
  v.A = g.split_ws; v.lda = 3 * K; v.K1 = K; v.A2 = nullptr; v.lda2 = 0; v.K2 = 0;
  if (v.n_split > 0) { v.Cb = nullptr; v.n_split = 0; }
  return (g.K1 % 64) == 0 && (g.K2 % 64) == 0 && gemm_tc_supported(v);
}

static int gemm_split_launch(const GemmArgs& g, int mode, cudaStream_t st) {
  const int K = g.K1 + g.K2;
  int r = split_rows((const float*)g.A, g.lda, g.K1, (const float*)g.A2, g.lda2, g.K2, g.M, g.split_ws, st);
  if (r != FB_OK) return r;
  GemmArgs v = g;
  v.A = g.split_ws; v.lda = 3 * K; v.K1 = K; v.A2 = nullptr; v.lda2 = 0; v.K2 = 0;
  v.nprod = gemm_mode_products(mode); v.exact_act = true;
  auto run = [&](const GemmArgs& a) {
    const int r4 = gemm_tc4_launch(a, st);
    return r4 != FB_ERR_UNSUPPORTED ? r4 : gemm_tc3_launch(a, st);
  };
  if (g.n_split <= 0) return run(v);
  // column-routed outputs are both fp32 in the split modes: two problems over the same split A operand
  GemmArgs lo = v, hi = v;
  lo.N = g.n_split; lo.Cb = nullptr; lo.n_split = 0;
  hi.N = g.N - g.n_split; hi.W = (const bf16*)g.W + (size_t)g.n_split * 3 * K; hi.bias = g.bias ? g.bias + g.n_split : nullptr;
  hi.C = (float*)g.Cb; hi.ldc = g.ldcb; hi.Cb = nullptr; hi.n_split = 0;
  r = run(lo);
  return r != FB_OK ? r : run(hi);
}

int gemm_launch(const GemmArgs& g, int mode, cudaStream_t st) {
  if (g.ldw != 0 && g.ldw != g.K1 + g.K2) return FB_ERR_UNSUPPORTED;   // a strided W is served by the multi-problem kernel only
  if (mode >= GEMM_SPLIT3) {
    // shapes the tcgen05 kernels cannot take (narrow N, K not a multiple of 64, device-side row counts with stored outputs) stay on
    // the FFMA kernel: same fp32 operands, W read from the fp32 arena by the caller's choice of pointer (see Run::mk)
    if (g.M < 1) return FB_OK;
    int r = FB_ERR_UNSUPPORTED;
    if (split_view_ok(g) && !(g.m_dev && (g.C || g.Cb))) r = gemm_split_launch(g, mode, st);
    if (r != FB_ERR_UNSUPPORTED) return r;
    if (!g.W_f32) return FB_ERR_UNSUPPORTED;
    GemmArgs f = g;
    f.W = g.W_f32;
    return gemm_simt_launch(f, false, st);
  }
  const bool bf16_mode = mode == GEMM_BF16;
  if (bf16_mode && gemm_tc_supported(g)) {
    if (tc_version() == 3) {
      // long (edge-level / pair-level) problems: CTA pairs (tcgen05 cta_group::2) -- weight tile stationary in shared memory when
      // K <= 512, both operands streaming otherwise (gemm_tc4.cu)
      const int r4 = gemm_tc4_launch(g, st);
      if (r4 != FB_ERR_UNSUPPORTED) return r4;
      // node-level problems: the multi-producer kernel (three TMA-issuing warps) as a one-problem launch; what it declines (row-dot,
      // dropout, device-side row counts) stays on v3
      const int r5 = gemm_tc5_launch(&g, 1, g.w_static, st);
      if (r5 != FB_ERR_UNSUPPORTED) return r5;
      const int r = gemm_tc3_launch(g, st);
      if (r != FB_ERR_UNSUPPORTED) return r;
    }
    if (g.drop.p > 0.f) return gemm_simt_launch(g, bf16_mode, st);   // only v3 / v4 / SIMT carry the dropout epilogue
    if (tc_version() >= 2 && gemm_tc2_shape_ok(g.N)) return gemm_tc2_launch(g, st);
    if (g.n_split > 0) return gemm_simt_launch(g, bf16_mode, st);   // v1 kernel has no column routing
    return gemm_tc_launch(g, st);
  }
  return gemm_simt_launch(g, bf16_mode, st);
}

int gemm_launch_pair(const GemmArgs& g0, const GemmArgs& g1, int mode, cudaStream_t st) {
#ifdef FB_DIAG
  static const bool group = [] { const char* e = getenv("FB_NO_GROUP"); return !(e && atoi(e)); }();
#else
  const bool group = true;
#endif
  const bool strided_w = (g0.ldw != 0 && g0.ldw != g0.K1 + g0.K2) || (g1.ldw != 0 && g1.ldw != g1.K1 + g1.K2);
  if (group && !strided_w && mode == GEMM_BF16 && tc_version() >= 2 && gemm_tc_supported(g0) && gemm_tc_supported(g1)) {
    int r = FB_ERR_UNSUPPORTED;
    if (tc_version() == 3 && g0.M > 0 && g1.M > 0) {
      const GemmArgs two[2] = {g0, g1};
      r = gemm_tc5_launch(two, 2, g0.w_static && g1.w_static, st);
    }
    if (r == FB_ERR_UNSUPPORTED && tc_version() == 3) r = gemm_tc3_launch_pair(g0, g1, st);
    if (r == FB_ERR_UNSUPPORTED && g0.drop.p <= 0.f && g1.drop.p <= 0.f) r = gemm_tc2_launch_pair(g0, g1, st);
    if (r != FB_ERR_UNSUPPORTED) return r;
  }
  const int r = gemm_launch(g0, mode, st);
  return r != FB_OK ? r : gemm_launch(g1, mode, st);
}

int gemm_launch_multi(const GemmArgs* g, int np, int mode, bool prefetch_w, cudaStream_t st) {
  if (np <= 0) return FB_OK;
#ifdef FB_DIAG
  static const bool multi = [] { const char* e = getenv("FB_NO_MULTI"); return !(e && atoi(e)); }();
#else
  const bool multi = true;
#endif
  if (multi && mode == GEMM_BF16) {
    // problems without rows drop out of the group
    GemmArgs live[4];
    int n = 0;
    bool fits = true;
    for (int i = 0; i < np; ++i) {
      if (g[i].M <= 0) continue;
      if (n == 4) { fits = false; break; }
      live[n++] = g[i];
    }
    if (fits && n > 0) {
      const int r = gemm_tc5_launch(live, n, prefetch_w, st);
      if (r != FB_ERR_UNSUPPORTED) return r;
    }
    if (fits && n == 0) return FB_OK;
  }
  for (int i = 0; i < np; ++i) {
    const int r = gemm_launch(g[i], mode, st);
    if (r != FB_OK) return r;
  }
  return FB_OK;
}

}  // namespace fb
