// GEMM dispatch: tcgen05 (bf16 operands, TMA-fed, TMEM accumulators) when the shape qualifies,
// otherwise the SIMT kernel.  tcgen05 has three kernels: v3 (persistent, double-buffered TMEM, TMA-store
// epilogue; default), v2 (same main loop, register/transposing epilogue; takes what v3 declines, FB_TC_V=2) and
// v1 (one tile per CTA; FB_TC_V=1) for A/B comparisons.
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

static int tc_version() {
  static int v = [] { const char* e = getenv("FB_TC_V"); return e ? atoi(e) : 3; }();
  return v;
}

int gemm_dot_tiles(int M, int N, int K, bool bf16_mode) {
  if (bf16_mode && gemm_tc_shape_ok(N, K)) {
    if (tc_version() >= 2 && gemm_tc2_shape_ok(N)) return gemm_tc2_dot_tiles(M, N);
    return gemm_tc_dot_tiles(N);
  }
  return gemm_simt_dot_tiles(N);
}

int gemm_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st) {
  if (bf16_mode && gemm_tc_supported(g)) {
    if (tc_version() == 3) {
      // long (edge-level / pair-level) problems: CTA pairs (tcgen05 cta_group::2), a third less operand traffic per MMA
      const int r4 = gemm_tc4_launch(g, st);
      if (r4 != FB_ERR_UNSUPPORTED) return r4;
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

int gemm_launch_pair(const GemmArgs& g0, const GemmArgs& g1, bool bf16_mode, cudaStream_t st) {
  static const bool group = [] { const char* e = getenv("FB_NO_GROUP"); return !(e && atoi(e)); }();
  if (group && bf16_mode && tc_version() >= 2 && gemm_tc_supported(g0) && gemm_tc_supported(g1)) {
    int r = tc_version() == 3 ? gemm_tc3_launch_pair(g0, g1, st) : FB_ERR_UNSUPPORTED;
    if (r == FB_ERR_UNSUPPORTED && g0.drop.p <= 0.f && g1.drop.p <= 0.f) r = gemm_tc2_launch_pair(g0, g1, st);
    if (r != FB_ERR_UNSUPPORTED) return r;
  }
  const int r = gemm_launch(g0, bf16_mode, st);
  return r != FB_OK ? r : gemm_launch(g1, bf16_mode, st);
}

}  // namespace fb
