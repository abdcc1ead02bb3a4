// GEMM dispatch: tcgen05 (bf16 operands, TMA-fed, TMEM accumulators) when the shape qualifies,
// otherwise the SIMT kernel.
#include "gemm.h"

namespace fb {

int gemm_dot_tiles(int N, int K, bool bf16_mode) {
  if (bf16_mode && gemm_tc_shape_ok(N, K)) return gemm_tc_dot_tiles(N);
  return gemm_simt_dot_tiles(N);
}

int gemm_launch(const GemmArgs& g, bool bf16_mode, cudaStream_t st) {
  if (bf16_mode && gemm_tc_supported(g)) return gemm_tc_launch(g, st);
  return gemm_simt_launch(g, bf16_mode, st);
}

}  // namespace fb
