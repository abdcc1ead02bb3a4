// tcgen05 GEMM (placeholder until the kernel lands: reports "unsupported" so dispatch stays on SIMT)
#include "gemm.h"
namespace fb {
bool gemm_tc_shape_ok(int, int) { return false; }
bool gemm_tc_supported(const GemmArgs&) { return false; }
int gemm_tc_launch(const GemmArgs&, cudaStream_t) { return FB_ERR_UNSUPPORTED; }
int gemm_tc_dot_tiles(int) { return 1; }
}  // namespace fb
