// Microbenchmark: L2 -> shared-memory bandwidth of ONE SM (and of all SMs together) through TMA as a function of the bytes kept in
// flight (ring depth x 16 KB boxes of 128 rows x 64 bf16, 128B swizzle -- the operand slabs of the tcgen05 GEMMs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// one thread per CTA drives a ring of `stages` boxes; box i of CTA b = rows [(b * boxes + i) * 128 ...), k-slab (i % kslabs)
// `warps` issuing threads (one per warp), each with its own ring of `stages` slots; a slot holds `per` boxes of 16 KB signalled on one barrier
__global__ void __launch_bounds__(128) bw_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map2, int stages, int per,
                                                 int boxes, int rows_total, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5;
  const int slot_bytes = per * 16384;
  uint64_t* bar = (uint64_t*)(smem + warps * stages * slot_bytes) + w * stages;
  uint8_t* ring = smem + w * stages * slot_bytes;
  if ((threadIdx.x & 31) == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int row_blocks = rows_total / 128;
    const int n = boxes / (per * warps);
    long long t0 = clock64();
    for (int i = 0; i < n + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&bar[s], ((i / stages) - 1) & 1);
      if (i < n) {
        mbar_expect(&bar[s], slot_bytes);
        for (int b = 0; b < per; ++b) {
          const long long id = ((long long)(blockIdx.x * warps + w) * n + i) * per + b;
          const int rb = (int)(id / 8 % row_blocks);
          tma2d((b & 1) ? &map2 : &map, &bar[s], ring + s * slot_bytes + b * 16384, (int)(id % 8) * 64, rb * 128);
        }
      }
    }
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 32768, cols = 512;                       // 32 MB of bf16: L2-resident after the first pass
  void* d; cudaMalloc(&d, (size_t)rows * cols * 2); cudaMemset(d, 1, (size_t)rows * cols * 2);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  if (((EncodeFn)f)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 16384 + 2048);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int boxes = 512;                                    // 8 MB per CTA
  void* d2; cudaMalloc(&d2, (size_t)rows * cols * 2); cudaMemset(d2, 1, (size_t)rows * cols * 2);
  CUtensorMap map2;
  ((EncodeFn)f)(&map2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d2, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  struct Cfg { int ctas, warps, per, stages; };
  const Cfg cfgs[] = {{1, 1, 1, 4}, {1, 1, 2, 4}, {1, 1, 4, 3}, {1, 2, 1, 4}, {1, 2, 2, 3}, {1, 4, 1, 3}, {1, 4, 2, 1},
                      {148, 1, 1, 4}, {148, 1, 2, 4}, {148, 2, 1, 4}, {148, 2, 2, 3}, {148, 4, 1, 3}};
  for (const Cfg& c : cfgs) {
    {
      const int ctas = c.ctas, stages = c.stages;
      float best = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        bw_kernel<<<ctas, 32 * c.warps, c.warps * stages * c.per * 16384 + 2048>>>(map, map2, stages, c.per, boxes, rows, cyc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
      }
      std::vector<long long> h(ctas); cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      const double bytes = (double)boxes * 16384;
      printf("{\"ctas\": %d, \"issuing_warps\": %d, \"boxes_per_barrier\": %d, \"stages\": %d, \"kb_in_flight\": %d, \"bytes_per_clk_per_sm\": %.1f, \"gbs_per_sm_event\": %.1f, \"tbs_total_event\": %.2f}\n",
             ctas, c.warps, c.per, stages, c.warps * stages * c.per * 16, bytes / (double)mx, bytes / (best * 1e-3) / 1e9, bytes * ctas / (best * 1e-3) / 1e12);
    }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  return 0;
}
