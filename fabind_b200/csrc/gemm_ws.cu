// Weight-stationary persistent tcgen05 GEMM (v3) for the long edge-level GEMMs (M >> N, K <= 512):
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// At 128x256 tiles the streaming kernel (gemm_tc2.cu) needs 48 KB of operands per 0.28 us of MMA time and SM,
// ~26 TB/s chip-wide, which the L2->SM fabric cannot deliver.  Here every CTA keeps ITS 128-column slice of W
// (all K: 128 x 512 bf16 = 128 KB) resident in shared memory for the whole kernel and only streams A tiles
// (16 KB per 64-wide K slab): operand traffic drops from 393 KB to 128 KB per 128x128 output tile, and the four
// CTAs that own the four column slices of the same row block run side by side, so three of the four reads of an
// A tile hit L2.
//
//   warp 0      TMA producer: the W slice once (8 slabs, one mbarrier), then A slabs into a 4-deep ring
//   warp 1      MMA issuer: tcgen05.mma 128x128x16, two TMEM accumulator stages
//   warps 2-9   epilogue (as in gemm_tc2.cu): bias / SiLU / ReLU / fused Linear(H,1) row-dot, XOR-swizzled 32x32
//               transpose through shared memory, row-contiguous stores
// Grid: ns * floor(SMs / ns) CTAs (ns = N / 128); CTA i owns column slice i % ns and row blocks i/ns, i/ns + G, ...
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace ws {
using namespace tc;

constexpr int BN = 128, STAGES = 4, KB_MAX = 8;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;   // 320

struct Params {
  int M, N, KB;
  const int* m_dev;
  const float* bias; int act;
  float* C; int ldc;
  bf16* Cb; int ldcb;
  const float* dotv; float* dot_out; int dot_stride;
};

struct Smem {
  static constexpr int W_BYTES = KB_MAX * BN * BK * 2;          // 131072: resident weight slice
  static constexpr int A_BYTES = BM * BK * 2;                   // 16384 per stage
  static constexpr int A_OFF = W_BYTES;
  static constexpr int XPOSE_OFF = A_OFF + STAGES * A_BYTES;    // EPI_WARPS x [32][32] floats (XOR swizzle)
  static constexpr int VEC_OFF = XPOSE_OFF + EPI_WARPS * 32 * 32 * 4;   // bias[128] | dot[128]
  static constexpr int BAR_OFF = VEC_OFF + 2 * BN * 4;          // wfull, full[S], empty[S], tfull[2], tempty[2], slot
  static constexpr int TOTAL = BAR_OFF + (1 + 2 * STAGES + 4) * 8 + 16 + 1024;
};
static_assert(Smem::TOTAL <= 232448, "shared memory budget");

__global__ void __launch_bounds__(THREADS, 1) gemm_ws_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_w, Params p, int ns) {
  using S = Smem;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* wfull = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* full = wfull + 1;
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  float* s_bias = (float*)(smem + S::VEC_OFF);
  float* s_dot = s_bias + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % ns, group = blockIdx.x / ns, n_groups = gridDim.x / ns;
  const int n0 = slice * BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(wfull, 1);
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int t = threadIdx.x; t < BN; t += THREADS) {
    s_bias[t] = p.bias ? p.bias[n0 + t] : 0.f;
    s_dot[t] = p.dotv ? p.dotv[n0 + t] : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // weights may be fetched before the previous kernel has finished; activations and the device row count may not
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(wfull, (uint32_t)(p.KB * BN * BK * 2));
    for (int kb = 0; kb < p.KB; ++kb) tma_load_2d(&map_w, wfull, smem + kb * (BN * BK * 2), kb * BK, n0);
  }
  pdl_wait();
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int m_tiles = (M + BM - 1) / BM;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int mt = group; mt < m_tiles; mt += n_groups) {
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full[s], S::A_BYTES);
          tma_load_2d(&map_a, &full[s], smem + S::A_OFF + s * S::A_BYTES, kb * BK, mt * BM);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(wfull, 0);
      int it = 0, lt = 0;
      for (int mt = group; mt < m_tiles; mt += n_groups, ++lt) {
        const int a = lt & 1;
        mbar_wait(&tempty[a], ((lt >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(smem + S::A_OFF + s * S::A_BYTES);
          const uint64_t bdesc = make_smem_desc(smem + kb * (BN * BK * 2));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[a]);
      }
    }
  } else {
    const int e = warp - 2, q = warp & 3, half = e >> 2;
    float* xp = (float*)(smem + S::XPOSE_OFF) + e * 32 * 32;
    int lt = 0;
    for (int mt = group; mt < m_tiles; mt += n_groups, ++lt) {
      const int a = lt & 1, m0 = mt * BM;
      mbar_wait(&tfull[a], (lt >> 1) & 1);
      tcgen05_fence_after();
      const int mrow = m0 + q * 32 + lane;
      float dsum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < BN / 2; cc += 32) {
        const int c = half * (BN / 2) + cc;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + s_bias[c + j];
          if (p.act == FB_ACT_SILU) x = __fdividef(x, 1.0f + __expf(-x));
          else if (p.act == FB_ACT_RELU) x = fmaxf(x, 0.0f);
          o[j] = x;
        }
        if (p.dotv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dsum = fmaf(s_dot[c + j], o[j], dsum);
        }
        if (p.C || p.Cb) {
#pragma unroll
          for (int j = 0; j < 32; ++j) xp[lane * 32 + (j ^ lane)] = o[j];   // XOR swizzle: conflict-free both ways
          __syncwarp();
          const int ncol = n0 + c;
          if (p.C) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int m = m0 + q * 32 + r;
              if (m < M) p.C[(size_t)m * p.ldc + ncol + lane] = xp[r * 32 + (lane ^ r)];
            }
          }
          if (p.Cb) {
            const int rr = lane >> 4, cp = (lane & 15) * 2;
#pragma unroll 8
            for (int r = 0; r < 32; r += 2) {
              const int row = r + rr, m = m0 + q * 32 + row;
              if (m < M) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(xp[row * 32 + (cp ^ row)], xp[row * 32 + ((cp + 1) ^ row)]);
                *reinterpret_cast<__nv_bfloat162*>(p.Cb + (size_t)m * p.ldcb + ncol + cp) = t;
              }
            }
          }
          __syncwarp();
        }
      }
      if (p.dotv && mrow < M) p.dot_out[(size_t)(slice * 2 + half) * p.dot_stride + mrow] = dsum;
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[a])) : "memory");
    }
  }
  // the weight loads must have landed before the CTA may exit (a CTA without row blocks never waits on them otherwise)
  if (warp == 0 && lane == 0) mbar_wait(wfull, 0);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

}  // namespace ws

bool gemm_ws_supported(const GemmArgs& g) {
  static int on = [] { const char* e = getenv("FB_WS"); return e ? atoi(e) : 1; }();
  if (!on) return false;
  const int K = g.K1 + g.K2;
  if (g.K2 > 0 || g.A2 || g.res || g.n_split > 0) return false;
  if (K % 64 || K > 64 * ws::KB_MAX || K < 64) return false;
  if (g.N % 128 || g.N / 128 > 8 || g.N / 128 < 2) return false;
  if (g.M < 16384) return false;
  return true;
}
int gemm_ws_dot_tiles(int N) { return 2 * (N / 128); }

int gemm_ws_launch(const GemmArgs& g, cudaStream_t st) {
  using namespace ws;
  static bool attr_set = false;
  static int num_sms = 0;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::TOTAL) != cudaSuccess) return FB_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const int ns = g.N / BN;
  CUtensorMap ma, mw;
  if (!tc_make_map(&ma, g.A, (uint64_t)g.M, (uint64_t)g.K1, (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (!tc_make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)g.K1, (uint64_t)g.K1, BN)) return FB_ERR_CUDA;
  Params p;
  p.M = g.M; p.N = g.N; p.KB = g.K1 / BK; p.m_dev = g.m_dev; p.bias = g.bias; p.act = g.act;
  p.C = g.C; p.ldc = g.ldc; p.Cb = (bf16*)g.Cb; p.ldcb = g.ldcb;
  p.dotv = g.dotv; p.dot_out = g.dot_out; p.dot_stride = g.dot_stride;
  const int m_tiles = (g.M + BM - 1) / BM;
  int groups = num_sms / ns;
  if (groups > m_tiles) groups = m_tiles;
  fb_launch(gemm_ws_kernel, dim3(groups * ns), dim3(THREADS), Smem::TOTAL, st, ma, mw, p, ns);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
