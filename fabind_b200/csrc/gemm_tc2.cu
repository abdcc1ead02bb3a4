// Persistent tcgen05 GEMM (v2):  C[M,N] = epilogue([A|A2][M,K] * W[N,K]^T), bf16 operands, fp32 accumulation.
//
// One CTA per SM loops over output tiles (n fastest, so CTAs running side by side share A rows in L2):
//   warp 0      TMA producer: 64-wide K slabs of A (128 rows) and W (BN rows) into a STAGES-deep smem ring
//   warp 1      MMA issuer: tcgen05.mma 128 x BN x 16 into one of TWO TMEM accumulator stages
//   warps 2-9   epilogue: tcgen05.ld of the finished accumulator while the next tile's MMAs run; bias /
//               activation / fused Linear(H,1) row-dot in the "thread = row" layout, then a 32x32 transpose
//               through padded shared memory so that global stores (and residual loads) are row-contiguous.
// The smem ring and its phases run continuously across tiles, so the tensor pipe only waits for TMA at the
// very first tile.
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

extern long long* g_tc_dbg;
bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace tc2 {
using namespace tc;

#define FB_DBG2(slot)                                                                     \
  do {                                                                                   \
    if (p.dbg && (blockIdx.x & 7) == 0) p.dbg[(blockIdx.x >> 3) * 8 + (slot)] = gtime(); \
  } while (0)

constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;   // 320
constexpr int MAX_N = 2304;                     // bias / dot vectors staged for the whole N
constexpr int MAX_N1 = 1024;                    // widest second problem of a grouped launch

struct Params {
  int M, N, KB1, KB2;
  int m_begin;             // first row of this problem in A / res / C / Cb (grouped launches share A)
  const int* m_dev;
  const float* bias; int act;
  const float* res; int ldres;
  float* C; int ldc;
  bf16* Cb; int ldcb;
  const float* dotv; float* dot_out; int dot_stride;
  int n_split;
  long long* dbg;
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;                       // full[S] empty[S] tfull[2] tempty[2] slot
  static constexpr int VEC_OFF = BAR_OFF + (2 * STAGES + 4) * 8 + 16;        // bias[MAX_N] | dotv[MAX_N]
  static constexpr int VEC1_OFF = VEC_OFF + 2 * MAX_N * 4;                   // bias of the second (grouped) problem
  static constexpr int XPOSE_OFF = VEC1_OFF + MAX_N1 * 4;                    // EPI_WARPS x [32][33] floats
  static constexpr int TOTAL = XPOSE_OFF + EPI_WARPS * 32 * 33 * 4 + 1024;   // + alignment slack
};

template <int BN, int STAGES, bool GROUPED>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_a2,
                                                               const __grid_constant__ CUtensorMap map_w,
                                                               const __grid_constant__ CUtensorMap map_w2,
                                                               const Params p, const Params p2) {
  // `q` (p2.M > 0) is an optional SECOND problem sharing A and K with `p` (different rows, weights, outputs):
  // its tiles are appended to p's, so two short GEMMs that do not depend on each other cost one launch.
  using S = Smem<BN, STAGES>;
  pdl_trigger();
  if (threadIdx.x == 0) FB_DBG2(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  float* s_bias = (float*)(smem + S::VEC_OFF);
  float* s_dot = s_bias + MAX_N;
  float* s_bias1 = (float*)(smem + S::VEC1_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.KB1 + p.KB2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
    if (p.KB2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a2) : "memory");
    if (GROUPED) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w2) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int t = threadIdx.x; t < p.N; t += THREADS) {
    s_bias[t] = p.bias ? p.bias[t] : 0.f;
    s_dot[t] = p.dotv ? p.dotv[t] : 0.f;
  }
  if (GROUPED) for (int t = threadIdx.x; t < p2.N; t += THREADS) s_bias1[t] = p2.bias ? p2.bias[t] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only weights and on-chip state; activations (and the device-side row count)
  // are produced by the previous kernel in the stream
  if (threadIdx.x == 0) FB_DBG2(1);
  pdl_wait();
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int n_tiles_n = p.N / BN;
  const int tiles0 = ((M + BM - 1) / BM) * n_tiles_n;
  const int ntn1 = GROUPED ? p2.N / BN : 1;
  const int n_tiles = tiles0 + (GROUPED ? ((p2.M + BM - 1) / BM) * ntn1 : 0);   // a CTA with no tile falls through
  // tile -> (problem, first row, first column)
  auto decode = [&](int tile, int& m0, int& n0) -> bool {
    if (!GROUPED || tile < tiles0) { m0 = (tile / n_tiles_n) * BM; n0 = (tile % n_tiles_n) * BN; return false; }
    const int t = tile - tiles0;
    m0 = p2.m_begin + (t / ntn1) * BM; n0 = (t % ntn1) * BN;
    return true;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int m0, n0;
        const bool second = decode(tile, m0, n0);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          mbar_expect_tx(&full[s], S::STAGE_BYTES);
          if (kb < p.KB1) tma_load_2d(&map_a, &full[s], a_dst, kb * BK, m0);
          else tma_load_2d(&map_a2, &full[s], a_dst, (kb - p.KB1) * BK, m0);
          tma_load_2d((GROUPED && second) ? &map_w2 : &map_w, &full[s], b_dst, kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const int a = lt & 1;
        mbar_wait(&tempty[a], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          if (it == 0) FB_DBG2(2);
          tcgen05_fence_after();
          const uint8_t* a_src = smem + s * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_src), bdesc = make_smem_desc(a_src + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        if (lt == 0) FB_DBG2(3);
        umma_commit(&tfull[a]);
      }
    }
  } else {
    // ===== epilogue =====
    const int e = warp - 2;                 // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // column half of the tile
    constexpr int COLS = BN / 2;            // columns per warp
    float* xp = (float*)(smem + S::XPOSE_OFF) + e * 32 * 33;
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int a = lt & 1;
      int m0, n0;
      const bool second = decode(tile, m0, n0);
      // per-tile copy of the epilogue description: plain kernel-parameter operands when not grouped
      struct { const float* res; int ldres, act, n_split; float* C; int ldc; bf16* Cb; int ldcb; } pp;
      if (GROUPED && second) pp = {p2.res, p2.ldres, p2.act, p2.n_split, p2.C, p2.ldc, p2.Cb, p2.ldcb};
      else pp = {p.res, p.ldres, p.act, p.n_split, p.C, p.ldc, p.Cb, p.ldcb};
      const float* sb = (GROUPED && second) ? s_bias1 : s_bias;
      const int m_end = (GROUPED && second) ? p2.m_begin + p2.M : M;
      // residual rows do not depend on the MMAs: for the 128-wide tiles (node-level GEMMs) fetch them into
      // registers, in the transposed "lane = column" layout, while the main loop of this tile is running
      constexpr bool PRE = (BN == 128);
      float rpre[PRE ? 2 : 1][PRE ? 32 : 1];
      if (PRE && pp.res) {
#pragma unroll
        for (int ch = 0; ch < (PRE ? 2 : 1); ++ch) {
#pragma unroll
          for (int r = 0; r < (PRE ? 32 : 1); ++r) {
            const int m = m0 + q * 32 + r;
            rpre[ch][r] = m < m_end ? pp.res[(size_t)m * pp.ldres + n0 + half * COLS + ch * 32 + lane] : 0.f;
          }
        }
      }
      mbar_wait(&tfull[a], (lt >> 1) & 1);
      if (lt == 0 && threadIdx.x == 64) FB_DBG2(4);
      tcgen05_fence_after();
      const int mrow = m0 + q * 32 + lane;       // row owned in the TMEM layout
      float dsum = 0.f;
#pragma unroll (BN == 128 ? 2 : 1)
      for (int cc = 0; cc < COLS; cc += 32) {
        const int c = half * COLS + cc;           // column inside the tile
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + sb[n0 + c + j];
          if (pp.act == FB_ACT_SILU) x = silu_fast(x);
          else if (pp.act == FB_ACT_RELU) x = fmaxf(x, 0.0f);
          o[j] = x;
        }
        if (p.dotv && !second) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dsum = fmaf(s_dot[n0 + c + j], o[j], dsum);
        }
        const int ncol0 = n0 + c;
        float* const outC = (pp.n_split > 0 && ncol0 >= pp.n_split) ? nullptr : pp.C;
        bf16* const outCb = (pp.n_split > 0 && ncol0 < pp.n_split) ? nullptr : pp.Cb;
        const int cb_shift = pp.n_split > 0 ? pp.n_split : 0;
        if (outC || outCb) {
          // transpose through smem: thread = row  ->  lane = column
#pragma unroll
          for (int j = 0; j < 32; ++j) xp[lane * 33 + j] = o[j];
          __syncwarp();
          const int ncol = n0 + c;
          if (outC || pp.res) {
            // C may alias res (in-place residual update of h): loads of a batch of rows are issued before any
            // store of that batch so that they pipeline instead of serialising behind may-alias stores
#pragma unroll (BN == 128 ? 4 : 1)
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float rv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m0 + q * 32 + r0 + i;
                if (PRE) rv[i] = pp.res ? rpre[(cc >> 5) & 1][(r0 + i) & 31] : 0.f;
                else rv[i] = (pp.res && m < m_end) ? pp.res[(size_t)m * pp.ldres + ncol + lane] : 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m0 + q * 32 + r0 + i;
                if (m < m_end) {
                  const float x = xp[(r0 + i) * 33 + lane] + rv[i];
                  if (outC) outC[(size_t)m * pp.ldc + ncol + lane] = x;
                  if (outCb) xp[(r0 + i) * 33 + lane] = x;     // keep the residual-added value for the bf16 copy
                }
              }
            }
            __syncwarp();
          }
          if (outCb) {
            const int rr = lane >> 4, cp = (lane & 15) * 2;
#pragma unroll 4
            for (int r = 0; r < 32; r += 2) {
              const int m = m0 + q * 32 + r + rr;
              if (m < m_end) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(xp[(r + rr) * 33 + cp], xp[(r + rr) * 33 + cp + 1]);
                *reinterpret_cast<__nv_bfloat162*>(outCb + (size_t)m * pp.ldcb + (ncol - cb_shift) + cp) = t;
              }
            }
          }
          __syncwarp();
        }
      }
      if (p.dotv && !second && mrow < m_end) {
        // two warps (column halves) share a row: partial index = 2 * n_tile + half
        p.dot_out[(size_t)((tile % n_tiles_n) * 2 + half) * p.dot_stride + mrow] = dsum;
      }
      if (lt == 0 && threadIdx.x == 64) FB_DBG2(5);
      // this warp is done reading the accumulator stage
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[a])) : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FB_DBG2(6);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

static void fill_params(Params& p, const GemmArgs& g, int m_begin) {
  // outputs are indexed by absolute row (m_begin + local row): rebase the row-0 pointers accordingly
  auto rebase = [&](const void* ptr, int ld, int esz) -> const void* {
    return ptr ? (const void*)((uintptr_t)ptr - (uintptr_t)m_begin * (size_t)ld * esz) : nullptr;
  };
  p.M = g.M; p.N = g.N; p.KB1 = g.K1 / BK; p.KB2 = g.K2 / BK; p.m_dev = g.m_dev; p.m_begin = m_begin;
  p.bias = g.bias; p.act = g.act;
  p.res = (const float*)rebase(g.res, g.ldres, 4); p.ldres = g.ldres;
  p.C = (float*)rebase(g.C, g.ldc, 4); p.ldc = g.ldc;
  p.Cb = (bf16*)rebase(g.Cb, g.ldcb, 2); p.ldcb = g.ldcb;
  p.dotv = g.dotv; p.dot_out = g.dot_out; p.dot_stride = g.dot_stride;
  p.n_split = g.n_split;
  p.dbg = g_tc_dbg;
}

// g1 (optional): second problem on rows [m_begin1, m_begin1 + g1->M) of the same A buffer
template <int BN, int STAGES>
static int launch(const GemmArgs& g, const GemmArgs* g1, int m_begin1, cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  static unsigned long long optin = 0;
  static int num_sms = 0;
  auto kern = g1 ? gemm_tc2_kernel<BN, STAGES, true> : gemm_tc2_kernel<BN, STAGES, false>;
  static unsigned long long optin_g = 0;
  if (!ensure_smem_optin(kern, S::TOTAL, g1 ? optin_g : optin)) return FB_ERR_CUDA;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  CUtensorMap ma, ma2, mw, mw2;
  const int K = g.K1 + g.K2;
  const int rows = g1 ? (m_begin1 + g1->M > g.M ? m_begin1 + g1->M : g.M) : g.M;
  if (!tc_make_map(&ma, g.A, (uint64_t)rows, (uint64_t)g.K1, (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (g.K2 > 0) {
    if (!tc_make_map(&ma2, g.A2, (uint64_t)g.M, (uint64_t)g.K2, (uint64_t)g.lda2, BM)) return FB_ERR_CUDA;
  } else {
    ma2 = ma;
  }
  if (!tc_make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)K, (uint64_t)K, BN)) return FB_ERR_CUDA;
  Params p, p2;
  fill_params(p, g, 0);
  int tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
  if (g1) {
    if (!tc_make_map(&mw2, g1->W, (uint64_t)g1->N, (uint64_t)K, (uint64_t)K, BN)) return FB_ERR_CUDA;
    fill_params(p2, *g1, m_begin1);
    tiles += ((g1->M + BM - 1) / BM) * (g1->N / BN);
  } else {
    mw2 = mw;
    p2 = p;
    p2.M = 0;
  }
  const int grid = tiles < num_sms ? tiles : num_sms;
  fb_launch(kern, dim3(grid), dim3(THREADS), S::TOTAL, st, ma, ma2, mw, mw2, p, p2);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace tc2

bool gemm_tc2_shape_ok(int N) { return N <= tc2::MAX_N; }
// 128x256 tiles once there are enough rows to fill the machine with them; 128x128 tiles (twice the CTAs, half
// the epilogue per CTA) for the short node-level GEMMs, which are latency-bound
int gemm_tc2_bn(int M, int N) { return ((N % 256) == 0 && M >= 16384) ? 256 : 128; }
int gemm_tc2_dot_tiles(int M, int N) { return 2 * (N / gemm_tc2_bn(M, N)); }

int gemm_tc2_launch(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0) return FB_OK;
  if (gemm_tc2_bn(g.M, g.N) == 256) return tc2::launch<256, 3>(g, nullptr, 0, st);
  return tc2::launch<128, 5>(g, nullptr, 0, st);
}

// Two independent GEMMs over disjoint row ranges of ONE activation buffer (same K, different weights and
// outputs) in one launch.  Returns FB_ERR_UNSUPPORTED when the pair cannot be grouped; the caller then launches
// the two problems one after the other.
int gemm_tc2_launch_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t st) {
  if (g0.M <= 0 || g1.M <= 0) return FB_ERR_UNSUPPORTED;
  if (g0.K2 || g1.K2 || g0.K1 != g1.K1 || g0.lda != g1.lda || g0.m_dev || g1.m_dev || g1.dotv) return FB_ERR_UNSUPPORTED;
  if (g0.N > tc2::MAX_N || g1.N > tc2::MAX_N1 || (g0.N % 128) || (g1.N % 128)) return FB_ERR_UNSUPPORTED;
  if (gemm_tc2_bn(g0.M, g0.N) != 128 || gemm_tc2_bn(g1.M, g1.N) != 128) return FB_ERR_UNSUPPORTED;
  const ptrdiff_t off = (const char*)g1.A - (const char*)g0.A;
  const ptrdiff_t row = (ptrdiff_t)g0.lda * 2;
  if (off < 0 || off % row) return FB_ERR_UNSUPPORTED;
  const ptrdiff_t m_begin1 = off / row;
  if (m_begin1 < g0.M) return FB_ERR_UNSUPPORTED;   // ranges must not overlap
  return tc2::launch<128, 5>(g0, &g1, (int)m_begin1, st);
}

}  // namespace fb
