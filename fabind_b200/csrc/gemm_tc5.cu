// Multi-problem persistent tcgen05 GEMM (v5): up to four INDEPENDENT problems  C_i = epilogue_i([A_i | A2_i] W_i^T)  in one launch.
//
// Why: the node-level part of a layer is a chain of small GEMMs (3.7k rows, K = 512): each launch is bounded by its fixed costs
// (launch + pipeline fill + epilogue + drain, ~8 us) and not by its FLOPs.  Folding consecutive Linear maps
// into pre-multiplied weights (forward.cu: "folded" sequences) turns dependent launches into independent problems over different
// operands, row ranges, K and epilogues; this kernel runs such a group as ONE persistent grid over the union of their 128x128 tiles.
//
// Structure = v3 (gemm_tc3.cu) with the producer role spread over three warps: warps 0-2 TMA producers (k-slab i of the CTA's tile
// sequence issued by producer i % 3 into ring slot i % 4), warp 3 MMA issuer (two TMEM accumulator stages), warps 4-11 epilogue
// (thread = row, swizzled staging boxes, TMA stores, residual through TMA).  Differences to v3:
//   * every problem brings its own operand / output descriptors, k-slab counts and epilogue (bias, activation, residual, fp32 / bf16 /
//     column-routed outputs); a tile index decodes to (problem, m0, n0) through the problems' tile offsets;
//   * the descriptors travel inside ONE __grid_constant__ parameter block (24 x 128 B);
//   * weight prefetch: the W slabs of a CTA's first tile are requested BEFORE griddepcontrol.wait -- weights do not depend on the
//     previous kernel of the stream, so their L2 -> smem latency overlaps that kernel's tail (opt-in per launch: the caller
//     guarantees the weights were not written by the immediately preceding kernels).
// No row-dot epilogue, no dropout, no device-side row counts: launches that need them stay on v3 / v4.
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
bool tc_make_map_out(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t rows, uint64_t cols, uint64_t ld);

extern long long* g_tc_dbg;

namespace tc5 {
using namespace tc;

// per-role wait accounting (diagnostic builds; scripts/dev/tc5_stalls.py): g_tc_dbg[cta * 8 + k], k = 0 MMA warp waits for operands
// (full), 1 MMA warp waits for a drained accumulator stage (tempty), 2 producer 0 waits for a free ring slot (empty), 3 epilogue warp 0
// waits for an accumulator (tfull), 4 epilogue warp 0 busy, 5 kernel total, 6 clocks spent in griddepcontrol.wait, 7 tiles of the CTA
#ifdef FB_DIAG
#define FB5_T0() const long long t0__ = clock64()
#define FB5_ACC(var) var += clock64() - t0__
#else
#define FB5_T0() do { } while (0)
#define FB5_ACC(var) do { } while (0)
#endif

constexpr int NPMAX = 4;
constexpr int BN = 128;
constexpr int STAGES = 4;
constexpr int NS = 3;                           // staging boxes per epilogue warp: two fp32 (also the residual landing zone) + one bf16
constexpr int EPI_WARPS = 8;
constexpr int PRODUCERS = 3;                    // TMA-issuing warps (see gemm_tc4.cu: one issuing thread is served at ~32 B/clk per SM)
constexpr int MMA_WARP = PRODUCERS;             // warp 3
constexpr int EPI_WARP0 = PRODUCERS + 1;        // warps 4..11
constexpr int THREADS = (PRODUCERS + 1 + EPI_WARPS) * 32;   // 384
constexpr int SLOT = 4096;                      // one staging box: 32 rows x 128 bytes

struct Prob {
  int M, N;                // rows / columns
  int KB1, KB2;            // 64-wide k-slabs taken from A / A2
  int tile_begin;          // index of this problem's first tile in the launch
  int ntn;                 // N / BN
  const float* bias; int act;
  int has_res, has_c, has_cb, n_split;
};
struct Params {
  CUtensorMap a[NPMAX], a2[NPMAX], w[NPMAX], c[NPMAX], cb[NPMAX], res[NPMAX];
  Prob q[NPMAX];
  int np, n_tiles, prefetch_w;
  long long* dbg;
};

struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;                    // 1024-aligned
  static constexpr int BAR_OFF = STAGING_OFF + EPI_WARPS * NS * SLOT;         // full[S] empty[S] tfull[2] tempty[2] res[8] slot
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4 + EPI_WARPS) * 8 + 16 + 1024;   // + alignment slack
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t sw_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1) gemm_tc5_kernel(const __grid_constant__ Params p) {
  using S = Smem;
  pdl_trigger();
#ifdef FB_DIAG
  const long long t_start = clock64();
  long long w_full = 0, w_tempty = 0, w_empty = 0, w_tfull = 0, e_busy = 0, w_pdl = 0;
  int n_my = 0;
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resbar = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(resbar + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_tiles;

  // tile -> (problem, first row, first column)
  auto decode = [&](int tile, int& m0, int& n0) -> int {
    int pi = 0;
#pragma unroll
    for (int i = 1; i < NPMAX; ++i)
      if (i < p.np && tile >= p.q[i].tile_begin) pi = i;
    const int local = tile - p.q[pi].tile_begin, ntn = p.q[pi].ntn;
    m0 = (local / ntn) * BM; n0 = (local % ntn) * BN;
    return pi;
  };

  // tile of this CTA in round r of the persistent loop: boustrophedon over the CTAs (even rounds ascending, odd rounds descending).
  // The launch orders the problems by descending K, so round 0 hands the heavy tiles to the low CTAs and the partial last rounds top
  // up the CTAs that carry the light ones (plain round-robin gave the heavy CTAs the extra tiles too: makespan 3 instead of 2 tile
  // units for node_mlp.2 + first projections, 6 instead of 5 for the stacked q|k|v group).  -1 = no tile in this round.
  const int G = gridDim.x, n_rounds = (n_tiles + G - 1) / G;
  auto tile_at = [&](int r) -> int {
    const int t = r * G + ((r & 1) ? G - 1 - (int)blockIdx.x : (int)blockIdx.x);
    return t < n_tiles ? t : -1;
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.np; ++i) {
      prefetch_map(&p.a[i]); prefetch_map(&p.w[i]);
      if (p.q[i].KB2) prefetch_map(&p.a2[i]);
    }
  }
  if (warp == EPI_WARP0 && lane == 0) {
    for (int i = 0; i < p.np; ++i) {
      if (p.q[i].has_c) prefetch_map(&p.c[i]);
      if (p.q[i].has_cb) prefetch_map(&p.cb[i]);
      if (p.q[i].has_res) prefetch_map(&p.res[i]);
    }
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
      for (int e = 0; e < EPI_WARPS; ++e) mbar_init(&resbar[e], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // weight slab of each producer's FIRST k-slab: requested before the wait on the previous grid (see the header)
  int pre_it = -1;
  if (warp < PRODUCERS && lane == 0 && p.prefetch_w) {
    int it = 0;
    for (int r = 0; r < n_rounds && pre_it < 0 && it < STAGES; ++r) {
      const int tile = tile_at(r);
      if (tile < 0) continue;
      int m0, n0;
      const int pi = decode(tile, m0, n0);
      const int KB = p.q[pi].KB1 + p.q[pi].KB2;
      for (int kb = 0; kb < KB && it < STAGES; ++kb, ++it) {
        if (it % PRODUCERS != warp) continue;
        mbar_expect_tx(&full[it], S::STAGE_BYTES);                  // first pass over the ring: slot it is free
        tma_load_2d(&p.w[pi], &full[it], smem + it * S::STAGE_BYTES + S::A_BYTES, kb * BK, n0);
        pre_it = it;
        break;
      }
    }
  }
  // everything above touched only on-chip state and the weights; activations are produced by the previous kernel in the stream
  { FB5_T0(); pdl_wait(); FB5_ACC(w_pdl); }

  if (warp < PRODUCERS) {
    // ===== TMA producers: producer `warp` issues the k-slabs it with it % PRODUCERS == warp of the CTA's tile sequence =====
    if (lane == 0) {
      int it = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int tile = tile_at(r);
        if (tile < 0) continue;
        int m0, n0;
        const int pi = decode(tile, m0, n0);
        const int KB1 = p.q[pi].KB1, KB = KB1 + p.q[pi].KB2;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          if (it % PRODUCERS != warp) continue;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          const bool armed = it == pre_it;          // barrier armed and W slab already in flight
          if (!armed) {
            { FB5_T0(); mbar_wait(&empty[s], ph ^ 1); FB5_ACC(w_empty); }
            mbar_expect_tx(&full[s], S::STAGE_BYTES);
          }
          if (kb < KB1) tma_load_2d(&p.a[pi], &full[s], a_dst, kb * BK, m0);
          else tma_load_2d(&p.a2[pi], &full[s], a_dst, (kb - KB1) * BK, m0);
          if (!armed) tma_load_2d(&p.w[pi], &full[s], a_dst + S::A_BYTES, kb * BK, n0);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0, lt = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int tile = tile_at(r);
        if (tile < 0) continue;
        int m0, n0;
        const int pi = decode(tile, m0, n0);
        const int KB = p.q[pi].KB1 + p.q[pi].KB2;
        const int a = lt & 1;
        { FB5_T0(); mbar_wait(&tempty[a], ((lt >> 1) & 1) ^ 1); FB5_ACC(w_tempty); }   // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          { FB5_T0(); mbar_wait(&full[s], ph); FB5_ACC(w_full); }
          tcgen05_fence_after();
          const uint8_t* a_src = smem + s * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_src), bdesc = make_smem_desc(a_src + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[a]);
        ++lt;
      }
    }
  } else {
    // ===== epilogue =====
    const int e = warp - EPI_WARP0;         // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // column half of the tile
    constexpr int COLS = BN / 2;            // columns per warp
    constexpr int NCH = COLS / 32;          // 32-column pieces per warp and tile
    uint8_t* const slots = smem + S::STAGING_OFF + e * NS * SLOT;
    uint64_t* const rbar = &resbar[e];
    uint32_t rphase = 0;
    int lt = -1;
    for (int r = 0; r < n_rounds; ++r) {
      const int tile = tile_at(r);
      if (tile < 0) continue;
      ++lt;
      const int a = lt & 1;
      int m0, n0;
      const int pi = decode(tile, m0, n0);
      const Prob& pp = p.q[pi];
      const CUtensorMap* mc = &p.c[pi];
      const CUtensorMap* mcb = &p.cb[pi];
      const CUtensorMap* mres = &p.res[pi];
      const int lrow0 = m0 + q * 32;                      // first row of this warp
      const bool rows_live = lrow0 < pp.M;                // warp-uniform
      const int colbase = n0 + half * COLS;
      const bool has_res = pp.has_res != 0;
      float bv[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) bv[ch] = pp.bias ? __ldg(pp.bias + colbase + ch * 32 + lane) : 0.f;
      // the staging boxes are free once every earlier bulk store of this warp has read its shared memory
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      if (has_res && rows_live && lane == 0) {
        mbar_expect_tx(rbar, NCH * SLOT);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) tma_load_2d(mres, rbar, slots + ch * SLOT, colbase + ch * 32, lrow0);
      }
      { FB5_T0(); mbar_wait(&tfull[a], (lt >> 1) & 1); FB5_ACC(w_tfull); }
#ifdef FB_DIAG
      const long long t_busy0 = clock64();
      ++n_my;
#endif
      tcgen05_fence_after();
      if (has_res && rows_live) { mbar_wait(rbar, rphase); rphase ^= 1; }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int c = half * COLS + ch * 32;      // column inside the tile
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
        if (ch == NCH - 1) {
          // accumulator stage drained: hand it back to the MMA warp before the stores
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[a])) : "memory");
        }
        if (!rows_live) continue;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + __shfl_sync(0xffffffffu, bv[ch], j);
          if (pp.act == FB_ACT_SILU) x = silu_fast(x);
          else if (pp.act == FB_ACT_RELU) x = fmaxf(x, 0.0f);
          o[j] = x;
        }
        const int ncol0 = n0 + c;
        const bool want_c = pp.has_c && !(pp.n_split > 0 && ncol0 >= pp.n_split);
        const bool want_cb = pp.has_cb && !(pp.n_split > 0 && ncol0 < pp.n_split);
        uint8_t* const fs = slots + ch * SLOT;
        uint8_t* const bs = slots + 2 * SLOT;
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(fs + sw_off(lane, j));
            o[4 * j] += r4.x; o[4 * j + 1] += r4.y; o[4 * j + 2] += r4.z; o[4 * j + 3] += r4.w;
          }
        }
        if (want_c) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(fs + sw_off(lane, j)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(mc, fs, ncol0, lrow0);
        }
        if (want_cb) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(o[8 * j], o[8 * j + 1]), t1 = __floats2bfloat162_rn(o[8 * j + 2], o[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(o[8 * j + 4], o[8 * j + 5]), t3 = __floats2bfloat162_rn(o[8 * j + 6], o[8 * j + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(bs + sw_off(lane, (ch & 1) * 4 + j)) = u;
          }
          if (ch & 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(mcb, bs, ncol0 - 32 - (pp.n_split > 0 ? pp.n_split : 0), lrow0);
          }
        }
      }
#ifdef FB_DIAG
      e_busy += clock64() - t_busy0;
#endif
    }
    // shared memory must stay valid until the last bulk stores have read it
    if (lane == 0) bulk_wait_read0();
  }
#ifdef FB_DIAG
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + (size_t)blockIdx.x * 8;
    if (warp == MMA_WARP) { d[0] = w_full; d[1] = w_tempty; }
    if (warp == 0) { d[2] = w_empty; d[6] = w_pdl; }
    if (warp == EPI_WARP0) { d[3] = w_tfull; d[4] = e_busy; d[7] = n_my; }
  }
#endif
  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
#ifdef FB_DIAG
  if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + 5] = clock64() - t_start;
#endif
}

}  // namespace tc5

// FB_ERR_UNSUPPORTED -> the caller launches the problems one after the other through gemm_launch
int gemm_tc5_launch(const GemmArgs* g, int np, bool prefetch_w, cudaStream_t st) {
  using namespace tc5;
  if (np < 1 || np > NPMAX) return FB_ERR_UNSUPPORTED;
  for (int i = 0; i < np; ++i) {
    const GemmArgs& a = g[i];
    if (a.M <= 0 || !gemm_tc_supported(a) || (a.N % BN) || a.dotv || a.m_dev || a.drop.p > 0.f || a.nprod) return FB_ERR_UNSUPPORTED;
    if (a.n_split > 0 && ((a.n_split % 64) || !a.C || !a.Cb)) return FB_ERR_UNSUPPORTED;
    if (a.res && !a.C && !a.Cb) return FB_ERR_UNSUPPORTED;
    if (a.res && a.n_split > 0) return FB_ERR_UNSUPPORTED;
    if (a.ldw != 0 && (a.ldw < a.K1 + a.K2 || (a.ldw % 8))) return FB_ERR_UNSUPPORTED;
  }
  static_assert(Smem::TOTAL <= 232448, "shared memory budget");
  static_assert(sizeof(Params) <= 4096, "kernel parameter budget");
  static unsigned long long optin = 0;
  static int num_sms = 0;
  if (!ensure_smem_optin(gemm_tc5_kernel, Smem::TOTAL, optin)) return FB_ERR_CUDA;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  Params p;
  int tiles = 0;
  for (int i = 0; i < NPMAX; ++i) {
    const GemmArgs& a = g[i < np ? i : 0];
    Prob& q = p.q[i];
    if (i >= np) {     // unused entries mirror problem 0 (never decoded: tile_begin beyond the last tile)
      p.a[i] = p.a[0]; p.a2[i] = p.a2[0]; p.w[i] = p.w[0]; p.c[i] = p.c[0]; p.cb[i] = p.cb[0]; p.res[i] = p.res[0];
      q = p.q[0]; q.tile_begin = 1 << 30;
      continue;
    }
    const int K = a.K1 + a.K2;
    if (!tc_make_map(&p.a[i], a.A, (uint64_t)a.M, (uint64_t)a.K1, (uint64_t)a.lda, BM)) return FB_ERR_CUDA;
    if (a.K2 > 0) {
      if (!tc_make_map(&p.a2[i], a.A2, (uint64_t)a.M, (uint64_t)a.K2, (uint64_t)a.lda2, BM)) return FB_ERR_CUDA;
    } else {
      p.a2[i] = p.a[i];
    }
    if (!tc_make_map(&p.w[i], a.W, (uint64_t)a.N, (uint64_t)K, (uint64_t)(a.ldw > 0 ? a.ldw : K), BN)) return FB_ERR_CUDA;
    p.c[i] = p.cb[i] = p.res[i] = p.a[i];   // placeholders for absent operands (never dereferenced)
    const int nc = a.n_split > 0 ? a.n_split : a.N, ncb = a.n_split > 0 ? a.N - a.n_split : a.N;
    if (a.C && !tc_make_map_out(&p.c[i], a.C, true, (uint64_t)a.M, (uint64_t)nc, (uint64_t)a.ldc)) return FB_ERR_CUDA;
    if (a.Cb && !tc_make_map_out(&p.cb[i], a.Cb, false, (uint64_t)a.M, (uint64_t)ncb, (uint64_t)a.ldcb)) return FB_ERR_CUDA;
    if (a.res && !tc_make_map_out(&p.res[i], a.res, true, (uint64_t)a.M, (uint64_t)nc, (uint64_t)a.ldres)) return FB_ERR_CUDA;
    q.M = a.M; q.N = a.N; q.KB1 = a.K1 / BK; q.KB2 = a.K2 / BK; q.tile_begin = tiles; q.ntn = a.N / BN;
    q.bias = a.bias; q.act = a.act;
    q.has_res = a.res != nullptr; q.has_c = a.C != nullptr; q.has_cb = a.Cb != nullptr; q.n_split = a.n_split;
    tiles += ((a.M + BM - 1) / BM) * q.ntn;
  }
  p.np = np; p.n_tiles = tiles; p.prefetch_w = prefetch_w ? 1 : 0;
  p.dbg = g_tc_dbg;
  const int grid = tiles < num_sms ? tiles : num_sms;
  fb_launch(gemm_tc5_kernel, dim3(grid), dim3(THREADS), Smem::TOTAL, st, p);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
