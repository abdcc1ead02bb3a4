// tcgen05 GEMM v4: CTA PAIRS (cta_group::2) for the long edge-level / pair-level GEMMs.
//
// Why: at a 128x256 tile per CTA (v3) every 64-wide k-slab needs (128 + 256) x 64 x 2 B = 48 KB of operands per 0.27 us of
// MMA time; chip-wide that is ~26 TB/s out of L2, about twice what the L2 -> SM fabric delivers, and v3 stalls at ~53 % of the
// measured bf16 peak with DRAM traffic equal to the algorithmic bytes.  Two CTAs of one TPC computing
// ONE 256x256 tile with tcgen05.mma.cta_group::2 each stage their own 128 rows of A and only HALF of the B (weight) slab:
// 32 KB per CTA and k-slab for the same MMA time, a third less operand traffic.
//
// Structure (per CTA): warps 0-2 = TMA producers (a CTA's 128 A rows + its 128 of the 256 W rows per k-slab, with the completion
// bytes of BOTH CTAs signalled on the LEADER's "full" barrier), warp 3 = MMA issuer (leader CTA only; one
// tcgen05.mma.cta_group::2 per 16-wide k-step, M = 256, N = 256; tcgen05.commit multicast releases the smem stage / publishes the
// accumulator in both CTAs), warps 4-11 = epilogue (each CTA drains its own 128 TMEM lanes; rows stored straight from registers).  The
// leader's "accumulator drained" barrier counts the epilogue warps of both CTAs (remote mbarrier arrive).
// THREE producer warps on three SM sub-partitions, k-slab i issued by producer i % 3 into ring slot i % 4: the bulk-tensor loads of
// ONE issuing thread are served at ~32 B/clk per SM however deep the ring (profiles/r2n_tma_issue_microbench_*: 23 B/clk with one
// box per barrier, 32 with two, 65 from two warps, 106 from four), and SIX ring stages: with four, the MMA warp spent 38-47 % of
// the kernel waiting for operands while the producer waited as long for free slots (profiles/r2p_tc4_role_stalls_4_stage_ring.txt:
// the slot round trip -- load, queue, MMA, commit, re-issue -- is ~3500 clk, four slots in flight sustain one k-slab per ~950 clk
// against ~750 clk of MMA time).  The 64 KB the two extra stages need were the TMA-store staging boxes of the epilogue.
// Persistent over 256x256 tiles; two TMEM accumulator stages (2 x 256 columns).  No residual, one stored output (the shapes this
// kernel is picked for never need more); outputs need 32-byte aligned rows (256-bit stores).
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

extern long long* g_tc_dbg;
bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace tc4 {
using namespace tc;

constexpr int BN = 256;                         // tile columns (the pair's MMA N)
constexpr int BNH = BN / 2;                     // W rows staged per CTA
constexpr int PM = 2 * BM;                      // tile rows of the pair
constexpr int EPI_WARPS = 8;
constexpr int PRODUCERS = 3;                    // TMA-issuing warps on three SM sub-partitions (the fourth hosts the MMA warp); 12 warps keep 168 registers per thread
constexpr int MMA_WARP = PRODUCERS;             // warp 3
constexpr int EPI_WARP0 = PRODUCERS + 1;        // warps 4..11
constexpr int THREADS = (PRODUCERS + 1 + EPI_WARPS) * 32;   // 384
constexpr int STAGES = 6;                       // 192 KB of operands in flight per CTA (see the header: the ring depth is what fed the MMAs too slowly)
constexpr int VEC = BN / 2;                     // columns per epilogue warp: its bias / row-dot vectors live in shared memory

struct Params {
  int M, N, KB1, KB2;
  int nprod, exact_act;     // split-precision mode (see gemm.h): products per k-block, KB1 = k-blocks of ONE plane
  const float* bias; int act;
  float* C; int ldc;
  bf16* Cb; int ldcb;
  const float* dotv; float* dot_out; int dot_stride;
  int n_split;
  const int* m_dev;
  DropCfg drop;
  int w_static;             // the weights were not written by the kernels just before this launch (weight-stationary form: prefetch)
  long long* dbg;
};

constexpr int KB_MAX = 8;                       // weight-stationary form: up to 8 resident k-slabs (K <= 512)

// WSTAT = weight-stationary form: the CTA's half of ONE 256-column weight tile (128 rows x K <= 512 = up to 128 KB) stays in shared
// memory for the life of the CTA and only A streams through the ring (16 KB slabs)
template <bool WSTAT>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2;
  static constexpr int NSTAGE = WSTAT ? 5 : STAGES;
  static constexpr int STAGE_BYTES = WSTAT ? A_BYTES : A_BYTES + B_BYTES;     // 16 / 32 KB
  static constexpr int W_OFF = 0;                                             // WSTAT: resident weight slabs
  static constexpr int RING_OFF = WSTAT ? KB_MAX * B_BYTES : 0;
  static constexpr int VEC_OFF = RING_OFF + NSTAGE * STAGE_BYTES;             // per epilogue warp: bias[VEC] | dotv[VEC] (fp32)
  static constexpr int BAR_OFF = VEC_OFF + EPI_WARPS * 2 * VEC * 4;           // full[S] empty[S] tfull[2] tempty[2] wfull slot
  static constexpr int TOTAL = BAR_OFF + (2 * NSTAGE + 5) * 8 + 16 + 1024;    // + alignment slack
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // relaxed arrive: barrier initialisation is published by fence.mbarrier_init.release.cluster, the tear-down sync orders nothing
  // but the life time of the barriers (a .release arrive costs a MEMBAR.ALL + ERRBAR per warp)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are signalled on the barrier at the same offset in the LEADER CTA of the pair
// (bit 24 of a shared::cluster address selects the CTA inside a pair; cleared = even rank)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      // .relaxed: the TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; a .release.cluster arrive
      // would add MEMBAR.ALL.GPU + ERRBAR (20 % of the epilogue's stall samples in ncu)
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// 32 bytes = one full sector per lane.  L2::evict_last: a row's 128-byte line is completed by two stores a chunk apart (and consumed
// from L2 by the next kernel); with the default policy the half-written lines were evicted early and written twice -- 75.5 MB of
// DRAM writes for 46 MB of output in the cold ncu capture (profiles/r2r_tc4_ncu_summary.txt)
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
// four consecutive floats of a read-only vector (16-byte load when the address allows it)
__device__ __forceinline__ float4 ldg4(const float* p) {
  if (((uintptr_t)p & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// SPLIT: split-precision instantiation (table-driven k-block walk + full-precision SiLU, see gemm_tc3.cu); ACT: activation of the
// epilogue as a template parameter, so that one instantiation carries one straight-line epilogue (the runtime select between three
// unrolled variants tripled the code the eight epilogue warps walk through: 18 % of their stall samples were instruction fetches,
// profiles/r2f_tc4_epilogue_before.txt)
#ifdef FB_DIAG
// per-role stall accounting (clock64 deltas summed per CTA): g_tc_dbg[cta * 8 + k], k = 0 MMA waits for data (full), 1 MMA waits for a
// drained accumulator (tempty), 2 producer 0 waits for a free slot (empty), 3 epilogue warp 0 waits for an accumulator (tfull),
// 4 epilogue warp 0 busy, 5 kernel total (thread 0)
#define FB_T0() const long long t0__ = clock64()
#define FB_ACC(var) var += clock64() - t0__
#else
#define FB_T0() do { } while (0)
#define FB_ACC(var) do { } while (0)
#endif

template <bool SPLIT, int ACT, bool WSTAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_w, const Params p) {
  using S = Smem<WSTAT>;
  constexpr int NST = S::NSTAGE;
  static_assert(!(SPLIT && WSTAT), "the weight-stationary form is a bf16-mode kernel");
  pdl_trigger();
#ifdef FB_DIAG
  const long long t_start = clock64();
  long long w_full = 0, w_tempty = 0, w_empty = 0, w_tfull = 0, e_busy = 0;
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int KB = SPLIT ? p.nprod * p.KB1 : p.KB1 + p.KB2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
    if (p.KB2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a2) : "memory");
  }
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      // tempty of the leader collects the epilogue warps of BOTH CTAs
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
      mbar_init(wfull, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();               // the peer's barriers are initialised before anything is signalled on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles_n = p.N / BN;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // tile schedule.  Streaming form: tile = pair, pair + n_pairs, ... over (row block, column tile) with the column tile fastest.
  // Weight-stationary form: the pair keeps column tile pair % n_tiles_n and walks the row blocks pair / n_tiles_n, + n_pairs /
  // n_tiles_n, ... (the launch sizes the grid to a multiple of n_tiles_n pairs).
  const int ppn = WSTAT ? n_pairs / n_tiles_n : 1;
  const int nt_fixed = WSTAT ? pair % n_tiles_n : 0;
  if (WSTAT && warp == 0 && lane == 0) {
    // resident weight tile: this CTA's 128 of the 256 rows, all k-slabs, completion on the leader's wfull.  Requested before the wait
    // on the previous grid when the caller vouches for the weights (w_static): they then arrive while that grid drains.
    if (p.w_static) {
      if (leader) mbar_expect_tx(wfull, 2 * KB * S::B_BYTES);
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d_pair(&map_w, wfull, smem + S::W_OFF + kb * S::B_BYTES, kb * BK, nt_fixed * BN + (int)rank * BNH);
    }
  }
  pdl_wait();
  if (WSTAT && warp == 0 && lane == 0 && !p.w_static) {
    if (leader) mbar_expect_tx(wfull, 2 * KB * S::B_BYTES);
    for (int kb = 0; kb < KB; ++kb)
      tma_load_2d_pair(&map_w, wfull, smem + S::W_OFF + kb * S::B_BYTES, kb * BK, nt_fixed * BN + (int)rank * BNH);
  }
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int n_rb = (M + PM - 1) / PM;
  const int n_tiles = n_rb * n_tiles_n;
  int my_tiles = 0;
  if (WSTAT) { const int rb0 = pair / n_tiles_n; my_tiles = rb0 < n_rb ? (n_rb - rb0 + ppn - 1) / ppn : 0; }
  else { my_tiles = pair < n_tiles ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0; }
  // lt-th tile of this pair -> (row block, column tile)
  auto sched = [&](int lt, int& rb, int& nt) {
    if (WSTAT) { rb = pair / n_tiles_n + lt * ppn; nt = nt_fixed; }
    else { const int tile = pair + lt * n_pairs; rb = tile / n_tiles_n; nt = tile % n_tiles_n; }
  };

  if (warp < PRODUCERS) {
    // ===== TMA producers (both CTAs): producer `warp` issues the k-slabs it, it + PRODUCERS, ... of the CTA's tile sequence =====
    if (lane == 0) {
      const int total = my_tiles * KB;
      for (int it = warp; it < total; it += PRODUCERS) {
        const int lt = it / KB, kb = it - lt * KB;
        int rb, nt;
        sched(lt, rb, nt);
        const int m0 = rb * PM + (int)rank * BM;
        const int n0 = nt * BN + (int)rank * BNH;
        const int s = it % NST;
        const uint32_t ph = (it / NST) & 1;
        { FB_T0(); mbar_wait(&empty[s], ph ^ 1); FB_ACC(w_empty); }
        uint8_t* a_dst = smem + S::RING_OFF + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        if (leader) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);     // bytes of both CTAs land on the leader's barrier
        if (SPLIT) {
          const int j = kb / p.KB1, r = kb - j * p.KB1;
          tma_load_2d_pair(&map_a, &full[s], a_dst, (split_plane_a(j, p.nprod) * p.KB1 + r) * BK, m0);
          tma_load_2d_pair(&map_w, &full[s], b_dst, (split_plane_w(j, p.nprod) * p.KB1 + r) * BK, n0);
          continue;
        }
        if (kb < p.KB1) tma_load_2d_pair(&map_a, &full[s], a_dst, kb * BK, m0);
        else tma_load_2d_pair(&map_a2, &full[s], a_dst, (kb - p.KB1) * BK, m0);
        if (!WSTAT) tma_load_2d_pair(&map_w, &full[s], b_dst, kb * BK, n0);
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
      int it = 0;
      if (WSTAT) mbar_wait(wfull, 0);
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int a = lt & 1;
        { FB_T0(); mbar_wait(&tempty[a], ((lt >> 1) & 1) ^ 1); FB_ACC(w_tempty); }   // both epilogues have drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % NST;
          const uint32_t ph = (it / NST) & 1;
          { FB_T0(); mbar_wait(&full[s], ph); FB_ACC(w_full); }
          tcgen05_fence_after();
          const uint8_t* a_src = smem + S::RING_OFF + s * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_src);
          const uint64_t bdesc = make_smem_desc(WSTAT ? smem + S::W_OFF + kb * S::B_BYTES : a_src + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16_pair(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty[s]);
        }
        umma_commit_pair(&tfull[a]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs, own 128 TMEM lanes) =====
    // thread = accumulator row, 32 columns per tcgen05.ld.  The per-column vectors (bias, row-dot weights) of this warp's 128 columns
    // sit in its own slice of shared memory and are read as warp-uniform 16-byte loads (one LDS.128 per four columns) -- the
    // lane-owns-a-column + shuffle form cost one SHFL per element and vector, a third of the epilogue's instructions.
    // bf16 SiLU: x silu = h + h tanh(h), h = (acc + b) / 2 = fma(acc, 0.5, b / 2): the stored vector is b / 2.
    const int e = warp - EPI_WARP0;         // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // column half of the tile
    constexpr int COLS = BN / 2;            // columns per warp
    constexpr int NCH = COLS / 32;          // 32-column pieces per warp and tile
    float* const vb = reinterpret_cast<float*>(smem + S::VEC_OFF) + e * 2 * VEC;
    float* const vd = vb + VEC;
    const bool use_dot = p.dotv != nullptr;
    constexpr bool HALF_BIAS = ACT == FB_ACT_SILU && !SPLIT;
    int vec_n0 = -1;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int a = lt & 1;
      int rb, nt;
      sched(lt, rb, nt);
      const int m0 = rb * PM + (int)rank * BM;
      const int n0 = nt * BN;
      const int lrow0 = m0 + q * 32;
      const bool rows_live = lrow0 < M;
      const int colbase = n0 + half * COLS;
      if (n0 != vec_n0) {                   // (warp-uniform) the vectors of this column range
        vec_n0 = n0;
        __syncwarp();
        float4 b4 = p.bias ? ldg4(p.bias + colbase + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (HALF_BIAS) { b4.x *= 0.5f; b4.y *= 0.5f; b4.z *= 0.5f; b4.w *= 0.5f; }
        reinterpret_cast<float4*>(vb)[lane] = b4;
        if (use_dot) reinterpret_cast<float4*>(vd)[lane] = ldg4(p.dotv + colbase + 4 * lane);
        __syncwarp();
      }
      { FB_T0(); mbar_wait(&tfull[a], (lt >> 1) & 1); FB_ACC(w_tfull); }
#ifdef FB_DIAG
      const long long t_busy0 = clock64();
#endif
      tcgen05_fence_after();
      float ds0 = 0.f, ds1 = 0.f, ds2 = 0.f, ds3 = 0.f;
      // TMEM loads are software-pipelined: the load of chunk ch+1 is in flight while chunk ch is processed
      uint32_t vv[2][32];
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + half * COLS);
      tmem_ld32_issue(tbase, vv[0]);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int c = half * COLS + ch * 32;      // column inside the tile
        uint32_t (&v)[32] = vv[ch & 1];
        tmem_ld32_wait(v);
        if (ch + 1 < NCH) tmem_ld32_issue(tbase + (uint32_t)((ch + 1) * 32), vv[(ch + 1) & 1]);
        if (ch == NCH - 1) {
          // accumulator stage drained: hand it back to the leader's MMA warp before the stores
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(&tempty[a], 0);
        }
        if (!rows_live) continue;
        float o[32];
        const float4* const b4p = reinterpret_cast<const float4*>(vb + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = b4p[j];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float acc = __uint_as_float(v[4 * j + t]);
            float x;
            if (HALF_BIAS) {
              const float h = fmaf(acc, 0.5f, bb[t]);
              float th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
              x = fmaf(h, th, h);
            } else {
              x = acc + bb[t];
              if (ACT == FB_ACT_SILU) x = silu(x);
              else if (ACT == FB_ACT_RELU) x = fmaxf(x, 0.0f);
            }
            o[4 * j + t] = x;
          }
        }
        if (p.drop.p > 0.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = drop_apply(o[j], p.drop, lrow0 + lane, n0 + c + j);
        }
        if (use_dot) {
          const float4* const d4p = reinterpret_cast<const float4*>(vd + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 d4 = d4p[j];
            ds0 = fmaf(d4.x, o[4 * j], ds0); ds1 = fmaf(d4.y, o[4 * j + 1], ds1);
            ds2 = fmaf(d4.z, o[4 * j + 2], ds2); ds3 = fmaf(d4.w, o[4 * j + 3], ds3);
          }
        }
        // rows go straight from registers to global memory: 32 consecutive columns of the thread's row = four (fp32) or two (bf16)
        // full 32-byte sectors per 256-bit store; the staging boxes of the TMA-store epilogue gave their 64 KB to the operand ring
        const int ncol0 = n0 + c;
        const bool want_c = p.C != nullptr && !(p.n_split > 0 && ncol0 >= p.n_split);
        const bool want_cb = p.Cb != nullptr && !(p.n_split > 0 && ncol0 < p.n_split);
        const int row = lrow0 + lane;
        if (want_c && row < M) {
          float* dst = p.C + (size_t)row * p.ldc + ncol0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = __float_as_uint(o[8 * j + t]);
            st_global_256(dst + 8 * j, w8);
          }
        }
        if (want_cb && row < M) {
          bf16* dst = p.Cb + (size_t)row * p.ldcb + (ncol0 - (p.n_split > 0 ? p.n_split : 0));
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              __nv_bfloat162 t2 = __floats2bfloat162_rn(o[16 * j + 2 * t], o[16 * j + 2 * t + 1]);
              w8[t] = *reinterpret_cast<uint32_t*>(&t2);
            }
            st_global_256(dst + 16 * j, w8);
          }
        }
      }
      if (use_dot && lrow0 + lane < M) {
        // two warps (column halves) share a row: partial index = 2 * n_tile + half
        p.dot_out[(size_t)(nt * 2 + half) * p.dot_stride + lrow0 + lane] = (ds0 + ds1) + (ds2 + ds3);
      }
#ifdef FB_DIAG
      e_busy += clock64() - t_busy0;
#endif
    }
  }
#ifdef FB_DIAG
  if (p.dbg && lane == 0) {
    long long* d = p.dbg + (size_t)blockIdx.x * 8;
    if (warp == MMA_WARP && leader) { d[0] = w_full; d[1] = w_tempty; }
    if (warp == 0) d[2] = w_empty;
    if (warp == EPI_WARP0) { d[3] = w_tfull; d[4] = e_busy; }
  }
#endif
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();               // no CTA of the pair leaves (or frees TMEM) while the other may still signal it
#ifdef FB_DIAG
  if (p.dbg && threadIdx.x == 0) p.dbg[(size_t)blockIdx.x * 8 + 5] = clock64() - t_start;
#endif
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

static int launch(const GemmArgs& g, cudaStream_t st) {
  static_assert(Smem<false>::TOTAL <= 232448 && Smem<true>::TOTAL <= 232448, "shared memory budget");
  static int num_sms = 0;
  const bool split = g.nprod > 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_tiles_n = g.N / BN, n_rb = (g.M + PM - 1) / PM;
  const int K = g.nprod ? 3 * g.K1 : g.K1 + g.K2;   // split precision: three bf16 planes of K1 columns each
  // weight-stationary form: bf16 mode, K <= 512, at least one pair per column tile
#ifdef FB_DIAG
  static const bool wstat_on = [] { const char* e = getenv("FB_WSTAT"); return !(e && atoi(e) == 0); }();
#else
  const bool wstat_on = true;
#endif
  const bool wstat = wstat_on && !split && K <= KB_MAX * BK && n_tiles_n <= num_sms / 2;
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, Params);
  static const Kern table[3][3] = {
      {gemm_tc4_kernel<false, FB_ACT_NONE, false>, gemm_tc4_kernel<false, FB_ACT_SILU, false>, gemm_tc4_kernel<false, FB_ACT_RELU, false>},
      {gemm_tc4_kernel<true, FB_ACT_NONE, false>, gemm_tc4_kernel<true, FB_ACT_SILU, false>, gemm_tc4_kernel<true, FB_ACT_RELU, false>},
      {gemm_tc4_kernel<false, FB_ACT_NONE, true>, gemm_tc4_kernel<false, FB_ACT_SILU, true>, gemm_tc4_kernel<false, FB_ACT_RELU, true>}};
  static unsigned long long optins[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (g.act < 0 || g.act > 2) return FB_ERR_BAD_ARG;
  const int v = wstat ? 2 : (split ? 1 : 0);
  Kern kern = table[v][g.act];
  const int smem_bytes = wstat ? Smem<true>::TOTAL : Smem<false>::TOTAL;
  if (!ensure_smem_optin(kern, smem_bytes, optins[v][g.act])) return FB_ERR_CUDA;
  CUtensorMap ma, ma2, mw;
  if (!tc_make_map(&ma, g.A, (uint64_t)g.M, (uint64_t)(g.nprod ? K : g.K1), (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (g.K2 > 0) {
    if (!tc_make_map(&ma2, g.A2, (uint64_t)g.M, (uint64_t)g.K2, (uint64_t)g.lda2, BM)) return FB_ERR_CUDA;
  } else {
    ma2 = ma;
  }
  if (!tc_make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)K, (uint64_t)K, BNH)) return FB_ERR_CUDA;
  Params p;
  p.M = g.M; p.N = g.N; p.KB1 = g.K1 / BK; p.KB2 = g.K2 / BK; p.nprod = g.nprod; p.exact_act = g.exact_act ? 1 : 0;
  p.bias = g.bias; p.act = g.act; p.C = g.C; p.ldc = g.ldc; p.Cb = (bf16*)g.Cb; p.ldcb = g.ldcb;
  p.dotv = g.dotv; p.dot_out = g.dot_out; p.dot_stride = g.dot_stride; p.n_split = g.n_split; p.m_dev = g.m_dev; p.drop = g.drop; p.dbg = g_tc_dbg;
  p.w_static = g.w_static ? 1 : 0;
  int pairs;
  if (wstat) {
    int ppn = (num_sms / 2) / n_tiles_n;       // pairs per column tile
    if (ppn > n_rb) ppn = n_rb;
    pairs = ppn * n_tiles_n;
  } else {
    const int tiles = n_rb * n_tiles_n;
    pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  }
  fb_launch(kern, dim3(2 * pairs), dim3(THREADS), smem_bytes, st, ma, ma2, mw, p);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace tc4

int gemm_tc2_bn(int M, int N);

// FB_ERR_UNSUPPORTED -> the caller goes on to the v3 kernel.  FB_TC4=0 disables the CTA-pair kernel (A/B comparisons, -DFB_DIAG builds).
int gemm_tc4_launch(const GemmArgs& g, cudaStream_t st) {
#ifdef FB_DIAG
  static const bool on = [] { const char* e = getenv("FB_TC4"); return !(e && atoi(e) == 0); }();
#else
  const bool on = true;
#endif
  if (!on || g.M <= 0) return FB_ERR_UNSUPPORTED;
  if (gemm_tc2_bn(g.M, g.N) != 256 || g.res || (g.C && g.Cb)) return FB_ERR_UNSUPPORTED;
  if (g.m_dev && (g.C || g.Cb)) return FB_ERR_UNSUPPORTED;
  if (g.n_split > 0 && (g.n_split % 64)) return FB_ERR_UNSUPPORTED;
  // 256-bit row stores: 32-byte aligned rows
  if (g.C && ((g.ldc % 8) || ((uintptr_t)g.C & 31))) return FB_ERR_UNSUPPORTED;
  if (g.Cb && ((g.nprod ? (g.ldcb % 8) : (g.ldcb % 16)) || ((uintptr_t)g.Cb & 31))) return FB_ERR_UNSUPPORTED;
  return tc4::launch(g, st);
}

}  // namespace fb
