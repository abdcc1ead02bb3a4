// tcgen05 GEMM v4: CTA PAIRS (cta_group::2) for the long edge-level / pair-level GEMMs.
//
// Why: at a 128x256 tile per CTA (v3) every 64-wide k-slab needs (128 + 256) x 64 x 2 B = 48 KB of operands per 0.27 us of
// MMA time; chip-wide that is ~26 TB/s out of L2, about twice what the L2 -> SM fabric delivers, and v3 stalls at ~53 % of the
// measured bf16 peak with DRAM traffic equal to the algorithmic bytes.  Two CTAs of one TPC computing
// ONE 256x256 tile with tcgen05.mma.cta_group::2 each stage their own 128 rows of A and only HALF of the B (weight) slab:
// 32 KB per CTA and k-slab for the same MMA time, a third less operand traffic.
//
// Structure (per CTA, same warp roles as v3): warp 0 = TMA producer (its 128 A rows + its 128 of the 256 W rows per k-slab, with
// the completion bytes of BOTH CTAs signalled on the LEADER's "full" barrier), warp 1 = MMA issuer (leader CTA only; one
// tcgen05.mma.cta_group::2 per 16-wide k-step, M = 256, N = 256; tcgen05.commit multicast releases the smem stage / publishes the
// accumulator in both CTAs), warps 2-9 = epilogue exactly as v3 (each CTA drains its own 128 TMEM lanes; TMA-store staging
// boxes).  The leader's "accumulator drained" barrier counts the epilogue warps of both CTAs (remote mbarrier arrive).
// Persistent over 256x256 tiles; two TMEM accumulator stages (2 x 256 columns).  No residual, one stored output (the shapes this
// kernel is picked for never need more).
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

extern long long* g_tc_dbg;
bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
bool tc_make_map_out(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t rows, uint64_t cols, uint64_t ld);

namespace tc4 {
using namespace tc;

constexpr int BN = 256;                         // tile columns (the pair's MMA N)
constexpr int BNH = BN / 2;                     // W rows staged per CTA
constexpr int PM = 2 * BM;                      // tile rows of the pair
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;   // 320
constexpr int SLOT = 4096;                      // one staging box: 32 rows x 128 bytes
constexpr int NS = 2;
constexpr int STAGES = 4;
constexpr int VEC = BN / 2;                     // columns per epilogue warp: its bias / row-dot vectors live in shared memory

struct Params {
  int M, N, KB1, KB2;
  int nprod, exact_act;     // split-precision mode (see gemm.h): products per k-block, KB1 = k-blocks of ONE plane
  const float* bias; int act;
  int has_c, has_cb;
  const float* dotv; float* dot_out; int dot_stride;
  int n_split;
  const int* m_dev;
  DropCfg drop;
};

struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BNH * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;                       // 32 KB
  static constexpr int STAGING_OFF = STAGES * STAGE_BYTES;
  static constexpr int VEC_OFF = STAGING_OFF + EPI_WARPS * NS * SLOT;         // per epilogue warp: bias[VEC] | dotv[VEC] (fp32)
  static constexpr int BAR_OFF = VEC_OFF + EPI_WARPS * 2 * VEC * 4;           // full[S] empty[S] tfull[2] tempty[2] slot
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;    // + alignment slack
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t sw_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
// four consecutive floats of a read-only vector (16-byte load when the address allows it)
__device__ __forceinline__ float4 ldg4(const float* p) {
  if (((uintptr_t)p & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// SPLIT: split-precision instantiation (table-driven k-block walk + full-precision SiLU, see gemm_tc3.cu); ACT: activation of the
// epilogue as a template parameter, so that one instantiation carries one straight-line epilogue (the runtime select between three
// unrolled variants tripled the code the eight epilogue warps walk through: 18 % of their stall samples were instruction fetches,
// profiles/r2f_tc4_epilogue_before.txt)
template <bool SPLIT, int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
                const __grid_constant__ CUtensorMap map_cb, const Params p) {
  using S = Smem;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int KB = SPLIT ? p.nprod * p.KB1 : p.KB1 + p.KB2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
    if (p.KB2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a2) : "memory");
  }
  if (warp == 2 && lane == 0) {
    if (p.has_c) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_c) : "memory");
    if (p.has_cb) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_cb) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      // tempty of the leader collects the epilogue warps of BOTH CTAs
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
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
  pdl_wait();
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int n_tiles_n = p.N / BN;
  const int n_tiles = ((M + PM - 1) / PM) * n_tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int m0 = (tile / n_tiles_n) * PM + (int)rank * BM;
        const int n0 = (tile % n_tiles_n) * BN + (int)rank * BNH;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
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
          tma_load_2d_pair(&map_w, &full[s], b_dst, kb * BK, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
      int it = 0, lt = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
        const int a = lt & 1;
        mbar_wait(&tempty[a], ((lt >> 1) & 1) ^ 1);   // both epilogues have drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tcgen05_fence_after();
          const uint8_t* a_src = smem + s * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_src), bdesc = make_smem_desc(a_src + S::A_BYTES);
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
    const int e = warp - 2;                 // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // column half of the tile
    constexpr int COLS = BN / 2;            // columns per warp
    constexpr int NCH = COLS / 32;          // 32-column pieces per warp and tile
    uint8_t* const slots = smem + S::STAGING_OFF + e * NS * SLOT;
    float* const vb = reinterpret_cast<float*>(smem + S::VEC_OFF) + e * 2 * VEC;
    float* const vd = vb + VEC;
    const bool use_dot = p.dotv != nullptr;
    constexpr bool HALF_BIAS = ACT == FB_ACT_SILU && !SPLIT;
    int lt = 0, vec_n0 = -1;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
      const int a = lt & 1;
      const int m0 = (tile / n_tiles_n) * PM + (int)rank * BM;
      const int n0 = (tile % n_tiles_n) * BN;
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
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      mbar_wait(&tfull[a], (lt >> 1) & 1);
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
        const int ncol0 = n0 + c;
        const bool want_c = p.has_c && !(p.n_split > 0 && ncol0 >= p.n_split);
        const bool want_cb = p.has_cb && !(p.n_split > 0 && ncol0 < p.n_split);
        uint8_t* const fs = slots + (ch & 1) * SLOT;
        uint8_t* const bs = slots + ((ch >> 1) & 1) * SLOT;
        if (want_c) {
          if (lane == 0) bulk_wait_read1();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(fs + sw_off(lane, j)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&map_c, fs, ncol0, lrow0);
        }
        if (want_cb) {
          if ((ch & 1) == 0) { if (lane == 0) bulk_wait_read1(); __syncwarp(); }
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
            if (lane == 0) tma_store_2d(&map_cb, bs, ncol0 - 32 - (p.n_split > 0 ? p.n_split : 0), lrow0);
          }
        }
      }
      if (use_dot && lrow0 + lane < M) {
        // two warps (column halves) share a row: partial index = 2 * n_tile + half
        p.dot_out[(size_t)((tile % n_tiles_n) * 2 + half) * p.dot_stride + lrow0 + lane] = (ds0 + ds1) + (ds2 + ds3);
      }
    }
    if (lane == 0) bulk_wait_read0();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();               // no CTA of the pair leaves (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

static int launch(const GemmArgs& g, cudaStream_t st) {
  using S = Smem;
  static_assert(S::TOTAL <= 232448, "shared memory budget");
  static int num_sms = 0;
  const bool split = g.nprod > 0;
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, Params);
  static const Kern table[2][3] = {
      {gemm_tc4_kernel<false, FB_ACT_NONE>, gemm_tc4_kernel<false, FB_ACT_SILU>, gemm_tc4_kernel<false, FB_ACT_RELU>},
      {gemm_tc4_kernel<true, FB_ACT_NONE>, gemm_tc4_kernel<true, FB_ACT_SILU>, gemm_tc4_kernel<true, FB_ACT_RELU>}};
  static unsigned long long optins[2][3] = {{0, 0, 0}, {0, 0, 0}};
  if (g.act < 0 || g.act > 2) return FB_ERR_BAD_ARG;
  Kern kern = table[split ? 1 : 0][g.act];
  if (!ensure_smem_optin(kern, S::TOTAL, optins[split ? 1 : 0][g.act])) return FB_ERR_CUDA;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  CUtensorMap ma, ma2, mw, mc, mcb;
  const int K = g.nprod ? 3 * g.K1 : g.K1 + g.K2;   // split precision: three bf16 planes of K1 columns each
  if (!tc_make_map(&ma, g.A, (uint64_t)g.M, (uint64_t)(g.nprod ? K : g.K1), (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (g.K2 > 0) {
    if (!tc_make_map(&ma2, g.A2, (uint64_t)g.M, (uint64_t)g.K2, (uint64_t)g.lda2, BM)) return FB_ERR_CUDA;
  } else {
    ma2 = ma;
  }
  if (!tc_make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)K, (uint64_t)K, BNH)) return FB_ERR_CUDA;
  mc = mcb = ma;   // placeholders for absent operands (never dereferenced)
  const int nc = g.n_split > 0 ? g.n_split : g.N, ncb = g.n_split > 0 ? g.N - g.n_split : g.N;
  if (g.C && !tc_make_map_out(&mc, g.C, true, (uint64_t)g.M, (uint64_t)nc, (uint64_t)g.ldc)) return FB_ERR_CUDA;
  if (g.Cb && !tc_make_map_out(&mcb, g.Cb, false, (uint64_t)g.M, (uint64_t)ncb, (uint64_t)g.ldcb)) return FB_ERR_CUDA;
  Params p;
  p.M = g.M; p.N = g.N; p.KB1 = g.K1 / BK; p.KB2 = g.K2 / BK; p.nprod = g.nprod; p.exact_act = g.exact_act ? 1 : 0;
  p.bias = g.bias; p.act = g.act; p.has_c = g.C != nullptr; p.has_cb = g.Cb != nullptr;
  p.dotv = g.dotv; p.dot_out = g.dot_out; p.dot_stride = g.dot_stride; p.n_split = g.n_split; p.m_dev = g.m_dev; p.drop = g.drop;
  const int tiles = ((g.M + PM - 1) / PM) * (g.N / BN);
  const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  fb_launch(kern, dim3(2 * pairs), dim3(THREADS), S::TOTAL, st, ma, ma2, mw, mc, mcb, p);
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
  return tc4::launch(g, st);
}

}  // namespace fb
