// Persistent tcgen05 GEMM (v3):  C[M,N] = epilogue([A|A2][M,K] * W[N,K]^T), bf16 operands, fp32 accumulation.
//
// Main loop as in v2 (one CTA per SM looping over tiles; warp 0 = TMA producer into a smem ring that runs across
// tiles, warp 1 = tcgen05.mma issuer into two TMEM accumulator stages).  The epilogue (warps 2-9) is rebuilt
// around the "thread = row" TMEM layout instead of transposing it away:
//   * every lane owns one output row and 32 consecutive columns per tcgen05.ld; it writes them as 16-byte pieces
//     into a 128B-swizzled staging box in shared memory (conflict-free: 8 rows hit 8 different 16-byte columns);
//   * one elected lane per warp hands the box to TMA (cp.async.bulk.tensor store): full-line, asynchronous
//     global writes that overlap the next tile's main loop, rows beyond M clipped by the tensor map;
//   * the fp32 residual tile arrives the same way (TMA load issued at the start of the tile, hidden behind the
//     main loop), is updated in place in shared memory and stored from there -- no residual registers;
//   * bias / row-dot vectors live in registers (lane = column) and are broadcast by shuffles: no staging loop
//     in the prologue, so the only work between CTA start and the first TMA is barrier init + TMEM alloc.
#include <cstdlib>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

extern long long* g_tc_dbg;
bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
bool tc_make_map_out(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t rows, uint64_t cols, uint64_t ld);

namespace tc3 {
using namespace tc;

#define FB_DBG3(slot)                                                                     \
  do {                                                                                   \
    if (p.dbg && (blockIdx.x & 7) == 0) p.dbg[(blockIdx.x >> 3) * 8 + (slot)] = gtime(); \
  } while (0)

#define FB_DBGX(slot)                                                                                  \
  do {                                                                                                 \
    if (p.dbg && lt == 0 && threadIdx.x == 64 && (blockIdx.x & 7) == 0) p.dbg[2048 + (blockIdx.x >> 3) * 16 + (slot)] = gtime(); \
  } while (0)

constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + EPI_WARPS * 32;   // 320
constexpr int SLOT = 4096;                      // one staging box: 32 rows x 128 bytes

struct Prob {
  int M, N;                // rows / columns of this problem
  int m_begin;             // first row of this problem in the (shared) A buffer
  const float* bias; int act;
  int has_res, has_c, has_cb;
  const float* dotv; float* dot_out; int dot_stride;
  int n_split;
  int exact_act;
  DropCfg drop;
};
struct Params {
  Prob q0, q1;
  int KB1, KB2;
  int nprod;               // split-precision mode: products per k-block (0 = plain bf16); then KB1 = k-blocks of ONE plane
  const int* m_dev;
  long long* dbg;
};

// NS = staging boxes per epilogue warp: 3 (two fp32 boxes that double as residual landing zone + one bf16 box) for the
// 128-wide tiles, 1 (re-used box by box) for the 256-wide tiles, which never carry a residual
template <int BN, int STAGES, int NS>
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
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte piece j (0..7) of row r (0..31) inside a 128B-swizzled box
__device__ __forceinline__ uint32_t sw_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

// SPLIT: split-precision instantiation (table-driven k-block walk in the producer, full-precision SiLU in the epilogue); a template
// parameter so that the bf16 instantiation carries neither (a runtime select between the two SiLU forms inside the unrolled epilogue
// cost 2.3x on the edge GEMM)
template <int BN, int STAGES, int NS, bool GROUPED, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w2,
                const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_cb,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_c1,
                const __grid_constant__ CUtensorMap map_cb1, const __grid_constant__ CUtensorMap map_res1, const Params p) {
  using S = Smem<BN, STAGES, NS>;
  pdl_trigger();
  if (threadIdx.x == 0) FB_DBG3(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* resbar = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(resbar + EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = SPLIT ? p.nprod * p.KB1 : p.KB1 + p.KB2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
    if (p.KB2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a2) : "memory");
    if (GROUPED) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w2) : "memory");
  }
  if (warp == 2 && lane == 0) {
    if (p.q0.has_c) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_c) : "memory");
    if (p.q0.has_cb) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_cb) : "memory");
    if (p.q0.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_res) : "memory");
    if (GROUPED) {
      if (p.q1.has_c) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_c1) : "memory");
      if (p.q1.has_cb) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_cb1) : "memory");
      if (p.q1.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_res1) : "memory");
    }
  }
  if (warp == 1) {
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
  // everything above touched only on-chip state; activations (and the device-side row count) are produced by the
  // previous kernel in the stream
  if (threadIdx.x == 0) FB_DBG3(1);
  pdl_wait();
  int M = p.q0.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int n_tiles_n = p.q0.N / BN;
  const int tiles0 = ((M + BM - 1) / BM) * n_tiles_n;
  const int ntn1 = GROUPED ? p.q1.N / BN : 1;
  const int n_tiles = tiles0 + (GROUPED ? ((p.q1.M + BM - 1) / BM) * ntn1 : 0);
  // tile -> (problem, first row in the A buffer, first column)
  auto decode = [&](int tile, int& m0, int& n0) -> bool {
    if (!GROUPED || tile < tiles0) { m0 = (tile / n_tiles_n) * BM; n0 = (tile % n_tiles_n) * BN; return false; }
    const int t = tile - tiles0;
    m0 = p.q1.m_begin + (t / ntn1) * BM; n0 = (t % ntn1) * BN;
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
          if (SPLIT) {
            // split precision: k-block kb = product j of plane pair (pa, pw), smallest terms first
            const int j = kb / p.KB1, r = kb - j * p.KB1;
            tma_load_2d(&map_a, &full[s], a_dst, (split_plane_a(j, p.nprod) * p.KB1 + r) * BK, m0);
            tma_load_2d(&map_w, &full[s], b_dst, (split_plane_w(j, p.nprod) * p.KB1 + r) * BK, n0);
            continue;
          }
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
          if (it == 0) FB_DBG3(2);
          tcgen05_fence_after();
          const uint8_t* a_src = smem + s * S::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_src), bdesc = make_smem_desc(a_src + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        if (lt == 0) FB_DBG3(3);
        umma_commit(&tfull[a]);
      }
    }
  } else {
    // ===== epilogue =====
    const int e = warp - 2;                 // 0..7
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // column half of the tile
    constexpr int COLS = BN / 2;            // columns per warp
    constexpr int NCH = COLS / 32;          // 32-column pieces per warp and tile
    static_assert(NS == 1 || NS == 2 || (NS == 3 && NCH == 2), "staging layout");
    uint8_t* const slots = smem + S::STAGING_OFF + e * NS * SLOT;
    uint64_t* const rbar = &resbar[e];
    uint32_t rphase = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int a = lt & 1;
      int m0, n0;
      const bool second = decode(tile, m0, n0);
      const Prob& pp = (GROUPED && second) ? p.q1 : p.q0;
      const CUtensorMap* mc = (GROUPED && second) ? &map_c1 : &map_c;
      const CUtensorMap* mcb = (GROUPED && second) ? &map_cb1 : &map_cb;
      const CUtensorMap* mres = (GROUPED && second) ? &map_res1 : &map_res;
      const int m_rows = (GROUPED && second) ? pp.M : M;             // rows of this problem
      const int lrow0 = m0 - ((GROUPED && second) ? pp.m_begin : 0) + q * 32;   // first row of this warp, problem-local
      const bool rows_live = lrow0 < m_rows;                         // warp-uniform
      const int colbase = n0 + half * COLS;
      const bool has_res = NS == 3 && pp.has_res;
      const bool use_dot = pp.dotv != nullptr && !(GROUPED && second);
      float bv[NCH], dv[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        bv[ch] = pp.bias ? __ldg(pp.bias + colbase + ch * 32 + lane) : 0.f;
        dv[ch] = use_dot ? __ldg(pp.dotv + colbase + ch * 32 + lane) : 0.f;
      }
      // the staging boxes are free once every earlier bulk store of this warp has read its shared memory
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      if (has_res && rows_live && lane == 0) {
        mbar_expect_tx(rbar, NCH * SLOT);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) tma_load_2d(mres, rbar, slots + ch * SLOT, colbase + ch * 32, lrow0);
      }
      mbar_wait(&tfull[a], (lt >> 1) & 1);
      if (lt == 0 && threadIdx.x == 64) FB_DBG3(4);
      tcgen05_fence_after();
      if (has_res && rows_live) { mbar_wait(rbar, rphase); rphase ^= 1; }
      float dsum = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int c = half * COLS + ch * 32;      // column inside the tile
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c), v);
        FB_DBGX(ch * 6 + 0);
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
          if (pp.act == FB_ACT_SILU) x = SPLIT ? silu(x) : silu_fast(x);
          else if (pp.act == FB_ACT_RELU) x = fmaxf(x, 0.0f);
          o[j] = x;
        }
        if (pp.drop.p > 0.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = drop_apply(o[j], pp.drop, lrow0 + lane, n0 + c + j);
        }
        if (use_dot) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dsum = fmaf(__shfl_sync(0xffffffffu, dv[ch], j), o[j], dsum);
        }
        FB_DBGX(ch * 6 + 1);
        const int ncol0 = n0 + c;
        const bool want_c = pp.has_c && !(pp.n_split > 0 && ncol0 >= pp.n_split);
        const bool want_cb = pp.has_cb && !(pp.n_split > 0 && ncol0 < pp.n_split);
        // NS == 2 (256-wide tiles, one stored output): two boxes used alternately, so a store only waits for the
        // one issued two boxes earlier
        uint8_t* const fs = slots + (NS == 3 ? ch : NS == 2 ? (ch & 1) : 0) * SLOT;
        uint8_t* const bs = slots + (NS == 3 ? 2 : NS == 2 ? ((ch >> 1) & 1) : 0) * SLOT;
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(fs + sw_off(lane, j));
            o[4 * j] += r4.x; o[4 * j + 1] += r4.y; o[4 * j + 2] += r4.z; o[4 * j + 3] += r4.w;
          }
        }
        if (want_c) {
          if (NS == 1) { if (lane == 0) bulk_wait_read0(); __syncwarp(); }
          if (NS == 2) { if (lane == 0) bulk_wait_read1(); __syncwarp(); }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(fs + sw_off(lane, j)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          FB_DBGX(ch * 6 + 2);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(mc, fs, ncol0, lrow0);
          FB_DBGX(ch * 6 + 3);
        }
        if (want_cb) {
          if (NS == 1 && (ch & 1) == 0) { if (lane == 0) bulk_wait_read0(); __syncwarp(); }
          if (NS == 2 && (ch & 1) == 0) { if (lane == 0) bulk_wait_read1(); __syncwarp(); }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(o[8 * j], o[8 * j + 1]), t1 = __floats2bfloat162_rn(o[8 * j + 2], o[8 * j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(o[8 * j + 4], o[8 * j + 5]), t3 = __floats2bfloat162_rn(o[8 * j + 6], o[8 * j + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(bs + sw_off(lane, (ch & 1) * 4 + j)) = u;
          }
          FB_DBGX(ch * 6 + 4);
          if (ch & 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_2d(mcb, bs, ncol0 - 32 - (pp.n_split > 0 ? pp.n_split : 0), lrow0);
          }
          FB_DBGX(ch * 6 + 5);
        }
      }
      if (use_dot && lrow0 + lane < m_rows) {
        // two warps (column halves) share a row: partial index = 2 * n_tile + half
        pp.dot_out[(size_t)((tile % n_tiles_n) * 2 + half) * pp.dot_stride + lrow0 + lane] = dsum;
      }
      if (lt == 0 && threadIdx.x == 64) FB_DBG3(5);
    }
    // shared memory must stay valid until the last bulk stores have read it
    if (lane == 0) bulk_wait_read0();
    lt = 0;
    FB_DBGX(12);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FB_DBG3(6);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

static bool fill_prob(Prob& q, const GemmArgs& g, int m_begin, CUtensorMap* mc, CUtensorMap* mcb, CUtensorMap* mres) {
  q.M = g.M; q.N = g.N; q.m_begin = m_begin;
  q.bias = g.bias; q.act = g.act;
  q.has_res = g.res != nullptr; q.has_c = g.C != nullptr; q.has_cb = g.Cb != nullptr;
  q.dotv = g.dotv; q.dot_out = g.dot_out; q.dot_stride = g.dot_stride;
  q.n_split = g.n_split;
  q.exact_act = g.exact_act ? 1 : 0;
  q.drop = g.drop;
  const int nc = g.n_split > 0 ? g.n_split : g.N, ncb = g.n_split > 0 ? g.N - g.n_split : g.N;
  if (g.C && !tc_make_map_out(mc, g.C, true, (uint64_t)g.M, (uint64_t)nc, (uint64_t)g.ldc)) return false;
  if (g.Cb && !tc_make_map_out(mcb, g.Cb, false, (uint64_t)g.M, (uint64_t)ncb, (uint64_t)g.ldcb)) return false;
  if (g.res && !tc_make_map_out(mres, g.res, true, (uint64_t)g.M, (uint64_t)nc, (uint64_t)g.ldres)) return false;
  return true;
}

// g1 (optional): second problem on rows [m_begin1, m_begin1 + g1->M) of the same A buffer
template <int BN, int STAGES, int NS>
static int launch(const GemmArgs& g, const GemmArgs* g1, int m_begin1, cudaStream_t st) {
  using S = Smem<BN, STAGES, NS>;
  static_assert(S::TOTAL <= 232448, "shared memory budget");
  static unsigned long long optin = 0, optin_g = 0;
  static int num_sms = 0;
  static unsigned long long optin_s = 0;
  const bool split = g.nprod > 0;
  auto kern = split ? gemm_tc3_kernel<BN, STAGES, NS, false, true>
                    : (g1 ? gemm_tc3_kernel<BN, STAGES, NS, true, false> : gemm_tc3_kernel<BN, STAGES, NS, false, false>);
  if (split && g1) return FB_ERR_UNSUPPORTED;
  if (!ensure_smem_optin(kern, S::TOTAL, split ? optin_s : (g1 ? optin_g : optin))) return FB_ERR_CUDA;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  CUtensorMap ma, ma2, mw, mw2, mc, mcb, mres, mc1, mcb1, mres1;
  // split precision: A and W hold three bf16 planes of K1 columns each
  const int K = g.nprod ? 3 * g.K1 : g.K1 + g.K2;
  const int rows = g1 ? (m_begin1 + g1->M > g.M ? m_begin1 + g1->M : g.M) : g.M;
  if (!tc_make_map(&ma, g.A, (uint64_t)rows, (uint64_t)(g.nprod ? K : g.K1), (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (g.K2 > 0) {
    if (!tc_make_map(&ma2, g.A2, (uint64_t)g.M, (uint64_t)g.K2, (uint64_t)g.lda2, BM)) return FB_ERR_CUDA;
  } else {
    ma2 = ma;
  }
  if (!tc_make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)K, (uint64_t)K, BN)) return FB_ERR_CUDA;
  mc = mcb = mres = ma;   // placeholders for absent operands (never dereferenced)
  Params p;
  if (!fill_prob(p.q0, g, 0, &mc, &mcb, &mres)) return FB_ERR_CUDA;
  p.KB1 = g.K1 / BK; p.KB2 = g.K2 / BK; p.nprod = g.nprod; p.m_dev = g.m_dev; p.dbg = g_tc_dbg;
  int tiles = ((g.M + BM - 1) / BM) * (g.N / BN);
  mc1 = mcb1 = mres1 = ma;
  if (g1) {
    if (!tc_make_map(&mw2, g1->W, (uint64_t)g1->N, (uint64_t)K, (uint64_t)K, BN)) return FB_ERR_CUDA;
    if (!fill_prob(p.q1, *g1, m_begin1, &mc1, &mcb1, &mres1)) return FB_ERR_CUDA;
    tiles += ((g1->M + BM - 1) / BM) * (g1->N / BN);
  } else {
    mw2 = mw;
    p.q1 = p.q0;
    p.q1.M = 0;
  }
  const int grid = tiles < num_sms ? tiles : num_sms;
  fb_launch(kern, dim3(grid), dim3(THREADS), S::TOTAL, st, ma, ma2, mw, mw2, mc, mcb, mres, mc1, mcb1, mres1, p);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// tile width: 256 once there are enough rows to fill the machine with such tiles and the epilogue is a plain one
// (no residual, at most one stored output); 128 for the short, latency-bound node-level GEMMs
static int pick_bn(const GemmArgs& g) {
  return ((g.N % 256) == 0 && g.M >= 16384 && !g.res && !(g.C && g.Cb)) ? 256 : 128;
}

}  // namespace tc3

int gemm_tc2_bn(int M, int N);

// FB_ERR_UNSUPPORTED -> the caller falls back to the v2 kernel
// FB_TC3_MAXM (diagnostic): problems with more rows go to the v2 kernel.  With ONE staging box per epilogue warp the
// 256-wide tiles of v3 lost to v2 on the long edge-level GEMMs (5.2 vs 3.4 ms/step); with two alternating boxes v3
// is ahead there too (2.6 vs 3.0 ms/step), so the default is "no limit".
static int tc3_max_m() {
#ifdef FB_DIAG
  static int m = [] { const char* e = getenv("FB_TC3_MAXM"); return e ? atoi(e) : 1 << 30; }();
  return m;
#else
  return 1 << 30;
#endif
}

int gemm_tc3_launch(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0) return FB_OK;
  if (g.M > tc3_max_m()) return FB_ERR_UNSUPPORTED;
  if (g.m_dev && (g.C || g.Cb)) return FB_ERR_UNSUPPORTED;            // rows beyond a device-side count must stay untouched
  if (g.n_split > 0 && (g.n_split % 64)) return FB_ERR_UNSUPPORTED;
  const int bn = tc3::pick_bn(g);
  if (g.dotv && bn != gemm_tc2_bn(g.M, g.N)) return FB_ERR_UNSUPPORTED;  // the partial count is planned from (M, N) alone
  if (bn == 256) return tc3::launch<256, 3, 2>(g, nullptr, 0, st);
  return tc3::launch<128, 4, 3>(g, nullptr, 0, st);
}

int gemm_tc3_launch_pair(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t st) {
  if (g0.M <= 0 || g1.M <= 0 || g0.nprod || g1.nprod) return FB_ERR_UNSUPPORTED;
  if (g0.K2 || g1.K2 || g0.K1 != g1.K1 || g0.lda != g1.lda || g0.m_dev || g1.m_dev || g1.dotv) return FB_ERR_UNSUPPORTED;
  if ((g0.N % 128) || (g1.N % 128)) return FB_ERR_UNSUPPORTED;
  if ((g0.n_split % 64) || (g1.n_split % 64)) return FB_ERR_UNSUPPORTED;
  if (gemm_tc2_bn(g0.M, g0.N) != 128 || gemm_tc2_bn(g1.M, g1.N) != 128) return FB_ERR_UNSUPPORTED;
  const ptrdiff_t off = (const char*)g1.A - (const char*)g0.A;
  const ptrdiff_t row = (ptrdiff_t)g0.lda * 2;
  if (off < 0 || off % row) return FB_ERR_UNSUPPORTED;
  const ptrdiff_t m_begin1 = off / row;
  if (m_begin1 < g0.M) return FB_ERR_UNSUPPORTED;   // ranges must not overlap
  return tc3::launch<128, 4, 3>(g0, &g1, (int)m_begin1, st);
}

}  // namespace fb
