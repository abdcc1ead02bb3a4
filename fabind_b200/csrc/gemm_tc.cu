// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue([A|A2][M,K] * W[N,K]^T), bf16 operands, fp32 accumulation.
//
//  * operands reach shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), 64-wide K slabs,
//    multi-stage mbarrier ring (full/empty);
//  * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16) with the accumulator in
//    TMEM; tcgen05.commit releases smem stages and signals the epilogue;
//  * four epilogue warps read the accumulator with tcgen05.ld (each warp owns its 32-lane TMEM
//    quarter = 32 output rows), apply bias / activation / residual, and write fp32 and/or bf16 rows,
//    or reduce the row against a vector (fused Linear(H,1) heads) without storing the tile.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue.
#include <cuda.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm.h"
#include "tc_common.cuh"

namespace fb {

namespace tc {

constexpr int NUM_THREADS = 192;

struct Params {
  int M, N, KB1, KB2;      // K slabs (of 64) taken from A and from A2
  const int* m_dev;
  const float* bias; int act;
  const float* res; int ldres;
  float* C; int ldc;
  bf16* Cb; int ldcb;
  const float* dotv; float* dot_out; int dot_stride;
  long long* dbg;          // optional phase timestamps (globaltimer ns), 8 per sampled CTA
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int VEC_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 16;       // staged bias[BN] | dotv[BN]
  static constexpr int TOTAL = VEC_OFF + 2 * BN * 4 + 1024;                 // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                               const __grid_constant__ CUtensorMap map_a2,
                                                               const __grid_constant__ CUtensorMap map_w, Params p) {
  using S = Smem<BN, STAGES>;
  pdl_entry();
  int M = p.M;
  if (p.m_dev) M = min(M, *p.m_dev);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;  // uniform for the whole CTA, before any barrier/TMEM state exists

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
  float* s_bias = (float*)(smem + S::VEC_OFF);
  float* s_dot = s_bias + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.KB1 + p.KB2;
  if (threadIdx.x == 0) FB_DBG(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_w) : "memory");
    if (p.KB2) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a2) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // TMEM: BN fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) FB_DBG(1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        mbar_expect_tx(&full[s], S::STAGE_BYTES);
        if (kb < p.KB1) tma_load_2d(&map_a, &full[s], a_dst, kb * BK, m0);
        else tma_load_2d(&map_a2, &full[s], a_dst, (kb - p.KB1) * BK, m0);
        tma_load_2d(&map_w, &full[s], b_dst, kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        if (kb == 0) FB_DBG(2);
        tcgen05_fence_after();
        const uint8_t* a_src = smem + s * S::STAGE_BYTES;
        const uint64_t adesc = make_smem_desc(a_src), bdesc = make_smem_desc(a_src + S::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing 16 bf16 (32 bytes) along K inside the swizzle atom = +2 in the 16-byte address field
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[s]);   // frees the smem stage once the MMAs above have read it
      }
      FB_DBG(3);
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), +32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int m = m0 + row;
    // stage the per-column vectors of this N tile in shared memory while the main loop runs
    for (int t = threadIdx.x - 64; t < BN; t += 128) {
      s_bias[t] = p.bias ? p.bias[n0 + t] : 0.f;
      s_dot[t] = p.dotv ? p.dotv[n0 + t] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
    mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) FB_DBG(4);
    tcgen05_fence_after();
    float dsum = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (m < M) {
        const int n = n0 + c;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + s_bias[c + j];
          // bf16 mode: fast exp/reciprocal (outputs are rounded to bf16 or feed fp32 sums at ~1e-6 rel)
          if (p.act == FB_ACT_SILU) x = silu_fast(x);
          else if (p.act == FB_ACT_RELU) x = fmaxf(x, 0.0f);
          o[j] = x;
        }
        if (p.res) {
          const float* r = p.res + (size_t)m * p.ldres + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 rv = *reinterpret_cast<const float4*>(r + j);
            o[j] += rv.x; o[j + 1] += rv.y; o[j + 2] += rv.z; o[j + 3] += rv.w;
          }
        }
        if (p.dotv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dsum = fmaf(s_dot[c + j], o[j], dsum);
        }
        if (p.C) {
          float* cp = p.C + (size_t)m * p.ldc + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
        if (p.Cb) {
          bf16* cb = p.Cb + (size_t)m * p.ldcb + n;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(o[j], o[j + 1]), t1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), t3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(cb + j) = u;
          }
        }
      }
    }
    if (p.dotv && m < M) p.dot_out[(size_t)blockIdx.x * p.dot_stride + m] = dsum;
    if (threadIdx.x == 64) FB_DBG(5);
    tcgen05_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) FB_DBG(6);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

// Encoded descriptors are cached: the operand / output buffers of a forward are the same every layer and iteration, and the driver
// call costs ~1 us of host time per descriptor (up to 24 per multi-problem launch).  Key = everything the descriptor depends on.
struct MapKey {
  const void* ptr; uint64_t rows, cols, ld; uint32_t kind;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && kind == o.kind; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full; h ^= h >> 29;
    h ^= (k.cols * 0x165667B19E3779F9ull) ^ (k.ld << 17) ^ ((uint64_t)k.kind << 51);
    return (size_t)(h ^ (h >> 32));
  }
};
template <typename F>
static bool cached_map(CUtensorMap* m, const MapKey& key, F encode) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *m = it->second; return true; }
  }
  if (!encode(m)) return false;
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *m);
  return true;
}

// 2-D bf16 tensor [rows, cols] with row stride ld (elements); box = box_rows x 64 columns, 128B swizzle
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  return cached_map(m, MapKey{ptr, rows, cols, ld, box_rows}, [&](CUtensorMap* out) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * 2};
    const cuuint32_t box[2] = {BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  });
}

constexpr int BN_SEL = 128, STAGES_SEL = 3;

}  // namespace tc

bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return tc::make_map(m, ptr, rows, cols, ld, box_rows);
}

// output / residual tiles of the v3 epilogue: box = 32 rows x 128 bytes (32 fp32 or 64 bf16 columns), 128B swizzle
bool tc_make_map_out(CUtensorMap* m, const void* ptr, bool is_f32, uint64_t rows, uint64_t cols, uint64_t ld) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  return tc::cached_map(m, tc::MapKey{ptr, rows, cols, ld, is_f32 ? 0x80000001u : 0x80000002u}, [&](CUtensorMap* out) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld * (is_f32 ? 4u : 2u)};
    const cuuint32_t box[2] = {is_f32 ? 32u : 64u, 32u};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  });
}

bool gemm_tc_shape_ok(int N, int K) { return N >= tc::BN_SEL && (N % tc::BN_SEL) == 0 && K >= 64 && (K % 64) == 0; }

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.n_split > 0 && (g.n_split % 128)) return false;
  if (g.ldw != 0 && (g.ldw % 8)) return false;            // (a strided W: only gemm_tc5.cu builds its tensor map with ldw)
  if (!gemm_tc_shape_ok(g.N, g.K1 + g.K2)) return false;
  if ((g.K1 % 64) || (g.K2 % 64) || g.K1 <= 0) return false;
  if ((g.lda % 8) || ((uintptr_t)g.A & 15) || ((uintptr_t)g.W & 15)) return false;
  if (g.A2 && ((g.lda2 % 8) || ((uintptr_t)g.A2 & 15))) return false;
  if (g.C && ((g.ldc % 4) || ((uintptr_t)g.C & 15))) return false;
  if (g.Cb && ((g.ldcb % 8) || ((uintptr_t)g.Cb & 15))) return false;
  if (g.res && ((g.ldres % 4) || ((uintptr_t)g.res & 15))) return false;
  return tc::encode_fn() != nullptr;
}

int gemm_tc_dot_tiles(int N) { return N / tc::BN_SEL; }

template <int BN, int STAGES>
static int launch_cfg(const GemmArgs& g, cudaStream_t st);

long long* g_tc_dbg = nullptr;   // set through fb_gemm_set_debug (development probe)

int gemm_tc_launch(const GemmArgs& g, cudaStream_t st) {
#ifdef FB_DIAG
  static int stages = [] { const char* e = getenv("FB_TC_STAGES"); return e ? atoi(e) : tc::STAGES_SEL; }();
#else
  const int stages = tc::STAGES_SEL;
#endif
  if (stages == 4) return launch_cfg<tc::BN_SEL, 4>(g, st);
  if (stages == 6) return launch_cfg<tc::BN_SEL, 6>(g, st);
  return launch_cfg<tc::BN_SEL, 3>(g, st);
}

template <int BN, int STAGES>
static int launch_cfg(const GemmArgs& g, cudaStream_t st) {
  using namespace tc;
  if (g.M <= 0) return FB_OK;
  using S = Smem<BN, STAGES>;
  static unsigned long long optin = 0;
  if (!ensure_smem_optin(gemm_tc_kernel<BN, STAGES>, S::TOTAL, optin)) return FB_ERR_CUDA;
  CUtensorMap ma, ma2, mw;
  const int K = g.K1 + g.K2;
  if (!make_map(&ma, g.A, (uint64_t)g.M, (uint64_t)g.K1, (uint64_t)g.lda, BM)) return FB_ERR_CUDA;
  if (g.K2 > 0) {
    if (!make_map(&ma2, g.A2, (uint64_t)g.M, (uint64_t)g.K2, (uint64_t)g.lda2, BM)) return FB_ERR_CUDA;
  } else {
    ma2 = ma;
  }
  if (!make_map(&mw, g.W, (uint64_t)g.N, (uint64_t)K, (uint64_t)K, BN)) return FB_ERR_CUDA;
  Params p;
  p.M = g.M; p.N = g.N; p.KB1 = g.K1 / BK; p.KB2 = g.K2 / BK; p.m_dev = g.m_dev;
  p.bias = g.bias; p.act = g.act; p.res = g.res; p.ldres = g.ldres; p.C = g.C; p.ldc = g.ldc;
  p.Cb = (bf16*)g.Cb; p.ldcb = g.ldcb; p.dotv = g.dotv; p.dot_out = g.dot_out; p.dot_stride = g.dot_stride;
  p.dbg = g_tc_dbg;
  dim3 grid(g.N / BN, (g.M + BM - 1) / BM);
  fb_launch(gemm_tc_kernel<BN, STAGES>, dim3(grid), dim3(NUM_THREADS), S::TOTAL, st, ma, ma2, mw, p);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
