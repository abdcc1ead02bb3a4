// Cross-attention core of the RowAttentionBlocks (FABind/fabind/models/cross_att.py:118-134, model_utils.py:21-38,96-133) on the
// 5th-generation tensor cores:  O[q, h*32+d] = sigmoid(G) * softmax_j(q_h . k_jh / sqrt(32) + pair_bias[q, j, h]) v_jh
// over the padded per-complex blocks (queries = one side of a complex, keys / values = the other side), 4 heads x 32 channels.
//
// One CTA = (complex, tile of 128 query rows).  Operands are the bf16 outputs of the stacked projection GEMMs:
//   * Q [128 x 128] and K [KT*128 x 128] tiles arrive by TMA (128B swizzle) and are used directly as tcgen05 operands: per head the
//     32 channels are a K-offset of 64 bytes inside the swizzle atom (the same descriptor arithmetic as a GEMM's 16-element k-steps);
//   * V is transposed through registers into a K-major [channel x key] tile (thread = key row), the B operand of P V;
//   * S_h = Q_h K_h^T (M = 128, N = keys padded to 32, K = 32) accumulates in TMEM; the softmax runs with thread = query row straight
//     off tcgen05.ld (no cross-thread reduction), adds the gated pair bias, writes the un-normalised probabilities as a bf16 A
//     operand into shared memory; O_h = P_h V_h (N = 32) accumulates in TMEM; the epilogue scales by 1 / l and sigmoid(G).
// Heads are processed in rounds of floor(256 / keys_padded) so that S never needs more than 256 TMEM columns (O takes 128 more).
// Warps 0-3: V transpose, softmax, epilogue (TMEM lane quarter = warp); warp 4: barriers, TMEM allocation, TMA and MMA issue.
// Key lists longer than 256 (whole proteins of the pocket stage) and compound sides above 128 rows stay on the SIMT kernel
// (layers.cu::row_attention_kernel), as do the fp32 / split-precision parity modes.
#include "gemm.h"
#include "layers.h"
#include "tc_common.cuh"

namespace fb {

bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace xa {
using namespace tc;

constexpr int NH = 4, DH = 32;
constexpr int PANEL = 128 * 128;          // bytes of one operand panel: 128 rows x 64 bf16 (128 B per row), 128B swizzle
constexpr int THREADS = 160;
constexpr int S_COL = 0, O_COL = 256;     // TMEM columns: scores [0, 256), outputs [256, 384)

struct Params {
  const int* c_off; const int* p_off; const int* pair_base;
  int q_is_prot, Nc;
  const bf16* QG; int ldqg, qcol, gcol;   // query-side projections, rows = side-local node index
  const bf16* KV; int ldkv, vcol;         // key-side projections (K tiles come through the tensor map)
  int kcol;
  const float* PB;                        // [pair rows, 4] gated pair bias of this block
  bf16* O; int ldo;                       // [N, 128], rows = internal node id
};

template <int KT>
struct Smem {
  static constexpr int QS = 0;
  static constexpr int KS = QS + 2 * PANEL;
  static constexpr int VT = KS + 2 * KT * PANEL;
  static constexpr int PS = VT + 2 * KT * PANEL;
  static constexpr int BAR = PS + 4 * PANEL;          // qk, v, s, p, o barriers + tmem slot
  static constexpr int TOTAL = BAR + 64 + 1024;       // + alignment slack
};

__device__ __forceinline__ uint32_t sw_off(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int KT>
__global__ void __launch_bounds__(THREADS, 1)
row_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const Params p) {
  using S = Smem<KT>;
  pdl_trigger();
  const int b = blockIdx.y, tile = blockIdx.x;
  // the layout arrays are inputs of the whole forward (host-built), not products of the previous kernel: safe before pdl_wait
  const int c_lo = p.c_off[b], nc1 = p.c_off[b + 1] - c_lo, p_lo = p.p_off[b], np1 = p.p_off[b + 1] - p_lo;
  const int n_q = p.q_is_prot ? np1 : nc1, n_k = p.q_is_prot ? nc1 : np1;
  const int q_lo = p.q_is_prot ? p_lo : c_lo, k_lo = p.q_is_prot ? c_lo : p_lo;
  if (tile * 128 >= n_q) return;
  const int q_side = p.q_is_prot ? p.Nc : 0, k_side = p.q_is_prot ? 0 : p.Nc;
  const int qrow0 = q_lo - q_side + tile * 128;          // first query row in the query-side buffer
  const int krow0 = k_lo - k_side;
  const int nk_pad = (n_k + 31) & ~31;
  const int hpr = nk_pad <= 64 ? 4 : (nk_pad <= 128 ? 2 : 1);   // heads per round
  const int rounds = NH / hpr;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Qs = smem + S::QS;
  uint8_t* Ks = smem + S::KS;
  uint8_t* Vt = smem + S::VT;
  uint8_t* Ps = smem + S::PS;
  uint64_t* bar_qk = (uint64_t*)(smem + S::BAR);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_slot = (uint32_t*)(bar_qk + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_k) : "memory");
      mbar_init(bar_qk, 1); mbar_init(bar_v, 128); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                              // the projections are written by the previous kernels of the stream

  if (warp == 4) {
    if (lane == 0) {
      // ---- operands: Q tile (2 panels of 64 channels) and the K tiles of the complex
      mbar_expect_tx(bar_qk, (2 + 2 * KT) * PANEL);
      for (int pa = 0; pa < 2; ++pa) tma_load_2d(&map_q, bar_qk, Qs + pa * PANEL, p.qcol + 64 * pa, qrow0);
      for (int kt = 0; kt < KT; ++kt)
        for (int pa = 0; pa < 2; ++pa) tma_load_2d(&map_k, bar_qk, Ks + (kt * 2 + pa) * PANEL, p.kcol + 64 * pa, krow0 + 128 * kt);
      mbar_wait(bar_qk, 0);
      tcgen05_fence_after();
      auto issue_scores = [&](int rd) {
        for (int hh = 0; hh < hpr; ++hh) {
          const int h = rd * hpr + hh;
          for (int kt = 0; kt < KT; ++kt) {
            const int n_kt = min(128, nk_pad - 128 * kt);
            if (n_kt <= 0) break;
            const uint64_t ad = make_smem_desc(Qs + (h >> 1) * PANEL) + (uint64_t)((h & 1) * 4);
            const uint64_t bd = make_smem_desc(Ks + (kt * 2 + (h >> 1)) * PANEL) + (uint64_t)((h & 1) * 4);
            const uint32_t d = tmem_base + (uint32_t)(S_COL + hh * nk_pad + 128 * kt);
#pragma unroll
            for (int ks = 0; ks < DH / UMMA_K; ++ks) umma_bf16(d, ad + 2 * ks, bd + 2 * ks, idesc_n(n_kt), ks != 0);
          }
        }
        umma_commit(bar_s);
      };
      issue_scores(0);
      mbar_wait(bar_v, 0);                 // V^T staged by the compute warps
      for (int rd = 0; rd < rounds; ++rd) {
        mbar_wait(bar_p, rd & 1);          // probabilities of this round are in shared memory, the score columns are free
        tcgen05_fence_after();
        for (int hh = 0; hh < hpr; ++hh) {
          const int h = rd * hpr + hh;
          const uint32_t d = tmem_base + (uint32_t)(O_COL + h * DH);
          for (int ks = 0; ks < nk_pad / UMMA_K; ++ks) {
            const int col0 = hh * nk_pad + UMMA_K * ks, key0 = UMMA_K * ks;
            const uint64_t ad = make_smem_desc(Ps + (col0 >> 6) * PANEL) + (uint64_t)(((col0 & 63) * 2) >> 4);
            const uint64_t bd = make_smem_desc(Vt + (key0 >> 6) * PANEL + h * DH * 128) + (uint64_t)(((key0 & 63) * 2) >> 4);
            umma_bf16(d, ad, bd, idesc_n(DH), ks != 0);
          }
        }
        umma_commit(bar_o);
        if (rd + 1 < rounds) issue_scores(rd + 1);
      }
    }
    __syncwarp();
  } else {
    // ---- compute warps: thread = key row (V transpose), then thread = query row
    const int r = threadIdx.x;             // 0..127
    for (int kt = 0; kt < KT; ++kt) {
      const int key = 128 * kt + r;
      uint32_t w[64];                      // the 128 bf16 values of this key row, two per word
      if (key < n_k) {
        const uint4* src = reinterpret_cast<const uint4*>(p.KV + (size_t)(krow0 + key) * p.ldkv + p.vcol);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint4 t = __ldg(src + i);
          w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) w[i] = 0u;
      }
      uint8_t* pan = Vt + (kt * 2 + (r >> 6)) * PANEL;
      const int kk = r & 63;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        const uint16_t e = (uint16_t)((c & 1) ? (w[c >> 1] >> 16) : (w[c >> 1] & 0xFFFFu));
        *reinterpret_cast<uint16_t*>(pan + c * 128 + (((kk >> 3) ^ (c & 7)) << 4) + (kk & 7) * 2) = e;
      }
    }
    fence_async_smem();
    mbar_arrive(bar_v);

    const int q_loc = tile * 128 + r;
    const bool q_ok = q_loc < n_q;
    const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const size_t pb_base = (size_t)p.pair_base[b] + (p.q_is_prot ? (size_t)q_loc * nc1 : (size_t)q_loc);
    const size_t pb_step = p.q_is_prot ? 1 : (size_t)nc1;
    float inv_l[NH];
    for (int rd = 0; rd < rounds; ++rd) {
      mbar_wait(bar_s, rd & 1);
      tcgen05_fence_after();
      if (rd > 0) mbar_wait(bar_o, (rd - 1) & 1);      // P V of the previous round has read the probability tile
      for (int hh = 0; hh < hpr; ++hh) {
        const int h = rd * hpr + hh;
        const uint32_t scol = tlane + (uint32_t)(S_COL + hh * nk_pad);
        float mx = -INFINITY;
        for (int c = 0; c < nk_pad / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(scol + 32 * c, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = 32 * c + j;
            if (key < n_k) {
              const float bias = q_ok ? __ldg(p.PB + (pb_base + key * pb_step) * 4 + h) : 0.f;
              mx = fmaxf(mx, fmaf(__uint_as_float(v[j]), scale, bias));
            }
          }
        }
        float l = 0.f;
        for (int c = 0; c < nk_pad / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(scol + 32 * c, v);
          float pr[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = 32 * c + j;
            float e = 0.f;
            if (key < n_k) {
              const float bias = q_ok ? __ldg(p.PB + (pb_base + key * pb_step) * 4 + h) : 0.f;
              e = __expf(fmaf(__uint_as_float(v[j]), scale, bias) - mx);
            }
            pr[j] = e;
            l += e;
          }
          const int col0 = hh * nk_pad + 32 * c;
          uint8_t* pan = Ps + (col0 >> 6) * PANEL;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(pr[8 * i], pr[8 * i + 1]), t1 = __floats2bfloat162_rn(pr[8 * i + 2], pr[8 * i + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(pr[8 * i + 4], pr[8 * i + 5]), t3 = __floats2bfloat162_rn(pr[8 * i + 6], pr[8 * i + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
            u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(pan + sw_off(r, ((col0 & 63) >> 3) + i)) = u;
          }
        }
        inv_l[h] = 1.0f / l;
      }
      tcgen05_fence_before();
      fence_async_smem();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (rounds - 1) & 1);
    tcgen05_fence_after();
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      uint32_t v[32];
      tmem_ld32(tlane + (uint32_t)(O_COL + h * DH), v);
      if (q_ok) {
        const uint4* gsrc = reinterpret_cast<const uint4*>(p.QG + (size_t)(qrow0 + r) * p.ldqg + p.gcol + h * DH);
        uint4* dst = reinterpret_cast<uint4*>(p.O + (size_t)(q_lo + q_loc) * p.ldo + h * DH);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 gq = __ldg(gsrc + i);
          const bf16* gb = reinterpret_cast<const bf16*>(&gq);
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[8 * i + j]) * inv_l[h] * sigmoidf(__bfloat162float(gb[j]));
          uint4 u;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(o[0], o[1]), t1 = __floats2bfloat162_rn(o[2], o[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(o[4], o[5]), t3 = __floats2bfloat162_rn(o[6], o[7]);
          u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
          u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
          dst[i] = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace xa

bool row_attention_tc_supported(int max_q, int max_k) { return max_q > 0 && max_k > 0 && max_k <= 256; }

// QG / KV: the bf16 outputs of the stacked projection GEMMs of the two sides (q_rows / k_rows rows, side-local row index);
// qcol / gcol / kcol / vcol: first column of Q, the gate, K and V (multiples of 64).  O: [N, 128] bf16, rows = internal node id.
int row_attention_tc(const GraphDev& g, int q_is_prot, int max_q, int max_k, const void* QG, int ldqg, int qcol, int gcol, int q_rows,
                     const void* KV, int ldkv, int kcol, int vcol, int k_rows, const float* PB, void* O, int ldo, cudaStream_t st) {
  if (!row_attention_tc_supported(max_q, max_k)) return FB_ERR_UNSUPPORTED;
  if ((qcol | gcol | kcol | vcol) & 63 || (ldqg & 7) || (ldkv & 7) || (ldo & 7)) return FB_ERR_BAD_ARG;
  CUtensorMap mq, mk;
  if (!tc_make_map(&mq, QG, (uint64_t)q_rows, (uint64_t)ldqg, (uint64_t)ldqg, 128)) return FB_ERR_CUDA;
  if (!tc_make_map(&mk, KV, (uint64_t)k_rows, (uint64_t)ldkv, (uint64_t)ldkv, 128)) return FB_ERR_CUDA;
  xa::Params p;
  p.c_off = g.c_off; p.p_off = g.p_off; p.pair_base = g.pair_base;
  p.q_is_prot = q_is_prot; p.Nc = g.Nc_tot;
  p.QG = (const bf16*)QG; p.ldqg = ldqg; p.qcol = qcol; p.gcol = gcol;
  p.KV = (const bf16*)KV; p.ldkv = ldkv; p.vcol = vcol; p.kcol = kcol;
  p.PB = PB; p.O = (bf16*)O; p.ldo = ldo;
  const dim3 grid((max_q + 127) / 128, g.B);
  static unsigned long long opt1 = 0, opt2 = 0;
  if (max_k <= 128) {
    if (!ensure_smem_optin(xa::row_attention_tc_kernel<1>, xa::Smem<1>::TOTAL, opt1)) return FB_ERR_CUDA;
    fb_launch(xa::row_attention_tc_kernel<1>, grid, dim3(xa::THREADS), xa::Smem<1>::TOTAL, st, mq, mk, p);
  } else {
    if (!ensure_smem_optin(xa::row_attention_tc_kernel<2>, xa::Smem<2>::TOTAL, opt2)) return FB_ERR_CUDA;
    fb_launch(xa::row_attention_tc_kernel<2>, grid, dim3(xa::THREADS), xa::Smem<2>::TOTAL, st, mq, mk, p);
  }
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
