// Cross-attention core of the RowAttentionBlocks (FABind/fabind/models/cross_att.py:118-134, model_utils.py:21-38,96-133) on the
// 5th-generation tensor cores:  O[q, h*32+d] = sigmoid(G) * softmax_j(q_h . k_jh / sqrt(32) + pair_bias[q, j, h]) v_jh
// over the padded per-complex blocks (queries = one side of a complex, keys / values = the other side), 4 heads x 32 channels.
//
// One CTA = (complex, tile of 128 query rows).  Operands are the bf16 outputs of the stacked projection GEMMs:
//   * K [KT*128 x 128] tiles arrive by TMA (128B swizzle) and are tcgen05 operands as they land;
//   * V is transposed through registers into a K-major [channel x key] tile, the B operand of P V;
//   * the A operand of the scores comes in two row layouts, chosen per CTA:
//       ROWS   (more than 32 queries: the protein side, large ligands): row = query; the Q tile arrives by TMA and a head's 32
//              channels are a K-offset of 64 bytes inside the swizzle atom; heads are processed in rounds of floor(256 / keys_padded);
//       PACKED (at most 32 queries: the usual compound side): row = (head, query) -- the four heads of every query fill the 128
//              TMEM lanes.  The A tile is block-diagonal (row (h, q) holds Q[q] in the channels of head h, zero elsewhere), so
//              ONE chain of K = 128 MMAs against the K tile gives all heads' scores, and one N = 128 MMA chain against V^T gives
//              every head's output in its diagonal 32-column block;
//   * S accumulates in TMEM; the softmax reads it with tcgen05.ld (thread = row), adds the gated pair bias, and writes the
//     un-normalised probabilities as a bf16 A operand into shared memory; O = P V accumulates in TMEM; the epilogue scales by
//     1 / l and sigmoid(G).
// 16 compute warps (TMEM lane quarter = warp & 3, slot = warp >> 2): a row's 32-column score chunks are dealt round-robin to the four
// slots; the per-chunk maxima and sums are exchanged through spare TMEM columns (tcgen05.st / ld), so the softmax costs each
// thread at most two chunks per round.  Warp 16: barriers, TMEM allocation, TMA and MMA issue.
// Key lists longer than 256 (whole proteins of the pocket stage) stay on the SIMT kernel (layers.cu::row_attention_kernel), as do
// the fp32 / split-precision parity modes.
#include "gemm.h"
#include "layers.h"
#include "tc_common.cuh"

namespace fb {

bool tc_make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

namespace xa {
using namespace tc;

constexpr int NH = 4, DH = 32;
constexpr int PANEL = 128 * 128;          // bytes of one operand panel: 128 rows x 64 bf16 (128 B per row), 128B swizzle
constexpr int CWARPS = 16, CTHREADS = CWARPS * 32, THREADS = CTHREADS + 32;
// TMEM columns: scores [0, 256), outputs [256, 384), chunk maxima [384, 392), chunk sums [392 + 8 * round, ...)
constexpr int S_COL = 0, O_COL = 256, RED_COL = 384, SUM_COL = 392;

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
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory"); }
// one fp32 per lane into / out of a TMEM column: the cross-warp exchange of softmax statistics
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
  u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
  return u;
}

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
  const bool packed = n_q <= 32;                          // row = (head, query): all four heads in one round
  const int hpr = packed ? 1 : (nk_pad <= 64 ? 4 : (nk_pad <= 128 ? 2 : 1));   // ROWS layout: heads per round (column blocks)
  const int rounds = packed ? 1 : NH / hpr;
  const int ncols = hpr * nk_pad;                         // score columns in use per round

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

  if (warp == CWARPS) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_q) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_k) : "memory");
      mbar_init(bar_qk, 1); mbar_init(bar_v, CTHREADS); mbar_init(bar_s, 1); mbar_init(bar_p, CTHREADS); mbar_init(bar_o, 1);
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

  if (warp == CWARPS) {
    if (lane == 0) {
      // ---- operands by TMA: the K tiles of the complex and (ROWS layout) the Q tile
      mbar_expect_tx(bar_qk, ((packed ? 0 : 2) + 2 * KT) * PANEL);
      if (!packed)
        for (int pa = 0; pa < 2; ++pa) tma_load_2d(&map_q, bar_qk, Qs + pa * PANEL, p.qcol + 64 * pa, qrow0);
      for (int kt = 0; kt < KT; ++kt)
        for (int pa = 0; pa < 2; ++pa) tma_load_2d(&map_k, bar_qk, Ks + (kt * 2 + pa) * PANEL, p.kcol + 64 * pa, krow0 + 128 * kt);
      mbar_wait(bar_qk, 0);
      mbar_wait(bar_v, 0);                 // V^T (and the block-diagonal A tile of the PACKED layout) staged by the compute warps
      tcgen05_fence_after();
      auto issue_scores = [&](int rd) {
        if (packed) {
          for (int kt = 0; kt < KT; ++kt) {
            const int n_kt = min(128, nk_pad - 128 * kt);
            if (n_kt <= 0) break;
            const uint32_t d = tmem_base + (uint32_t)(S_COL + 128 * kt);
#pragma unroll
            for (int ks = 0; ks < 128 / UMMA_K; ++ks) {
              const uint64_t ad = make_smem_desc(Qs + (ks >> 2) * PANEL) + (uint64_t)(2 * (ks & 3));
              const uint64_t bd = make_smem_desc(Ks + (kt * 2 + (ks >> 2)) * PANEL) + (uint64_t)(2 * (ks & 3));
              umma_bf16(d, ad, bd, idesc_n(n_kt), ks != 0);
            }
          }
        } else {
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
        }
        umma_commit(bar_s);
      };
      issue_scores(0);
      for (int rd = 0; rd < rounds; ++rd) {
        mbar_wait(bar_p, rd & 1);          // probabilities of this round are in shared memory, the score columns are free
        tcgen05_fence_after();
        if (packed) {
          // O[(h, q), (h', d)] = sum_j P[(h, q), j] V[j, (h', d)]: only the diagonal blocks h' = h are read back
          const uint32_t d = tmem_base + (uint32_t)O_COL;
          for (int ks = 0; ks < nk_pad / UMMA_K; ++ks) {
            const int key0 = UMMA_K * ks;
            const uint64_t ad = make_smem_desc(Ps + (key0 >> 6) * PANEL) + (uint64_t)(((key0 & 63) * 2) >> 4);
            const uint64_t bd = make_smem_desc(Vt + (key0 >> 6) * PANEL) + (uint64_t)(((key0 & 63) * 2) >> 4);
            umma_bf16(d, ad, bd, idesc_n(128), ks != 0);
          }
        } else {
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
        }
        umma_commit(bar_o);
        if (rd + 1 < rounds) issue_scores(rd + 1);
      }
    }
    __syncwarp();
  } else {
    const int t = threadIdx.x;             // 0 .. 511
    const int quarter = warp & 3, slot = warp >> 2;
    const int row = quarter * 32 + lane;   // TMEM lane = row of the score / output tiles
    // ---- stage V^T: thread = (key row, 32-channel block)
    {
      const int kr = t & 127, cb = t >> 7;
      for (int kt = 0; kt < KT; ++kt) {
        const int key = 128 * kt + kr;
        uint32_t w[16];                    // 32 bf16 channels of this key row, two per word
        if (key < n_k) {
          const uint4* src = reinterpret_cast<const uint4*>(p.KV + (size_t)(krow0 + key) * p.ldkv + p.vcol + 32 * cb);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = __ldg(src + i);
            w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = 0u;
        }
        uint8_t* pan = Vt + (kt * 2 + (kr >> 6)) * PANEL;
        const int kk = kr & 63;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int ch = 32 * cb + c;
          const uint16_t e = (uint16_t)((c & 1) ? (w[c >> 1] >> 16) : (w[c >> 1] & 0xFFFFu));
          *reinterpret_cast<uint16_t*>(pan + ch * 128 + (((kk >> 3) ^ (ch & 7)) << 4) + (kk & 7) * 2) = e;
        }
      }
    }
    // ---- PACKED layout: the block-diagonal A tile, row (h, q) = Q[q] in the channels of head h
    if (packed) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = t + CTHREADS * i;  // 2048 chunks of 16 bytes: (row, 8-channel chunk)
        const int r = idx >> 4, j = idx & 15;
        const int q = r & 31;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if ((j >> 2) == (r >> 5) && q < n_q)
          u = __ldg(reinterpret_cast<const uint4*>(p.QG + (size_t)(qrow0 + q) * p.ldqg + p.qcol + 8 * j));
        *reinterpret_cast<uint4*>(Qs + (j >> 3) * PANEL + sw_off(r, j & 7)) = u;
      }
    }
    fence_async_smem();
    mbar_arrive(bar_v);

    const int q_loc = packed ? (row & 31) : tile * 128 + row;
    const bool q_ok = q_loc < n_q;
    const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const size_t pb_base = (size_t)p.pair_base[b] + (p.q_is_prot ? (size_t)q_loc * nc1 : (size_t)q_loc);
    const size_t pb_step = p.q_is_prot ? 1 : (size_t)nc1;
    for (int rd = 0; rd < rounds; ++rd) {
      mbar_wait(bar_s, rd & 1);
      tcgen05_fence_after();
      if (rd > 0) mbar_wait(bar_o, (rd - 1) & 1);      // P V of the previous round has read the probability tile
      // this slot's score chunks: 32-column chunk cc = slot and slot + 4 of the round's columns
      float sv[2][32];
      float cmax[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cc = slot + 4 * u;
        cmax[u] = -INFINITY;
        if (32 * cc < ncols) {             // warp-uniform
          const int cb = (32 * cc) / nk_pad, key0 = 32 * cc - cb * nk_pad;
          const int h = packed ? (row >> 5) : rd * hpr + cb;
          uint32_t v[32];
          tmem_ld32(tlane + (uint32_t)(S_COL + 32 * cc), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = key0 + j;
            float s = -INFINITY;
            if (key < n_k) {
              const float bias = q_ok ? __ldg(p.PB + (pb_base + key * pb_step) * 4 + h) : 0.f;
              s = fmaf(__uint_as_float(v[j]), scale, bias);
            }
            sv[u][j] = s;
            cmax[u] = fmaxf(cmax[u], s);
          }
        }
        tmem_st1(tlane + (uint32_t)(RED_COL + cc), cmax[u]);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      compute_sync();
      tcgen05_fence_after();
      float red[8];
      tmem_ld8(tlane + (uint32_t)RED_COL, red);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cc = slot + 4 * u;
        float csum = 0.f;
        if (32 * cc < ncols) {
          const int cb = (32 * cc) / nk_pad;
          float m = -INFINITY;             // maximum over the chunks of the same column block (= the same head of this row)
#pragma unroll
          for (int c2 = 0; c2 < 8; ++c2)
            if (32 * c2 < ncols && (32 * c2) / nk_pad == cb) m = fmaxf(m, red[c2]);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = sv[u][j] == -INFINITY ? 0.f : __expf(sv[u][j] - m);
            sv[u][j] = e;                  // in place: the probabilities replace the scores
            csum += e;
          }
          const int col0 = 32 * cc;
          uint8_t* pan = Ps + (col0 >> 6) * PANEL;
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(pan + sw_off(row, ((col0 & 63) >> 3) + i)) = pack8(sv[u] + 8 * i);
        }
        tmem_st1(tlane + (uint32_t)(SUM_COL + 8 * rd + cc), csum);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      fence_async_smem();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (rounds - 1) & 1);
    tcgen05_fence_after();
    // ---- epilogue: this thread's 32 output channels = head `slot` (ROWS) / the diagonal block of the row's head (PACKED)
    const int h = slot;
    if (!packed || (row >> 5) == h) {      // warp-uniform: a warp's 32 rows share one head in the PACKED layout
      const int rd_h = packed ? 0 : h / hpr, cb_h = packed ? 0 : h % hpr;
      float sums[8];
      tmem_ld8(tlane + (uint32_t)(SUM_COL + 8 * rd_h), sums);
      float l = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < 8; ++c2)
        if (32 * c2 < ncols && (32 * c2) / nk_pad == cb_h) l += sums[c2];
      const float inv_l = 1.0f / l;
      uint32_t v[32];
      tmem_ld32(tlane + (uint32_t)(O_COL + h * DH), v);
      if (q_ok) {
        const uint4* gsrc = reinterpret_cast<const uint4*>(p.QG + (size_t)(qrow0 + (packed ? q_loc : row)) * p.ldqg + p.gcol + h * DH);
        uint4* dst = reinterpret_cast<uint4*>(p.O + (size_t)(q_lo + q_loc) * p.ldo + h * DH);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 gq = __ldg(gsrc + i);
          const bf16* gb = reinterpret_cast<const bf16*>(&gq);
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = __uint_as_float(v[8 * i + j]) * inv_l * sigmoidf(__bfloat162float(gb[j]));
          dst[i] = pack8(o);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == CWARPS) {
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
