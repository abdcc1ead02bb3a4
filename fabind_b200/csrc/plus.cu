// Kernels that only the FABind+ weight layout needs (FABind_plus/fabind/models/model_utils.py:10-74: every MLP is
// LayerNorm -> Linear -> ReLU -> Linear (-> ReLU); models/cross_att.py:43-45: the pair embedding is updated on ALL
// pair rows and carried to the next layer).  The LayerNorm in front of a per-edge MLP is folded through the node-level
// hoisting of its first Linear:
//     W1 LN(z) + b1 = rstd_e * ( (W1*gamma) z  -  mu_e * (W1*gamma) 1 ) + (W1 beta + b1)
// with z = [h_row | h_col | radial]: (W1*gamma) z is a sum of two per-NODE projections (one node GEMM) plus a rank-1
// radial term, and the per-EDGE statistics mu_e / rstd_e come from per-node sums of h and h^2.
// All kernels: one warp per row/edge, 16-byte accesses along the feature dimension.  T = float (fp32 parity mode) / bf16.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "layers.h"

namespace fb {

static inline int warp_grid_p(long long n_rows) { return (int)((n_rows * 32 + 255) / 256); }

// ------------------------------------------------------------------------------------------------
// per-row statistics: out[r] = { sum_f x, sum_f x^2, sum_f x*w }   (w optional)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void row_stats_kernel(const T* __restrict__ x, int ld, int M, int H, const float* __restrict__ w, float* __restrict__ out) {
  pdl_entry();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= M) return;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int f = lane * 4; f < H; f += 128) {
    const float4 v = ld4(x + (size_t)r * ld + f);
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (w) { const float4 q = ld4(w + f); s3 += (v.x * q.x + v.y * q.y) + (v.z * q.z + v.w * q.w); }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
  if (lane == 0) { out[3 * r] = s1; out[3 * r + 1] = s2; out[3 * r + 2] = s3; }
}

int row_stats(const void* x, int ld, int M, int H, const float* w, float* out, bool typed_bf16, cudaStream_t st) {
  if (M <= 0) return FB_OK;
  if (H & 3) return FB_ERR_UNSUPPORTED;
  if (typed_bf16) fb_launch(row_stats_kernel<bf16>, dim3(warp_grid_p(M)), dim3(256), 0, st, (const bf16*)x, ld, M, H, w, out);
  else fb_launch(row_stats_kernel<float>, dim3(warp_grid_p(M)), dim3(256), 0, st, (const float*)x, ld, M, H, w, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the concatenation [x1 (H1 cols, type T1) | x2 (H2 cols, type T2)] of each row -> out (type TO)
// (torch.nn.LayerNorm: biased variance, eps inside the square root).  Two-pass in registers.  x2 may be null.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;   // up to 8 * 128 = 1024 features per row
template <typename T1, typename T2, typename TO>
__global__ void ln_rows_kernel(const T1* __restrict__ x1, int ld1, int H1, const T2* __restrict__ x2, int ld2, int H2, int M,
                               const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                               TO* __restrict__ out, int ldo) {
  pdl_entry();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= M) return;
  const int D = H1 + H2;
  float4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int f = (k * 32 + lane) * 4;
    if (f < D) {
      v[k] = f < H1 ? ld4(x1 + (size_t)r * ld1 + f) : ld4(x2 + (size_t)r * ld2 + (f - H1));
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float mu = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int f = (k * 32 + lane) * 4;
    if (f < D) {
      const float a = v[k].x - mu, b = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int f = (k * 32 + lane) * 4;
    if (f < D) {
      const float4 gm = ld4(gamma + f), bt = ld4(beta + f);
      st4(out + (size_t)r * ldo + f, make_float4(fmaf((v[k].x - mu) * rstd, gm.x, bt.x), fmaf((v[k].y - mu) * rstd, gm.y, bt.y),
                                                 fmaf((v[k].z - mu) * rstd, gm.z, bt.z), fmaf((v[k].w - mu) * rstd, gm.w, bt.w)));
    }
  }
}

// x1: fp32 (x1_typed == false) or typed; x2 (optional): typed; out: typed
int ln_rows(const void* x1, bool x1_typed, int ld1, int H1, const void* x2, int ld2, int H2, int M, const float* gamma,
            const float* beta, float eps, void* out, int ldo, bool bf16_mode, cudaStream_t st) {
  if (M <= 0) return FB_OK;
  if ((H1 & 3) || (H2 & 3) || H1 + H2 > LN_MAXV * 128) return FB_ERR_UNSUPPORTED;
  const dim3 grid(warp_grid_p(M)), blk(256);
  if (!bf16_mode) {
    fb_launch(ln_rows_kernel<float, float, float>, grid, blk, 0, st, (const float*)x1, ld1, H1, (const float*)x2, ld2, H2, M, gamma, beta, eps, (float*)out, ldo);
  } else if (x1_typed) {
    fb_launch(ln_rows_kernel<bf16, bf16, bf16>, grid, blk, 0, st, (const bf16*)x1, ld1, H1, (const bf16*)x2, ld2, H2, M, gamma, beta, eps, (bf16*)out, ldo);
  } else {
    fb_launch(ln_rows_kernel<float, bf16, bf16>, grid, blk, 0, st, (const float*)x1, ld1, H1, (const bf16*)x2, ld2, H2, M, gamma, beta, eps, (bf16*)out, ldo);
  }
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// FABind+ GCL edge stage 1 (P/models/egnn.py:44-57 with MLPwithLastAct, model_utils.py:32-53):
//   A1[e, j] = ReLU( rstd_e * (Pr[row, j] + Pc[col, j] + rn_e * w_rad[j] - mu_e * g[j]) + c0[j] ),   j < Dp
// mu_e / rstd_e: LayerNorm statistics of [h_row | h_col | rn_e] (D = 2H+1 values) from the per-node sums.
// P is [N, 2*Dp] (row part | col part), Dp = D rounded up to 64 (padded columns are zero in every operand).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gcl_edge_pre_plus_kernel(int E, int H, int Dp, const int* __restrict__ erow, const int* __restrict__ ecol,
                                         const int* __restrict__ node_cplx, const T* __restrict__ P, const float* __restrict__ hstat,
                                         const float* __restrict__ rad, const float* __restrict__ norm,
                                         const float* __restrict__ w_rad, const float* __restrict__ gsum, const float* __restrict__ c0,
                                         float eps, T* __restrict__ A1, DropCfg dc, const int* __restrict__ emap) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  const int e = warp, r = erow[e], c = ecol[e];
  const float rn = rad[emap ? emap[e] : e] / radial_norm(norm, node_cplx[r]);
  const float invD = 1.0f / (float)(2 * H + 1);
  const float mu = (hstat[3 * r] + hstat[3 * c] + rn) * invD;
  // E[(x-mu)^2] = E[x^2] - mu^2, accumulated in fp32 from the per-node sums
  const float var = fmaxf((hstat[3 * r + 1] + hstat[3 * c + 1] + rn * rn) * invD - mu * mu, 0.f);
  const float rstd = rsqrtf(var + eps);
  const T* pr = P + (size_t)r * 2 * Dp;
  const T* pc = P + (size_t)c * 2 * Dp + Dp;
  for (int f = lane * 8; f < Dp; f += 256) {
    float a[8], b[8], w[8], gg[8], cc[8], o[8];
    ld8(pr + f, a); ld8(pc + f, b); ld8(w_rad + f, w); ld8(gsum + f, gg); ld8(c0 + f, cc);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaxf(fmaf(rstd, a[i] + b[i] + fmaf(rn, w[i], -mu * gg[i]), cc[i]), 0.f);
    if (dc.p > 0.f) {   // edge_mlp.dropout1 (P/models/model_utils.py:50); padded columns are zero either way
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = drop_apply(o[i], dc, e, f + i);
    }
    st8(A1 + (size_t)e * Dp + f, o);
  }
}

int gcl_edge_pre_plus(int E, int H, int Dp, const int* erow, const int* ecol, const int* node_cplx, const void* P,
                      const float* hstat, const float* rad, const float* norm, const float* w_rad, const float* gsum,
                      const float* c0, float eps, void* A1, bool bf16_mode, cudaStream_t st, DropCfg drop, const int* emap) {
  if (E <= 0) return FB_OK;
  if (Dp & 7) return FB_ERR_UNSUPPORTED;
  if (bf16_mode) fb_launch(gcl_edge_pre_plus_kernel<bf16>, dim3(warp_grid_p(E)), dim3(256), 0, st, E, H, Dp, erow, ecol, node_cplx, (const bf16*)P, hstat, rad, norm, w_rad, gsum, c0, eps, (bf16*)A1, drop, emap);
  else fb_launch(gcl_edge_pre_plus_kernel<float>, dim3(warp_grid_p(E)), dim3(256), 0, st, E, H, Dp, erow, ecol, node_cplx, (const float*)P, hstat, rad, norm, w_rad, gsum, c0, eps, (float*)A1, drop, emap);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// FABind+ pair-transition input (P/models/cross_att.py:43-45):
//   Zl[pair, :] = LayerNorm( pair[pair, :] + W_o32 (p32[prot] * c32[comp]) + b_o32 )       for EVERY pair row
// One warp per pair row; W_o32 ([H, 32], fp32) staged in shared memory transposed to [32][H] so that a lane reads its
// features with 16-byte accesses.  pc32: [N, ld32] with the 32 interaction channels of protein rows at column 0 and
// of compound rows at column 32 (the stacked q|k GEMM writes them there).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_complex_p(const int* __restrict__ pair_base, int B, int pair) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pair_base[mid] <= pair) lo = mid; else hi = mid;
  }
  return lo;
}

// R pair rows per warp share every shared-memory read of W_o32 (the kernel is bound by those reads, not by HBM).
constexpr int PZ_R = 4;
// FULL: H == VEC * 128 exactly, so every "feature < H" guard is compiled out (no BSSY/BSYNC pairs inside the k loop).
template <typename T, int VEC, bool FULL>
__global__ void __launch_bounds__(256) pair_zin_plus_kernel(GraphDev g, int P_total, int H_rt, const T* __restrict__ pair,
                                                            const float* __restrict__ pc32, int ld32, const float* __restrict__ Wo /*[32,H] (pre-transposed)*/,
                                                            const float* __restrict__ bo, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, T* __restrict__ Zl) {
  const int H = FULL ? VEC * 128 : H_rt;
  extern __shared__ float zs[];   // [32][H] W_o32^T
  for (int i = threadIdx.x * 4; i < H * 32; i += blockDim.x * 4) *reinterpret_cast<float4*>(zs + i) = ld4(Wo + i);
  pdl_entry();
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const float invH = 1.0f / (float)H;
  for (int pr0 = warp * PZ_R; pr0 < P_total; pr0 += n_warps * PZ_R) {
    float t[PZ_R];
    float4 z[PZ_R][VEC];
#pragma unroll
    for (int r = 0; r < PZ_R; ++r) {
      const int pr = min(pr0 + r, P_total - 1);   // tail rows recompute the last row (never stored)
      const int b = find_complex_p(g.pair_base, g.B, pr);
      const int nc1 = g.c_off[b + 1] - g.c_off[b];
      const int loc = pr - g.pair_base[b];
      const int pi = g.p_off[b] + loc / nc1, ci = g.c_off[b] + loc % nc1;
      t[r] = pc32[(size_t)pi * ld32 + lane] * pc32[(size_t)ci * ld32 + 32 + lane];   // lane = interaction channel
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int f = (i * 32 + lane) * 4;
        if (FULL || f < H) {
          const float4 p4 = ld4(pair + (size_t)pr * H + f), b4 = ld4(bo + f);
          z[r][i] = make_float4(p4.x + b4.x, p4.y + b4.y, p4.z + b4.z, p4.w + b4.w);
        }
      }
    }
#pragma unroll 2
    for (int k = 0; k < 32; ++k) {
      float tk[PZ_R];
#pragma unroll
      for (int r = 0; r < PZ_R; ++r) tk[r] = __shfl_sync(0xffffffffu, t[r], k);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int f = (i * 32 + lane) * 4;
        if (FULL || f < H) {
          const float4 w4 = *reinterpret_cast<const float4*>(zs + k * H + f);
#pragma unroll
          for (int r = 0; r < PZ_R; ++r) {
            z[r][i].x = fmaf(tk[r], w4.x, z[r][i].x); z[r][i].y = fmaf(tk[r], w4.y, z[r][i].y);
            z[r][i].z = fmaf(tk[r], w4.z, z[r][i].z); z[r][i].w = fmaf(tk[r], w4.w, z[r][i].w);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < PZ_R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int f = (i * 32 + lane) * 4;
        if (FULL || f < H) s += (z[r][i].x + z[r][i].y) + (z[r][i].z + z[r][i].w);
      }
      const float mu = warp_sum(s) * invH;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int f = (i * 32 + lane) * 4;
        if (FULL || f < H) {
          const float a = z[r][i].x - mu, bq = z[r][i].y - mu, c = z[r][i].z - mu, d = z[r][i].w - mu;
          q += (a * a + bq * bq) + (c * c + d * d);
        }
      }
      const float rstd = rsqrtf(warp_sum(q) * invH + eps);
      if (pr0 + r < P_total) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const int f = (i * 32 + lane) * 4;
          if (FULL || f < H) {
            const float4 gm = ld4(gamma + f), bt = ld4(beta + f);
            st4(Zl + (size_t)(pr0 + r) * H + f,
                make_float4(fmaf((z[r][i].x - mu) * rstd, gm.x, bt.x), fmaf((z[r][i].y - mu) * rstd, gm.y, bt.y),
                            fmaf((z[r][i].z - mu) * rstd, gm.z, bt.z), fmaf((z[r][i].w - mu) * rstd, gm.w, bt.w)));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 mode, H % 32 == 0: the same operation with the 32-channel product W_o32 t on the tensor cores (mma.sync m16n8k16, bf16
// operands, fp32 accumulation): the SIMT kernel above spends 16k FMA per pair row (1.6 GFMA per launch at B = 16) plus the
// shared-memory reads of W_o32 and ran at 139 us per launch, 17 % of a FABind+ forward; here a warp owns 16 pair rows, the
// product costs 8 mma per 32 output features, and the kernel is bound by its HBM traffic (pair read, Zl written).
// Two passes over the feature chunks: pass 1 accumulates the LayerNorm statistics of z = pair + W_o32 t + b, pass 2 recomputes z
// (the pair row comes back from L1/L2) and stores the normalised row -- z itself (512 values x 2 rows per lane) never has to live
// in registers.  Feature <-> MMA column mapping: tile j (8 columns), column n: feature (j/4)*32 + (n/2)*8 + (j%4)*2 + (n%2), so
// that the accumulator fragments of four consecutive tiles give every lane 8 CONSECUTIVE features (16-byte pair loads / stores).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(256) pair_zin_plus_mma_kernel(GraphDev g, int P_total, int H, const bf16* __restrict__ pair,
                                                                const float* __restrict__ pc32, int ld32,
                                                                const float* __restrict__ WoT /*[32,H]*/, const float* __restrict__ bo,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float eps, bf16* __restrict__ Zl) {
  extern __shared__ __align__(16) unsigned char pz_smem[];
  const int n_tiles = H >> 3;
  uint2* bfrag = reinterpret_cast<uint2*>(pz_smem);                      // [n_tiles][2 k-steps][32 lanes]
  float* sb = reinterpret_cast<float*>(bfrag + n_tiles * 64);             // bo | gamma | beta, H floats each
  for (int i = threadIdx.x; i < n_tiles * 64; i += blockDim.x) {
    const int ln = i & 31, s = (i >> 5) & 1, j = i >> 6;
    const int gq = ln >> 2, t = ln & 3;
    const int f = (j >> 2) * 32 + (gq >> 1) * 8 + (j & 3) * 2 + (gq & 1);   // feature of MMA column n = gq of tile j
    const int k0 = 16 * s + 2 * t;
    bfrag[i] = make_uint2(pack_bf16x2(WoT[(size_t)k0 * H + f], WoT[(size_t)(k0 + 1) * H + f]),
                          pack_bf16x2(WoT[(size_t)(k0 + 8) * H + f], WoT[(size_t)(k0 + 9) * H + f]));
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sb[i] = bo[i]; sb[H + i] = gamma[i]; sb[2 * H + i] = beta[i]; }
  pdl_entry();
  __syncthreads();
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const float invH = 1.0f / (float)H;
  const int n_chunks = H >> 5;
  for (int pr0 = warp * 16; pr0 < P_total; pr0 += n_warps * 16) {
    // this lane's two pair rows (fragment rows gq and gq + 8) and their interaction operand t = p32[prot] * c32[comp]
    uint32_t a[2][4];
    const bf16* prow[2];
    int rows[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pr = min(pr0 + gq + 8 * h, P_total - 1);
      rows[h] = pr;
      const int b = find_complex_p(g.pair_base, g.B, pr);
      const int nc1 = g.c_off[b + 1] - g.c_off[b];
      const int loc = pr - g.pair_base[b];
      const float* pp = pc32 + (size_t)(g.p_off[b] + loc / nc1) * ld32;
      const float* cc = pc32 + (size_t)(g.c_off[b] + loc % nc1) * ld32 + 32;
      prow[h] = pair + (size_t)pr * H;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int k0 = 16 * s + 2 * t;
        const float2 p0 = *reinterpret_cast<const float2*>(pp + k0), c0 = *reinterpret_cast<const float2*>(cc + k0);
        const float2 p1 = *reinterpret_cast<const float2*>(pp + k0 + 8), c1 = *reinterpret_cast<const float2*>(cc + k0 + 8);
        a[s][h] = pack_bf16x2(p0.x * c0.x, p0.y * c0.y);          // a0 (row gq) / a1 (row gq + 8): k = k0, k0 + 1
        a[s][2 + h] = pack_bf16x2(p1.x * c1.x, p1.y * c1.y);      // a2 / a3: k = k0 + 8, k0 + 9
      }
    }
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
    float mu[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 2
      for (int m = 0; m < n_chunks; ++m) {
        float acc[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) { acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.f; }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) mma_bf16_16816(acc[jj], a[s], bfrag[((4 * m + jj) * 2 + s) * 32 + lane]);
        const int f0 = 32 * m + 8 * t;           // this lane's 8 consecutive features of the chunk
        const float4 b0 = *reinterpret_cast<const float4*>(sb + f0), b1 = *reinterpret_cast<const float4*>(sb + f0 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float pv[8];
          ld8(prow[h] + f0, pv);
          float z[8];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            z[2 * jj] = pv[2 * jj] + bb[2 * jj] + acc[jj][2 * h];
            z[2 * jj + 1] = pv[2 * jj + 1] + bb[2 * jj + 1] + acc[jj][2 * h + 1];
          }
          if (pass == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1[h] += z[i]; s2[h] = fmaf(z[i], z[i], s2[h]); }
          } else if (pr0 + gq + 8 * h < P_total) {
            const float4 g0 = *reinterpret_cast<const float4*>(sb + H + f0), g1 = *reinterpret_cast<const float4*>(sb + H + f0 + 4);
            const float4 t0 = *reinterpret_cast<const float4*>(sb + 2 * H + f0), t1 = *reinterpret_cast<const float4*>(sb + 2 * H + f0 + 4);
            const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf((z[i] - mu[h]) * rstd[h], gm[i], bt[i]);
            st8(Zl + (size_t)rows[h] * H + f0, o);
          }
        }
      }
      if (pass == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // the four lanes of a quad hold disjoint feature subsets of the same two rows
          s1[h] += __shfl_xor_sync(0xffffffffu, s1[h], 1); s1[h] += __shfl_xor_sync(0xffffffffu, s1[h], 2);
          s2[h] += __shfl_xor_sync(0xffffffffu, s2[h], 1); s2[h] += __shfl_xor_sync(0xffffffffu, s2[h], 2);
          mu[h] = s1[h] * invH;
          rstd[h] = rsqrtf(fmaxf(s2[h] * invH - mu[h] * mu[h], 0.f) + eps);
        }
      }
    }
  }
}

int pair_zin_plus(const GraphDev& g, int P_total, int H, const void* pair, const float* pc32, int ld32, const float* Wo,
                  const float* bo, const float* gamma, const float* beta, float eps, void* Zl, bool bf16_mode, cudaStream_t st) {
  if (P_total <= 0) return FB_OK;
  if ((H & 3) || H > 512) return FB_ERR_UNSUPPORTED;
#ifdef FB_DIAG
  static const bool use_mma = [] { const char* e = getenv("FB_PZ_MMA"); return !(e && atoi(e) == 0); }();
#else
  const bool use_mma = true;
#endif
  if (bf16_mode && (H & 31) == 0 && use_mma) {
    const int smem_m = (H >> 3) * 64 * 8 + 3 * H * 4;
    const int grid_m = std::max(1, std::min(148 * 4, (P_total + 127) / 128));
    static unsigned long long done_m = 0;
    if (!ensure_smem_optin(pair_zin_plus_mma_kernel, 512 / 8 * 64 * 8 + 3 * 512 * 4, done_m)) return FB_ERR_CUDA;
    fb_launch(pair_zin_plus_mma_kernel, dim3(grid_m), dim3(256), smem_m, st, g, P_total, H, (const bf16*)pair, pc32, ld32, Wo, bo, gamma,
              beta, eps, (bf16*)Zl);
    count_launch(1);
    FB_CHECK_LAUNCH();
    return FB_OK;
  }
  const int smem = H * 32 * 4;
  const int grid = std::max(1, std::min(148 * 2, (P_total + 8 * PZ_R - 1) / (8 * PZ_R)));
  static unsigned long long done[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define FB_PZ_(T, VEC, FULL, slot)                                                                              \
  do {                                                                                                          \
    if (!ensure_smem_optin(pair_zin_plus_kernel<T, VEC, FULL>, 512 * 32 * 4, done[slot])) return FB_ERR_CUDA;    \
    fb_launch(pair_zin_plus_kernel<T, VEC, FULL>, dim3(grid), dim3(256), smem, st, g, P_total, H, (const T*)pair, pc32, ld32, Wo, bo, \
              gamma, beta, eps, (T*)Zl);                                                                        \
  } while (0)
#define FB_PZ(T, VEC, slot) do { if (H == (VEC) * 128) FB_PZ_(T, VEC, true, slot); else FB_PZ_(T, VEC, false, (slot) + 1); } while (0)
  if (bf16_mode) { if (H <= 128) FB_PZ(bf16, 1, 0); else if (H <= 256) FB_PZ(bf16, 2, 2); else FB_PZ(bf16, 4, 4); }
  else { if (H <= 128) FB_PZ(float, 1, 6); else if (H <= 256) FB_PZ(float, 2, 8); else FB_PZ(float, 4, 10); }
#undef FB_PZ_
#undef FB_PZ
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// pb_dense[pair] = sum of the row-dot partial tiles + constant, for every pair row (attn_bias_proj on the new pair embedding)
__global__ void pair_bias_all_kernel(int P_total, const float* __restrict__ dot, int tiles, int stride, const float* __restrict__ cst,
                                     float* __restrict__ pb_dense) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P_total) return;
  float s = 0.f;
  for (int t = 0; t < tiles; ++t) s += dot[(size_t)t * stride + i];
  pb_dense[i] = s + cst[0];
}

int pair_bias_all(int P_total, const float* dot, int tiles, int stride, const float* cst, float* pb_dense, cudaStream_t st) {
  if (P_total <= 0) return FB_OK;
  fb_launch(pair_bias_all_kernel, dim3((P_total + 255) / 256), dim3(256), 0, st, P_total, dot, tiles, stride, cst, pb_dense);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// in-place dropout of the residual stream (fp32) and its typed copy: MCAttEGNN `h = self.dropout(h)` in front of linear_out
// (P/models/egnn.py:428)
template <typename T>
__global__ void dropout_rows_kernel(float* __restrict__ x, T* __restrict__ xT, int M, int H, DropCfg dc) {
  pdl_entry();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= M) return;
  for (int f = lane * 4; f < H; f += 128) {
    float4 v = ld4(x + (size_t)r * H + f);
    v.x = drop_apply(v.x, dc, r, f); v.y = drop_apply(v.y, dc, r, f + 1);
    v.z = drop_apply(v.z, dc, r, f + 2); v.w = drop_apply(v.w, dc, r, f + 3);
    st4(x + (size_t)r * H + f, v);
    if (xT && (void*)xT != (void*)x) st4(xT + (size_t)r * H + f, v);
  }
}

int dropout_rows(float* x, void* xT, int M, int H, bool bf16_mode, DropCfg drop, cudaStream_t st) {
  if (M <= 0 || drop.p <= 0.f) return FB_OK;
  if (H & 3) return FB_ERR_UNSUPPORTED;
  if (bf16_mode) fb_launch(dropout_rows_kernel<bf16>, dim3(warp_grid_p(M)), dim3(256), 0, st, x, (bf16*)xT, M, H, drop);
  else fb_launch(dropout_rows_kernel<float>, dim3(warp_grid_p(M)), dim3(256), 0, st, x, (float*)xT, M, H, drop);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// packed pair rows [P_total, H] (typed) -> the reference's dense zero-padded [B, max_p, max_c, H] fp32 block
// (to_dense_batch layout, P/models/att_model.py:223); the caller zero-fills `out` first
template <typename T>
__global__ void pair_unpack_kernel(GraphDev g, int P_total, int H, int max_p, int max_c, const T* __restrict__ pair, float* __restrict__ out) {
  pdl_entry();
  const int pr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (pr >= P_total) return;
  const int b = find_complex_p(g.pair_base, g.B, pr);
  const int nc1 = g.c_off[b + 1] - g.c_off[b];
  const int loc = pr - g.pair_base[b];
  const int pi = loc / nc1, ci = loc % nc1;
  float* dst = out + (((size_t)b * max_p + pi) * max_c + ci) * H;
  for (int f = lane * 4; f < H; f += 128) st4(dst + f, ld4(pair + (size_t)pr * H + f));
}

int pair_unpack(const GraphDev& g, int P_total, int H, int max_p, int max_c, const void* pair, float* out, bool bf16_mode, cudaStream_t st) {
  if (P_total <= 0) return FB_OK;
  if (bf16_mode) fb_launch(pair_unpack_kernel<bf16>, dim3(warp_grid_p(P_total)), dim3(256), 0, st, g, P_total, H, max_p, max_c, (const bf16*)pair, out);
  else fb_launch(pair_unpack_kernel<float>, dim3(warp_grid_p(P_total)), dim3(256), 0, st, g, P_total, H, max_p, max_c, (const float*)pair, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
