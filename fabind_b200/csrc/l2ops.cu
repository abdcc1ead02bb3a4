// Kernels for the L2 wrapper around the docking stack (reference: FABind/fabind/models/model.py):
// feature/coordinate assembly, LayerNorm, soft / hard pocket centre, pocket mask, ligand placement and the
// pairwise-distance head.  C ABI at the bottom (declared in include/fabind_b200.h).
#include "../../include/fabind_b200.h"
#include <type_traits>

#include "common.cuh"

namespace fb {

// out[i, :] = scale * src_{kind[i]}[idx[i], :]     (kind 0..3; used for [glb_c | atoms | glb_p | residues] assembly)
__global__ void assemble_rows_kernel(float* __restrict__ out, int M, int D, const uint8_t* __restrict__ kind,
                                     const int* __restrict__ idx, const float* __restrict__ s0, const float* __restrict__ s1,
                                     const float* __restrict__ s2, const float* __restrict__ s3, float scale) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int k = kind[warp];
  const float* src = k == 0 ? s0 : (k == 1 ? s1 : (k == 2 ? s2 : s3));
  const float* row = src ? src + (size_t)idx[warp] * D : nullptr;
  for (int f = lane; f < D; f += 32) out[(size_t)warp * D + f] = row ? scale * row[f] : 0.f;
}

// LayerNorm over the last dimension (torch.nn.LayerNorm, eps inside the sqrt, biased variance)
__global__ void layernorm_kernel(const float* __restrict__ x, int M, int D, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* __restrict__ out) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = x + (size_t)warp * D;
  float s = 0.f;
  for (int f = lane; f < D; f += 32) s += r[f];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int f = lane; f < D; f += 32) { const float d = r[f] - mean; v = fmaf(d, d, v); }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + eps);
  for (int f = lane; f < D; f += 32) out[(size_t)warp * D + f] = (r[f] - mean) * rstd * gamma[f] + beta[f];
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// Pocket centre per complex.  mode 0: model.forward, eval (model.py:146-158): weights = softmax over
// {log(1-p), log p} / tau of the clamped sigmoid (gumbel_softmax_no_random, utils.py:687-699), optional
// straight-through hard one-hot.  mode 1: model.inference (model.py:423-437): mean of the residues whose
// rounded sigmoid is 1, soft weights (unclamped) when none is.
// noise (optional, mode 0 only): [n_res, 2] gumbel samples added to {log(1-p), log p} before the softmax = F.gumbel_softmax of
// the FABind+ train()-mode forward (P/models/model.py:136-137)
__global__ void __launch_bounds__(256) pocket_center_kernel(const float* __restrict__ logit, const float* __restrict__ xyz,
                                                            const int* __restrict__ off, float tau, int hard, int mode,
                                                            float* __restrict__ centers, const float* __restrict__ noise) {
  pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.x, lo = off[b], hi = off[b + 1];
  float sw = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float p = 1.0f / (1.0f + expf(-logit[i]));
    float p1 = p, p0 = 1.0f - p;
    if (mode == 0) {
      p1 = fminf(fmaxf(p1, 1e-6f), 1.0f - 1e-6f);
      p0 = fminf(fmaxf(p0, 1e-6f), 1.0f - 1e-6f);
    }
    float l0 = logf(p0), l1 = logf(p1);
    if (noise) { l0 += noise[2 * i]; l1 += noise[2 * i + 1]; }
    l0 /= tau; l1 /= tau;
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    float w = e1 / (e0 + e1);
    if (hard) { const float yh = (l1 > l0) ? 1.0f : 0.0f; w = (yh - w) + w; }
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    sw += w; sx = fmaf(w, x, sx); sy = fmaf(w, y, sy); sz = fmaf(w, z, sz);
    if (mode == 1 && rintf(p) == 1.0f) { cnt += 1.f; hx += x; hy += y; hz += z; }
  }
  sw = block_sum(sw, red); sx = block_sum(sx, red); sy = block_sum(sy, red); sz = block_sum(sz, red);
  if (mode == 1) {
    cnt = block_sum(cnt, red); hx = block_sum(hx, red); hy = block_sum(hy, red); hz = block_sum(hz, red);
  }
  if (threadIdx.x == 0) {
    if (mode == 1 && cnt > 0.f) { centers[3 * b] = hx / cnt; centers[3 * b + 1] = hy / cnt; centers[3 * b + 2] = hz / cnt; }
    else { centers[3 * b] = sx / sw; centers[3 * b + 1] = sy / sw; centers[3 * b + 2] = sz / sw; }
  }
}

// keep[i] = |xyz_i - centre_b| < radius  with the reference's fp32 evaluation order
// sqrt((dx*dx + dy*dy) + dz*dz)  (utils.py:147-158: torch.sum over the last dim is (x+y)+z, no FMA);
// a complex with fewer than 5 kept residues keeps its first 100 instead (model.py:199-202).
__global__ void __launch_bounds__(256) pocket_mask_kernel(const float* __restrict__ xyz, const int* __restrict__ off,
                                                          const float* __restrict__ centers, float radius,
                                                          uint8_t* __restrict__ keep, int* __restrict__ less5) {
  pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.x, lo = off[b], hi = off[b + 1];
  const float cx = centers[3 * b], cy = centers[3 * b + 1], cz = centers[3 * b + 2];
  float n = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float dx = __fsub_rn(xyz[3 * i], cx), dy = __fsub_rn(xyz[3 * i + 1], cy), dz = __fsub_rn(xyz[3 * i + 2], cz);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const bool k = d < radius;
    keep[i] = k ? 1 : 0;
    n += k ? 1.f : 0.f;
  }
  n = block_sum(n, red);
  if (n < 5.f) {
    for (int i = lo + threadIdx.x; i < min(hi, lo + 100); i += blockDim.x) keep[i] = 1;
    if (threadIdx.x == 0) less5[b] = 1;
  } else if (threadIdx.x == 0) less5[b] = 0;
}

// ligand start pose (model.py:227): lig - mean(lig) + mean(kept pocket residues), per complex
__global__ void __launch_bounds__(256) ligand_place_kernel(const float* __restrict__ lig, const int* __restrict__ coff,
                                                           const float* __restrict__ pocket, const int* __restrict__ poff,
                                                           float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.x;
  float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = coff[b] + threadIdx.x; i < coff[b + 1]; i += blockDim.x) { m[0] += lig[3 * i]; m[1] += lig[3 * i + 1]; m[2] += lig[3 * i + 2]; }
  for (int i = poff[b] + threadIdx.x; i < poff[b + 1]; i += blockDim.x) { m[3] += pocket[3 * i]; m[4] += pocket[3 * i + 1]; m[5] += pocket[3 * i + 2]; }
  for (int k = 0; k < 6; ++k) m[k] = block_sum(m[k], red);
  const float nl = (float)(coff[b + 1] - coff[b]), np = (float)(poff[b + 1] - poff[b]);
  for (int i = coff[b] + threadIdx.x; i < coff[b + 1]; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) out[3 * i + k] = (lig[3 * i + k] - m[k] / nl) + m[3 + k] / np;
  }
}

__device__ __forceinline__ int seg_find(const int* __restrict__ base, int B, int v) {
  int lo = 0, hi = B;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (base[mid] <= v) lo = mid; else hi = mid; }
  return lo;
}

// head operand (model.py:355): z[pair,:] = LN(pocket_i) * LN(atom_j), pairs = complex-major, residue, atom
template <typename T>
__global__ void head_outer_kernel(const float* __restrict__ P, const float* __restrict__ Cm, const int* __restrict__ poff,
                                  const int* __restrict__ coff, const int* __restrict__ qoff, int B, int H, T* __restrict__ Z) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= qoff[B]) return;
  const int b = seg_find(qoff, B, warp), loc = warp - qoff[b], nc = coff[b + 1] - coff[b];
  const int pi = poff[b] + loc / nc, ci = coff[b] + loc % nc;
  for (int f = lane * 4; f < H; f += 128) {
    const float4 a = ld4(P + (size_t)pi * H + f), c = ld4(Cm + (size_t)ci * H + f);
    st4(Z + (size_t)warp * H + f, make_float4(a.x * c.x, a.y * c.y, a.z * c.z, a.w * c.w));
  }
}

// y_pred = 10 * sigmoid(dot + b2)   (model.py:358-361);  y_coords = clamp(scale * |pocket_i - atom_j|, 0, 10) (:349,363-365)
__global__ void head_finish_kernel(const float* __restrict__ dot, int tiles, int stride, const float* __restrict__ b2,
                                   const float* __restrict__ pxyz, const float* __restrict__ lxyz,
                                   const int* __restrict__ poff, const int* __restrict__ coff, const int* __restrict__ qoff, int B,
                                   float scale, float cap, float* __restrict__ y_pred, float* __restrict__ y_coords) {
  pdl_entry();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qoff[B]) return;
  float s = b2[0];
  for (int t = 0; t < tiles; ++t) s += dot[(size_t)t * stride + q];
  y_pred[q] = cap / (1.0f + expf(-s));
  const int b = seg_find(qoff, B, q), loc = q - qoff[b], nc = coff[b + 1] - coff[b];
  const int pi = poff[b] + loc / nc, ci = coff[b] + loc % nc;
  const float dx = pxyz[3 * pi] - lxyz[3 * ci], dy = pxyz[3 * pi + 1] - lxyz[3 * ci + 1], dz = pxyz[3 * pi + 2] - lxyz[3 * ci + 2];
  y_coords[q] = fminf(fmaxf(scale * sqrtf(dx * dx + dy * dy + dz * dz), 0.f), cap);
}

// flat pair distances min(|a_i - b_j|, cap) per complex (model.py:286-287: the dis_map label)
__global__ void pair_dist_kernel(const float* __restrict__ pxyz, const float* __restrict__ lxyz, const int* __restrict__ poff,
                                 const int* __restrict__ coff, const int* __restrict__ qoff, int B, float cap,
                                 float* __restrict__ out) {
  pdl_entry();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qoff[B]) return;
  const int b = seg_find(qoff, B, q), loc = q - qoff[b], nc = coff[b + 1] - coff[b];
  const int pi = poff[b] + loc / nc, ci = coff[b] + loc % nc;
  const float dx = pxyz[3 * pi] - lxyz[3 * ci], dy = pxyz[3 * pi + 1] - lxyz[3 * ci + 1], dz = pxyz[3 * pi + 2] - lxyz[3 * ci + 2];
  out[q] = fminf(sqrtf(dx * dx + dy * dy + dz * dz), cap);
}

// out[m] = sum_t dot[t, m] + bias[0]   (finishes a fused Linear(H,1) row-dot epilogue)
__global__ void dot_finish_kernel(const float* __restrict__ dot, int tiles, int stride, int M, const float* __restrict__ bias,
                                  float* __restrict__ out) {
  pdl_entry();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = bias ? bias[0] : 0.f;
  for (int t = 0; t < tiles; ++t) s += dot[(size_t)t * stride + m];
  out[m] = s;
}


// ---- FABind+ wrapper (FABind_plus/fabind/models/model.py::FABindPlus) -------------------------------------------------------
// LayerNorm gathered from an arbitrary row list of a [*, D] fp32 table, output fp32 or bf16 (the A operand of the distance
// head: rows pair[b, 1+i, 1+j, :] of the dense pair embedding, P/models/model.py:379-384)
template <typename TO>
__global__ void layernorm_rows_kernel(const float* __restrict__ x, const int* __restrict__ rows, int M, int D,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, TO* __restrict__ out) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = x + (size_t)(rows ? rows[warp] : warp) * D;
  float s = 0.f;
  for (int f = lane; f < D; f += 32) s += r[f];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int f = lane; f < D; f += 32) { const float d = r[f] - mean; v = fmaf(d, d, v); }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + eps);
  for (int f = lane; f < D; f += 32) {
    const float y = (r[f] - mean) * rstd * gamma[f] + beta[f];
    if constexpr (std::is_same<TO, float>::value) out[(size_t)warp * D + f] = y;
    else out[(size_t)warp * D + f] = __float2bfloat16(y);
  }
}

// out[b, :] = sum of rows off[b]..off[b+1] of src (the ligand-atom sum in front of the pocket radius head, model.py:110-114)
__global__ void segment_sum_rows_kernel(const float* __restrict__ src, int D, const int* __restrict__ off, float* __restrict__ out) {
  pdl_entry();
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < D; f += blockDim.x) {
    float s = 0.f;
    for (int i = off[b]; i < off[b + 1]; ++i) s += src[(size_t)i * D + f];
    out[(size_t)b * D + f] = s;
  }
}

// FABind+ pocket crop: radius_pred = relu(head output); crop radius = radius_pred * buffer (buffer <= 2) or + buffer, floored at
// min_radius, or the fixed radius when fixed_radius >= 0 (model.py:223-231); then the same predicate / "<5 -> first 100" rule
__global__ void __launch_bounds__(256) pocket_mask_r_kernel(const float* __restrict__ xyz, const int* __restrict__ off,
                                                            const float* __restrict__ centers, const float* __restrict__ radius_raw,
                                                            float buffer, float min_radius, float fixed_radius,
                                                            uint8_t* __restrict__ keep, int* __restrict__ less5,
                                                            float* __restrict__ radius_pred) {
  pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.x, lo = off[b], hi = off[b + 1];
  const float rp = fmaxf(radius_raw[b], 0.f);
  float radius = buffer <= 2.0f ? __fmul_rn(rp, buffer) : __fadd_rn(rp, buffer);
  if (radius < min_radius) radius = min_radius;
  if (fixed_radius >= 0.f) radius = fixed_radius;
  if (threadIdx.x == 0) radius_pred[b] = rp;
  const float cx = centers[3 * b], cy = centers[3 * b + 1], cz = centers[3 * b + 2];
  float n = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float dx = __fsub_rn(xyz[3 * i], cx), dy = __fsub_rn(xyz[3 * i + 1], cy), dz = __fsub_rn(xyz[3 * i + 2], cz);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const bool k = d < radius;
    keep[i] = k ? 1 : 0;
    n += k ? 1.f : 0.f;
  }
  n = block_sum(n, red);
  if (n < 5.f) {
    for (int i = lo + threadIdx.x; i < min(hi, lo + 100); i += blockDim.x) keep[i] = 1;
    if (threadIdx.x == 0) less5[b] = 1;
  } else if (threadIdx.x == 0) less5[b] = 0;
}

// per-segment mean of [n,3] rows and the rows re-centred on it (pocket_coords - pocket_coords.mean(0), model.py:255-258)
__global__ void __launch_bounds__(256) center_rows3_kernel(const float* __restrict__ xyz, const int* __restrict__ off,
                                                           float* __restrict__ centered, float* __restrict__ mean) {
  pdl_entry();
  __shared__ float red[32];
  const int b = blockIdx.x;
  float m[3] = {0.f, 0.f, 0.f};
  for (int i = off[b] + threadIdx.x; i < off[b + 1]; i += blockDim.x) { m[0] += xyz[3 * i]; m[1] += xyz[3 * i + 1]; m[2] += xyz[3 * i + 2]; }
  const float n = (float)(off[b + 1] - off[b]);
  for (int k = 0; k < 3; ++k) m[k] = block_sum(m[k], red) / n;
  if (threadIdx.x < 3) mean[3 * b + threadIdx.x] = m[threadIdx.x];
  for (int i = off[b] + threadIdx.x; i < off[b + 1]; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) centered[3 * i + k] = xyz[3 * i + k] - m[k];
  }
}

// out = xyz + sign * shift[segment]   (data.coords -= centre, model.py:257; prediction + pocket_center_bias, model.py:684)
__global__ void shift_rows3_kernel(const float* __restrict__ xyz, const int* __restrict__ off, int B, const float* __restrict__ shift,
                                   float sign, float* __restrict__ out) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= off[B]) return;
  const int b = seg_find(off, B, i);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * i + k] = xyz[3 * i + k] + sign * shift[3 * b + k];
}

}  // namespace fb

using namespace fb;

extern "C" {

int32_t fb_assemble_rows(float* out, int32_t M, int32_t D, const uint8_t* kind, const int32_t* idx, const float* s0,
                         const float* s1, const float* s2, const float* s3, float scale, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(assemble_rows_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, out, M, D, kind, idx, s0, s1, s2, s3, scale);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_layernorm(const float* x, int32_t M, int32_t D, const float* gamma, const float* beta, float eps, float* out,
                     void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(layernorm_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, M, D, gamma, beta, eps, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pocket_center(const float* logit, const float* xyz, const int32_t* prot_off, int32_t B, float tau, int32_t hard,
                         int32_t mode, float* centers, void* stream) {
  fb_launch(pocket_center_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, logit, xyz, prot_off, tau, hard, mode, centers,
            (const float*)nullptr);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pocket_center_gumbel(const float* logit, const float* noise, const float* xyz, const int32_t* prot_off, int32_t B, float tau,
                                int32_t hard, float* centers, void* stream) {
  if (!noise) return FB_ERR_BAD_ARG;
  fb_launch(pocket_center_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, logit, xyz, prot_off, tau, hard, 0, centers, noise);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pocket_mask(const float* xyz, const int32_t* prot_off, int32_t B, const float* centers, float radius, uint8_t* keep,
                       int32_t* less5, void* stream) {
  fb_launch(pocket_mask_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, xyz, prot_off, centers, radius, keep, less5);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_ligand_place(const float* lig, const int32_t* comp_off, const float* pocket, const int32_t* pocket_off, int32_t B,
                        float* out, void* stream) {
  fb_launch(ligand_place_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, lig, comp_off, pocket, pocket_off, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_head_outer(const float* pocket_ln, const float* comp_ln, const int32_t* pocket_off, const int32_t* comp_off,
                      const int32_t* pair_off, int32_t B, int32_t n_pairs, int32_t H, void* Z, int32_t bf16_mode, void* stream) {
  if (n_pairs <= 0) return FB_OK;
  if (H & 3) return FB_ERR_UNSUPPORTED;
  const dim3 grid(((long long)n_pairs * 32 + 255) / 256);
  if (bf16_mode) fb_launch(head_outer_kernel<bf16>, grid, dim3(256), 0, (cudaStream_t)stream, pocket_ln, comp_ln, pocket_off, comp_off, pair_off, B, H, (bf16*)Z);
  else fb_launch(head_outer_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, pocket_ln, comp_ln, pocket_off, comp_off, pair_off, B, H, (float*)Z);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_head_finish_cap(const float* dot, int32_t tiles, int32_t stride, const float* b2, const float* pocket_xyz,
                           const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off, const int32_t* pair_off,
                           int32_t B, int32_t n_pairs, float scale, float cap, float* y_pred, float* y_coords, void* stream) {
  if (n_pairs <= 0) return FB_OK;
  fb_launch(head_finish_kernel, dim3((n_pairs + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dot, tiles, stride, b2, pocket_xyz,
            lig_xyz, pocket_off, comp_off, pair_off, B, scale, cap, y_pred, y_coords);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_head_finish(const float* dot, int32_t tiles, int32_t stride, const float* b2, const float* pocket_xyz,
                       const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off, const int32_t* pair_off,
                       int32_t B, int32_t n_pairs, float scale, float* y_pred, float* y_coords, void* stream) {
  return fb_head_finish_cap(dot, tiles, stride, b2, pocket_xyz, lig_xyz, pocket_off, comp_off, pair_off, B, n_pairs, scale, 10.0f,
                            y_pred, y_coords, stream);
}

int32_t fb_layernorm_rows(const float* x, const int32_t* rows, int32_t M, int32_t D, const float* gamma, const float* beta, float eps,
                          void* out, int32_t out_bf16, void* stream) {
  if (M <= 0) return FB_OK;
  if (out_bf16) fb_launch(layernorm_rows_kernel<__nv_bfloat16>, dim3((M * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, rows, M, D, gamma, beta, eps, (__nv_bfloat16*)out);
  else fb_launch(layernorm_rows_kernel<float>, dim3((M * 32 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, rows, M, D, gamma, beta, eps, (float*)out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_segment_sum_rows(const float* src, int32_t D, const int32_t* off, int32_t B, float* out, void* stream) {
  if (B <= 0) return FB_OK;
  fb_launch(segment_sum_rows_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, src, D, off, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pocket_mask_r(const float* xyz, const int32_t* prot_off, int32_t B, const float* centers, const float* radius_raw,
                         float buffer, float min_radius, float fixed_radius, uint8_t* keep, int32_t* less5, float* radius_pred,
                         void* stream) {
  fb_launch(pocket_mask_r_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, xyz, prot_off, centers, radius_raw, buffer, min_radius,
            fixed_radius, keep, less5, radius_pred);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_center_rows3(const float* xyz, const int32_t* off, int32_t B, float* centered, float* mean, void* stream) {
  if (B <= 0) return FB_OK;
  fb_launch(center_rows3_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, xyz, off, centered, mean);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_shift_rows3(const float* xyz, const int32_t* off, int32_t B, int32_t n_rows, const float* shift, float sign, float* out,
                       void* stream) {
  if (n_rows <= 0) return FB_OK;
  fb_launch(shift_rows3_kernel, dim3((n_rows + 255) / 256), dim3(256), 0, (cudaStream_t)stream, xyz, off, B, shift, sign, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pair_dist(const float* pocket_xyz, const float* lig_xyz, const int32_t* pocket_off, const int32_t* comp_off,
                     const int32_t* pair_off, int32_t B, int32_t n_pairs, float cap, float* out, void* stream) {
  if (n_pairs <= 0) return FB_OK;
  fb_launch(pair_dist_kernel, dim3((n_pairs + 255) / 256), dim3(256), 0, (cudaStream_t)stream, pocket_xyz, lig_xyz, pocket_off, comp_off,
            pair_off, B, cap, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_dot_finish(const float* dot, int32_t tiles, int32_t stride, int32_t M, const float* bias, float* out, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(dot_finish_kernel, dim3((M + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dot, tiles, stride, M, bias, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // extern "C"
