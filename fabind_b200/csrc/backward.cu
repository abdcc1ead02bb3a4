// Reverse-pass primitives of the training path (fp32; BASELINE config 5: forward + backward + gradient all-reduce).
// Each kernel is one launch of the hand-derived reverse pass that tests/emulate_backward.py specifies (and pins against
// autograd and the unmodified reference); the data-gradient GEMMs `dX = dY W` run through fb_gemm on a transposed weight.
// Scatter directions use fp32 atomics (order-dependent in the last bits, inside the 1e-4 training tolerance); the
// CSR-by-source variant that removes them is the planned follow-up.  C ABI at the bottom.
#include "../../include/fabind_b200.h"

#include "common.cuh"

namespace fb {

__device__ __forceinline__ float act_grad(float z, int kind) {
  if (kind == FB_ACT_SILU) {
    const float s = 1.0f / (1.0f + expf(-z));
    return s * (1.0f + z * (1.0f - s));
  }
  if (kind == FB_ACT_RELU) return z > 0.f ? 1.0f : 0.f;
  return 1.0f;
}
__device__ __forceinline__ float act_value(float z, int kind) {
  if (kind == FB_ACT_SILU) return z / (1.0f + expf(-z));
  if (kind == FB_ACT_RELU) return fmaxf(z, 0.f);
  return z;
}

// Y = act(Z)   (re-materialises an activation from its saved pre-activation)
__global__ void act_fwd_kernel(const float* __restrict__ Z, float* __restrict__ Y, long long n, int kind) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Y[i] = act_value(Z[i], kind);
}

// dZ = dY * act'(Z)
__global__ void act_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ dY, float* __restrict__ dZ, long long n,
                               int kind) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dZ[i] = dY[i] * act_grad(Z[i], kind);
}

// dZ[m, n] = u[m] * v[n] * act'(Z[m, n])   (reverse of a Linear(H,1) head behind an activation: s = act(Z) . v)
__global__ void outer_act_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ u, const float* __restrict__ v,
                                     float* __restrict__ dZ, int M, int N, int kind) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float um = u[warp];
  const float* z = Z + (size_t)warp * N;
  float* d = dZ + (size_t)warp * N;
  for (int f = lane; f < N; f += 32) d[f] = um * v[f] * act_grad(z[f], kind);
}

// out[n] (+)= sum_m w[m] * A[m, n]   (bias gradients; w == null: plain column sums; w = radial: rank-1 column gradients)
// grid.x tiles the columns (32 per CTA), grid.y splits the rows; partial sums are combined with atomics into a zeroed or
// caller-accumulated buffer.
__global__ void colsum_kernel(const float* __restrict__ A, int lda, int M, int N, const float* __restrict__ w,
                              float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;   // 8 row lanes x 32 columns
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float s = 0.f;
  if (col < N)
    for (int m = m0 + ty; m < m1; m += 8) s = fmaf(w ? w[m] : 1.0f, A[(size_t)m * lda + col], s);
  red[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    atomicAdd(&out[col], t);
  }
}

// out[m] = sum_n A[m, n] * v[n]   (one warp per row)
__global__ void rowdot_kernel(const float* __restrict__ A, int lda, int M, int N, const float* __restrict__ v,
                              float* __restrict__ out) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = A + (size_t)warp * lda;
  float s = 0.f;
  for (int f = lane; f < N; f += 32) s = fmaf(r[f], v[f], s);
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}

// dst[idx[e], :D] += src[e, :D]   (reverse of a row gather; one warp per source row, fp32 atomics)
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int E, int D,
                                        float* __restrict__ dst, int ldd) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  const float* s = src + (size_t)warp * lds;
  float* d = dst + (size_t)idx[warp] * ldd;
  for (int f = lane; f < D; f += 32) atomicAdd(&d[f], s[f]);
}

// dst[e, :D] += src[idx[e], :D]   (reverse of a segment sum: every edge receives its destination node's gradient)
__global__ void gather_add_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int E, int D,
                                       float* __restrict__ dst, int ldd) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= E) return;
  const float* s = src + (size_t)idx[warp] * lds;
  float* d = dst + (size_t)warp * ldd;
  for (int f = lane; f < D; f += 32) d[f] += s[f];
}

// dW[n, k] (+)= sum_m dY[m, n] * X[m, k]     (weight gradient: reduction over the rows)
// CTA = 32 x 32 tile of dW and one slice of the rows (grid.z); 256 threads, each owns a 2 x 2 micro-tile; operands are staged
// 32 rows at a time in shared memory; slices are combined with atomics (dW zeroed or accumulated by the caller).
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx,
                                                    int M, int N, int K, float* __restrict__ dW, int ldw) {
  pdl_entry();
  __shared__ float sY[32][33], sX[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int rows_per = ((M + gridDim.z - 1) / gridDim.z + 31) & ~31;
  const int m_lo = blockIdx.z * rows_per, m_hi = min(M, m_lo + rows_per);
  const int tn = (threadIdx.x >> 4) * 2, tk = (threadIdx.x & 15) * 2;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  for (int m0 = m_lo; m0 < m_hi; m0 += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int r = i >> 5, c = i & 31, m = m0 + r;
      sY[r][c] = (m < m_hi && n0 + c < N) ? dY[(size_t)m * ldy + n0 + c] : 0.f;
      sX[r][c] = (m < m_hi && k0 + c < K) ? X[(size_t)m * ldx + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float y0 = sY[r][tn], y1 = sY[r][tn + 1], x0 = sX[r][tk], x1 = sX[r][tk + 1];
      a00 = fmaf(y0, x0, a00); a01 = fmaf(y0, x1, a01);
      a10 = fmaf(y1, x0, a10); a11 = fmaf(y1, x1, a11);
    }
    __syncthreads();
  }
  const int n = n0 + tn, k = k0 + tk;
  if (n < N && k < K) atomicAdd(&dW[(size_t)n * ldw + k], a00);
  if (n < N && k + 1 < K) atomicAdd(&dW[(size_t)n * ldw + k + 1], a01);
  if (n + 1 < N && k < K) atomicAdd(&dW[(size_t)(n + 1) * ldw + k], a10);
  if (n + 1 < N && k + 1 < K) atomicAdd(&dW[(size_t)(n + 1) * ldw + k + 1], a11);
}

// Reverse of the clamped coordinate step  x_new[i] = x[i] + clamp(sum_{e: row(e)=i} (x[i]-x[col(e)]) * s[e] / cnt[i], +-cmax)
// (egnn.py:85-98 mean aggregation: cnt = max(degree,1); egnn.py:228-233 interfacial sum: cnt == null).
// One lane per edge: ds[e] = <g_i, d_e>,  dx[i] += g_i s_e,  dx[col] -= g_i s_e  with g_i = dx_new[i] * [|step_i| <= cmax] / cnt_i.
// dx must hold dx_new on entry (the identity path of x_new = x + ...).
__global__ void coord_step_bwd_kernel(const float* __restrict__ x, const int* __restrict__ row, const int* __restrict__ col, int E,
                                      const float* __restrict__ s, const float* __restrict__ step, const float* __restrict__ cnt,
                                      float cmax, const float* __restrict__ dx_new, float* __restrict__ dx,
                                      float* __restrict__ ds) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = row[e], j = col[e];
  const float inv = cnt ? 1.0f / fmaxf(cnt[i], 1.0f) : 1.0f;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float g = (fabsf(step[3 * i + a]) <= cmax ? dx_new[3 * i + a] : 0.f) * inv;
    const float d = x[3 * i + a] - x[3 * j + a];
    acc = fmaf(g, d, acc);
    const float t = g * s[e];
    atomicAdd(&dx[3 * i + a], t);
    atomicAdd(&dx[3 * j + a], -t);
  }
  ds[e] = acc;
}

// Reverse of coord2radial with the per-complex norm (egnn.py:767-787): rn_e = d2_e / nrm_b, nrm_b = sqrt(sum_e d2_e^2).
// pass 1: dot[b] = sum_e drn_e d2_e ; pass 2: dd2_e = drn_e / nrm_b - d2_e dot_b / nrm_b^3, dx[row] += 2 d dd2, dx[col] -= ...
__global__ void radial_bwd_dot_kernel(const float* __restrict__ x, const int* __restrict__ row, const int* __restrict__ col, int E,
                                      const int* __restrict__ cplx, const float* __restrict__ drn, float* __restrict__ dot) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = row[e], j = col[e];
  const float d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
  atomicAdd(&dot[cplx[i]], drn[e] * (d0 * d0 + d1 * d1 + d2 * d2));
}
__global__ void radial_bwd_apply_kernel(const float* __restrict__ x, const int* __restrict__ row, const int* __restrict__ col, int E,
                                        const int* __restrict__ cplx, const float* __restrict__ nrm, const float* __restrict__ drn,
                                        const float* __restrict__ dot, float* __restrict__ dx) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = row[e], j = col[e], b = cplx[i];
  const float d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
  const float q = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float n = nrm[b];
  const float dd2 = drn[e] / n - q * dot[b] / (n * n * n);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float t = 2.0f * d[a] * dd2;
    atomicAdd(&dx[3 * i + a], t);
    atomicAdd(&dx[3 * j + a], -t);
  }
}

// Reverse of the LAS step (egnn.py:433-449): x_new[j] = x[j] + clamp(step * sum_{(i,j)} 4 (|x_i-x_j|^2 - ref_ij) (x_i-x_j), +-lcl).
// acc = the unclamped step of the forward; dx must hold dx_new on entry.
__global__ void las_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xref, const int* __restrict__ a_idx,
                               const int* __restrict__ b_idx, int E, const float* __restrict__ acc, float step_size, float lcl,
                               const float* __restrict__ dx_new, float* __restrict__ dx) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = a_idx[e], j = b_idx[e];
  float d[3], f[3], cur = 0.f, ref = 0.f, fd = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    d[a] = x[3 * i + a] - x[3 * j + a];
    const float r = xref[3 * i + a] - xref[3 * j + a];
    cur = fmaf(d[a], d[a], cur);
    ref = fmaf(r, r, ref);
    f[a] = (fabsf(acc[3 * j + a]) <= lcl ? dx_new[3 * j + a] : 0.f) * step_size;
    fd = fmaf(f[a], d[a], fd);
  }
  const float diff = cur - ref;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float t = 4.0f * diff * f[a] + 8.0f * fd * d[a];
    atomicAdd(&dx[3 * i + a], t);
    atomicAdd(&dx[3 * j + a], -t);
  }
}

static inline int grid_1d(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fb

using namespace fb;

extern "C" {

int32_t fb_act_fwd(const float* Z, float* Y, int64_t n, int32_t act, void* stream) {
  if (n <= 0) return FB_OK;
  fb_launch(act_fwd_kernel, dim3(grid_1d(n, 256)), dim3(256), 0, (cudaStream_t)stream, Z, Y, (long long)n, (int)act);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_act_bwd(const float* Z, const float* dY, float* dZ, int64_t n, int32_t act, void* stream) {
  if (n <= 0) return FB_OK;
  fb_launch(act_bwd_kernel, dim3(grid_1d(n, 256)), dim3(256), 0, (cudaStream_t)stream, Z, dY, dZ, (long long)n, (int)act);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_outer_act_bwd(const float* Z, const float* u, const float* v, float* dZ, int32_t M, int32_t N, int32_t act, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  fb_launch(outer_act_bwd_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, Z, u, v, dZ, (int)M,
            (int)N, (int)act);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_colsum(const float* A, int32_t lda, int32_t M, int32_t N, const float* w, float* out, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  int split = (M + 255) / 256;
  if (split > 64) split = 64;
  fb_launch(colsum_kernel, dim3((N + 31) / 32, split), dim3(256), 0, (cudaStream_t)stream, A, (int)lda, (int)M, (int)N, w, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_rowdot(const float* A, int32_t lda, int32_t M, int32_t N, const float* v, float* out, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(rowdot_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, A, (int)lda, (int)M,
            (int)N, v, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_scatter_add_rows(const float* src, int32_t lds, const int32_t* idx, int32_t E, int32_t D, float* dst, int32_t ldd,
                            void* stream) {
  if (E <= 0 || D <= 0) return FB_OK;
  fb_launch(scatter_add_rows_kernel, dim3((int)(((long long)E * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, src,
            (int)lds, idx, (int)E, (int)D, dst, (int)ldd);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_gather_add_rows(const float* src, int32_t lds, const int32_t* idx, int32_t E, int32_t D, float* dst, int32_t ldd,
                           void* stream) {
  if (E <= 0 || D <= 0) return FB_OK;
  fb_launch(gather_add_rows_kernel, dim3((int)(((long long)E * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, src,
            (int)lds, idx, (int)E, (int)D, dst, (int)ldd);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_gemm_wgrad(const float* dY, int32_t ldy, const float* X, int32_t ldx, int32_t M, int32_t N, int32_t K, float* dW,
                      int32_t ldw, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return FB_OK;
  const int tiles = ((N + 31) / 32) * ((K + 31) / 32);
  int split = (2 * 148 + tiles - 1) / tiles;          // about two waves of CTAs
  const int max_split = (M + 127) / 128;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  fb_launch(wgrad_kernel, dim3((N + 31) / 32, (K + 31) / 32, split), dim3(256), 0, (cudaStream_t)stream, dY, (int)ldy, X, (int)ldx,
            (int)M, (int)N, (int)K, dW, (int)ldw);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_coord_step_bwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const float* s, const float* step,
                          const float* cnt, float cmax, const float* dx_new, float* dx, float* ds, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(coord_step_bwd_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, row, col, (int)E, s, step, cnt,
            cmax, dx_new, dx, ds);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_radial_bwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const int32_t* node_cplx, const float* nrm,
                      const float* drn, float* dot_zeroed, float* dx, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(radial_bwd_dot_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, row, col, (int)E, node_cplx, drn,
            dot_zeroed);
  fb_launch(radial_bwd_apply_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, row, col, (int)E, node_cplx, nrm,
            drn, (const float*)dot_zeroed, dx);
  count_launch(2);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_las_bwd(const float* x, const float* xref, const int32_t* a_idx, const int32_t* b_idx, int32_t E, const float* acc,
                   float step_size, float lcl, const float* dx_new, float* dx, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(las_bwd_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, xref, a_idx, b_idx, (int)E, acc, step_size,
            lcl, dx_new, dx);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // extern "C"
