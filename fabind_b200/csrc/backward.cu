// Reverse-pass primitives of the training path (fp32; BASELINE config 5: forward + backward + gradient all-reduce).
// Each kernel is one launch of the hand-derived reverse pass that tests/emulate_backward.py specifies (and pins against
// autograd and the unmodified reference); the data-gradient GEMMs `dX = dY W` run through fb_gemm on a transposed weight.
// Scatter directions use fp32 atomics (order-dependent in the last bits, inside the 1e-4 training tolerance); the
// CSR-by-source variant that removes them is the planned follow-up.  C ABI at the bottom.
#include "../../include/fabind_b200.h"

#include "common.cuh"
#include "layers.h"

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

// Training-mode twin of gcl_edge_pre (layers.cu): first edge-MLP Linear hoisted per node, pre-activation KEPT for the reverse pass.
//   Z1[e, :] = Pn[row[e], 0:H] + Pn[col[e], H:2H] + rn[e] * w_rad + b1        A1 = act(Z1)  (fp32, and bf16 for the next GEMM)
// One pass instead of a memset, two gather-adds, two rank-1 updates, the activation and the bf16 conversion (1.15 GB -> 0.23 GB of
// traffic per sub-layer at B = 16).  One warp per edge, lane = 4 consecutive features.
__global__ void __launch_bounds__(256) edge_pre_train_kernel(const float* __restrict__ Pn, const int* __restrict__ row,
                                                             const int* __restrict__ col, int E, int H, const float* __restrict__ rn,
                                                             const float* __restrict__ w_rad, const float* __restrict__ b1,
                                                             float* __restrict__ Z1, float* __restrict__ A1, bf16* __restrict__ A16,
                                                             int kind) {
  pdl_entry();
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= E) return;
  const float* pr = Pn + (size_t)row[e] * 2 * H;
  const float* pc = Pn + (size_t)col[e] * 2 * H + H;
  const float r = rn[e];
  for (int f = lane * 4; f < H; f += 128) {
    const float4 a = ld4(pr + f), b = ld4(pc + f), w = ld4(w_rad + f), bb = ld4(b1 + f);
    float4 z;
    z.x = fmaf(r, w.x, a.x + b.x) + bb.x; z.y = fmaf(r, w.y, a.y + b.y) + bb.y;
    z.z = fmaf(r, w.z, a.z + b.z) + bb.z; z.w = fmaf(r, w.w, a.w + b.w) + bb.w;
    const float4 y = make_float4(act_value(z.x, kind), act_value(z.y, kind), act_value(z.z, kind), act_value(z.w, kind));
    st4(Z1 + (size_t)e * H + f, z);
    st4(A1 + (size_t)e * H + f, y);
    if (A16) st4(A16 + (size_t)e * H + f, y);
  }
}

// Y = drop(act(Z)) in fp32 and, optionally, bf16 (the operand of the next GEMM): activation, the reference's nn.Dropout behind it
// (egnn.py:82) and the conversion in one pass.  Row-major [M, N], N a multiple of 4.
__global__ void __launch_bounds__(256) act_drop_kernel(const float* __restrict__ Z, int M, int N, int kind, DropCfg dc,
                                                       float* __restrict__ Y, bf16* __restrict__ Y16) {
  pdl_entry();
  const long long total4 = (long long)M * N / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i * 4;
    const int m = (int)(o / N), n = (int)(o - (long long)m * N);
    const float4 z = ld4(Z + o);
    float4 y = make_float4(act_value(z.x, kind), act_value(z.y, kind), act_value(z.z, kind), act_value(z.w, kind));
    if (dc.p > 0.f) {
      y.x = drop_apply(y.x, dc, m, n); y.y = drop_apply(y.y, dc, m, n + 1);
      y.z = drop_apply(y.z, dc, m, n + 2); y.w = drop_apply(y.w, dc, m, n + 3);
    }
    st4(Y + o, y);
    if (Y16) st4(Y16 + o, y);
  }
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

// Reverse-pass twins of the fused forward kernels: the gradient tensor goes out in fp32 AND bf16 (the operand of the data-gradient
// GEMM), and the nn.Dropout in front of the activation's output is applied to the incoming gradient in the same pass.
//   act_bwd_drop:    dZ = drop(dY) * act'(Z)                          (dY: gradient of the DROPPED activation, egnn.py:82)
//   outer_act_bwd2:  dZ[m, n] = u[m] v[n] act'(Z[m, n])               (same as outer_act_bwd_kernel)
__global__ void __launch_bounds__(256) act_bwd_drop_kernel(const float* __restrict__ Z, const float* __restrict__ dY, int M, int N, int kind,
                                                           DropCfg dc, float* __restrict__ dZ, bf16* __restrict__ dZ16) {
  pdl_entry();
  const long long total4 = (long long)M * N / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i * 4;
    const int m = (int)(o / N), n = (int)(o - (long long)m * N);
    const float4 z = ld4(Z + o);
    float4 g = ld4(dY + o);
    if (dc.p > 0.f) {
      g.x = drop_apply(g.x, dc, m, n); g.y = drop_apply(g.y, dc, m, n + 1);
      g.z = drop_apply(g.z, dc, m, n + 2); g.w = drop_apply(g.w, dc, m, n + 3);
    }
    const float4 d = make_float4(g.x * act_grad(z.x, kind), g.y * act_grad(z.y, kind), g.z * act_grad(z.z, kind), g.w * act_grad(z.w, kind));
    st4(dZ + o, d);
    if (dZ16) st4(dZ16 + o, d);
  }
}

__global__ void __launch_bounds__(256) outer_act_bwd2_kernel(const float* __restrict__ Z, const float* __restrict__ u, const float* __restrict__ v,
                                                             float* __restrict__ dZ, bf16* __restrict__ dZ16, int M, int N, int kind) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float um = u[warp];
  const size_t base = (size_t)warp * N;
  for (int f = lane * 4; f < N; f += 128) {
    const float4 z = ld4(Z + base + f), vv = ld4(v + f);
    const float4 d = make_float4(um * vv.x * act_grad(z.x, kind), um * vv.y * act_grad(z.y, kind), um * vv.z * act_grad(z.z, kind),
                                 um * vv.w * act_grad(z.w, kind));
    st4(dZ + base + f, d);
    if (dZ16) st4(dZ16 + base + f, d);
  }
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
// per-complex sums over an edge list: the edges are ordered by destination row, so a warp's 32 edges almost always belong to ONE
// complex -- summed in the warp first, one atomic per warp (one atomic per EDGE onto the 16 per-complex addresses serialised in L2:
// 40 us per launch at 44.9k edges).  Every lane of the warp calls this (inactive lanes with active = false).
__device__ __forceinline__ void atomic_add_per_complex(float* __restrict__ base, int key, float v, bool active) {
  if (!active) { key = -1; v = 0.f; }
  const int k0 = __shfl_sync(0xffffffffu, key, 0);
  const unsigned same = __ballot_sync(0xffffffffu, key == k0 || key < 0);
  if (same == 0xffffffffu) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && k0 >= 0) atomicAdd(base + k0, v);
  } else if (active) {
    atomicAdd(base + key, v);
  }
}

__global__ void radial_bwd_dot_kernel(const float* __restrict__ x, const int* __restrict__ row, const int* __restrict__ col, int E,
                                      const int* __restrict__ cplx, const float* __restrict__ drn, float* __restrict__ dot) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < E;
  int b = -1;
  float v = 0.f;
  if (active) {
    const int i = row[e], j = col[e];
    const float d0 = x[3 * i] - x[3 * j], d1 = x[3 * i + 1] - x[3 * j + 1], d2 = x[3 * i + 2] - x[3 * j + 2];
    b = cplx[i];
    v = drn[e] * (d0 * d0 + d1 * d1 + d2 * d2);
  }
  atomic_add_per_complex(dot, b, v, active);
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

// ---- second group: MC_Att_L reverse (row attention, interfacial attention, pair path) ----------------------------------------

// out[m] = sum_n A[m,n] B[m,n]
__global__ void rowdot2_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M, int N,
                               float* __restrict__ out) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* a = A + (size_t)warp * lda;
  const float* b = Bm + (size_t)warp * ldb;
  float s = 0.f;
  for (int f = lane; f < N; f += 32) s = fmaf(a[f], b[f], s);
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}

// A[m,n] = A[m,n] * u[m]   (mode 0)      A[m,n] += u[m] * v[n]   (mode 1)
__global__ void rows_update_kernel(float* __restrict__ A, int lda, int M, int N, const float* __restrict__ u,
                                   const float* __restrict__ v, int mode) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  float* a = A + (size_t)warp * lda;
  const float um = u[warp];
  if (mode == 0) { for (int f = lane; f < N; f += 32) a[f] *= um; }
  else { for (int f = lane; f < N; f += 32) a[f] = fmaf(um, v[f], a[f]); }
}

// c = a * b (op 0), c = a + b (op 1), c += a * b (op 2)
__global__ void vec_op_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, long long n, int op) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (op == 0) c[i] = a[i] * b[i];
    else if (op == 1) c[i] = a[i] + b[i];
    else c[i] = fmaf(a[i], b[i], c[i]);
  }
}

// segment softmax reverse (scatter_softmax over the destination row, egnn.py:221): dlogit_e = alpha_e (dalpha_e - sum_{e' in row} alpha dalpha)
__global__ void softmax_seg_bwd_sum_kernel(const float* __restrict__ alpha, const float* __restrict__ dalpha,
                                           const int* __restrict__ row, int E, float* __restrict__ t) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) atomicAdd(&t[row[e]], alpha[e] * dalpha[e]);
}
__global__ void softmax_seg_bwd_apply_kernel(const float* __restrict__ alpha, const float* __restrict__ dalpha,
                                             const int* __restrict__ row, int E, const float* __restrict__ t,
                                             float* __restrict__ dlogit) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) dlogit[e] = alpha[e] * (dalpha[e] - t[row[e]]);
}

// gated pair bias reverse (model_utils.py:96-133 `linear(pair) * sigmoid(linear_g(pair))`): raw [P, ld] holds per block
// (8 columns each) 4 values then 4 gates; dPB [P, nblk, 4] -> draw [P, ld] (columns >= 8 nblk zeroed)
__global__ void pair_bias_gate_bwd_kernel(const float* __restrict__ raw, int ld, long long P, int nblk, const float* __restrict__ dPB,
                                          float* __restrict__ draw) {
  pdl_entry();
  const long long total = P * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / ld;
    const int c = (int)(i - p * ld);
    float o = 0.f;
    if (c < 8 * nblk) {
      const int blk = c >> 3, h = c & 3, is_gate = (c >> 2) & 1;
      const float* rp = raw + p * ld + blk * 8;
      const float sg = 1.0f / (1.0f + expf(-rp[4 + h]));
      const float d = dPB[(p * nblk + blk) * 4 + h];
      o = is_gate ? d * rp[h] * sg * (1.0f - sg) : d * sg;
    }
    draw[i] = o;
  }
}

// reverse of the per-complex outer product pair0 operand  outer[b, i, j, :] = pp[i, :] * cc[j, :]  (model_utils.py:216-220):
// one CTA per protein-side row i: dpp[i,:] = sum_j dO[i,j,:] cc[j,:]  (written),  dcc[j,:] += dO[i,j,:] pp[i,:]  (atomics)
__global__ void pair_outer_bwd_kernel(const float* __restrict__ dO, const float* __restrict__ pc, int H, const int* __restrict__ c_off,
                                      const int* __restrict__ p_off, const int* __restrict__ pair_base,
                                      const int* __restrict__ node_cplx, int p_begin, float* __restrict__ dpc) {
  pdl_entry();
  const int i = p_begin + blockIdx.x;            // internal row of a protein-side node
  const int b = node_cplx[i];
  const int nc1 = c_off[b + 1] - c_off[b];
  const size_t base = (size_t)pair_base[b] + (size_t)(i - p_off[b]) * nc1;
  for (int f = threadIdx.x; f < H; f += blockDim.x) {
    const float pv = pc[(size_t)i * H + f];
    float acc = 0.f;
    for (int j = 0; j < nc1; ++j) {
      const float d = dO[(base + j) * H + f];
      acc = fmaf(d, pc[(size_t)(c_off[b] + j) * H + f], acc);
      atomicAdd(&dpc[(size_t)(c_off[b] + j) * H + f], d * pv);
    }
    dpc[(size_t)i * H + f] = acc;
  }
}

// RowAttentionBlock core reverse (forward: layers.cu::row_attention_kernel; cross_att.py:118-134, model_utils.py:21-38):
//   O[q, h*32+d] = sigmoid(G) * sum_j a_j v_j[d],  a = softmax_j(q.k_j / sqrt(32) + bias_j)
// One CTA per (complex, head, tile of q_tile queries): the keys / values of the head and their gradient accumulators live in shared
// memory (n_k <= KC <= RB_MAXK, sized to the launch), every warp walks queries of the tile, probabilities are recomputed.  Writes dQ, dG,
// dPB (one owner each) and, once per CTA, ADDS its dK / dV to global memory (the caller zeroes them) -- with one CTA per (complex, head)
// and 140 KB of shared memory the launch was 64 CTAs at one per SM (156 us at B = 16).
constexpr int RB_MAXK = 256;
constexpr int RB_WARPS = 8;
constexpr int RB_SMEM_FLOATS = 4 * RB_MAXK * 33 + RB_WARPS * (2 * RB_MAXK + 64);
__global__ void __launch_bounds__(RB_WARPS * 32) row_attention_bwd_kernel(
    const int* __restrict__ c_off, const int* __restrict__ p_off, const int* __restrict__ pair_base, int q_is_prot,
    const float* __restrict__ Q, int ldq, const float* __restrict__ G, int ldg, const float* __restrict__ Kb, int ldk,
    const float* __restrict__ Vb, int ldv, const float* __restrict__ PB, const float* __restrict__ dO, int ldo,
    float* __restrict__ dQ, int lddq, float* __restrict__ dG, int lddg, float* __restrict__ dK, int lddk,
    float* __restrict__ dV, int lddv, float* __restrict__ dPB, int KC, int q_tile) {
  pdl_entry();
  extern __shared__ float rb_smem[];
  const int b = blockIdx.x, head = blockIdx.y;
  const int c_lo = c_off[b], nc1 = c_off[b + 1] - c_lo, p_lo = p_off[b], np1 = p_off[b + 1] - p_lo;
  const int n_q = q_is_prot ? np1 : nc1, n_k = q_is_prot ? nc1 : np1;
  const int q_lo = q_is_prot ? p_lo : c_lo, k_lo = q_is_prot ? c_lo : p_lo;
  const int q_begin = blockIdx.z * q_tile, q_end = min(n_q, q_begin + q_tile);
  if (q_begin >= n_q) return;            // block-uniform
  float* sK = rb_smem;                   // [n_k][33]
  float* sV = sK + KC * 33;
  float* sdK = sV + KC * 33;
  float* sdV = sdK + KC * 33;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pa = sdV + KC * 33 + warp * (2 * KC + 64);              // probabilities of the warp's current query
  float* pd = pa + KC;                                           // da, then dlogit
  float* wq = pd + KC;                                           // scaled query, 32 channels
  float* wdo = wq + 32;                                          // gradient w.r.t. the un-gated output, 32 channels
  const float scale = 0.17677669529663687f;                      // 1/sqrt(32)
  for (int i = threadIdx.x; i < n_k * 32; i += blockDim.x) {
    const int j = i >> 5, d = i & 31;
    sK[j * 33 + d] = Kb[(size_t)(k_lo + j) * ldk + head * 32 + d];
    sV[j * 33 + d] = Vb[(size_t)(k_lo + j) * ldv + head * 32 + d];
    sdK[j * 33 + d] = 0.f;
    sdV[j * 33 + d] = 0.f;
  }
  __syncthreads();
  for (int ql = q_begin + warp; ql < q_end; ql += RB_WARPS) {
    const int qn = q_lo + ql;
    const float qd = Q[(size_t)qn * ldq + head * 32 + lane] * scale;      // lane = channel
    const float gd = G[(size_t)qn * ldg + head * 32 + lane];
    const float sg = 1.0f / (1.0f + expf(-gd));
    const float dod = dO[(size_t)qn * ldo + head * 32 + lane];
    const float do_d = dod * sg;
    wq[lane] = qd;
    wdo[lane] = do_d;
    __syncwarp();
    const size_t pb0 = (size_t)pair_base[b];
    // scores and softmax: lanes over keys
    float mx = -INFINITY;
    for (int j = lane; j < n_k; j += 32) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) sc = fmaf(wq[d], sK[j * 33 + d], sc);
      const size_t pair = pb0 + (q_is_prot ? ((size_t)ql * nc1 + j) : ((size_t)j * nc1 + ql));
      sc += PB[pair * 4 + head];
      pa[j] = sc;
      mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n_k; j += 32) { const float e = expf(pa[j] - mx); pa[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    // da_j = <do, v_j>,  sum_j a_j da_j
    float dsum = 0.f;
    for (int j = lane; j < n_k; j += 32) {
      const float a = pa[j] * inv;
      float da = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) da = fmaf(wdo[d], sV[j * 33 + d], da);
      pa[j] = a;
      pd[j] = da;
      dsum = fmaf(a, da, dsum);
    }
    dsum = warp_sum(dsum);
    for (int j = lane; j < n_k; j += 32) {
      const float dl = pa[j] * (pd[j] - dsum);
      pd[j] = dl;
      const size_t pair = pb0 + (q_is_prot ? ((size_t)ql * nc1 + j) : ((size_t)j * nc1 + ql));
      dPB[pair * 4 + head] = dl;
    }
    __syncwarp();
    // lane = channel: o (for the gate gradient), dq, and the accumulation into dk / dv
    float od = 0.f, dq_d = 0.f;
    for (int j = 0; j < n_k; ++j) {
      const float a = pa[j], dl = pd[j];
      od = fmaf(a, sV[j * 33 + lane], od);
      dq_d = fmaf(dl, sK[j * 33 + lane], dq_d);
      atomicAdd(&sdV[j * 33 + lane], a * do_d);
      atomicAdd(&sdK[j * 33 + lane], dl * qd);
    }
    dG[(size_t)qn * lddg + head * 32 + lane] = dod * od * sg * (1.0f - sg);
    dQ[(size_t)qn * lddq + head * 32 + lane] = dq_d * scale;
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_k * 32; i += blockDim.x) {
    const int j = i >> 5, d = i & 31;
    atomicAdd(&dK[(size_t)(k_lo + j) * lddk + head * 32 + d], sdK[j * 33 + d]);
    atomicAdd(&dV[(size_t)(k_lo + j) * lddv + head * 32 + d], sdV[j * 33 + d]);
  }
}

// Same reverse for long key lists (n_k > RB_MAXK, e.g. whole proteins in the pocket stage): keys / values are read from global
// memory (L2-resident: <= 1500 x 128 floats per side and complex), dK / dV accumulate with global atomics into buffers the caller
// zeroed; one warp per query, grid = (complex, head, query tile).  n_k <= RBB_MAXK.
constexpr int RBB_MAXK = 2048;
__global__ void __launch_bounds__(RB_WARPS * 32) row_attention_bwd_big_kernel(
    const int* __restrict__ c_off, const int* __restrict__ p_off, const int* __restrict__ pair_base, int q_is_prot,
    const float* __restrict__ Q, int ldq, const float* __restrict__ G, int ldg, const float* __restrict__ Kb, int ldk,
    const float* __restrict__ Vb, int ldv, const float* __restrict__ PB, const float* __restrict__ dO, int ldo,
    float* __restrict__ dQ, int lddq, float* __restrict__ dG, int lddg, float* __restrict__ dK, int lddk,
    float* __restrict__ dV, int lddv, float* __restrict__ dPB) {
  pdl_entry();
  extern __shared__ float rb_smem[];
  const int b = blockIdx.x, head = blockIdx.y;
  const int c_lo = c_off[b], nc1 = c_off[b + 1] - c_lo, p_lo = p_off[b], np1 = p_off[b + 1] - p_lo;
  const int n_q = q_is_prot ? np1 : nc1, n_k = q_is_prot ? nc1 : np1;
  const int q_lo = q_is_prot ? p_lo : c_lo, k_lo = q_is_prot ? c_lo : p_lo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ql = blockIdx.z * RB_WARPS + warp;
  if (ql >= n_q) return;                                         // warp-uniform; no block-level barrier below
  float* pa = rb_smem + warp * (2 * RBB_MAXK + 64);
  float* pd = pa + RBB_MAXK;
  float* wq = pd + RBB_MAXK;
  float* wdo = wq + 32;
  const float scale = 0.17677669529663687f;
  const int qn = q_lo + ql;
  const float qd = Q[(size_t)qn * ldq + head * 32 + lane] * scale;
  const float gd = G[(size_t)qn * ldg + head * 32 + lane];
  const float sg = 1.0f / (1.0f + expf(-gd));
  const float dod = dO[(size_t)qn * ldo + head * 32 + lane];
  const float do_d = dod * sg;
  wq[lane] = qd;
  wdo[lane] = do_d;
  __syncwarp();
  const size_t pb0 = (size_t)pair_base[b];
  float mx = -INFINITY;
  for (int j = lane; j < n_k; j += 32) {
    const float* kr = Kb + (size_t)(k_lo + j) * ldk + head * 32;
    float sc = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) sc = fmaf(wq[d], kr[d], sc);
    const size_t pair = pb0 + (q_is_prot ? ((size_t)ql * nc1 + j) : ((size_t)j * nc1 + ql));
    sc += PB[pair * 4 + head];
    pa[j] = sc;
    mx = fmaxf(mx, sc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n_k; j += 32) { const float e = expf(pa[j] - mx); pa[j] = e; sum += e; }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  float dsum = 0.f;
  for (int j = lane; j < n_k; j += 32) {
    const float* vr = Vb + (size_t)(k_lo + j) * ldv + head * 32;
    const float a = pa[j] * inv;
    float da = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) da = fmaf(wdo[d], vr[d], da);
    pa[j] = a;
    pd[j] = da;
    dsum = fmaf(a, da, dsum);
  }
  dsum = warp_sum(dsum);
  for (int j = lane; j < n_k; j += 32) {
    const float dl = pa[j] * (pd[j] - dsum);
    pd[j] = dl;
    const size_t pair = pb0 + (q_is_prot ? ((size_t)ql * nc1 + j) : ((size_t)j * nc1 + ql));
    dPB[pair * 4 + head] = dl;
  }
  __syncwarp();
  float od = 0.f, dq_d = 0.f;
  for (int j = 0; j < n_k; ++j) {
    const float a = pa[j], dl = pd[j];
    od = fmaf(a, Vb[(size_t)(k_lo + j) * ldv + head * 32 + lane], od);
    dq_d = fmaf(dl, Kb[(size_t)(k_lo + j) * ldk + head * 32 + lane], dq_d);
    atomicAdd(&dV[(size_t)(k_lo + j) * lddv + head * 32 + lane], a * do_d);
    atomicAdd(&dK[(size_t)(k_lo + j) * lddk + head * 32 + lane], dl * qd);
  }
  dG[(size_t)qn * lddg + head * 32 + lane] = dod * od * sg * (1.0f - sg);
  dQ[(size_t)qn * lddq + head * 32 + lane] = dq_d * scale;
}

// ---- training-mode forward pieces: the sub-steps whose intermediates the reverse pass needs and the fused inference kernels do
// not keep (unclamped coordinate steps, per-complex radial norms, attention probabilities).  GPU parity tests gated behind
// FB_EXPERIMENTAL until they have run on a B200; the forward orchestration over them is validated on the CPU. ------------------

// d[e] = x[row] - x[col], d2[e] = |d|^2
__global__ void edge_diff_kernel(const float* __restrict__ x, const int* __restrict__ row, const int* __restrict__ col, int E,
                                 float* __restrict__ d, float* __restrict__ d2) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = row[e], j = col[e];
  const float a = x[3 * i] - x[3 * j], b = x[3 * i + 1] - x[3 * j + 1], c = x[3 * i + 2] - x[3 * j + 2];
  d[3 * e] = a; d[3 * e + 1] = b; d[3 * e + 2] = c;
  d2[e] = a * a + b * b + c * c;
}
// coord2radial, norm_type per_sample (egnn.py:775-779): S[b] = sum_e d2^2 ; rn = d2 / sqrt(S[b]) ; nrm[b] = sqrt(S[b])
__global__ void radial_sum_kernel(const float* __restrict__ d2, const int* __restrict__ row, const int* __restrict__ cplx, int E,
                                  float* __restrict__ S) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < E;
  atomic_add_per_complex(S, active ? cplx[row[e]] : -1, active ? d2[e] * d2[e] : 0.f, active);
}
__global__ void radial_norm_kernel(const float* __restrict__ d2, const int* __restrict__ row, const int* __restrict__ cplx, int E,
                                   const float* __restrict__ S, int B, float* __restrict__ rn, float* __restrict__ nrm) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < B) nrm[e] = sqrtf(S[e]);
  if (e < E) rn[e] = d2[e] / sqrtf(S[cplx[row[e]]]);
}
// step = sum / max(cnt, 1) (cnt == null: sum) ; x_new = x + clamp(step, +-cmax)   (egnn.py:85-98, :228-233)
__global__ void coord_apply_kernel(const float* __restrict__ x, const float* __restrict__ sum, const float* __restrict__ cnt, int N,
                                   float cmax, float* __restrict__ step, float* __restrict__ x_new) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * N) return;
  const float s = cnt ? sum[i] / fmaxf(cnt[i / 3], 1.0f) : sum[i];
  step[i] = s;
  x_new[i] = x[i] + fminf(fmaxf(s, -cmax), cmax);
}
// scatter_softmax over destination rows stored CSR (edges of a row contiguous): one warp per row
__global__ void softmax_seg_fwd_kernel(const float* __restrict__ logit, const int* __restrict__ rowptr, int n_rows,
                                       float* __restrict__ alpha) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_rows) return;
  const int lo = rowptr[warp], hi = rowptr[warp + 1];
  float mx = -INFINITY;
  for (int e = lo + lane; e < hi; e += 32) mx = fmaxf(mx, logit[e]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int e = lo + lane; e < hi; e += 32) sum += expf(logit[e] - mx);
  sum = warp_sum(sum);
  for (int e = lo + lane; e < hi; e += 32) alpha[e] = expf(logit[e] - mx) / sum;
}
// LAS step, unclamped part (egnn.py:433-449): acc[j] += step * 4 (|x_i-x_j|^2 - |ref_i-ref_j|^2) (x_i - x_j)
__global__ void las_acc_kernel(const float* __restrict__ x, const float* __restrict__ xref, const int* __restrict__ a_idx,
                               const int* __restrict__ b_idx, int E, float step_size, float* __restrict__ acc) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int i = a_idx[e], j = b_idx[e];
  float d[3], cur = 0.f, ref = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    d[a] = x[3 * i + a] - x[3 * j + a];
    const float r = xref[3 * i + a] - xref[3 * j + a];
    cur = fmaf(d[a], d[a], cur);
    ref = fmaf(r, r, ref);
  }
  const float f = 4.0f * (cur - ref) * step_size;
#pragma unroll
  for (int a = 0; a < 3; ++a) atomicAdd(&acc[3 * j + a], f * d[a]);
}
// InteractionModule outer product operand (model_utils.py:216-220): outer[pair(i,j), :] = pc[i, :] * pc[j, :]; one CTA per protein row
__global__ void pair_outer_fwd_kernel(const float* __restrict__ pc, int H, const int* __restrict__ c_off, const int* __restrict__ p_off,
                                      const int* __restrict__ pair_base, const int* __restrict__ node_cplx, int p_begin,
                                      float* __restrict__ outer) {
  pdl_entry();
  const int i = p_begin + blockIdx.x;
  const int b = node_cplx[i];
  const int nc1 = c_off[b + 1] - c_off[b];
  const size_t base = (size_t)pair_base[b] + (size_t)(i - p_off[b]) * nc1;
  for (int f = threadIdx.x; f < H; f += blockDim.x) {
    const float pv = pc[(size_t)i * H + f];
    for (int j = 0; j < nc1; ++j) outer[(base + j) * H + f] = pv * pc[(size_t)(c_off[b] + j) * H + f];
  }
}
// gated pair bias (model_utils.py:96-133): PB[p, blk, h] = raw[p, 8 blk + h] * sigmoid(raw[p, 8 blk + 4 + h])
__global__ void pair_bias_gate_fwd_kernel(const float* __restrict__ raw, int ld, long long P, int nblk, float* __restrict__ PB) {
  pdl_entry();
  const long long total = P * nblk * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i & 3);
    const long long pb = i >> 2;
    const int blk = (int)(pb % nblk);
    const long long p = pb / nblk;
    const float* rp = raw + p * ld + blk * 8;
    PB[i] = rp[h] / (1.0f + expf(-rp[4 + h]));
  }
}

// ---- FABind+ layout: LayerNorm MLPs (P/models/model_utils.py:10-74) and the LayerNorm folded through the node-level hoisting
// (DESIGN section 1).  Compiled and bound; orchestration CPU-validated; GPU parity tests gated (FB_EXPERIMENTAL). ---------------

// LayerNorm reverse, statistics recomputed: xhat = (x - mean) rstd (written for the gamma gradient),
// dx = rstd (g - mean(g) - xhat mean(g xhat)),  g = dy * gamma.  One warp per row.
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy, int M,
                                     int D, float eps, float* __restrict__ dx, float* __restrict__ xhat) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = x + (size_t)warp * D;
  const float* g = dy + (size_t)warp * D;
  float s = 0.f;
  for (int f = lane; f < D; f += 32) s += r[f];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int f = lane; f < D; f += 32) { const float d = r[f] - mean; v = fmaf(d, d, v); }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)D + eps);
  float a = 0.f, b = 0.f;
  for (int f = lane; f < D; f += 32) {
    const float xh = (r[f] - mean) * rstd, gg = g[f] * gamma[f];
    a += gg; b = fmaf(gg, xh, b);
  }
  a = warp_sum(a) / (float)D; b = warp_sum(b) / (float)D;
  for (int f = lane; f < D; f += 32) {
    const float xh = (r[f] - mean) * rstd;
    xhat[(size_t)warp * D + f] = xh;
    dx[(size_t)warp * D + f] = rstd * (g[f] * gamma[f] - a - xh * b);
  }
}
// per-row statistics of the folded LayerNorm: s1 = sum h, s2 = sum h^2, s3 = sum h w (w optional)
__global__ void row_stats_kernel(const float* __restrict__ h, int ld, int M, int D, const float* __restrict__ w, float* __restrict__ s1,
                                 float* __restrict__ s2, float* __restrict__ s3) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* r = h + (size_t)warp * ld;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int f = lane; f < D; f += 32) { const float t = r[f]; a += t; b = fmaf(t, t, b); if (w) c = fmaf(t, w[f], c); }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if (lane == 0) { s1[warp] = a; s2[warp] = b; if (w) s3[warp] = c; }
}
// reverse: dh[m, :] += ds1[m] + 2 h[m, :] ds2[m] + w[:] ds3[m]
__global__ void row_stats_bwd_kernel(const float* __restrict__ h, int ld, int M, int D, const float* __restrict__ w,
                                     const float* __restrict__ ds1, const float* __restrict__ ds2, const float* __restrict__ ds3,
                                     float* __restrict__ dh, int lddh) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float a = ds1[warp], b = 2.0f * ds2[warp], c = (w && ds3) ? ds3[warp] : 0.f;
  for (int f = lane; f < D; f += 32)
    dh[(size_t)warp * lddh + f] += a + b * h[(size_t)warp * ld + f] + (w ? c * w[f] : 0.f);
}
// folded LayerNorm statistics of an edge row z_e = [gathered node rows | rn_e * w]:
//   mu = (A1 + rn a0) / D,  ex2 = (A2 + 2 rn A3 + rn^2 a1) / D,  var = ex2 - mu^2,  rstd = rsqrt(max(var, 0) + eps)
// (context edges: A1 = s1[row]+s1[col], A2 = s2[row]+s2[col], A3 = 0, a0 = a1 = 1;  interfacial coordinate head: A_k = s_k[col],
//  a0 = sum v_r, a1 = sum v_r^2)
__global__ void folded_stats_fwd_kernel(const float* __restrict__ A1, const float* __restrict__ A2, const float* __restrict__ A3,
                                        const float* __restrict__ rn, float a0, float a1, float D, float eps, int E,
                                        float* __restrict__ mu, float* __restrict__ var_raw, float* __restrict__ rstd) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float r = rn[e];
  const float m = (A1[e] + r * a0) / D;
  const float ex2 = (A2[e] + 2.0f * r * (A3 ? A3[e] : 0.f) + r * r * a1) / D;
  const float v = ex2 - m * m;
  mu[e] = m; var_raw[e] = v; rstd[e] = rsqrtf(fmaxf(v, 0.f) + eps);
}
__global__ void folded_stats_bwd_kernel(const float* __restrict__ A3, const float* __restrict__ rn, float a0, float a1, float D, int E,
                                        const float* __restrict__ mu, const float* __restrict__ var_raw, const float* __restrict__ rstd,
                                        const float* __restrict__ drstd, const float* __restrict__ dmu_in, float* __restrict__ dA1,
                                        float* __restrict__ dA2, float* __restrict__ dA3, float* __restrict__ drn,
                                        float* __restrict__ da /* [2] accumulated, may be null */) {
  pdl_entry();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float p0 = 0.f, p1 = 0.f;
  if (e < E) {
    const float r = rn[e], rs = rstd[e];
    const float dvar = var_raw[e] >= 0.f ? drstd[e] * (-0.5f) * rs * rs * rs : 0.f;
    const float dmu = dmu_in[e] - 2.0f * mu[e] * dvar;
    dA1[e] = dmu / D;
    dA2[e] = dvar / D;
    if (dA3) dA3[e] = 2.0f * r * dvar / D;
    drn[e] += a0 * dmu / D + (2.0f * (A3 ? A3[e] : 0.f) + 2.0f * r * a1) * dvar / D;
    p0 = r * dmu / D; p1 = r * r * dvar / D;
  }
  if (da) {
    p0 = warp_sum(p0); p1 = warp_sum(p1);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&da[0], p0); atomicAdd(&da[1], p1); }
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

int32_t fb_edge_pre_train(const float* Pn, const int32_t* row, const int32_t* col, int32_t E, int32_t H, const float* rn,
                          const float* w_rad, const float* b1, float* Z1, float* A1, void* A16, int32_t act, void* stream) {
  if (E <= 0) return FB_OK;
  if (H <= 0 || (H & 3) || !Pn || !row || !col || !rn || !w_rad || !b1 || !Z1 || !A1) return FB_ERR_BAD_ARG;
  fb_launch(edge_pre_train_kernel, dim3((int)(((long long)E * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, Pn, (const int*)row,
            (const int*)col, (int)E, (int)H, rn, w_rad, b1, Z1, A1, (bf16*)A16, (int)act);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_act_drop(const float* Z, int32_t M, int32_t N, int32_t act, float p, uint32_t seed, uint32_t site, int32_t row0,
                    int32_t colonly, float* Y, void* Y16, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if ((N & 3) || !Z || !Y || !(p >= 0.f && p < 1.f)) return FB_ERR_BAD_ARG;
  const DropCfg dc = p > 0.f ? make_drop(p, seed, site, row0, colonly) : DropCfg();
  fb_launch(act_drop_kernel, dim3(grid_1d((long long)M * N / 4, 256)), dim3(256), 0, (cudaStream_t)stream, Z, (int)M, (int)N, (int)act, dc, Y,
            (bf16*)Y16);
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

int32_t fb_act_bwd_drop(const float* Z, const float* dY, int32_t M, int32_t N, int32_t act, float p, uint32_t seed, uint32_t site,
                        int32_t row0, int32_t colonly, float* dZ, void* dZ16, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if ((N & 3) || !Z || !dY || !dZ || !(p >= 0.f && p < 1.f)) return FB_ERR_BAD_ARG;
  const DropCfg dc = p > 0.f ? make_drop(p, seed, site, row0, colonly) : DropCfg();
  fb_launch(act_bwd_drop_kernel, dim3(grid_1d((long long)M * N / 4, 256)), dim3(256), 0, (cudaStream_t)stream, Z, dY, (int)M, (int)N, (int)act,
            dc, dZ, (bf16*)dZ16);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_outer_act_bwd2(const float* Z, const float* u, const float* v, float* dZ, void* dZ16, int32_t M, int32_t N, int32_t act,
                          void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if ((N & 3) || !Z || !u || !v || !dZ) return FB_ERR_BAD_ARG;
  fb_launch(outer_act_bwd2_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, Z, u, v, dZ, (bf16*)dZ16,
            (int)M, (int)N, (int)act);
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

int32_t fb_rowdot2(const float* A, int32_t lda, const float* B, int32_t ldb, int32_t M, int32_t N, float* out, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(rowdot2_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, A, (int)lda, B, (int)ldb,
            (int)M, (int)N, out);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_rows_update(float* A, int32_t lda, int32_t M, int32_t N, const float* u, const float* v, int32_t mode, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if (mode != 0 && mode != 1) return FB_ERR_BAD_ARG;
  fb_launch(rows_update_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, A, (int)lda, (int)M,
            (int)N, u, v, (int)mode);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// dst[n, m] = bf16(src[m, n]) for m < M, 0 for M <= m < Mp: the K-major bf16 operand of a weight-gradient GEMM (reduction over the rows
// of the activation), 32 x 32 tiles through shared memory so that both the fp32 reads and the bf16 writes are row-contiguous
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const float* __restrict__ src, int ld, int M, int N, bf16* __restrict__ dst, int Mp) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const int tiles_m = (Mp + 31) / 32, tiles_n = (N + 31) / 32;
  for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int m0 = (t % tiles_m) * 32, n0 = (t / tiles_m) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i, n = n0 + tx;
      tile[ty + 8 * i][tx] = (m < M && n < N) ? src[(size_t)m * ld + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i, m = m0 + tx;
      if (n < N && m < Mp) dst[(size_t)n * Mp + m] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

int32_t fb_transpose_bf16(const float* src, int32_t ld, int32_t M, int32_t N, void* dst, int32_t Mp, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if (Mp < M) return FB_ERR_BAD_ARG;
  const long long tiles = (long long)((Mp + 31) / 32) * ((N + 31) / 32);
  fb_launch(transpose_bf16_kernel, dim3((int)(tiles < 148 * 16 ? tiles : 148 * 16)), dim3(256), 0, (cudaStream_t)stream, src, (int)ld, (int)M,
            (int)N, (bf16*)dst, (int)Mp);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// Batched form for the weight arena (training step: every matrix slot needs its transposed bf16 twin once per optimizer step, the
// operand of the data-gradient GEMMs): desc[4 i .. 4 i + 3] = {source element offset, rows, cols, destination element offset}, slot i
// = arena[off : off + rows * cols] as [rows, cols] fp32 -> dst[doff : doff + rows * cols] as [cols, rows] bf16; tile_begin[i] = first
// 32 x 32 tile of slot i in the launch's tile list (tile_begin[n] = total).  One launch instead of one per slot.
__global__ void __launch_bounds__(256) transpose_slots_kernel(const float* __restrict__ arena, const long long* __restrict__ desc,
                                                              const int* __restrict__ tile_begin, int n, bf16* __restrict__ dst) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int total = tile_begin[n];
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int lo = 0, hi = n;                       // tile_begin[lo] <= t < tile_begin[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_begin[mid] <= t) lo = mid; else hi = mid; }
    const long long off = desc[4 * lo], doff = desc[4 * lo + 3];
    const int R = (int)desc[4 * lo + 1], Cc = (int)desc[4 * lo + 2];
    const int tiles_r = (R + 31) / 32, local = t - tile_begin[lo];
    const int r0 = (local % tiles_r) * 32, c0 = (local / tiles_r) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      tile[ty + 8 * i][tx] = (r < R && c < Cc) ? arena[off + (long long)r * Cc + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;
      if (c < Cc && r < R) dst[doff + (long long)c * R + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

int32_t fb_transpose_slots_bf16(const float* arena, const int64_t* desc, const int32_t* tile_begin, int32_t n, int32_t n_tiles, void* dst,
                                void* stream) {
  if (n <= 0 || n_tiles <= 0) return FB_OK;
  if (!arena || !desc || !tile_begin || !dst) return FB_ERR_BAD_ARG;
  fb_launch(transpose_slots_kernel, dim3(n_tiles < 148 * 16 ? n_tiles : 148 * 16), dim3(256), 0, (cudaStream_t)stream, arena,
            (const long long*)desc, (const int*)tile_begin, (int)n, (bf16*)dst);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// dst[m, n] = keep(seed, site, row0 + m, n) ? src[m, n] / (1 - p) : 0  -- the library's counter-based dropout mask (common.cuh) as a
// stand-alone op: the training-mode forward applies it where the inference path has it fused into an epilogue, and the reverse
// pass applies the SAME mask to the incoming gradient (the mask is a pure function of its coordinates, nothing is stored).
__global__ void dropout_apply_kernel(const float* __restrict__ src, float* __restrict__ dst, int ld, int M, int N, DropCfg dc) {
  pdl_entry();
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    dst[(size_t)m * ld + n] = drop_apply(src[(size_t)m * ld + n], dc, m, n);
  }
}

int32_t fb_dropout_apply(const float* src, float* dst, int32_t ld, int32_t M, int32_t N, float p, uint32_t seed, uint32_t site,
                         int32_t row0, int32_t colonly, void* stream) {
  if (M <= 0 || N <= 0) return FB_OK;
  if (!(p >= 0.f && p < 1.f)) return FB_ERR_BAD_ARG;
  const DropCfg dc = make_drop(p, seed, site, row0, colonly);
  if (p == 0.f) {
    if (src != dst) cudaMemcpy2DAsync(dst, sizeof(float) * ld, src, sizeof(float) * ld, sizeof(float) * N, M, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return FB_OK;
  }
  fb_launch(dropout_apply_kernel, dim3(grid_1d((long long)M * N, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, (int)ld, (int)M, (int)N, dc);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_vec_op(const float* a, const float* b, float* c, int64_t n, int32_t op, void* stream) {
  if (n <= 0) return FB_OK;
  if (op < 0 || op > 2) return FB_ERR_BAD_ARG;
  fb_launch(vec_op_kernel, dim3(grid_1d(n, 256)), dim3(256), 0, (cudaStream_t)stream, a, b, c, (long long)n, (int)op);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_softmax_seg_bwd(const float* alpha, const float* dalpha, const int32_t* row, int32_t E, float* t_zeroed, float* dlogit,
                           void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(softmax_seg_bwd_sum_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, alpha, dalpha, row, (int)E, t_zeroed);
  fb_launch(softmax_seg_bwd_apply_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, alpha, dalpha, row, (int)E,
            (const float*)t_zeroed, dlogit);
  count_launch(2);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pair_bias_gate_bwd(const float* raw, int32_t ld, int64_t P, int32_t nblk, const float* dPB, float* draw, void* stream) {
  if (P <= 0) return FB_OK;
  if (8 * nblk > ld) return FB_ERR_BAD_ARG;
  fb_launch(pair_bias_gate_bwd_kernel, dim3(grid_1d(P * ld, 256)), dim3(256), 0, (cudaStream_t)stream, raw, (int)ld, (long long)P,
            (int)nblk, dPB, draw);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pair_outer_bwd(const float* dO, const float* pc, int32_t H, const int32_t* c_off, const int32_t* p_off,
                          const int32_t* pair_base, const int32_t* node_cplx, int32_t p_begin, int32_t n_p_rows, float* dpc,
                          void* stream) {
  if (n_p_rows <= 0) return FB_OK;
  fb_launch(pair_outer_bwd_kernel, dim3(n_p_rows), dim3(128), 0, (cudaStream_t)stream, dO, pc, (int)H, c_off, p_off, pair_base, node_cplx,
            (int)p_begin, dpc);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_row_attention_bwd(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t q_is_prot,
                             int32_t max_q, int32_t max_k, const float* Q, int32_t ldq, const float* G, int32_t ldg, const float* K, int32_t ldk,
                             const float* V, int32_t ldv, const float* PB, const float* dO, int32_t ldo, float* dQ, int32_t lddq,
                             float* dG, int32_t lddg, float* dK, int32_t lddk, float* dV, int32_t lddv, float* dPB, void* stream) {
  if (B <= 0) return FB_OK;
  if (max_k > RB_MAXK) {
    // long key lists: global-memory variant (dK / dV must be zeroed by the caller: they are accumulated with atomics)
    if (max_k > RBB_MAXK || max_q <= 0) return FB_ERR_UNSUPPORTED;
    static unsigned long long done_big = 0;
    const int smem = RB_WARPS * (2 * RBB_MAXK + 64) * 4;
    if (!ensure_smem_optin(row_attention_bwd_big_kernel, smem, done_big)) return FB_ERR_CUDA;
    fb_launch(row_attention_bwd_big_kernel, dim3(B, 4, (max_q + RB_WARPS - 1) / RB_WARPS), dim3(RB_WARPS * 32), smem, (cudaStream_t)stream,
              c_off, p_off, pair_base, (int)q_is_prot, Q, (int)ldq, G, (int)ldg, K, (int)ldk, V, (int)ldv, PB, dO, (int)ldo, dQ, (int)lddq,
              dG, (int)lddg, dK, (int)lddk, dV, (int)lddv, dPB);
    count_launch(1);
    FB_CHECK_LAUNCH();
    return FB_OK;
  }
  static unsigned long long done = 0;
  if (!ensure_smem_optin(row_attention_bwd_kernel, RB_SMEM_FLOATS * 4, done)) return FB_ERR_CUDA;
  if (max_q <= 0) return FB_OK;
  // shared memory sized to the keys that exist; queries split over grid.z: two per warp (16 per CTA) for long query lists, one per
  // warp for short ones.  dK / dV are ACCUMULATED (atomics): the caller zeroes them, as for the long-key variant above.
  const int KC = (max_k + 31) & ~31;
  const int q_tile = max_q > 64 ? 2 * RB_WARPS : RB_WARPS;
  const int smem = (4 * KC * 33 + RB_WARPS * (2 * KC + 64)) * 4;
  fb_launch(row_attention_bwd_kernel, dim3(B, 4, (max_q + q_tile - 1) / q_tile), dim3(RB_WARPS * 32), smem, (cudaStream_t)stream, c_off, p_off,
            pair_base, (int)q_is_prot, Q, (int)ldq, G, (int)ldg, K, (int)ldk, V, (int)ldv, PB, dO, (int)ldo, dQ, (int)lddq, dG, (int)lddg, dK,
            (int)lddk, dV, (int)lddv, dPB, KC, q_tile);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_radial_fwd(const float* x, const int32_t* row, const int32_t* col, int32_t E, const int32_t* node_cplx, int32_t B,
                      float* S_zeroed, float* d, float* d2, float* rn, float* nrm, void* stream) {
  if (E <= 0) return FB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = ((E > B ? E : B) + 255) / 256;
  fb_launch(edge_diff_kernel, dim3((E + 255) / 256), dim3(256), 0, st, x, row, col, (int)E, d, d2);
  fb_launch(radial_sum_kernel, dim3((E + 255) / 256), dim3(256), 0, st, (const float*)d2, row, node_cplx, (int)E, S_zeroed);
  fb_launch(radial_norm_kernel, dim3(g), dim3(256), 0, st, (const float*)d2, row, node_cplx, (int)E, (const float*)S_zeroed, (int)B, rn, nrm);
  count_launch(3);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_coord_apply(const float* x, const float* sum, const float* cnt, int32_t N, float cmax, float* step, float* x_new, void* stream) {
  if (N <= 0) return FB_OK;
  fb_launch(coord_apply_kernel, dim3((3 * N + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, sum, cnt, (int)N, cmax, step, x_new);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_softmax_seg_fwd(const float* logit, const int32_t* rowptr, int32_t n_rows, float* alpha, void* stream) {
  if (n_rows <= 0) return FB_OK;
  fb_launch(softmax_seg_fwd_kernel, dim3((int)(((long long)n_rows * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, logit, rowptr,
            (int)n_rows, alpha);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_las_acc(const float* x, const float* xref, const int32_t* a_idx, const int32_t* b_idx, int32_t E, float step_size,
                   float* acc_zeroed, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(las_acc_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x, xref, a_idx, b_idx, (int)E, step_size, acc_zeroed);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pair_outer_fwd(const float* pc, int32_t H, const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base,
                          const int32_t* node_cplx, int32_t p_begin, int32_t n_p_rows, float* outer, void* stream) {
  if (n_p_rows <= 0) return FB_OK;
  fb_launch(pair_outer_fwd_kernel, dim3(n_p_rows), dim3(128), 0, (cudaStream_t)stream, pc, (int)H, c_off, p_off, pair_base, node_cplx,
            (int)p_begin, outer);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_pair_bias_gate_fwd(const float* raw, int32_t ld, int64_t P, int32_t nblk, float* PB, void* stream) {
  if (P <= 0) return FB_OK;
  if (8 * nblk > ld) return FB_ERR_BAD_ARG;
  fb_launch(pair_bias_gate_fwd_kernel, dim3(grid_1d(P * nblk * 4, 256)), dim3(256), 0, (cudaStream_t)stream, raw, (int)ld, (long long)P,
            (int)nblk, PB);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_row_attention_fwd(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t q_is_prot,
                             int32_t max_q, int32_t max_k, const float* Q, int32_t ldq, const float* G, int32_t ldg, const float* K,
                             int32_t ldk, const float* V, int32_t ldv, const float* PB, float* O, int32_t ldo, void* stream) {
  GraphDev g;
  g.B = B; g.c_off = c_off; g.p_off = p_off; g.pair_base = pair_base;
  return row_attention(g, q_is_prot, max_q, max_k, Q, ldq, G, ldg, K, ldk, V, ldv, PB, O, ldo, false, (cudaStream_t)stream);
}

int32_t fb_layernorm_bwd(const float* x, const float* gamma, const float* dy, int32_t M, int32_t D, float eps, float* dx, float* xhat,
                         void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(layernorm_bwd_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, gamma, dy, (int)M,
            (int)D, eps, dx, xhat);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_row_stats(const float* h, int32_t ld, int32_t M, int32_t D, const float* w, float* s1, float* s2, float* s3, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(row_stats_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, h, (int)ld, (int)M, (int)D,
            w, s1, s2, s3);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_row_stats_bwd(const float* h, int32_t ld, int32_t M, int32_t D, const float* w, const float* ds1, const float* ds2,
                         const float* ds3, float* dh, int32_t lddh, void* stream) {
  if (M <= 0) return FB_OK;
  fb_launch(row_stats_bwd_kernel, dim3((int)(((long long)M * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, h, (int)ld, (int)M,
            (int)D, w, ds1, ds2, ds3, dh, (int)lddh);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_folded_stats_fwd(const float* A1, const float* A2, const float* A3, const float* rn, float a0, float a1, float D, float eps,
                            int32_t E, float* mu, float* var_raw, float* rstd, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(folded_stats_fwd_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, A1, A2, A3, rn, a0, a1, D, eps, (int)E, mu,
            var_raw, rstd);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int32_t fb_folded_stats_bwd(const float* A3, const float* rn, float a0, float a1, float D, int32_t E, const float* mu, const float* var_raw,
                            const float* rstd, const float* drstd, const float* dmu_in, float* dA1, float* dA2, float* dA3, float* drn,
                            float* da, void* stream) {
  if (E <= 0) return FB_OK;
  fb_launch(folded_stats_bwd_kernel, dim3((E + 255) / 256), dim3(256), 0, (cudaStream_t)stream, A3, rn, a0, a1, D, (int)E, mu, var_raw, rstd,
            drstd, dmu_in, dA1, dA2, dA3, drn, da);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // extern "C"
