// Ligand post-optimisation on the GPU (reference: FABind/fabind/utils/post_optim_utils.py:8-64, called once per ligand on the
// CPU by fabind_inference.py:285-316 -- 1000 Adam steps each, the dominant cost of end-to-end inference once the model is fast).
//
//   minimise  sum_{(i,j) in LAS} | |x_i - x_j| - |r_i - r_j| |  +  2 sum_{i,j} relu(1.22 - |x_i - x_j|)      (LAS mask given)
//             sum_{i,j} | |x_i - x_j| - |r_i - r_j| |                                                          (rigid: no mask)
//   over ORDERED pairs, with torch.optim.Adam(lr = 0.1, betas = (0.9, 0.999), eps = 1e-8), x_0 = predicted coordinates.
//
// One CTA per ligand, the whole optimisation in ONE launch: coordinates, reference coordinates, Adam moments and the LAS bit
// matrix live in shared memory; four threads share an atom (each sums a quarter of the j loop, two shuffles combine them), the
// step counter runs inside the kernel.  Distances are evaluated from coordinate differences (exact near zero); the reference's
// torch.cdist switches to the |a|^2 + |b|^2 - 2ab form above 25 atoms, whose fp32 cancellation error (~1e-4 A at |x| ~ 50 A)
// makes its own trajectory depend on the ligand's absolute position -- parity there is at the level of the outcome, see tests.
#include "../../include/fabind_b200.h"
#include "common.cuh"

namespace fb {

constexpr int PO_TPA = 4;            // threads per atom
constexpr int PO_THREADS = 256;      // 64 atoms in flight per pass

__global__ void __launch_bounds__(PO_THREADS)
post_optimize_kernel(const float* __restrict__ ref, const float* __restrict__ pred, const int* __restrict__ atom_off,
                     const int* __restrict__ las, const int* __restrict__ las_off, int n_las_total, int epochs, float lr,
                     int use_mask, int max_n, float* __restrict__ out, float* __restrict__ out_loss, float* __restrict__ out_rmsd) {
  pdl_entry();
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int a0 = atom_off[b], n = atom_off[b + 1] - a0;
  const int words = (max_n + 31) >> 5;
  float* x = sm;                       // [max_n][3]
  float* xn = x + 3 * max_n;           // next iterate
  float* r = xn + 3 * max_n;           // reference conformer
  float* m = r + 3 * max_n;            // Adam first moment
  float* v = m + 3 * max_n;            // Adam second moment
  unsigned* bits = reinterpret_cast<unsigned*>(v + 3 * max_n);   // [max_n][words] LAS adjacency
  __shared__ float red[PO_THREADS / 32];
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
    x[i] = pred[3 * a0 + i]; r[i] = ref[3 * a0 + i]; m[i] = 0.f; v[i] = 0.f;
  }
  for (int i = threadIdx.x; i < n * words; i += blockDim.x) bits[i] = 0u;
  __syncthreads();
  if (use_mask) {
    // to_dense_adj(LAS_edge_index): adj[e0, e1] = 1 (post_optim_utils.py:39); edges carry ligand-local atom ids
    for (int e = las_off[b] + threadIdx.x; e < las_off[b + 1]; e += blockDim.x) {
      const int i = las[e], j = las[n_las_total + e];
      if (i >= 0 && i < n && j >= 0 && j < n) atomicOr(&bits[i * words + (j >> 5)], 1u << (j & 31));
    }
  }
  __syncthreads();
  const int grp = threadIdx.x / PO_TPA, sub = threadIdx.x % PO_TPA;
  double p1 = 1.0, p2 = 1.0;           // beta^t in double, as python computes the bias corrections
  float loss_last = 0.f;
  for (int t = 1; t <= epochs; ++t) {
    p1 *= 0.9; p2 *= 0.999;
    const float step = (float)((double)lr / (1.0 - p1));
    const float bc2s = (float)sqrt(1.0 - p2);
    float lsum = 0.f;
    for (int i0 = 0; i0 < n; i0 += PO_THREADS / PO_TPA) {
      const int i = i0 + grp;
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (i < n) {
        const float xi0 = x[3 * i], xi1 = x[3 * i + 1], xi2 = x[3 * i + 2];
        const float ri0 = r[3 * i], ri1 = r[3 * i + 1], ri2 = r[3 * i + 2];
        for (int j = sub; j < n; j += PO_TPA) {
          const float d0 = xi0 - x[3 * j], d1 = xi1 - x[3 * j + 1], d2 = xi2 - x[3 * j + 2];
          const float d = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
          const float c0 = ri0 - r[3 * j], c1 = ri1 - r[3 * j + 1], c2 = ri2 - r[3 * j + 2];
          const float err = d - sqrtf(c0 * c0 + c1 * c1 + c2 * c2);
          const float sg = (err > 0.f) ? 1.f : ((err < 0.f) ? -1.f : 0.f);
          float w;   // w_ij + w_ji: the ordered pairs (i,j) and (j,i) both move x_i along (x_i - x_j) / d
          if (use_mask) {
            const float mij = (float)((bits[i * words + (j >> 5)] >> (j & 31)) & 1u);
            const float mji = (float)((bits[j * words + (i >> 5)] >> (i & 31)) & 1u);
            const float rep = d < 1.22f ? 1.f : 0.f;
            w = sg * (mij + mji) - 4.0f * rep;
            lsum += fabsf(err) * mij + 2.0f * fmaxf(1.22f - d, 0.f);
          } else {
            w = 2.0f * sg;
            lsum += fabsf(err);
          }
          if (d > 0.f) {
            const float s = w / d;
            g0 = fmaf(s, d0, g0); g1 = fmaf(s, d1, g1); g2 = fmaf(s, d2, g2);
          }
        }
      }
#pragma unroll
      for (int o = 1; o < PO_TPA; o <<= 1) {
        g0 += __shfl_xor_sync(0xffffffffu, g0, o); g1 += __shfl_xor_sync(0xffffffffu, g1, o); g2 += __shfl_xor_sync(0xffffffffu, g2, o);
      }
      if (i < n && sub == 0) {
        const float g[3] = {g0, g1, g2};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float mm = m[3 * i + k], vv = v[3 * i + k];
          mm = mm + 0.1f * (g[k] - mm);                                  // exp_avg.lerp_(grad, 1 - beta1)
          vv = vv * 0.999f + (float)(1.0 - 0.999) * g[k] * g[k];         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
          m[3 * i + k] = mm; v[3 * i + k] = vv;
          const float denom = sqrtf(vv) / bc2s + 1e-8f;
          xn[3 * i + k] = x[3 * i + k] - step * (mm / denom);
        }
      }
    }
    if (t == epochs) {   // the reference reports the loss evaluated BEFORE the last step (post_optim_utils.py:50-58)
      lsum = warp_sum(lsum);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) x[i] = xn[i];
    if (t == epochs && threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < PO_THREADS / 32; ++w) s += red[w];
      loss_last = s;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) out[3 * a0 + i] = x[i];
  // compute_RMSD(reference, x): sqrt(mean_i |r_i - x_i|^2)  (post_optim_utils.py:5-6,60)
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float e0 = r[3 * i] - x[3 * i], e1 = r[3 * i + 1] - x[3 * i + 1], e2 = r[3 * i + 2] - x[3 * i + 2];
    s += e0 * e0 + e1 * e1 + e2 * e2;
  }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < PO_THREADS / 32; ++w) tot += red[w];
    if (out_rmsd) out_rmsd[b] = n > 0 ? sqrtf(tot / (float)n) : 0.f;
    if (out_loss) out_loss[b] = loss_last;
  }
}

}  // namespace fb

using namespace fb;

extern "C" int32_t fb_post_optimize(const float* ref_coords, const float* pred_coords, const int32_t* atom_off, int32_t B,
                                    int32_t max_atoms, const int32_t* las_edges, const int32_t* las_off, int32_t n_las_total,
                                    int32_t epochs, float lr, float* out_coords, float* out_loss, float* out_rmsd, void* stream) {
  if (B <= 0) return FB_OK;
  if (!ref_coords || !pred_coords || !atom_off || !out_coords || max_atoms <= 0 || epochs < 0) return FB_ERR_BAD_ARG;
  const int words = (max_atoms + 31) / 32;
  const size_t smem = (size_t)15 * max_atoms * sizeof(float) + (size_t)max_atoms * words * sizeof(unsigned);
  if (smem > 200 * 1024) return FB_ERR_UNSUPPORTED;      // ~ 600 atoms
  static unsigned long long optin = 0;
  if (!ensure_smem_optin(post_optimize_kernel, 200 * 1024, optin)) return FB_ERR_CUDA;
  fb_launch(post_optimize_kernel, dim3(B), dim3(PO_THREADS), smem, (cudaStream_t)stream, ref_coords, pred_coords, atom_off, las_edges,
            las_off, n_las_total, epochs, lr, las_edges != nullptr ? 1 : 0, max_atoms, out_coords, out_loss, out_rmsd);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}
