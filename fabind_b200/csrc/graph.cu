// Radius-graph construction for the docking stack (replaces ComplexGraph.construct_edges,
// reference FABind/fabind/models/att_model.py:38-128).
//
// One warp per row node; candidates are the nodes of the same complex.  Two passes (count, fill)
// around an exclusive scan give deterministic, row-sorted edge lists without atomics.
//
// Distance predicate: the reference tests `torch.norm(xi - xj, dim=-1) <= cutoff` in fp32.  On the
// reference's CPU path that norm is sqrt(fma(dz,dz, fma(dy,dy, dx*dx))) (verified bit-for-bit on
// 2e6 random triples), which is what edge_dist() spells out with IEEE round-to-nearest intrinsics so
// that borderline pairs land on the same side of the cutoff.
#include "graph.h"

namespace fb {

__device__ __forceinline__ float edge_dist(const float* __restrict__ x, int i, int j) {
  const float dx = __fsub_rn(x[3 * i + 0], x[3 * j + 0]);
  const float dy = __fsub_rn(x[3 * i + 1], x[3 * j + 1]);
  const float dz = __fsub_rn(x[3 * i + 2], x[3 * j + 2]);
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// category of the ordered pair (r, c), r != c, same complex:
//  0 protein-protein within intra cutoff | 1 global-normal (same segment) | 2 global-global
//  3 compound<->protein within inter cutoff | -1 no edge
__device__ __forceinline__ int edge_category(const float* __restrict__ x, const uint8_t* __restrict__ flags,
                                             int r, int c, float intra, float inter) {
  const uint8_t fr = flags[r], fc = flags[c];
  const bool sr = fr & 1, sc = fc & 1, gr = fr & 2, gc = fc & 2;
  if (!gr && !gc) {
    if (sr == sc) {
      if (!sr) return -1;
      return edge_dist(x, r, c) <= intra ? 0 : -1;
    }
    return edge_dist(x, r, c) <= inter ? 3 : -1;
  }
  if (sr == sc) return 1;
  return (gr && gc) ? 2 : -1;
}

// ---------------------------------------------------------------------------------------------
// internal (type-sorted) graph: ctx CSR (bond edges first, then geometric edges by ascending col)
// and inter CSR.  MODE bit0: produce ctx, bit1: produce inter.  FILL: second pass.
// ---------------------------------------------------------------------------------------------
template <int MODE, bool FILL>
__global__ void __launch_bounds__(256) graph_rows_kernel(GraphDev g, const float* __restrict__ x,
                                                         float intra, float inter) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= g.N) return;
  const int r = warp;
  const int b = g.node_cplx[r];
  const unsigned lt = (1u << lane) - 1u;
  int n_ctx = 0, n_int = 0;
  int base_ctx = 0, base_int = 0;
  if (FILL) {
    if (MODE & 1) base_ctx = g.ctx_rowptr[r];
    if (MODE & 2) base_int = g.int_rowptr[r];
  }
  const bool r_prot = g.node_flags[r] & 1;
  if (MODE & 1) {
    // bond edges (compound_edge_index, converted to internal ids) whose source is r, input order
    for (int e0 = 0; e0 < g.n_bond; e0 += 32) {
      const int e = e0 + lane;
      const bool hit = e < g.n_bond && g.bond_row[e] == r;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const int pos = base_ctx + n_ctx + __popc(m & lt);
        if (pos < g.ctx_cap) { g.ctx_row[pos] = r; g.ctx_col[pos] = g.bond_col[e]; }
      }
      n_ctx += __popc(m);
    }
  }
  const int c_lo = g.c_off[b], c_hi = g.c_off[b + 1], p_lo = g.p_off[b], p_hi = g.p_off[b + 1];
  const int nc1 = c_hi - c_lo;
  const int total = (c_hi - c_lo) + (p_hi - p_lo);
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    int c = -1, cat = -1;
    if (t < total) {
      c = t < nc1 ? c_lo + t : p_lo + (t - nc1);
      if (c != r) cat = edge_category(x, g.node_flags, r, c, intra, inter);
    }
    if (MODE & 1) {
      const bool hit = cat >= 0 && cat <= 2;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const int pos = base_ctx + n_ctx + __popc(m & lt);
        if (pos < g.ctx_cap) { g.ctx_row[pos] = r; g.ctx_col[pos] = c; }
      }
      n_ctx += __popc(m);
    }
    if (MODE & 2) {
      bool hit = cat == 3;
      if (FILL && *g.int_fallback) {
        // reference fallback (att_model.py:85-86): no inter edge in the whole batch -> the first
        // candidate pair (first ligand atom, first residue of complex 0) in both directions
        const int fa = g.fb_atom, fr = g.fb_res;
        hit = (r == fa && c == fr) || (r == fr && c == fa);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const int pos = base_int + n_int + __popc(m & lt);
        g.int_row[pos] = r;
        g.int_col[pos] = c;
        const int ci = r_prot ? c : r, pi = r_prot ? r : c;
        g.int_pair[pos] = g.pair_base[b] + (pi - p_lo) * nc1 + (ci - c_lo);
      }
      n_int += __popc(m);
    }
  }
  if (!FILL && lane == 0) {
    if (MODE & 1) g.ctx_deg[r] = n_ctx;
    if (MODE & 2) g.int_deg[r] = n_int;
  }
}

// exclusive scan of deg[0..n) into rowptr[0..n]; single block.  For the inter list it also applies
// the zero-edge fallback (sets the flag and gives the two designated rows degree 1).
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ deg, int* __restrict__ rowptr,
                                                    int n, int* fallback, int fa, int fr, int* total_out) {
  pdl_entry();
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  __shared__ int fb_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int pass = 0; pass < 2; ++pass) {
    if (tid == 0) carry_s = 0;
    __syncthreads();
    const bool use_fb = pass == 1;
    for (int i0 = 0; i0 < n; i0 += 1024) {
      const int i = i0 + tid;
      int v = 0;
      if (i < n) v = use_fb ? ((i == fa || i == fr) ? 1 : 0) : deg[i];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) warp_tot[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += t;
        }
        warp_tot[lane] = w;
      }
      __syncthreads();
      const int carry = carry_s;
      const int excl = carry + (wid ? warp_tot[wid - 1] : 0) + inc - v;
      if (i < n) rowptr[i] = excl;
      __syncthreads();
      if (tid == 1023) carry_s = carry + warp_tot[31];
      __syncthreads();
    }
    if (tid == 0) {
      rowptr[n] = carry_s;
      fb_s = (fallback != nullptr && pass == 0 && carry_s == 0) ? 1 : 0;
      if (fallback != nullptr && pass == 0) *fallback = fb_s;
    }
    __syncthreads();
    if (!fb_s) break;
  }
  if (tid == 0 && total_out) *total_out = rowptr[n];   // (per-iteration edge statistics of the forward: saves a 4-byte copy node)
}

__global__ void convert_edges_kernel(const long long* __restrict__ e, int n_e, const int* __restrict__ inv,
                                     int* __restrict__ row, int* __restrict__ col) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_e) {
    row[i] = inv[(int)e[i]];
    col[i] = inv[(int)e[n_e + i]];
  }
}

// CSR over destination (index 1) of the LAS pairs: lasr_rowptr via count+scan, stable fill
template <bool FILL>
__global__ void __launch_bounds__(256) las_rows_kernel(GraphDev g) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= g.N) return;
  const int r = warp;
  const unsigned lt = (1u << lane) - 1u;
  int n = 0;
  const int base = FILL ? g.las_rowptr[r] : 0;
  {
    for (int e0 = 0; e0 < g.n_las; e0 += 32) {
      const int e = e0 + lane;
      const bool hit = e < g.n_las && g.las_dst[e] == r;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) g.las_csr_src[base + n + __popc(m & lt)] = g.las_src[e];
      n += __popc(m);
    }
  }
  if (!FILL && lane == 0) g.las_deg[r] = n;
}

static inline int warp_grid(int n_rows) { return (n_rows * 32 + 255) / 256; }

// moving rows (flag bit2) in ascending order + exclusive scan of their context degrees; single block.  Also publishes the two edge
// counts the host reads back (counts[0] = E_ctx, counts[1] = E_mv).
__global__ void __launch_bounds__(1024) mv_index_kernel(GraphDev g) {
  pdl_entry();
  __shared__ int tot_c[32], tot_d[32];
  __shared__ int carry_c, carry_d;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry_c = 0; carry_d = 0; }
  __syncthreads();
  for (int i0 = 0; i0 < g.N; i0 += 1024) {
    const int i = i0 + tid;
    const int f = (i < g.N && (g.node_flags[i] & 4)) ? 1 : 0;
    const int d = f ? g.ctx_rowptr[i + 1] - g.ctx_rowptr[i] : 0;
    int ic = f, id = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, ic, o), td = __shfl_up_sync(0xffffffffu, id, o);
      if (lane >= o) { ic += tc; id += td; }
    }
    if (lane == 31) { tot_c[wid] = ic; tot_d[wid] = id; }
    __syncthreads();
    if (wid == 0) {
      int wc = tot_c[lane], wd = tot_d[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tc = __shfl_up_sync(0xffffffffu, wc, o), td = __shfl_up_sync(0xffffffffu, wd, o);
        if (lane >= o) { wc += tc; wd += td; }
      }
      tot_c[lane] = wc; tot_d[lane] = wd;
    }
    __syncthreads();
    const int pos = carry_c + (wid ? tot_c[wid - 1] : 0) + ic - f;
    const int off = carry_d + (wid ? tot_d[wid - 1] : 0) + id - d;
    if (f && pos < g.n_mv) { g.mv_rows[pos] = i; g.mv_rowptr[pos] = off; }
    __syncthreads();
    if (tid == 1023) { carry_c += tot_c[31]; carry_d += tot_d[31]; }
    __syncthreads();
  }
  if (tid == 0) {
    g.mv_rowptr[g.n_mv] = carry_d;
    g.counts[0] = g.ctx_rowptr[g.N];
    g.counts[1] = carry_c == g.n_mv ? carry_d : 0;    // a host / device disagreement on the number of moving rows switches the subset off
  }
}

// compact edge lists of the moving rows: one warp per row
__global__ void __launch_bounds__(256) mv_fill_kernel(GraphDev g) {
  pdl_entry();
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= g.n_mv) return;
  const int r = g.mv_rows[j], lo = g.ctx_rowptr[r], n = g.ctx_rowptr[r + 1] - lo, dst = g.mv_rowptr[j];
  for (int k = lane; k < n && dst + k < g.mv_cap; k += 32) {
    g.mv_erow[dst + k] = r;
    g.mv_ecol[dst + k] = g.ctx_col[lo + k];
    g.mv_emap[dst + k] = lo + k;
  }
}

// Host-supplied edge counts (fb_model_params.layout_flag): the dataloader counted the context edges on the CPU, so the host sized
// the edge-level scratch and the GEMM row counts without reading the device.  One block checks the claim against the device-side
// counts; a wrong claim raises the flag and clamps both CSR row pointers into the claimed sizes, so that every later consumer stays
// inside the buffers the host sized (results are then garbage by contract, and flagged).
__global__ void __launch_bounds__(1024) verify_counts_kernel(GraphDev g, int e_ctx, int e_mv, int* __restrict__ flag) {
  pdl_entry();
  if (g.counts[0] == e_ctx && g.counts[1] == e_mv) return;
  if (threadIdx.x == 0) *flag = 1;
  for (int i = threadIdx.x; i <= g.N; i += blockDim.x) g.ctx_rowptr[i] = min(g.ctx_rowptr[i], e_ctx);
  for (int i = threadIdx.x; i <= g.n_mv; i += blockDim.x) g.mv_rowptr[i] = min(g.mv_rowptr[i], e_mv);
}

int graph_verify_counts(const GraphDev& g, int e_ctx, int e_mv, int* flag, cudaStream_t st) {
  fb_launch(verify_counts_kernel, dim3(1), dim3(1024), 0, st, g, e_ctx, e_mv, flag);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_mv_index(const GraphDev& g, cudaStream_t st) {
  fb_launch(mv_index_kernel, dim3(1), dim3(1024), 0, st, g);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_mv_fill(const GraphDev& g, cudaStream_t st) {
  if (g.n_mv <= 0) return FB_OK;
  fb_launch(mv_fill_kernel, dim3(warp_grid(g.n_mv)), dim3(256), 0, st, g);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_prepare_static(const GraphDev& g, const long long* bonds, const long long* las,
                         cudaStream_t st) {
  if (g.n_bond > 0) fb_launch(convert_edges_kernel, dim3((g.n_bond + 255) / 256), dim3(256), 0, st, bonds, g.n_bond, g.inv, g.bond_row, g.bond_col);
  if (g.n_las > 0) fb_launch(convert_edges_kernel, dim3((g.n_las + 255) / 256), dim3(256), 0, st, las, g.n_las, g.inv, g.las_src, g.las_dst);
  fb_launch(las_rows_kernel<false>, dim3(warp_grid(g.N)), dim3(256), 0, st, g);
  fb_launch(scan_kernel, dim3(1), dim3(1024), 0, st, g.las_deg, g.las_rowptr, g.N, nullptr, -1, -1, nullptr);
  fb_launch(las_rows_kernel<true>, dim3(warp_grid(g.N)), dim3(256), 0, st, g);
  count_launch(3 + (g.n_bond > 0) + (g.n_las > 0));
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_count_ctx(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st) {
  fb_launch(graph_rows_kernel<1, false>, dim3(warp_grid(g.N)), dim3(256), 0, st, g, x, intra, inter);
  fb_launch(scan_kernel, dim3(1), dim3(1024), 0, st, g.ctx_deg, g.ctx_rowptr, g.N, nullptr, -1, -1, nullptr);
  count_launch(2);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_fill_ctx(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st) {
  fb_launch(graph_rows_kernel<1, true>, dim3(warp_grid(g.N)), dim3(256), 0, st, g, x, intra, inter);
  count_launch(1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_build_inter(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st, int* total_out) {
  fb_launch(graph_rows_kernel<2, false>, dim3(warp_grid(g.N)), dim3(256), 0, st, g, x, intra, inter);
  fb_launch(scan_kernel, dim3(1), dim3(1024), 0, st, g.int_deg, g.int_rowptr, g.N, g.int_fallback, g.fb_atom, g.fb_res, total_out);
  fb_launch(graph_rows_kernel<2, true>, dim3(warp_grid(g.N)), dim3(256), 0, st, g, x, intra, inter);
  count_launch(3);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// reference-order edge lists (API parity for ComplexGraph.construct_edges): nodes in the caller's
// order, candidates = the contiguous node range of the same complex, four category lists each
// sorted by (row, col).
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) ref_rows_kernel(int N, const int* __restrict__ cplx, const int* __restrict__ off,
                                                       const uint8_t* __restrict__ flags, const float* __restrict__ x,
                                                       float intra, float inter, int* __restrict__ deg /*[4][N]*/,
                                                       const int* __restrict__ rowptr /*[4][N+1]*/,
                                                       const int* __restrict__ cat_base /*[4]*/, const int* fallback,
                                                       long long* __restrict__ ctx_out, int e_ctx,
                                                       long long* __restrict__ int_out, int e_int) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const int r = warp, b = cplx[r];
  const int lo = off[b], hi = off[b + 1];
  const unsigned lt = (1u << lane) - 1u;
  int n[4] = {0, 0, 0, 0};
  for (int c0 = lo; c0 < hi; c0 += 32) {
    const int c = c0 + lane;
    int cat = -1;
    if (c < hi && c != r) cat = edge_category(x, flags, r, c, intra, inter);
    if (FILL && *fallback) {
      const int fa = off[0] + 1;
      int fr = -1;  // first non-global protein node of complex 0
      for (int k = off[0]; k < off[1]; ++k) if ((flags[k] & 3) == 1) { fr = k; break; }
      cat = (c < hi && ((r == fa && c == fr) || (r == fr && c == fa))) ? 3 : (cat == 3 ? -1 : cat);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool hit = cat == k;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (FILL && hit) {
        const int pos = rowptr[k * (N + 1) + r] + n[k] + __popc(m & lt);
        if (k < 3) {
          ctx_out[cat_base[k] + pos] = r;
          ctx_out[e_ctx + cat_base[k] + pos] = c;
        } else {
          int_out[pos] = r;
          int_out[e_int + pos] = c;
        }
      }
      n[k] += __popc(m);
    }
  }
  if (!FILL && lane == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) deg[k * N + r] = n[k];
  }
}

__global__ void ref_fallback_fix_kernel(int N, int* deg, const int* off, const uint8_t* flags) {
  pdl_entry();
  // single thread: if no inter candidate survived anywhere, give the two fallback rows degree 1
  int fa = off[0] + 1, fr = -1;
  for (int k = off[0]; k < off[1]; ++k) if ((flags[k] & 3) == 1) { fr = k; break; }
  if (fr >= 0) { deg[3 * N + fa] = 1; deg[3 * N + fr] = 1; }
}

int graph_ref_count(int N, const int* cplx, const int* off, const uint8_t* flags, const float* x,
                    float intra, float inter, int* deg, int* rowptr, int* fallback, cudaStream_t st) {
  fb_launch(ref_rows_kernel<false>, dim3(warp_grid(N)), dim3(256), 0, st, N, cplx, off, flags, x, intra, inter, deg, nullptr, nullptr,
                                                        nullptr, nullptr, 0, nullptr, 0);
  for (int k = 0; k < 4; ++k)
    fb_launch(scan_kernel, dim3(1), dim3(1024), 0, st, deg + (size_t)k * N, rowptr + (size_t)k * (N + 1), N,
                                    k == 3 ? fallback : nullptr, -1, -1, nullptr);
  count_launch(5);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

int graph_ref_fill(int N, const int* cplx, const int* off, const uint8_t* flags, const float* x,
                   float intra, float inter, int* deg, int* rowptr, const int* cat_base, int* fallback,
                   int fallback_host, long long* ctx_out, int e_ctx, long long* int_out, int e_int,
                   cudaStream_t st) {
  if (fallback_host) {
    fb_launch(ref_fallback_fix_kernel, dim3(1), dim3(1), 0, st, N, deg, off, flags);
    fb_launch(scan_kernel, dim3(1), dim3(1024), 0, st, deg + (size_t)3 * N, rowptr + (size_t)3 * (N + 1), N, nullptr, -1, -1, nullptr);
  }
  fb_launch(ref_rows_kernel<true>, dim3(warp_grid(N)), dim3(256), 0, st, N, cplx, off, flags, x, intra, inter, deg, rowptr, cat_base,
                                                       fallback, ctx_out, e_ctx, int_out, e_int);
  count_launch(fallback_host ? 3 : 1);
  FB_CHECK_LAUNCH();
  return FB_OK;
}

}  // namespace fb
