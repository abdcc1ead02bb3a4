// Device-side description of one batch of complexes in the INTERNAL node order:
// all compound-side nodes (glb_c + ligand atoms) of all complexes first, then all protein-side nodes
// (glb_p + residues).  Inside a side, nodes keep the caller's order, so the dense per-complex blocks
// the reference builds with to_dense_batch (egnn.py:260-265) are plain row slices here.
#pragma once
#include "common.cuh"

namespace fb {

struct GraphDev {
  int N = 0, B = 0, Nc_tot = 0;
  int n_bond = 0, n_las = 0;
  int fb_atom = 0, fb_res = 0;          // internal ids used by the zero-inter-edge fallback
  // layout (host-built, read-only on device)
  const int* perm = nullptr;            // [N] internal -> caller index
  const int* inv = nullptr;             // [N] caller -> internal index
  const int* node_cplx = nullptr;       // [N]
  const uint8_t* node_flags = nullptr;  // [N] bit0 protein side, bit1 global node, bit2 moves between iterations
  const int* c_off = nullptr;           // [B+1] compound-side node range of each complex
  const int* p_off = nullptr;           // [B+1] protein-side node range (absolute internal ids)
  const int* pair_base = nullptr;       // [B+1] first pair row of each complex (pair = prot_local * nc1 + comp_local)
  // bond / LAS lists in internal ids
  int* bond_row = nullptr; int* bond_col = nullptr;   // [n_bond]
  int* las_src = nullptr; int* las_dst = nullptr;     // [n_las]
  int* las_deg = nullptr; int* las_rowptr = nullptr;  // [N], [N+1] CSR over destination
  int* las_csr_src = nullptr;                         // [n_las]
  // context graph (static per forward)
  int* ctx_deg = nullptr; int* ctx_rowptr = nullptr;  // [N], [N+1]
  int* ctx_row = nullptr; int* ctx_col = nullptr;     // [E_ctx]
  // interface graph (rebuilt every refinement iteration)
  int* int_deg = nullptr; int* int_rowptr = nullptr;  // [N], [N+1]
  int* int_row = nullptr; int* int_col = nullptr; int* int_pair = nullptr;  // [cap_int]
  int* int_fallback = nullptr;                        // [1]
  float* xtmp = nullptr;                              // [3N] coordinates in internal order (count pass)
  // moving rows (node_flags bit2) and the context edges INTO them, compacted: the out_layer of a non-final iteration only has to
  // produce the coordinates of these rows (att_model.py:232-236)
  int n_mv = 0;                                       // number of moving rows (host-known)
  int* mv_rows = nullptr;                             // [n_mv] internal ids, ascending
  int* mv_rowptr = nullptr;                           // [n_mv + 1] CSR over the compact edge list
  int* mv_erow = nullptr; int* mv_ecol = nullptr;     // [E_mv] endpoints
  int* mv_emap = nullptr;                             // [E_mv] index of the edge in the full context list
  int* counts = nullptr;                              // [2] E_ctx, E_mv (read by the host after fb_graph_static)
  // capacities of ctx_row / ctx_col and of the compact moving-rows lists: every fill is bounded by them.  They only bind when the
  // host SUPPLIED the counts (fb_model_params.layout_flag, a dataloader-side layout) and was wrong; see graph_verify_counts
  int ctx_cap = 0x7fffffff, mv_cap = 0x7fffffff;
};

int graph_prepare_static(const GraphDev& g, const long long* bonds, const long long* las, cudaStream_t st);
int graph_count_ctx(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st);
int graph_fill_ctx(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st);
// moving rows: index + compact row pointers (after graph_count_ctx), compact edge lists (after graph_fill_ctx)
int graph_mv_index(const GraphDev& g, cudaStream_t st);
int graph_mv_fill(const GraphDev& g, cudaStream_t st);
// host-supplied counts (after graph_mv_index): device counts != (e_ctx, e_mv) -> *flag = 1 and ctx_rowptr / mv_rowptr clamped to them
int graph_verify_counts(const GraphDev& g, int e_ctx, int e_mv, int* flag, cudaStream_t st);
// total_out (optional, device): receives the number of interface edges of this build
int graph_build_inter(const GraphDev& g, const float* x, float intra, float inter, cudaStream_t st, int* total_out = nullptr);

int graph_ref_count(int N, const int* cplx, const int* off, const uint8_t* flags, const float* x,
                    float intra, float inter, int* deg, int* rowptr, int* fallback, cudaStream_t st);
int graph_ref_fill(int N, const int* cplx, const int* off, const uint8_t* flags, const float* x,
                   float intra, float inter, int* deg, int* rowptr, const int* cat_base, int* fallback,
                   int fallback_host, long long* ctx_out, int e_ctx, long long* int_out, int e_int,
                   cudaStream_t st);

}  // namespace fb
