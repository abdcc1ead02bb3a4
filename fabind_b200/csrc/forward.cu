// Host-side orchestration of the docking stack: weight-arena layout, scratch planning and the launch
// sequence of EfficientMCAttModel.forward.  No allocation, no synchronisation: everything is enqueued
// on the caller's stream.
#include <atomic>
#include <cstdlib>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>

#include "../../include/fabind_b200.h"
#include "gemm.h"
#include "layers.h"

namespace fb {
extern long long* g_tc_dbg;

// ------------------------------------------------------------------------------------------------
// bookkeeping: launch counter and optional per-category CUDA-event timing (bench.py roofline leg)
// ------------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches += n; }
// Diagnostic environment knobs exist only in -DFB_DIAG builds (python -m fabind_b200.build --diag); the release library reads no
// environment variable.  FB_PDL=0 disables programmatic dependent launch; FB_SKIP_CATS=<bitmask> (scripts/dev/skip_probe.sh) drops
// every launch of the masked categories so the in-situ cost of a category can be read off as a difference of step times (results
// are garbage).
#ifdef FB_DIAG
bool pdl_enabled() {
  static bool on = [] { const char* e = getenv("FB_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
static int skip_mask() {
  static int m = [] { const char* e = getenv("FB_SKIP_CATS"); return e ? atoi(e) : 0; }();
  return m;
}
static int dup_mask() {
  static int m = [] { const char* e = getenv("FB_KDUP"); return e ? atoi(e) : 0; }();
  return m;
}
#else
bool pdl_enabled() { return true; }
static constexpr int skip_mask() { return 0; }
static constexpr int dup_mask() { return 0; }
#endif

struct ProfSpan { cudaEvent_t a, b; int cat; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_spans;       // recorded spans of the current profile window
static std::vector<ProfSpan> g_pool;        // reusable events
static int g_open_cat = -1;
static cudaEvent_t g_open_a, g_open_b;
static double g_flops[CAT_COUNT] = {0};     // 2 M N K of the GEMMs enqueued per category while profiling (M = row CAPACITY for
                                            // problems with a device-side row count)
static void prof_flops(int cat, const GemmArgs& a) {
  if (g_prof_on && cat >= 0 && cat < CAT_COUNT) g_flops[cat] += 2.0 * a.M * a.N * (a.K1 + a.K2);
}

void prof_begin(int cat, cudaStream_t st) {
  if (!g_prof_on) return;
  if (g_pool.empty()) {
    ProfSpan s; s.cat = 0;
    cudaEventCreate(&s.a); cudaEventCreate(&s.b);
    g_pool.push_back(s);
  }
  ProfSpan s = g_pool.back(); g_pool.pop_back();
  g_open_cat = cat; g_open_a = s.a; g_open_b = s.b;
  cudaEventRecord(g_open_a, st);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_open_cat < 0) return;
  cudaEventRecord(g_open_b, st);
  g_spans.push_back({g_open_a, g_open_b, g_open_cat});
  g_open_cat = -1;
}

constexpr int HD = 128;  // RowAttentionBlock: 4 heads x 32 channels (cross_att.py:98)
static inline int pb_cols(int L) { return (16 * L + 127) / 128 * 128; }
constexpr int QKX = 128; // extra columns of the stacked q|k GEMM: inter_layer linear_p (32) | linear_c (32) | pad

// ------------------------------------------------------------------------------------------------
// weight arena layout
// ------------------------------------------------------------------------------------------------
struct Slot { std::string name; int64_t rows, cols, off; };

struct GclW { int64_t e1_rc, e1_rad, e1_b, e2_w, e2_b, c1_w, c1_b, c2_w, n1_w, n1_b, n2_w, n2_b, f_e1b; };
struct AttW {
  int64_t ca_c_w, ca_c_b, ca_p_w, ca_p_b, ca_p2_w, o_p_w, o_p_b, o_c_w, o_c_b;
  int64_t tp1_w, tp1_b, tp2_w, tp2_b, tc1_w, tc1_b, tc2_w, tc2_b;
  int64_t pt1_w, pt1_b, pt2v, pt_c;
  int64_t qk_w, qk_b, k_r, v_r, ac1_b, ac2_w, ac_u;
  // derived ("folded") slots, filled by fb_derive_weights from the slots above (names start with f_: the packers skip them)
  int64_t f_cac_w, f_cac_b, f_cap_w, f_cap_b, f_l3_w, f_l3_b, f_l5_w, f_l5_b, f_qkc_w, f_qkc_b;
};
// FABind+ layout (LayerNorm -> Linear -> ReLU -> Linear [-> ReLU] MLPs, P/models/model_utils.py:10-74)
static inline int dp_of(int H) { return (2 * H + 1 + 63) / 64 * 64; }   // 2H+1 edge-MLP features padded to a multiple of 64
struct GclPW {
  int64_t e1_rc, e1_rad, e1_g, e1_c0, e2_w, e2_b;          // edge_mlp: LN folded into the hoisted first Linear
  int64_t cl_g, cl_b, c1_w, c1_b, c2_w;                     // coord_mlp: explicit LN, Linear+ReLU, row-dot
  int64_t nl_g, nl_b, n1_w, n1_b, n2_w, n2_b;               // node_mlp: explicit LN over [h | agg]
};
struct AttPW {
  int64_t ca_c_w, ca_c_b, ca_p_w, ca_p_b, ca_p2_w, o_p_w, o_p_b, o_c_w, o_c_b;
  int64_t tpl_g, tpl_b, tp1_w, tp1_b, tp2_w, tp2_b, tcl_g, tcl_b, tc1_w, tc1_b, tc2_w, tc2_b;
  int64_t pb_w, pb_b;                                       // pair-bias projections of THIS layer (pair changes per layer)
  int64_t zo_w, zo_b, zl_g, zl_b, pt1_w, pt1_b, pt2_w, pt2_b, wb, pt_c;
  int64_t qk_w, qk_b, k_r, v_r, ac_c0, ac2_w, ac_u, ac_g, ac_r;
};
struct ModelW {
  int64_t in_w, in_b, out_w, out_b, il_p_w, il_p_b, il_c_w, il_c_b, il_o_w, il_o_b, pb_w, pb_b;
  std::vector<GclW> gcl;  // n_layers + 1 (last = out_layer)
  std::vector<AttW> att;
  std::vector<GclPW> gclp;
  std::vector<AttPW> attp;
  int flavour = 0;
  int64_t total = 0;
  std::vector<Slot> slots;
};

static void build_weights(int H, int L, ModelW& w) {
  int64_t off = 0;
  auto add = [&](const std::string& name, int64_t rows, int64_t cols) {
    const int64_t o = off;
    w.slots.push_back({name, rows, cols, o});
    off += (rows * cols + 63) / 64 * 64;  // 256-byte aligned slots
    return o;
  };
  w.in_w = add("in_w", H, H); w.in_b = add("in_b", 1, H);
  w.out_w = add("out_w", H, H); w.out_b = add("out_b", 1, H);
  w.il_p_w = add("il_p_w", H, H); w.il_p_b = add("il_p_b", 1, H);
  w.il_c_w = add("il_c_w", H, H); w.il_c_b = add("il_c_b", 1, H);
  w.il_o_w = add("il_o_w", H, H); w.il_o_b = add("il_o_b", 1, H);
  // pair-bias projections of all layers, zero-padded to a multiple of 128 outputs so the GEMM tiles on tcgen05
  w.pb_w = add("pb_w", pb_cols(L), H); w.pb_b = add("pb_b", 1, pb_cols(L));
  for (int i = 0; i <= L; ++i) {
    const std::string p = i < L ? "gcl" + std::to_string(i) + "." : std::string("out.");
    GclW g;
    g.e1_rc = add(p + "e1_rc", 2 * H, H); g.e1_rad = add(p + "e1_rad", 1, H); g.e1_b = add(p + "e1_b", 1, H);
    g.e2_w = add(p + "e2_w", H, H); g.e2_b = add(p + "e2_b", 1, H);
    g.c1_w = add(p + "c1_w", H, H); g.c1_b = add(p + "c1_b", 1, H); g.c2_w = add(p + "c2_w", 1, H);
    g.n1_w = add(p + "n1_w", H, 2 * H); g.n1_b = add(p + "n1_b", 1, H);
    g.n2_w = add(p + "n2_w", H, H); g.n2_b = add(p + "n2_b", 1, H);
    g.f_e1b = 0;
    w.gcl.push_back(g);
  }
  for (int i = 0; i < L; ++i) {
    const std::string p = "att" + std::to_string(i) + ".";
    AttW a;
    a.ca_c_w = add(p + "ca_c_w", 4 * HD, H); a.ca_c_b = add(p + "ca_c_b", 1, 4 * HD);
    a.ca_p_w = add(p + "ca_p_w", 2 * HD, H); a.ca_p_b = add(p + "ca_p_b", 1, 2 * HD);
    a.ca_p2_w = add(p + "ca_p2_w", 2 * HD, H);
    a.o_p_w = add(p + "o_p_w", H, HD); a.o_p_b = add(p + "o_p_b", 1, H);
    a.o_c_w = add(p + "o_c_w", H, HD); a.o_c_b = add(p + "o_c_b", 1, H);
    a.tp1_w = add(p + "tp1_w", 2 * H, H); a.tp1_b = add(p + "tp1_b", 1, 2 * H);
    a.tp2_w = add(p + "tp2_w", H, 2 * H); a.tp2_b = add(p + "tp2_b", 1, H);
    a.tc1_w = add(p + "tc1_w", 2 * H, H); a.tc1_b = add(p + "tc1_b", 1, 2 * H);
    a.tc2_w = add(p + "tc2_w", H, 2 * H); a.tc2_b = add(p + "tc2_b", 1, H);
    a.pt1_w = add(p + "pt1_w", 2 * H, H + 64); a.pt1_b = add(p + "pt1_b", 1, 2 * H);
    a.pt2v = add(p + "pt2v", 1, 2 * H); a.pt_c = add(p + "pt_c", 1, 1);
    // one stacked node GEMM: q | k | inter32_p | inter32_c | pad  ||  v | vc   (vc = coord_mlp.0 applied to v, folded)
    a.qk_w = add(p + "qk_w", 4 * H + QKX, H); a.qk_b = add(p + "qk_b", 1, 4 * H + QKX); a.k_r = add(p + "k_r", 1, H);
    a.v_r = add(p + "v_r", 1, H);
    a.ac1_b = add(p + "ac1_b", 1, H);
    a.ac2_w = add(p + "ac2_w", 1, H); a.ac_u = add(p + "ac_u", 1, H);
    w.att.push_back(a);
  }
  // ---- DERIVED slots, all behind the last base slot (a packer / the training path works on the base prefix of the arena) ----
  // [e1_b | 0]: the bias of the stacked per-node projection GEMM -- the edge kernel then adds nothing but the radial term
  for (int i = 0; i <= L; ++i) {
    const std::string p = i < L ? "gcl" + std::to_string(i) + "." : std::string("out.");
    w.gcl[i].f_e1b = add(p + "f_e1b", 1, 2 * H);
  }
  // folded projections (see Run::run_att_folded): consecutive Linear maps of the cross-attention block collapsed into
  // pre-multiplied weights over K-concatenated operands, so that dependent launches become independent problems of one launch
  //   f_cac / f_cap : [W_ca | W_ca W_n2]            the block's first projections from [h | T1] of the preceding MC_E_GCL
  //   f_l3          : [W_p2 | W_p2 W_op] (2 HD rows), [W_tp1 | W_tp1 W_op] (2H rows)     k/v of the new p, p transition hidden from [h_p | O_p]
  //   f_l5          : [W_tc1 | W_tc1 W_oc]          c transition hidden from [h_c | O_c]
  //   f_qkc         : [W_qk | W_qk W_tc2]           stacked q|k|v|vc of the compound rows from [h_c | TH_c]
  for (int i = 0; i < L; ++i) {
    const std::string p = "att" + std::to_string(i) + ".";
    AttW& a = w.att[i];
    a.f_cac_w = add(p + "f_cac_w", 4 * HD, 2 * H); a.f_cac_b = add(p + "f_cac_b", 1, 4 * HD);
    a.f_cap_w = add(p + "f_cap_w", 2 * HD, 2 * H); a.f_cap_b = add(p + "f_cap_b", 1, 2 * HD);
    a.f_l3_w = add(p + "f_l3_w", 2 * HD + 2 * H, H + HD); a.f_l3_b = add(p + "f_l3_b", 1, 2 * HD + 2 * H);
    a.f_l5_w = add(p + "f_l5_w", 2 * H, H + HD); a.f_l5_b = add(p + "f_l5_b", 1, 2 * H);
    a.f_qkc_w = add(p + "f_qkc_w", 4 * H + QKX, 3 * H); a.f_qkc_b = add(p + "f_qkc_b", 1, 4 * H + QKX);
  }
  w.total = off;
}

// out[r, 0:K) = Wa[r, :],  out[r, K + j) = sum_k Wa[r, k] Wb[k, j],  ob[r] = ba[r] + sum_k Wa[r, k] bb[k]   (fp64 accumulation,
// rounded once: the same arithmetic as the packer's float64 derivations).  Runs when the weights change, not per forward.
__global__ void __launch_bounds__(256) fold_weights_kernel(const float* __restrict__ Wa, const float* __restrict__ ba, int R, int K,
                                                           const float* __restrict__ Wb, const float* __restrict__ bb, int J,
                                                           float* __restrict__ out, float* __restrict__ ob) {
  const int j = blockIdx.x * 32 + threadIdx.x;          // 0 .. K + J (copy columns first), one extra column for the bias
  const int r = blockIdx.y * 8 + threadIdx.y;
  if (r >= R) return;
  const int ldo = K + J;
  if (j < K) { out[(size_t)r * ldo + j] = Wa[(size_t)r * K + j]; return; }
  const int jj = j - K;
  if (jj > J) return;
  double acc = 0.0;
  if (jj < J) {
    for (int k = 0; k < K; ++k) acc += (double)Wa[(size_t)r * K + k] * (double)Wb[(size_t)k * J + jj];
    out[(size_t)r * ldo + j] = (float)acc;
  } else {
    for (int k = 0; k < K; ++k) acc += (double)Wa[(size_t)r * K + k] * (double)bb[k];
    ob[r] = (float)(acc + (ba ? (double)ba[r] : 0.0));
  }
}

static int fold_weights(float* w, int64_t wa, int64_t ba, int R, int K, int64_t wb, int64_t bb, int J, int64_t out, int64_t ob,
                        cudaStream_t st) {
  const dim3 grid((K + J + 1 + 31) / 32, (R + 7) / 8), block(32, 8);
  fold_weights_kernel<<<grid, block, 0, st>>>(w + wa, ba >= 0 ? w + ba : nullptr, R, K, w + wb, w + bb, J, w + out, w + ob);
  return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_CUDA;
}

static void build_weights_plus(int H, int L, ModelW& w) {
  int64_t off = 0;
  auto add = [&](const std::string& name, int64_t rows, int64_t cols) {
    const int64_t o = off;
    w.slots.push_back({name, rows, cols, o});
    off += (rows * cols + 63) / 64 * 64;
    return o;
  };
  const int Dp = dp_of(H);
  w.flavour = 1;
  w.in_w = add("in_w", H, H); w.in_b = add("in_b", 1, H);
  w.out_w = add("out_w", H, H); w.out_b = add("out_b", 1, H);
  w.il_p_w = add("il_p_w", H, H); w.il_p_b = add("il_p_b", 1, H);
  w.il_c_w = add("il_c_w", H, H); w.il_c_b = add("il_c_b", 1, H);
  w.il_o_w = add("il_o_w", H, H); w.il_o_b = add("il_o_b", 1, H);
  w.pb_w = w.pb_b = 0;
  for (int i = 0; i <= L; ++i) {
    const std::string p = i < L ? "gcl" + std::to_string(i) + "." : std::string("out.");
    GclPW g;
    g.e1_rc = add(p + "e1_rc", 2 * Dp, H); g.e1_rad = add(p + "e1_rad", 1, Dp); g.e1_g = add(p + "e1_g", 1, Dp);
    g.e1_c0 = add(p + "e1_c0", 1, Dp);
    g.e2_w = add(p + "e2_w", H, Dp); g.e2_b = add(p + "e2_b", 1, H);
    g.cl_g = add(p + "cl_g", 1, H); g.cl_b = add(p + "cl_b", 1, H);
    g.c1_w = add(p + "c1_w", H, H); g.c1_b = add(p + "c1_b", 1, H); g.c2_w = add(p + "c2_w", 1, H);
    g.nl_g = add(p + "nl_g", 1, 2 * H); g.nl_b = add(p + "nl_b", 1, 2 * H);
    g.n1_w = add(p + "n1_w", 2 * H, 2 * H); g.n1_b = add(p + "n1_b", 1, 2 * H);
    g.n2_w = add(p + "n2_w", H, 2 * H); g.n2_b = add(p + "n2_b", 1, H);
    w.gclp.push_back(g);
  }
  for (int i = 0; i < L; ++i) {
    const std::string p = "att" + std::to_string(i) + ".";
    AttPW a;
    a.ca_c_w = add(p + "ca_c_w", 4 * HD, H); a.ca_c_b = add(p + "ca_c_b", 1, 4 * HD);
    a.ca_p_w = add(p + "ca_p_w", 2 * HD, H); a.ca_p_b = add(p + "ca_p_b", 1, 2 * HD);
    a.ca_p2_w = add(p + "ca_p2_w", 2 * HD, H);
    a.o_p_w = add(p + "o_p_w", H, HD); a.o_p_b = add(p + "o_p_b", 1, H);
    a.o_c_w = add(p + "o_c_w", H, HD); a.o_c_b = add(p + "o_c_b", 1, H);
    a.tpl_g = add(p + "tpl_g", 1, H); a.tpl_b = add(p + "tpl_b", 1, H);
    a.tp1_w = add(p + "tp1_w", H, H); a.tp1_b = add(p + "tp1_b", 1, H);
    a.tp2_w = add(p + "tp2_w", H, H); a.tp2_b = add(p + "tp2_b", 1, H);
    a.tcl_g = add(p + "tcl_g", 1, H); a.tcl_b = add(p + "tcl_b", 1, H);
    a.tc1_w = add(p + "tc1_w", H, H); a.tc1_b = add(p + "tc1_b", 1, H);
    a.tc2_w = add(p + "tc2_w", H, H); a.tc2_b = add(p + "tc2_b", 1, H);
    a.pb_w = add(p + "pb_w", 128, H); a.pb_b = add(p + "pb_b", 1, 128);
    a.zo_w = add(p + "zo_w", 32, H) /* inter_layer.linear_out.weight TRANSPOSED */; a.zo_b = add(p + "zo_b", 1, H);
    a.zl_g = add(p + "zl_g", 1, H); a.zl_b = add(p + "zl_b", 1, H);
    a.pt1_w = add(p + "pt1_w", H, H); a.pt1_b = add(p + "pt1_b", 1, H);
    a.pt2_w = add(p + "pt2_w", H, H); a.pt2_b = add(p + "pt2_b", 1, H);
    a.wb = add(p + "wb", 1, H); a.pt_c = add(p + "pt_c", 1, 1);
    a.qk_w = add(p + "qk_w", 4 * H + QKX, H); a.qk_b = add(p + "qk_b", 1, 4 * H + QKX); a.k_r = add(p + "k_r", 1, H);
    a.v_r = add(p + "v_r", 1, H);
    a.ac_c0 = add(p + "ac_c0", 1, H); a.ac2_w = add(p + "ac2_w", 1, H); a.ac_u = add(p + "ac_u", 1, H);
    a.ac_g = add(p + "ac_g", 1, H); a.ac_r = add(p + "ac_r", 1, 2);
    w.attp.push_back(a);
  }
  w.total = off;
}

static const ModelW& weights_for(int H, int L, int flavour = 0) {
  struct Key { int H, L, f; ModelW* w; };
  static std::vector<Key> cache;
  for (auto& kv : cache)
    if (kv.H == H && kv.L == L && kv.f == flavour) return *kv.w;
  ModelW* w = new ModelW();
  if (flavour == 1) build_weights_plus(H, L, *w); else build_weights(H, L, *w);
  cache.push_back({H, L, flavour, w});
  return *w;
}

// ------------------------------------------------------------------------------------------------
// scratch arena (dry run = size query)
// ------------------------------------------------------------------------------------------------
struct Arena {
  char* base; size_t cap; size_t off = 0; size_t peak = 0; bool dry; bool ok = true;
  Arena(void* b, size_t c, bool d) : base((char*)b), cap(c), dry(d) {}
  void* take(size_t bytes) {
    const size_t a = (off + 255) / 256 * 256;
    off = a + bytes;
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)(a + 256);  // non-null fake
    if (off > cap) { ok = false; return nullptr; }
    return base + a;
  }
  template <typename T> T* get(size_t n) { return (T*)take(n * sizeof(T)); }
};

struct GraphBufs { GraphDev g; };

static void plan_graph(const fb_model_params& p, Arena& a, GraphDev& g) {
  g.N = p.N; g.B = p.B; g.Nc_tot = p.Nc_tot; g.n_bond = p.n_bond; g.n_las = p.n_las;
  g.fb_atom = p.fb_atom; g.fb_res = p.fb_res;
  g.perm = p.perm; g.inv = p.inv; g.node_cplx = p.node_cplx; g.node_flags = p.node_flags;
  g.c_off = p.c_off; g.p_off = p.p_off; g.pair_base = p.pair_base;
  g.bond_row = a.get<int>(p.n_bond + 1); g.bond_col = a.get<int>(p.n_bond + 1);
  g.las_src = a.get<int>(p.n_las + 1); g.las_dst = a.get<int>(p.n_las + 1);
  g.las_deg = a.get<int>(p.N); g.las_rowptr = a.get<int>(p.N + 1); g.las_csr_src = a.get<int>(p.n_las + 1);
  g.ctx_deg = a.get<int>(p.N); g.ctx_rowptr = a.get<int>(p.N + 1);
  g.int_deg = a.get<int>(p.N); g.int_rowptr = a.get<int>(p.N + 1);
  g.int_fallback = a.get<int>(1);
  g.xtmp = a.get<float>(3 * (size_t)p.N);
  g.n_mv = p.n_mv > 0 && p.n_mv <= p.N ? p.n_mv : 0;
  g.mv_rows = a.get<int>((size_t)g.n_mv + 1); g.mv_rowptr = a.get<int>((size_t)g.n_mv + 1);
  g.counts = a.get<int>(2);
}

struct Bufs {
  // edges
  int *ctx_row, *ctx_col, *int_row, *int_col, *int_pair, *mv_erow, *mv_ecol, *mv_emap;
  // coordinates
  float *x_state, *xa, *xb, *xl;
  // node features
  float *Hin32, *h, *h2, *h0, *pc, *CAc, *CAp, *CAp2, *QK, *Hfin;
  void *HinT, *hT, *hT2, *h0T, *Pn0, *agg, *T1, *O, *TH, *VT, *Pn, *VCT;
  void *CAcT, *CApT, *CAp2T;   // bf16 projections of the cross-attention block (tcgen05 attention core)
  // pair
  void *P0, *A0, *Zg, *T64; float *PBraw, *PB, *pb_dense, *dotU;
  // edge
  float *radc, *normc, *radi, *normi, *dotE, *lgt, *sde;
  // FABind+ only
  float *hstat, *vstat, *dotP; void *M2, *TH2, *Tn, *PairA, *PairB, *Zl, *Zh; void *A1, *M;
  void* split_ws; size_t split_bytes;
};

static void plan_main(const fb_model_params& p, Arena& a, Bufs& b) {
  const size_t N = p.N, H = p.hidden, E = p.E_ctx > 0 ? p.E_ctx : 1, P = p.P_total, L = p.n_layers;
  const size_t capI = p.cap_int > 0 ? p.cap_int : 2, capU = capI / 2 + 1;
  const bool bf = p.bf16_mode == FB_PREC_BF16;
  const bool plus = p.flavour == FB_FLAVOUR_PLUS;
  const int gmode = p.bf16_mode;
  const size_t TS = bf ? 2 : 4;
  const size_t Nc = p.Nc_tot, Np = N - Nc;
  const size_t tilesH = gemm_dot_tiles((int)E, (int)H, (int)H, gmode), tiles2H = gemm_dot_tiles((int)(capI / 2), (int)(2 * H), (int)H, gmode);
  const size_t Dp = dp_of((int)H);
  b.ctx_row = a.get<int>(E); b.ctx_col = a.get<int>(E);
  b.int_row = a.get<int>(capI); b.int_col = a.get<int>(capI); b.int_pair = a.get<int>(capI);
  {
    const size_t Emv = p.E_ctx_mv > 0 ? p.E_ctx_mv : 1;
    b.mv_erow = a.get<int>(Emv); b.mv_ecol = a.get<int>(Emv); b.mv_emap = a.get<int>(Emv);
  }
  b.x_state = a.get<float>(3 * N); b.xa = a.get<float>(3 * N); b.xb = a.get<float>(3 * N); b.xl = a.get<float>(3 * N);
  b.Hin32 = a.get<float>(N * H);
  b.HinT = bf ? a.take(N * H * TS) : (void*)b.Hin32;
  b.h = a.get<float>(N * H);
  b.hT = bf ? a.take(N * H * TS) : (void*)b.h;
  // second residual stream: the folded sequences write the new h next to the old one, which other problems of the same launch
  // still read as their A operand (Run::run_att_folded)
  b.h2 = b.h; b.hT2 = b.hT;
  b.h0 = nullptr; b.h0T = nullptr; b.Pn0 = nullptr;
  if (!plus) {
    b.h2 = a.get<float>(N * H); b.hT2 = bf ? a.take(N * H * TS) : (void*)b.h2;
    // iteration-invariant head of the stack: linear_in(H) and the first layer's per-node edge-MLP projections of it
    b.h0 = a.get<float>(N * H); b.h0T = bf ? a.take(N * H * TS) : (void*)b.h0;
    b.Pn0 = a.take(N * 2 * H * TS);
  }
  b.Hfin = a.get<float>(N * H);
  b.pc = a.get<float>(N * H);
  b.P0 = a.take(P * H * TS);
  b.PB = a.get<float>(P * (plus ? 1 : L) * 8);
  b.pb_dense = a.get<float>(P);
  // per-sub-layer temporaries
  b.Pn = a.take(N * 2 * (plus ? Dp : H) * TS);
  b.radc = a.get<float>(E); b.normc = a.get<float>((size_t)p.B * RAD_SLICES);
  b.A1 = a.take(E * (plus ? Dp : H) * TS); b.M = a.take(E * H * TS);
  if (plus) {
    b.hstat = a.get<float>(3 * N); b.vstat = a.get<float>(3 * N);
    b.M2 = a.take(E * H * TS); b.TH2 = a.take(N * 2 * H * TS); b.Tn = a.take(N * H * TS);
    b.PairA = a.take(P * H * TS); b.PairB = a.take(P * H * TS); b.Zl = a.take(P * H * TS); b.Zh = a.take(P * H * TS);
    b.dotP = a.get<float>((size_t)gemm_dot_tiles((int)P, (int)H, (int)H, gmode) * P);
  }
  {
    const size_t tiles_small = gemm_dot_tiles(1, (int)H, (int)H, gmode);   // short edge lists (compound rows only) take the narrow tiles
    b.dotE = a.get<float>((tilesH > tiles_small ? tilesH : tiles_small) * E);
  }
  b.agg = a.take(N * H * TS); b.T1 = a.take(N * H * TS);
  b.CAc = a.get<float>((Nc + 1) * 4 * HD); b.CAp = a.get<float>((Np + 1) * 2 * HD); b.CAp2 = a.get<float>((Np + 1) * 2 * HD);
  b.CAcT = b.CApT = b.CAp2T = nullptr;
  if (bf) { b.CAcT = a.take((Nc + 1) * 4 * HD * 2); b.CApT = a.take((Np + 1) * 2 * HD * 2); b.CAp2T = a.take((Np + 1) * 2 * HD * 2); }
  b.O = a.take(N * HD * TS); b.TH = a.take(N * 2 * H * TS);
  b.Zg = a.take(capU * H * TS); b.T64 = a.take(capU * 64 * TS);
  b.dotU = a.get<float>(tiles2H * capU);
  b.radi = a.get<float>(capI); b.lgt = a.get<float>(capI); b.sde = a.get<float>(capI); b.normi = a.get<float>((size_t)p.B * RAD_SLICES);
  b.QK = a.get<float>(N * (2 * H + QKX));
  b.VT = a.take(N * 2 * H * TS);   // [N, 2H] typed: v | vc
  b.VCT = nullptr;
  // pair0 construction temporaries (alive only before the iteration loop, but kept simple: own space)
  b.A0 = a.take(P * H * TS);
  b.PBraw = a.get<float>(P * (plus ? 128 : pb_cols((int)L)));
  // split-precision modes: scratch for the three bf16 planes of the largest A operand (edge rows x Dp / H, pair rows x H,
  // unique interface pairs x (H + 64), node rows x 2H)
  b.split_ws = nullptr; b.split_bytes = 0;
  if (gmode >= FB_PREC_SPLIT3) {
    size_t el = E * (plus ? Dp : H);
    if (P * H > el) el = P * H;
    if (capU * (H + 64) > el) el = capU * (H + 64);
    if (N * 2 * H > el) el = N * 2 * H;
    b.split_bytes = el * 3 * sizeof(bf16);
    b.split_ws = a.take(b.split_bytes);
  }
}

// ------------------------------------------------------------------------------------------------
// launch sequence
// ------------------------------------------------------------------------------------------------
struct Run {
  const fb_model_params& p;
  const ModelW& w;
  GraphDev g;
  Bufs b;
  cudaStream_t st;
  bool bf;           // bf16 activations (FB_PREC_BF16)
  int gmode = 0;     // precision mode of the GEMMs (FB_PREC_*); the split modes keep fp32 activations
  int H, N, Nc, Np;
  size_t TS;
  int rc = FB_OK;

  const void* W(int64_t off) const {
    if (gmode >= FB_PREC_SPLIT3) return (const void*)((const bf16*)p.w16 + 3 * off);
    return bf ? (const void*)((const bf16*)p.w16 + off) : (const void*)(p.w32 + off);
  }
  const float* F(int64_t off) const { return p.w32 + off; }
  void* at(void* base, size_t elem) const { return (char*)base + elem * TS; }
  const void* at(const void* base, size_t elem) const { return (const char*)base + elem * TS; }
  void chk(int r) { if (rc == FB_OK && r != FB_OK) rc = r; }
  // category-tagged launch of a non-GEMM stage
  // kid (diagnostic builds): kernel id for FB_KDUP=<bitmask> -- the masked kernels are launched TWICE (all of them are idempotent:
  // outputs never alias inputs), so the in-situ cost of one instance can be read off as a difference of step times with the results
  // intact.  0 gcl_edge_pre, 1 gcl_node, 2 row_attention, 3 pair_gather, 4 inter_logit + inter_aggregate is NOT idempotent (h += ...)
  // and has no id, 5 radial, 6 las_step
  template <typename F> void stage(int cat, F f, int kid = -1) {
    if (skip_mask() >> cat & 1) return;
    prof_begin(cat, st); chk(f());
    if (kid >= 0 && (dup_mask() >> kid & 1)) chk(f());
    prof_end(st);
  }
  int gemm_cat = CAT_GEMM_NODE;

  // ---- dropout sites of the FABind+ stack (one per nn.Dropout of the reference; tests/emulate_packed.py and the patched
  // reference of scripts/make_golden.py use the same numbering): site = (layer + 1) * 32 + local, layer = -1 for the stack's
  // own dropout, n_layers for out_layer; the iteration enters through the seed
  enum Site { S_EDGE1 = 0, S_EDGE2, S_GCOORD, S_NODE1, S_NODE2, S_PATT, S_CATT, S_PTR1, S_PTR2, S_CTR1, S_CTR2, S_PAIR1, S_PAIR2,
              S_AGG, S_ACOORD, S_IN, S_OUT };
  int cur_it = 0, cur_layer = -1;
  DropCfg dr(int local, int row0 = 0) const {
    return make_drop(p.dropout_p, p.dropout_seed + 0x632BE5ABu * (uint32_t)cur_it, (uint32_t)((cur_layer + 1) * 32 + local), row0,
                     p.dropout_colonly);
  }
  static GemmArgs wd(GemmArgs a, const DropCfg& d) { a.drop = d; return a; }

  // arguments of  C = act(A W^T + b)  with the usual optional extras
  GemmArgs mk(const void* A, int lda, int K, int64_t w_off, int Nout, int64_t b_off, int act, int M, float* C, int ldc,
              void* Cb, int ldcb, const float* res = nullptr, int ldres = 0, const void* A2 = nullptr, int lda2 = 0,
              int K2 = 0, int64_t dotv_off = -1, float* dot_out = nullptr, int dot_stride = 0, const int* m_dev = nullptr,
              int n_split = 0) {
    GemmArgs a;
    a.A = A; a.lda = lda; a.K1 = K; a.A2 = A2; a.lda2 = lda2; a.K2 = K2;
    a.W = W(w_off); a.bias = b_off >= 0 ? F(b_off) : nullptr; a.act = act;
    a.W_f32 = p.w32 + w_off; a.split_ws = b.split_ws; a.split_ws_bytes = b.split_bytes;
    a.w_static = true;       // the arena is written before the forward starts, never by its kernels
    a.res = res; a.ldres = ldres; a.C = C; a.ldc = ldc;
    a.Cb = bf ? Cb : nullptr; a.ldcb = ldcb;
    if (!bf && Cb != nullptr && (void*)C != Cb) {  // fp32 mode: the typed output IS the fp32 output
      if (C == nullptr) { a.C = (float*)Cb; a.ldc = ldcb; }
    }
    a.dotv = dotv_off >= 0 ? F(dotv_off) : nullptr; a.dot_out = dot_out; a.dot_stride = dot_stride;
    a.M = M; a.N = Nout; a.m_dev = m_dev; a.n_split = n_split;
    if (n_split > 0) { a.C = C; a.ldc = ldc; a.Cb = Cb; a.ldcb = ldcb; }   // both outputs are live in either precision
    return a;
  }
  void gemm(const GemmArgs& a) {
    if (a.M <= 0 || (skip_mask() >> gemm_cat & 1)) return;
    prof_begin(gemm_cat, st);
    prof_flops(gemm_cat, a);
    chk(gemm_launch(a, gmode, st));
    prof_end(st);
  }
  template <typename... Ts> void gemm(const void* A, Ts... ts) { gemm(mk(A, ts...)); }
  // compound-side + protein-side problems of one stage (disjoint row ranges of the same activation buffer)
  void gemm_pair(const GemmArgs& c, const GemmArgs& pr) {
    if (c.M <= 0 || pr.M <= 0 || (skip_mask() >> gemm_cat & 1)) { gemm(c); gemm(pr); return; }
    prof_begin(gemm_cat, st);
    prof_flops(gemm_cat, c); prof_flops(gemm_cat, pr);
    chk(gemm_launch_pair(c, pr, gmode, st));
    prof_end(st);
  }

  // independent problems of one stage (folded sequences): one multi-problem tcgen05 launch in bf16 mode, else one after the other
  void gemm_multi(const GemmArgs* g, int n) {
    if (skip_mask() >> gemm_cat & 1) return;
    prof_begin(gemm_cat, st);
    for (int i = 0; i < n; ++i) prof_flops(gemm_cat, g[i]);
    chk(gemm_launch_multi(g, n, gmode, /*prefetch_w=*/true, st));
    prof_end(st);
  }

  // Folded sequences (v1 layout, no dropout, SIMT attention core): node_mlp.2 -> first projections of the cross-attention block,
  // linear_o -> k/v of the other side -> transition.linear_1, transition.linear_2 -> q|k|v of the interfacial attention are exact
  // compositions of Linear maps; with the pre-multiplied weights of fb_derive_weights the ten dependent node-level GEMM launches of
  // a layer become six (run_gcl + run_att_folded).  Off: the launch sequence of round 1 (every Linear its own launch).
  bool fold_on() const {
#ifdef FB_DIAG
    static const bool on = [] { const char* e = getenv("FB_FOLD"); return !(e && atoi(e) == 0); }();
    if (!on) return false;
#endif
    return w.flavour == 0 && p.dropout_p <= 0.f && !xa_on() && p.n_layers > 0;
  }
  bool mv_ready = false;   // compact edge lists of the moving rows are filled (forward(): after graph_fill_ctx)
  bool ca_ready = false;   // the block's first projections (CAc, CAp) were produced by the preceding run_gcl

  // the attention core of both RowAttentionBlocks runs on tcgen05 (xatt_tc.cu) in bf16 mode when asked for (fb_model_params.attn_tc)
  // and the per-complex blocks fit its tiles (keys <= 256 per complex); otherwise (and in the fp32 / split-precision parity modes)
  // on the SIMT kernel, which is the faster of the two at these block sizes (measured: DESIGN.md section 5)
  bool xa_on() const {
    return bf && p.attn_tc && row_attention_tc_supported(p.max_p, p.max_c) && row_attention_tc_supported(p.max_c, p.max_p);
  }
  int att_p(const float* PBs) {
    if (xa_on())
      return row_attention_tc(g, 1, p.max_p, p.max_c, b.CApT, 2 * HD, 0, HD, Np, b.CAcT, 4 * HD, 0, HD, Nc, PBs, b.O, HD, st);
    const float* CApv = b.CAp - (size_t)Nc * 2 * HD;  // virtual base indexed by internal node id
    return row_attention(g, 1, p.max_p, p.max_c, CApv, 2 * HD, CApv + HD, 2 * HD, b.CAc, 4 * HD, b.CAc + HD, 4 * HD, PBs, b.O, HD, bf, st);
  }
  int att_c(const float* PBs) {
    if (xa_on())
      return row_attention_tc(g, 0, p.max_c, p.max_p, b.CAcT, 4 * HD, 2 * HD, 3 * HD, Nc, b.CAp2T, 2 * HD, 0, HD, Np, PBs, b.O, HD, st);
    const float* CAp2v = b.CAp2 - (size_t)Nc * 2 * HD;
    return row_attention(g, 0, p.max_c, p.max_p, b.CAc + 2 * HD, 4 * HD, b.CAc + 3 * HD, 4 * HD, CAp2v, 2 * HD, CAp2v + HD, 2 * HD, PBs, b.O, HD,
                         bf, st);
  }
  // projections of the block: fp32 outputs for the SIMT attention kernel, bf16 outputs (TMA-loadable operands) for the tcgen05 one
  GemmArgs proj(const void* A, int64_t w_off, int Nout, int64_t b_off, int M, float* C32, void* C16) {
    return xa_on() ? mk(A, H, H, w_off, Nout, b_off, FB_ACT_NONE, M, nullptr, 0, C16, Nout)
                   : mk(A, H, H, w_off, Nout, b_off, FB_ACT_NONE, M, C32, Nout, nullptr, 0);
  }

  // hin32 / hin16 / Pn_pre (folded path only): the layer's input h lives in another buffer (the iteration-invariant linear_in(H)) and
  // its per-node edge-MLP projections are already there; b.h / b.hT are then only written (through the second residual stream)
  void run_gcl(const GclW& gw, const float* x_in, float* x_out, bool need_h, const AttW* fold_att = nullptr, const float* hin32 = nullptr,
               const void* hin16 = nullptr, const void* Pn_pre = nullptr) {
    const float* h_in = hin32 ? hin32 : b.h;
    const void* hT_in = hin16 ? hin16 : b.hT;
    const void* Pn = Pn_pre ? Pn_pre : b.Pn;
    // coordinates only (out_layer of a non-final iteration: H is discarded and only the masked nodes keep their update,
    // att_model.py:232-236): the edge MLP runs on the compact list of context edges INTO the moving rows (graph_mv_*: ligand atoms
    // and the two global nodes, ~1/8 of the edges) and the coordinate step on those rows
    const bool mv_only = !need_h && mv_ready && p.dropout_p <= 0.f;
    const int E = mv_only ? p.E_ctx_mv : p.E_ctx;
    const int n_rows = mv_only ? g.n_mv : N;
    const int* erow = mv_only ? g.mv_erow : g.ctx_row;
    const int* ecol = mv_only ? g.mv_ecol : g.ctx_col;
    stage(CAT_GRAPH_MISC, [&] { return radial(g, g.ctx_rowptr, g.ctx_row, g.ctx_col, x_in, b.radc, b.normc, st); }, 5);
    gemm_cat = CAT_GEMM_NODE;
    // the first edge-MLP Linear per node: P[:, 0:H] = h W_row^T + b1, P[:, H:2H] = h W_col^T (bias [b1 | 0] = derived slot f_e1b)
    if (!Pn_pre) gemm(hT_in, H, H, gw.e1_rc, 2 * H, gw.f_e1b, FB_ACT_NONE, N, nullptr, 0, b.Pn, 2 * H);
    stage(CAT_EDGE_ELEMWISE, [&] {
      return gcl_edge_pre(E, H, erow, ecol, g.node_cplx, Pn, b.radc, b.normc, F(gw.e1_rad), nullptr, b.A1, bf, st,
                          mv_only ? g.mv_emap : nullptr);
    }, 0);
    // (profiling category = kernel class: the short edge lists of the moving-rows form run on the node-level kernels)
    gemm_cat = mv_only ? CAT_GEMM_NODE : CAT_GEMM_EDGE;
    // training-mode dropout of the v1 stack (dropout_p > 0): edge_mlp output (egnn.py:82), node_mlp output (egnn.py:106)
    gemm(wd(mk(b.A1, H, H, gw.e2_w, H, gw.e2_b, FB_ACT_SILU, E, nullptr, 0, b.M, H), dr(S_EDGE2)));
    const int tiles = gemm_dot_tiles(E, H, H, gmode);
    gemm(b.M, H, H, gw.c1_w, H, gw.c1_b, FB_ACT_SILU, E, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, 0, gw.c2_w, b.dotE, E);
    stage(CAT_EDGE_ELEMWISE, [&] {
      return gcl_node(n_rows, H, mv_only ? g.mv_rowptr : g.ctx_rowptr, ecol, b.M, b.dotE, tiles, E, x_in, p.coord_clamp,
                      need_h ? b.agg : nullptr, x_out, bf, st, mv_only ? g.mv_rows : nullptr, mv_only ? nullptr : &g);
    }, 1);
    gemm_cat = CAT_GEMM_NODE;
    if (need_h) {
      gemm(hT_in, H, H, gw.n1_w, H, gw.n1_b, FB_ACT_SILU, N, nullptr, 0, b.T1, H, nullptr, 0, b.agg, H, H);
      if (fold_att && fold_on()) {
        // node_mlp.2 + residual into the SECOND residual stream, and the cross-attention block's first projections of the new h
        // from [h | T1] in the same launch:  h' W^T = h W^T + T1 (W W_n2)^T + W b_n2
        const AttW& aw = *fold_att;
        const size_t op = (size_t)Nc * H;
        const GemmArgs g[3] = {
            mk(hT_in, H, H, aw.f_cac_w, 4 * HD, aw.f_cac_b, FB_ACT_NONE, Nc, b.CAc, 4 * HD, nullptr, 0, nullptr, 0, b.T1, H, H),
            mk(at(hT_in, op), H, H, aw.f_cap_w, 2 * HD, aw.f_cap_b, FB_ACT_NONE, Np, b.CAp, 2 * HD, nullptr, 0, nullptr, 0, at(b.T1, op), H, H),
            mk(b.T1, H, H, gw.n2_w, H, gw.n2_b, FB_ACT_NONE, N, b.h2, H, b.hT2, H, h_in, H)};
        gemm_multi(g, 3);
        std::swap(b.h, b.h2); std::swap(b.hT, b.hT2);
        ca_ready = true;
      } else {
        gemm(wd(mk(b.T1, H, H, gw.n2_w, H, gw.n2_b, FB_ACT_NONE, N, b.h, H, b.hT, H, b.h, H), dr(S_NODE2)));
      }
    }
  }

  // cross-attention block + stacked q|k|v projections of run_att as THREE multi-problem launches (see fold_on).  Two residual
  // streams: "o" = (b.h, b.hT), current on entry and on exit, "a" = (b.h2, b.hT2); a problem never writes rows of a stream that another
  // problem of the same launch reads as its A operand:
  //   L3 (after attention(p)):  TH_p  = relu([h_p|O_p] f_l3[2HD:]^T)        CAp2 = [h_p|O_p] f_l3[:2HD]^T        h_p: o -> a (linear_o + residual)
  //   L5 (after attention(c)):  h_p: a -> o (p transition linear_2 + residual)     TH_c = relu([h_c|O_c] f_l5^T)     h_c: o -> a
  //   L6:                       q|k|v(c) = [h_c(a)|TH_c] f_qkc^T        h_c: a -> o (c transition linear_2 + residual)        q|k|v(p) = h_p(o) qk^T
  void run_att_folded(const AttW& aw, int layer) {
    const size_t P = p.P_total;
    const size_t op = (size_t)Nc * H;
    float *h_o = b.h, *h_a = b.h2;
    void *hT_o = b.hT, *hT_a = b.hT2;
    void* Op = at(b.O, (size_t)Nc * HD);
    void* THp = at(b.TH, (size_t)Nc * 2 * H);
    gemm_cat = CAT_GEMM_NODE;
    if (!ca_ready)
      gemm_pair(proj(hT_o, aw.ca_c_w, 4 * HD, aw.ca_c_b, Nc, b.CAc, b.CAcT), proj(at(hT_o, op), aw.ca_p_w, 2 * HD, aw.ca_p_b, Np, b.CAp, b.CApT));
    ca_ready = false;
    stage(CAT_ATTENTION, [&] { return att_p(b.PB + (size_t)(layer * 2 + 0) * P * 4); }, 2);
    {
      const GemmArgs g[3] = {
          mk(at(hT_o, op), H, H, aw.f_l3_w + (int64_t)2 * HD * (H + HD), 2 * H, aw.f_l3_b + 2 * HD, FB_ACT_RELU, Np, nullptr, 0, THp, 2 * H,
             nullptr, 0, Op, HD, HD),
          mk(at(hT_o, op), H, H, aw.f_l3_w, 2 * HD, aw.f_l3_b, FB_ACT_NONE, Np, b.CAp2, 2 * HD, nullptr, 0, nullptr, 0, Op, HD, HD),
          mk(Op, HD, HD, aw.o_p_w, H, aw.o_p_b, FB_ACT_NONE, Np, h_a + op, H, at(hT_a, op), H, h_o + op, H)};
      gemm_multi(g, 3);
    }
    stage(CAT_ATTENTION, [&] { return att_c(b.PB + (size_t)(layer * 2 + 1) * P * 4); }, 2);
    {
      const GemmArgs g[3] = {
          mk(THp, 2 * H, 2 * H, aw.tp2_w, H, aw.tp2_b, FB_ACT_NONE, Np, h_o + op, H, at(hT_o, op), H, h_a + op, H),
          mk(hT_o, H, H, aw.f_l5_w, 2 * H, aw.f_l5_b, FB_ACT_RELU, Nc, nullptr, 0, b.TH, 2 * H, nullptr, 0, b.O, HD, HD),
          mk(b.O, HD, HD, aw.o_c_w, H, aw.o_c_b, FB_ACT_NONE, Nc, h_a, H, hT_a, H, h_o, H)};
      gemm_multi(g, 3);
    }
    {
      const int ldqk = 2 * H + QKX;
      const GemmArgs g[3] = {
          mk(hT_a, H, H, aw.f_qkc_w, ldqk + 2 * H, aw.f_qkc_b, FB_ACT_NONE, Nc, b.QK, ldqk, b.VT, 2 * H, nullptr, 0, b.TH, 2 * H, 2 * H, -1,
             nullptr, 0, nullptr, ldqk),
          mk(b.TH, 2 * H, 2 * H, aw.tc2_w, H, aw.tc2_b, FB_ACT_NONE, Nc, h_o, H, hT_o, H, h_a, H),
          mk(at(hT_o, op), H, H, aw.qk_w, ldqk + 2 * H, aw.qk_b, FB_ACT_NONE, Np, b.QK + (size_t)Nc * ldqk, ldqk, at(b.VT, (size_t)Nc * 2 * H),
             2 * H, nullptr, 0, nullptr, 0, 0, -1, nullptr, 0, nullptr, ldqk)};
      gemm_multi(g, 3);
    }
  }

  void run_att(const AttW& aw, int layer, const float* x_in, float* x_out, float* att = nullptr) {
    const size_t P = p.P_total;
    float* hp = b.h + (size_t)Nc * H;              // protein-side rows
    void* hTp = at(b.hT, (size_t)Nc * H);
    gemm_cat = CAT_GEMM_NODE;
    const int ldqk = 2 * H + QKX;
    if (fold_on()) {
      run_att_folded(aw, layer);
    } else {
    // --- cross attention (cross_att.py:24-54) on the per-complex blocks
    if (!ca_ready)
      gemm_pair(proj(b.hT, aw.ca_c_w, 4 * HD, aw.ca_c_b, Nc, b.CAc, b.CAcT), proj(hTp, aw.ca_p_w, 2 * HD, aw.ca_p_b, Np, b.CAp, b.CApT));
    ca_ready = false;
    stage(CAT_ATTENTION, [&] { return att_p(b.PB + (size_t)(layer * 2 + 0) * P * 4); });
    // RowAttentionBlock dropout on the attention output (cross_att.py:128), row index = internal node id
    gemm(wd(mk(at(b.O, (size_t)Nc * HD), HD, HD, aw.o_p_w, H, aw.o_p_b, FB_ACT_NONE, Np, hp, H, hTp, H, hp, H), dr(S_PATT, Nc)));
    gemm(proj(hTp, aw.ca_p2_w, 2 * HD, -1, Np, b.CAp2, b.CAp2T));
    stage(CAT_ATTENTION, [&] { return att_c(b.PB + (size_t)(layer * 2 + 1) * P * 4); });
    gemm(wd(mk(b.O, HD, HD, aw.o_c_w, H, aw.o_c_b, FB_ACT_NONE, Nc, b.h, H, b.hT, H, b.h, H), dr(S_CATT)));
    // transitions (model_utils.py:171-175), residual
    void* THp = at(b.TH, (size_t)Nc * 2 * H);
    gemm_pair(mk(b.hT, H, H, aw.tc1_w, 2 * H, aw.tc1_b, FB_ACT_RELU, Nc, nullptr, 0, b.TH, 2 * H),
              mk(hTp, H, H, aw.tp1_w, 2 * H, aw.tp1_b, FB_ACT_RELU, Np, nullptr, 0, THp, 2 * H));
    gemm_pair(mk(b.TH, 2 * H, 2 * H, aw.tc2_w, H, aw.tc2_b, FB_ACT_NONE, Nc, b.h, H, b.hT, H, b.h, H),
              mk(THp, 2 * H, 2 * H, aw.tp2_w, H, aw.tp2_b, FB_ACT_NONE, Np, hp, H, hTp, H, hp, H));
    // q | k of the interfacial attention stacked with the 32-channel interaction projections
    // (cross_att.py:22,51: linear_p on protein rows, linear_c on compound rows) in ONE node GEMM
    gemm(b.hT, H, H, aw.qk_w, ldqk + 2 * H, aw.qk_b, FB_ACT_NONE, N, b.QK, ldqk, b.VT, 2 * H, nullptr, 0, nullptr, 0, 0, -1,
         nullptr, 0, nullptr, ldqk);
    }
    // --- pair path on the unique inter pairs only
    const int capU = p.cap_int / 2;
    const int* u_dev = g.int_rowptr + Nc;  // number of compound->protein edges
    stage(CAT_ATTENTION, [&] { return pair_gather(g, capU, H, b.P0, b.QK + 2 * H, ldqk, b.Zg, b.T64, bf, st); }, 3);
    const int tiles2 = gemm_dot_tiles(capU, 2 * H, H, gmode);
    gemm_cat = CAT_GEMM_PAIR;
    gemm(b.Zg, H, H, aw.pt1_w, 2 * H, aw.pt1_b, FB_ACT_RELU, capU, nullptr, 0, nullptr, 0, nullptr, 0, b.T64, 64, 64,
         aw.pt2v, b.dotU, capU, u_dev);
    gemm_cat = CAT_GEMM_NODE;
    // FB_PB_FOLD=0 (A/B): separate pair_bias_finish launch + dense scatter instead of reading the row-dot partials in inter_logit
#ifdef FB_DIAG
    static const bool pb_fold = [] { const char* e = getenv("FB_PB_FOLD"); return !(e && atoi(e) == 0); }();
#else
    const bool pb_fold = true;
#endif
    if (!pb_fold) stage(CAT_ATTENTION, [&] { return pair_bias_finish(g, capU, b.dotU, tiles2, capU, F(aw.pt_c), b.pb_dense, st); });
    // --- interfacial attention (egnn.py:186-252)
    stage(CAT_GRAPH_MISC, [&] { return radial(g, g.int_rowptr, g.int_row, g.int_col, x_in, b.radi, b.normi, st); }, 5);
    stage(CAT_ATTENTION, [&] {
      return inter_attention(g, p.cap_int, H, b.QK, ldqk, b.QK + H, ldqk, b.VT, at(b.VT, (size_t)H), 2 * H, F(aw.k_r), F(aw.v_r), F(aw.ac_u), F(aw.ac1_b), F(aw.ac2_w), b.radi,
                             b.normi, b.pb_dense, x_in, p.coord_clamp, b.h, bf ? b.hT : nullptr, x_out, att, b.lgt, b.sde, bf, st,
                             nullptr, nullptr, nullptr, 1e-5f, DropCfg(), dr(S_AGG) /* egnn.py:236 */, pb_fold ? b.dotU : nullptr, tiles2, capU, F(aw.pt_c));
    });
  }

  // ---- FABind+ layout -------------------------------------------------------------------------------------------
  static constexpr float LN_EPS = 1e-5f;   // torch.nn.LayerNorm default (P/models/model_utils.py:15,37,60)

  // MC_E_GCL (P/models/egnn.py:44-115)
  void run_gcl_plus(const GclPW& gw, const float* x_in, float* x_out, bool need_h) {
    const int Dp = dp_of(H);
    // coordinates only (out_layer of a non-final iteration): the edges INTO the moving rows, as in run_gcl
    const bool mv_only = !need_h && mv_ready && p.dropout_p <= 0.f;
    const int E = mv_only ? p.E_ctx_mv : p.E_ctx;
    const int n_rows = mv_only ? g.n_mv : N;
    const int* erow = mv_only ? g.mv_erow : g.ctx_row;
    const int* ecol = mv_only ? g.mv_ecol : g.ctx_col;
    stage(CAT_GRAPH_MISC, [&] { return radial(g, g.ctx_rowptr, g.ctx_row, g.ctx_col, x_in, b.radc, b.normc, st); });
    // per-node sums of h and h^2: the LayerNorm statistics of [h_row | h_col | radial] are assembled per edge
    stage(CAT_EDGE_ELEMWISE, [&] { return row_stats(b.h, H, N, H, nullptr, b.hstat, false, st); });
    gemm_cat = CAT_GEMM_NODE;
    gemm(b.hT, H, H, gw.e1_rc, 2 * Dp, -1, FB_ACT_NONE, N, nullptr, 0, b.Pn, 2 * Dp);
    stage(CAT_EDGE_ELEMWISE, [&] {
      return gcl_edge_pre_plus(E, H, Dp, erow, ecol, g.node_cplx, b.Pn, b.hstat, b.radc, b.normc, F(gw.e1_rad), F(gw.e1_g),
                               F(gw.e1_c0), LN_EPS, b.A1, bf, st, dr(S_EDGE1), mv_only ? g.mv_emap : nullptr);
    });
    gemm_cat = CAT_GEMM_EDGE;
    gemm(wd(mk(b.A1, Dp, Dp, gw.e2_w, H, gw.e2_b, FB_ACT_RELU, E, nullptr, 0, b.M, H), dr(S_EDGE2)));
    // coord_mlp = MLPwoBias: LayerNorm on the edge message, Linear + ReLU, Linear(H,1) as the row-dot epilogue
    stage(CAT_EDGE_ELEMWISE, [&] { return ln_rows(b.M, true, H, H, nullptr, 0, 0, E, F(gw.cl_g), F(gw.cl_b), LN_EPS, b.M2, H, bf, st); });
    const int tiles = gemm_dot_tiles(E, H, H, gmode);
    gemm(wd(mk(b.M2, H, H, gw.c1_w, H, gw.c1_b, FB_ACT_RELU, E, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, 0, gw.c2_w, b.dotE, E),
            dr(S_GCOORD)));
    stage(CAT_EDGE_ELEMWISE, [&] {
      return gcl_node(n_rows, H, mv_only ? g.mv_rowptr : g.ctx_rowptr, ecol, b.M, b.dotE, tiles, E, x_in, p.coord_clamp,
                      need_h ? b.agg : nullptr, x_out, bf, st, mv_only ? g.mv_rows : nullptr, mv_only ? nullptr : &g);
    });
    gemm_cat = CAT_GEMM_NODE;
    if (need_h) {
      // node_mlp = MLPwithLastAct on [h | agg], residual
      stage(CAT_EDGE_ELEMWISE, [&] { return ln_rows(b.h, false, H, H, b.agg, H, H, N, F(gw.nl_g), F(gw.nl_b), LN_EPS, b.TH, 2 * H, bf, st); });
      gemm(wd(mk(b.TH, 2 * H, 2 * H, gw.n1_w, 2 * H, gw.n1_b, FB_ACT_RELU, N, nullptr, 0, b.TH2, 2 * H), dr(S_NODE1)));
      gemm(wd(mk(b.TH2, 2 * H, 2 * H, gw.n2_w, H, gw.n2_b, FB_ACT_RELU, N, b.h, H, b.hT, H, b.h, H), dr(S_NODE2)));
    }
  }

  // MC_Att_L (P/models/egnn.py:269-300) with CrossAttentionModule (P/models/cross_att.py:20-45); the pair embedding is
  // read from pair_in and the updated one written to pair_out (every pair row, carried to the next layer)
  void run_att_plus(const AttPW& aw, const void* pair_in, void* pair_out, const float* x_in, float* x_out, float* att = nullptr) {
    const size_t P = p.P_total;
    float* hp = b.h + (size_t)Nc * H;
    void* hTp = at(b.hT, (size_t)Nc * H);
    // gated pair biases of the two RowAttentionBlocks from the CURRENT pair embedding (cross_att.py:80-82)
    gemm_cat = CAT_GEMM_PAIR0;
    gemm(pair_in, H, H, aw.pb_w, 128, aw.pb_b, FB_ACT_NONE, (int)P, b.PBraw, 128, nullptr, 0);
    stage(CAT_EDGE_ELEMWISE, [&] { return pair_bias_gate((int)P, 1, b.PBraw, 128, b.PB, st); });
    gemm_cat = CAT_GEMM_NODE;
    gemm_pair(proj(b.hT, aw.ca_c_w, 4 * HD, aw.ca_c_b, Nc, b.CAc, b.CAcT), proj(hTp, aw.ca_p_w, 2 * HD, aw.ca_p_b, Np, b.CAp, b.CApT));
    stage(CAT_ATTENTION, [&] { return att_p(b.PB); });
    gemm(wd(mk(at(b.O, (size_t)Nc * HD), HD, HD, aw.o_p_w, H, aw.o_p_b, FB_ACT_NONE, Np, hp, H, hTp, H, hp, H), dr(S_PATT, Nc)));
    gemm(proj(hTp, aw.ca_p2_w, 2 * HD, -1, Np, b.CAp2, b.CAp2T));
    stage(CAT_ATTENTION, [&] { return att_c(b.PB + P * 4); });
    gemm(wd(mk(b.O, HD, HD, aw.o_c_w, H, aw.o_c_b, FB_ACT_NONE, Nc, b.h, H, b.hT, H, b.h, H), dr(S_CATT)));
    // transitions = MLPwithLastAct (LN, Linear+ReLU, Linear+ReLU), residual
    void* Tnp = at(b.Tn, (size_t)Nc * H);
    void* THp = at(b.TH, (size_t)Nc * H);
    stage(CAT_EDGE_ELEMWISE, [&] { return ln_rows(b.h, false, H, H, nullptr, 0, 0, Nc, F(aw.tcl_g), F(aw.tcl_b), LN_EPS, b.Tn, H, bf, st); });
    stage(CAT_EDGE_ELEMWISE, [&] { return ln_rows(hp, false, H, H, nullptr, 0, 0, Np, F(aw.tpl_g), F(aw.tpl_b), LN_EPS, Tnp, H, bf, st); });
    gemm_pair(wd(mk(b.Tn, H, H, aw.tc1_w, H, aw.tc1_b, FB_ACT_RELU, Nc, nullptr, 0, b.TH, H), dr(S_CTR1)),
              wd(mk(Tnp, H, H, aw.tp1_w, H, aw.tp1_b, FB_ACT_RELU, Np, nullptr, 0, THp, H), dr(S_PTR1, Nc)));
    gemm_pair(wd(mk(b.TH, H, H, aw.tc2_w, H, aw.tc2_b, FB_ACT_RELU, Nc, b.h, H, b.hT, H, b.h, H), dr(S_CTR2)),
              wd(mk(THp, H, H, aw.tp2_w, H, aw.tp2_b, FB_ACT_RELU, Np, hp, H, hTp, H, hp, H), dr(S_PTR2, Nc)));
    // q | k | inter32_p | inter32_c | pad  ||  v | vc   (vc = (coord_mlp.linear1 * gamma) applied to v, folded)
    const int ldqk = 2 * H + QKX;
    gemm(b.hT, H, H, aw.qk_w, ldqk + 2 * H, aw.qk_b, FB_ACT_NONE, N, b.QK, ldqk, b.VT, 2 * H, nullptr, 0, nullptr, 0, 0, -1,
         nullptr, 0, nullptr, ldqk);
    // pair <- MLPwithLastAct(pair + inter32(p, c)) on every pair row; attn_bias_proj of the result as the row-dot epilogue
    stage(CAT_ATTENTION, [&] {
      return pair_zin_plus(g, (int)P, H, pair_in, b.QK + 2 * H, ldqk, F(aw.zo_w), F(aw.zo_b), F(aw.zl_g), F(aw.zl_b), LN_EPS, b.Zl, bf, st);
    });
    gemm_cat = CAT_GEMM_PAIR;
    gemm(wd(mk(b.Zl, H, H, aw.pt1_w, H, aw.pt1_b, FB_ACT_RELU, (int)P, nullptr, 0, b.Zh, H), dr(S_PAIR1)));
    const int tilesP = gemm_dot_tiles((int)P, H, H, gmode);
    gemm(wd(mk(b.Zh, H, H, aw.pt2_w, H, aw.pt2_b, FB_ACT_RELU, (int)P, nullptr, 0, pair_out, H, nullptr, 0, nullptr, 0, 0, aw.wb, b.dotP,
               (int)P), dr(S_PAIR2)));
    gemm_cat = CAT_GEMM_NODE;
    stage(CAT_ATTENTION, [&] { return pair_bias_all((int)P, b.dotP, tilesP, (int)P, F(aw.pt_c), b.pb_dense, st); });
    // interfacial attention; coordinate head = MLPwoBias on v_e with its LayerNorm folded (per-node sums of v)
    stage(CAT_GRAPH_MISC, [&] { return radial(g, g.int_rowptr, g.int_row, g.int_col, x_in, b.radi, b.normi, st); });
    stage(CAT_ATTENTION, [&] { return row_stats(b.VT, 2 * H, N, H, F(aw.v_r), b.vstat, bf, st); });
    stage(CAT_ATTENTION, [&] {
      return inter_attention(g, p.cap_int, H, b.QK, ldqk, b.QK + H, ldqk, b.VT, at(b.VT, (size_t)H), 2 * H, F(aw.k_r), F(aw.v_r), F(aw.ac_u), F(aw.ac_c0),
                             F(aw.ac2_w), b.radi, b.normi, b.pb_dense, x_in, p.coord_clamp, b.h, bf ? b.hT : nullptr, x_out, att, b.lgt,
                             b.sde, bf, st, F(aw.ac_g), F(aw.ac_r), b.vstat, LN_EPS, dr(S_ACOORD), dr(S_AGG));
    });
  }

  void tap(int slot, const float* x) {
    if (p.trace_h) cudaMemcpyAsync(p.trace_h + (size_t)slot * N * H, b.h, sizeof(float) * (size_t)N * H, cudaMemcpyDeviceToDevice, st);
    if (p.trace_x) cudaMemcpyAsync(p.trace_x + (size_t)slot * N * 3, x, sizeof(float) * (size_t)N * 3, cudaMemcpyDeviceToDevice, st);
  }

  // one MCAttEGNN pass (or a subset of its steps) on caller-supplied graphs
  void forward_egnn(const fb_egnn_extra& e) {
    const size_t P = p.P_total;
    stage(CAT_GRAPH_MISC, [&] { return permute_in(g, p.H_in, p.X_in, p.X_las, H, b.Hin32, b.HinT, bf, b.x_state, b.xl, st); });
    auto cp = [&](void* dst, const void* src, size_t bytes) { if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st); };
    cp(g.ctx_rowptr, e.ctx_rowptr, sizeof(int) * (N + 1));
    cp(g.ctx_row, e.ctx_row, sizeof(int) * p.E_ctx); cp(g.ctx_col, e.ctx_col, sizeof(int) * p.E_ctx);
    cp(g.int_rowptr, e.int_rowptr, sizeof(int) * (N + 1));
    cp(g.int_row, e.int_row, sizeof(int) * e.E_int); cp(g.int_col, e.int_col, sizeof(int) * e.E_int);
    cp(g.int_pair, e.int_pair, sizeof(int) * e.E_int);
    gemm_cat = CAT_GEMM_NODE;
    if (e.steps & FB_STEP_LINEAR_IN) gemm(b.HinT, H, H, w.in_w, H, w.in_b, FB_ACT_NONE, N, b.h, H, b.hT, H);
    else chk(convert_copy(b.Hin32, (size_t)N * H, b.h, bf ? b.hT : nullptr, bf, st));
    const bool plus = w.flavour == 1;
    if ((e.steps & FB_STEP_ATT) && p.n_layers > 0) {
      chk(convert_copy(e.pair0, P * H, nullptr, b.P0, bf, st));
      if (!plus) {
        const int pbc = pb_cols(p.n_layers);
        gemm_cat = CAT_GEMM_PAIR0;
        gemm(b.P0, H, H, w.pb_w, pbc, w.pb_b, FB_ACT_NONE, (int)P, b.PBraw, pbc, nullptr, 0);
        stage(CAT_EDGE_ELEMWISE, [&] { return pair_bias_gate((int)P, p.n_layers, b.PBraw, pbc, b.PB, st); });
        gemm_cat = CAT_GEMM_NODE;
      }
    }
    const float* xc = b.x_state;
    float* bufs[2] = {b.xa, b.xb};
    int k = 0;
    const void* pair_cur = b.P0;   // FABind+: propagated layer to layer (P/models/egnn.py:380-392)
    for (int l = 0; l < p.n_layers; ++l) {
      cur_layer = l;
      if (e.steps & FB_STEP_GCL) {
        if (plus) run_gcl_plus(w.gclp[l], xc, bufs[k], true);
        else run_gcl(w.gcl[l], xc, bufs[k], true, (e.steps & FB_STEP_ATT) ? &w.att[l] : nullptr);
        xc = bufs[k]; k ^= 1;
      }
      if (e.steps & FB_STEP_ATT) {
        float* att_l = e.att_out ? e.att_out + (size_t)l * e.E_int : nullptr;
        if (plus) {
          void* pair_next = (l & 1) ? b.PairB : b.PairA;
          run_att_plus(w.attp[l], pair_cur, pair_next, xc, bufs[k], att_l);
          pair_cur = pair_next;
        } else {
          run_att(w.att[l], l, xc, bufs[k], att_l);
        }
        xc = bufs[k]; k ^= 1;
      }
      if (e.steps & FB_STEP_LAS) { chk(las_step(g, xc, b.xl, p.las_step, p.las_clamp, bufs[k], st)); xc = bufs[k]; k ^= 1; }
    }
    if (e.steps & FB_STEP_OUT_LAYER) {
      cur_layer = p.n_layers;
      if (plus) run_gcl_plus(w.gclp[p.n_layers], xc, bufs[k], true); else run_gcl(w.gcl[p.n_layers], xc, bufs[k], true);
      xc = bufs[k]; k ^= 1;
    }
    if (plus && (e.steps & FB_STEP_ATT) && p.pair_out)
      chk(pair_unpack(g, (int)P, H, p.max_p, p.max_c, pair_cur, p.pair_out, bf, st));
    if (e.steps & FB_STEP_LINEAR_OUT) {
      gemm(b.hT, H, H, w.out_w, H, w.out_b, FB_ACT_NONE, N, b.Hfin, H, nullptr, 0);
      chk(permute_out_h(g, b.Hfin, H, p.H_out, st));
    } else {
      chk(permute_out_h(g, b.h, H, p.H_out, st));
    }
    chk(permute_out_x(g, xc, p.X_out, st));
  }

  void forward() {
    const size_t P = p.P_total;
    stage(CAT_GRAPH_MISC, [&] { return permute_in(g, p.H_in, p.X_in, p.X_las, H, b.Hin32, b.HinT, bf, b.x_state, b.xl, st); });
    // pair_embed0 = InteractionModule(H_p, H_c)  (att_model.py:198-206, model_utils.py:219-222)
    gemm_cat = CAT_GEMM_NODE;
    gemm_pair(mk(b.HinT, H, H, w.il_c_w, H, w.il_c_b, FB_ACT_NONE, Nc, b.pc, H, nullptr, 0),
              mk(at(b.HinT, (size_t)Nc * H), H, H, w.il_p_w, H, w.il_p_b, FB_ACT_NONE, Np, b.pc + (size_t)Nc * H, H, nullptr, 0));
    stage(CAT_EDGE_ELEMWISE, [&] { return pair_outer(g, (int)P, H, b.pc, b.A0, bf, st); });
    gemm_cat = CAT_GEMM_PAIR0;
    gemm(b.A0, H, H, w.il_o_w, H, w.il_o_b, FB_ACT_NONE, (int)P, nullptr, 0, b.P0, H);
    // gated pair biases of all RowAttentionBlocks at once (pair0 is layer- and iteration-invariant in v1)
    const bool plus = w.flavour == 1;
    if (p.n_layers > 0 && !plus) {
      const int pbc = pb_cols(p.n_layers);
      gemm(b.P0, H, H, w.pb_w, pbc, w.pb_b, FB_ACT_NONE, (int)P, b.PBraw, pbc, nullptr, 0);
      stage(CAT_EDGE_ELEMWISE, [&] { return pair_bias_gate((int)P, p.n_layers, b.PBraw, pbc, b.PB, st); });
    }
    gemm_cat = CAT_GEMM_NODE;
    // context graph: protein coordinates are reset every iteration, so it is built once
    if (p.layout_flag) {
      // host-supplied counts (a dataloader-side layout): entries a wrong claim would leave unwritten must still be valid node ids
      if (p.E_ctx > 0) {
        chk(cudaMemsetAsync(g.ctx_row, 0, (size_t)p.E_ctx * sizeof(int), st) == cudaSuccess ? FB_OK : FB_ERR_CUDA);
        chk(cudaMemsetAsync(g.ctx_col, 0, (size_t)p.E_ctx * sizeof(int), st) == cudaSuccess ? FB_OK : FB_ERR_CUDA);
      }
      if (p.E_ctx_mv > 0) {
        chk(cudaMemsetAsync(g.mv_erow, 0, (size_t)p.E_ctx_mv * sizeof(int), st) == cudaSuccess ? FB_OK : FB_ERR_CUDA);
        chk(cudaMemsetAsync(g.mv_ecol, 0, (size_t)p.E_ctx_mv * sizeof(int), st) == cudaSuccess ? FB_OK : FB_ERR_CUDA);
        chk(cudaMemsetAsync(g.mv_emap, 0, (size_t)p.E_ctx_mv * sizeof(int), st) == cudaSuccess ? FB_OK : FB_ERR_CUDA);
      }
    }
    stage(CAT_GRAPH_MISC, [&] { return graph_fill_ctx(g, b.x_state, p.intra_cutoff, p.inter_cutoff, st); });
    if (g.n_mv > 0 && p.E_ctx_mv > 0 && p.E_ctx_mv < p.E_ctx && p.n_iter > 1) {
      stage(CAT_GRAPH_MISC, [&] { return graph_mv_fill(g, st); });
      mv_ready = true;
    }
    // iteration-invariant head (folded path): every iteration restarts from linear_in(H) (att_model.py:227-231 feeds the same H
    // each time), so it and the first layer's per-node projections of it are computed once per forward
    const bool hoist = fold_on() && p.n_iter > 1 && b.h0 != nullptr;
    if (hoist) {
      gemm_cat = CAT_GEMM_NODE;
      gemm(b.HinT, H, H, w.in_w, H, w.in_b, FB_ACT_NONE, N, b.h0, H, b.h0T, H);
      gemm(b.h0T, H, H, w.gcl[0].e1_rc, 2 * H, w.gcl[0].f_e1b, FB_ACT_NONE, N, nullptr, 0, b.Pn0, 2 * H);
    }
    for (int it = 0; it < p.n_iter; ++it) {
      const bool last = it == p.n_iter - 1;
      stage(CAT_GRAPH_MISC, [&] { return graph_build_inter(g, b.x_state, p.intra_cutoff, p.inter_cutoff, st, p.stats ? p.stats + it : nullptr); });
      gemm_cat = CAT_GEMM_NODE;
      cur_it = it; cur_layer = -1;
      if (!hoist) gemm(wd(mk(b.HinT, H, H, w.in_w, H, w.in_b, FB_ACT_NONE, N, b.h, H, b.hT, H), dr(S_IN)));
      const float* xc = b.x_state;
      float* bufs[2] = {b.xa, b.xb};
      int k = 0;
      const void* pair_cur = b.P0;   // FABind+: every iteration restarts from pair_embed0 (P/models/att_model.py:209-218)
      for (int l = 0; l < p.n_layers; ++l) {
        cur_layer = l;
        if (plus) run_gcl_plus(w.gclp[l], xc, bufs[k], true);
        else if (hoist && l == 0) run_gcl(w.gcl[l], xc, bufs[k], true, &w.att[l], b.h0, b.h0T, b.Pn0);
        else run_gcl(w.gcl[l], xc, bufs[k], true, &w.att[l]);
        xc = bufs[k]; k ^= 1;
        if (last) tap(2 * l, xc);
        if (plus) {
          void* pair_next = (l & 1) ? b.PairB : b.PairA;
          run_att_plus(w.attp[l], pair_cur, pair_next, xc, bufs[k]);
          pair_cur = pair_next;
        } else {
          run_att(w.att[l], l, xc, bufs[k]);
        }
        xc = bufs[k]; k ^= 1;
        if (last) tap(2 * l + 1, xc);
        stage(CAT_GRAPH_MISC, [&] { return las_step(g, xc, b.xl, p.las_step, p.las_clamp, bufs[k], st); }, 6);
        xc = bufs[k]; k ^= 1;
      }
      // the out-layer node update and linear_out only matter on the last iteration
      // (att_model.py:232: non-final iterations discard H)
      cur_layer = p.n_layers;
      if (plus) run_gcl_plus(w.gclp[p.n_layers], xc, bufs[k], last); else run_gcl(w.gcl[p.n_layers], xc, bufs[k], last);
      xc = bufs[k];
      if (last && plus && p.pair_out)
        stage(CAT_GRAPH_MISC, [&] { return pair_unpack(g, (int)P, H, p.max_p, p.max_c, pair_cur, p.pair_out, bf, st); });
      if (last) {
        gemm_cat = CAT_GEMM_NODE;
        cur_layer = -1;
        if (p.dropout_p > 0.f) stage(CAT_EDGE_ELEMWISE, [&] { return dropout_rows(b.h, b.hT, N, H, bf, dr(S_OUT), st); });
        gemm(b.hT, H, H, w.out_w, H, w.out_b, FB_ACT_NONE, N, b.Hfin, H, nullptr, 0);
        stage(CAT_GRAPH_MISC, [&] { return permute_out_h(g, b.Hfin, H, p.H_out, st); });
      }
      stage(CAT_GRAPH_MISC, [&] { return masked_update_x(g, b.x_state, xc, last ? p.X_out : nullptr, st); });
    }
  }
};

}  // namespace fb

using namespace fb;

static bool params_ok(const fb_model_params* p) {
  return p && p->N > 0 && p->B > 0 && p->hidden > 0 && (p->hidden % 8) == 0 && p->hidden <= 512 && p->n_layers >= 0 &&
         p->n_iter >= 1 && p->Nc_tot > 0 && p->Nc_tot <= p->N && (p->flavour == FB_FLAVOUR_V1 || p->flavour == FB_FLAVOUR_PLUS) &&
         p->dropout_p >= 0.f && p->dropout_p < 1.f && p->bf16_mode >= FB_PREC_FP32 && p->bf16_mode <= FB_PREC_SPLIT6;
}

extern "C" {

int32_t fb_abi_version(void) { return FB_ABI_VERSION; }

#ifndef FB_SOURCE_HASH
#define FB_SOURCE_HASH 0
#endif
int64_t fb_source_hash(void) { return (int64_t)FB_SOURCE_HASH; }

int64_t fb_launch_count(void) { return (int64_t)g_launches.load(); }

int32_t fb_prof_enable(int32_t on) {
  g_prof_on = on != 0;
  return FB_OK;
}

int32_t fb_prof_read(double* ms, int64_t* spans, int32_t n_cat) {
  // requires the recorded work to have completed (caller synchronises the stream first)
  for (int i = 0; i < n_cat; ++i) { ms[i] = 0.0; spans[i] = 0; }
  for (auto& s : g_spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.a, s.b) != cudaSuccess) return FB_ERR_CUDA;
    if (s.cat < n_cat) { ms[s.cat] += t; spans[s.cat] += 1; }
    g_pool.push_back(s);
  }
  g_spans.clear();
  return FB_OK;
}

int32_t fb_prof_flops(double* flops, int32_t n_cat) {
  for (int i = 0; i < n_cat; ++i) flops[i] = i < CAT_COUNT ? g_flops[i] : 0.0;
  for (int i = 0; i < CAT_COUNT; ++i) g_flops[i] = 0.0;
  return FB_OK;
}

int32_t fb_weight_slot_count(int32_t hidden, int32_t n_layers) { return (int32_t)weights_for(hidden, n_layers).slots.size(); }

int32_t fb_weight_slot_info(int32_t hidden, int32_t n_layers, int32_t i, char* name, int32_t name_cap, int64_t* rows,
                            int64_t* cols, int64_t* offset) {
  const ModelW& w = weights_for(hidden, n_layers);
  if (i < 0 || i >= (int)w.slots.size()) return FB_ERR_BAD_ARG;
  const Slot& s = w.slots[i];
  if (name && name_cap > 0) { std::strncpy(name, s.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (rows) *rows = s.rows;
  if (cols) *cols = s.cols;
  if (offset) *offset = s.off;
  return FB_OK;
}

int64_t fb_weight_arena_elems(int32_t hidden, int32_t n_layers) { return weights_for(hidden, n_layers).total; }

int32_t fb_weight_slot_count_f(int32_t hidden, int32_t n_layers, int32_t flavour) {
  return (int32_t)weights_for(hidden, n_layers, flavour).slots.size();
}

int32_t fb_weight_slot_info_f(int32_t hidden, int32_t n_layers, int32_t flavour, int32_t i, char* name, int32_t name_cap, int64_t* rows,
                              int64_t* cols, int64_t* offset) {
  const ModelW& w = weights_for(hidden, n_layers, flavour);
  if (i < 0 || i >= (int)w.slots.size()) return FB_ERR_BAD_ARG;
  const Slot& s = w.slots[i];
  if (name && name_cap > 0) { std::strncpy(name, s.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (rows) *rows = s.rows;
  if (cols) *cols = s.cols;
  if (offset) *offset = s.off;
  return FB_OK;
}

int64_t fb_weight_arena_elems_f(int32_t hidden, int32_t n_layers, int32_t flavour) { return weights_for(hidden, n_layers, flavour).total; }

int32_t fb_derive_weights(float* w32, int32_t hidden, int32_t n_layers, int32_t flavour, void* stream) {
  if (!w32 || hidden <= 0 || n_layers < 0 || (flavour != FB_FLAVOUR_V1 && flavour != FB_FLAVOUR_PLUS)) return FB_ERR_BAD_ARG;
  if (flavour != FB_FLAVOUR_V1) return FB_OK;           // the FABind+ sequences are not folded
  const ModelW& w = weights_for(hidden, n_layers, flavour);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = hidden;
  for (int l = 0; l <= n_layers; ++l) {
    const GclW& g = w.gcl[l];
    if (cudaMemcpyAsync(w32 + g.f_e1b, w32 + g.e1_b, sizeof(float) * H, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
        cudaMemsetAsync(w32 + g.f_e1b + H, 0, sizeof(float) * H, st) != cudaSuccess)
      return FB_ERR_CUDA;
  }
  for (int l = 0; l < n_layers; ++l) {
    const AttW& a = w.att[l];
    const GclW& g = w.gcl[l];
    int r = fold_weights(w32, a.ca_c_w, a.ca_c_b, 4 * HD, H, g.n2_w, g.n2_b, H, a.f_cac_w, a.f_cac_b, st);
    if (r == FB_OK) r = fold_weights(w32, a.ca_p_w, a.ca_p_b, 2 * HD, H, g.n2_w, g.n2_b, H, a.f_cap_w, a.f_cap_b, st);
    if (r == FB_OK) r = fold_weights(w32, a.ca_p2_w, -1, 2 * HD, H, a.o_p_w, a.o_p_b, HD, a.f_l3_w, a.f_l3_b, st);
    if (r == FB_OK) r = fold_weights(w32, a.tp1_w, a.tp1_b, 2 * H, H, a.o_p_w, a.o_p_b, HD, a.f_l3_w + (int64_t)2 * HD * (H + HD), a.f_l3_b + 2 * HD, st);
    if (r == FB_OK) r = fold_weights(w32, a.tc1_w, a.tc1_b, 2 * H, H, a.o_c_w, a.o_c_b, HD, a.f_l5_w, a.f_l5_b, st);
    if (r == FB_OK) r = fold_weights(w32, a.qk_w, a.qk_b, 4 * H + QKX, H, a.tc2_w, a.tc2_b, 2 * H, a.f_qkc_w, a.f_qkc_b, st);
    if (r != FB_OK) return r;
  }
  return FB_OK;
}

int64_t fb_graph_workspace_bytes(const fb_model_params* p) {
  if (!params_ok(p)) return FB_ERR_BAD_ARG;
  Arena a(nullptr, 0, true);
  GraphDev g;
  plan_graph(*p, a, g);
  return (int64_t)a.peak + 256;
}

int64_t fb_model_workspace_bytes(const fb_model_params* p) {
  if (!params_ok(p)) return FB_ERR_BAD_ARG;
  Arena a(nullptr, 0, true);
  Bufs b;
  plan_main(*p, a, b);
  return (int64_t)a.peak + 256;
}

int32_t fb_graph_static(const fb_model_params* p, void* stream) {
  if (!params_ok(p)) return FB_ERR_BAD_ARG;
  Arena a(p->ws_graph, p->ws_graph_bytes, false);
  GraphDev g;
  plan_graph(*p, a, g);
  if (!a.ok) return FB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int r = graph_prepare_static(g, (const long long*)p->bonds, (const long long*)p->las, st);
  if (r != FB_OK) return r;
  // the count pass needs coordinates in the internal order
  r = permute_x(g, p->X_in, g.xtmp, st);
  if (r != FB_OK) return r;
  r = graph_count_ctx(g, g.xtmp, p->intra_cutoff, p->inter_cutoff, st);
  if (r != FB_OK) return r;
  r = graph_mv_index(g, st);
  if (r != FB_OK || !p->layout_flag) return r;
  // ABI 6: the host supplied E_ctx / E_ctx_mv (no read-back): check the claim on the device
  return graph_verify_counts(g, p->E_ctx, p->E_ctx_mv, p->layout_flag, st);
}

const int32_t* fb_graph_counts_ptr(const fb_model_params* p) {
  Arena a(p->ws_graph, p->ws_graph_bytes, false);
  GraphDev g;
  plan_graph(*p, a, g);
  return g.counts;
}

const int32_t* fb_graph_ctx_count_ptr(const fb_model_params* p) {
  Arena a(p->ws_graph, p->ws_graph_bytes, false);
  GraphDev g;
  plan_graph(*p, a, g);
  return g.ctx_rowptr + p->N;
}

int32_t fb_model_forward(const fb_model_params* p, void* stream) {
  if (!params_ok(p) || p->E_ctx < 0 || (p->bf16_mode != FB_PREC_FP32 && !p->w16)) return FB_ERR_BAD_ARG;
  const ModelW& w = weights_for(p->hidden, p->n_layers, p->flavour);
  Run r{*p, w};
  Arena ag(p->ws_graph, p->ws_graph_bytes, false);
  plan_graph(*p, ag, r.g);
  Arena am(p->ws_main, p->ws_main_bytes, false);
  plan_main(*p, am, r.b);
  if (!ag.ok || !am.ok) return FB_ERR_WORKSPACE;
  r.g.ctx_row = r.b.ctx_row; r.g.ctx_col = r.b.ctx_col;
  r.g.int_row = r.b.int_row; r.g.int_col = r.b.int_col; r.g.int_pair = r.b.int_pair;
  r.g.mv_erow = r.b.mv_erow; r.g.mv_ecol = r.b.mv_ecol; r.g.mv_emap = r.b.mv_emap;
  r.g.ctx_cap = p->E_ctx; r.g.mv_cap = p->E_ctx_mv;     // the sizes plan_main gave the lists
  r.st = (cudaStream_t)stream;
  r.bf = p->bf16_mode == FB_PREC_BF16; r.gmode = p->bf16_mode;
  r.H = p->hidden; r.N = p->N; r.Nc = p->Nc_tot; r.Np = p->N - p->Nc_tot;
  r.TS = r.bf ? 2 : 4;
  r.forward();
  if (r.rc != FB_OK) return r.rc;
  return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_CUDA;
}

int32_t fb_egnn_forward(const fb_model_params* p, const fb_egnn_extra* e, void* stream) {
  if (!p || !e || p->N <= 0 || p->B <= 0 || p->hidden <= 0 || (p->hidden % 8) || p->hidden > 512 || p->E_ctx < 0) return FB_ERR_BAD_ARG;
  if ((e->steps & FB_STEP_ATT) && (!e->pair0 || p->Nc_tot <= 0 || p->Nc_tot >= p->N)) return FB_ERR_BAD_ARG;
  if (e->E_int > p->cap_int) return FB_ERR_BAD_ARG;
  if (p->flavour != FB_FLAVOUR_V1 && p->flavour != FB_FLAVOUR_PLUS) return FB_ERR_BAD_ARG;
  if (p->bf16_mode < FB_PREC_FP32 || p->bf16_mode > FB_PREC_SPLIT6 || (p->bf16_mode != FB_PREC_FP32 && !p->w16)) return FB_ERR_BAD_ARG;
  if (p->dropout_p != 0.f) return FB_ERR_UNSUPPORTED;           // sampling mode goes through fb_model_forward
  const ModelW& w = weights_for(p->hidden, p->n_layers, p->flavour);
  Run r{*p, w};
  Arena ag(p->ws_graph, p->ws_graph_bytes, false);
  plan_graph(*p, ag, r.g);
  Arena am(p->ws_main, p->ws_main_bytes, false);
  plan_main(*p, am, r.b);
  if (!ag.ok || !am.ok) return FB_ERR_WORKSPACE;
  r.g.ctx_row = r.b.ctx_row; r.g.ctx_col = r.b.ctx_col;
  r.g.int_row = r.b.int_row; r.g.int_col = r.b.int_col; r.g.int_pair = r.b.int_pair;
  r.g.mv_erow = r.b.mv_erow; r.g.mv_ecol = r.b.mv_ecol; r.g.mv_emap = r.b.mv_emap;
  r.st = (cudaStream_t)stream;
  r.bf = p->bf16_mode == FB_PREC_BF16; r.gmode = p->bf16_mode;
  r.H = p->hidden; r.N = p->N; r.Nc = p->Nc_tot; r.Np = p->N - p->Nc_tot;
  r.TS = r.bf ? 2 : 4;
  r.forward_egnn(*e);
  if (r.rc != FB_OK) return r.rc;
  return cudaGetLastError() == cudaSuccess ? FB_OK : FB_ERR_CUDA;
}

int32_t fb_edges_ref_count(int32_t N, const int32_t* cplx, const int32_t* off, const uint8_t* flags, const float* x,
                           float intra, float inter, int32_t* ws, void* stream) {
  int* deg = ws;
  int* rowptr = ws + (size_t)4 * N;
  int* fallback = rowptr + (size_t)4 * (N + 1);
  return graph_ref_count(N, cplx, off, flags, x, intra, inter, deg, rowptr, fallback, (cudaStream_t)stream);
}

int32_t fb_edges_ref_fill(int32_t N, const int32_t* cplx, const int32_t* off, const uint8_t* flags, const float* x,
                          float intra, float inter, int32_t* ws, const int32_t* counts_host, int64_t* ctx_out,
                          int64_t* inter_out, void* stream) {
  int* deg = ws;
  int* rowptr = ws + (size_t)4 * N;
  int* fallback = rowptr + (size_t)4 * (N + 1);
  int* cat_base = fallback + 1;
  const int host_base[4] = {0, counts_host[0], counts_host[0] + counts_host[1], 0};
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemcpyAsync(cat_base, host_base, sizeof(host_base), cudaMemcpyHostToDevice, st);
  const int e_ctx = counts_host[0] + counts_host[1] + counts_host[2];
  return graph_ref_fill(N, cplx, off, flags, x, intra, inter, deg, rowptr, cat_base, fallback, counts_host[4],
                        (long long*)ctx_out, e_ctx, (long long*)inter_out, counts_host[3], st);
}

static GemmArgs gemm_args_from(const fb_gemm_params* q) {
  GemmArgs a;
  a.A = q->A; a.lda = q->lda; a.K1 = q->K1; a.A2 = q->A2; a.lda2 = q->lda2; a.K2 = q->K2; a.W = q->W;
  a.bias = q->bias; a.act = q->act; a.res = q->res; a.ldres = q->ldres; a.C = q->C; a.ldc = q->ldc;
  a.Cb = q->Cb; a.ldcb = q->ldcb; a.dotv = q->dotv; a.dot_out = q->dot_out; a.dot_stride = q->dot_stride;
  a.M = q->M; a.N = q->N; a.m_dev = q->m_dev; a.n_split = q->n_split; a.ldw = q->ldw;
  a.drop = make_drop(q->drop_p, q->drop_seed, q->drop_site, q->drop_row0, q->drop_colonly);
  a.split_ws = q->split_ws; a.split_ws_bytes = q->split_ws_bytes; a.W_f32 = q->W_f32;
  return a;
}
static bool prec_ok(int m) { return m >= FB_PREC_FP32 && m <= FB_PREC_SPLIT6; }

int32_t fb_gemm(const fb_gemm_params* q, void* stream) {
  if (!q) return FB_ERR_BAD_ARG;
  const GemmArgs a = gemm_args_from(q);
  if (!prec_ok(q->bf16_mode)) return FB_ERR_BAD_ARG;
  if (q->force_simt) {
    if (q->bf16_mode >= FB_PREC_SPLIT3) return FB_ERR_BAD_ARG;
    return gemm_simt_launch(a, q->bf16_mode == FB_PREC_BF16, (cudaStream_t)stream);
  }
  return gemm_launch(a, q->bf16_mode, (cudaStream_t)stream);
}

int32_t fb_gemm_pair(const fb_gemm_params* q0, const fb_gemm_params* q1, void* stream) {
  if (!q0 || !q1 || q0->bf16_mode != q1->bf16_mode || !prec_ok(q0->bf16_mode)) return FB_ERR_BAD_ARG;
  return gemm_launch_pair(gemm_args_from(q0), gemm_args_from(q1), q0->bf16_mode, (cudaStream_t)stream);
}

int32_t fb_gemm_multi(const fb_gemm_params* q, int32_t n, int32_t prefetch_w, void* stream) {
  if (!q || n < 1 || n > 4) return FB_ERR_BAD_ARG;
  GemmArgs a[4];
  for (int i = 0; i < n; ++i) {
    if (q[i].bf16_mode != q[0].bf16_mode || !prec_ok(q[i].bf16_mode) || q[i].force_simt) return FB_ERR_BAD_ARG;
    a[i] = gemm_args_from(&q[i]);
  }
  return gemm_launch_multi(a, n, q[0].bf16_mode, prefetch_w != 0, (cudaStream_t)stream);
}

int32_t fb_gemm_set_debug(int64_t* dbg) {
  fb::g_tc_dbg = (long long*)dbg;
  return FB_OK;
}

int32_t fb_gemm_dot_tiles(int32_t M, int32_t N, int32_t K, int32_t bf16_mode, int32_t force_simt) {
  if (force_simt) return gemm_simt_dot_tiles(N);
  return gemm_dot_tiles(M, N, K, bf16_mode);
}

int32_t fb_row_attention_tc(const int32_t* c_off, const int32_t* p_off, const int32_t* pair_base, int32_t B, int32_t Nc_tot,
                            int32_t q_is_prot, int32_t max_q, int32_t max_k, const void* QG, int32_t ldqg, int32_t qcol, int32_t gcol,
                            int32_t q_rows, const void* KV, int32_t ldkv, int32_t kcol, int32_t vcol, int32_t k_rows, const float* PB,
                            void* O, int32_t ldo, void* stream) {
  if (!c_off || !p_off || !pair_base || B <= 0 || !QG || !KV || !PB || !O) return FB_ERR_BAD_ARG;
  GraphDev g{};
  g.B = B; g.Nc_tot = Nc_tot; g.c_off = c_off; g.p_off = p_off; g.pair_base = pair_base;
  return row_attention_tc(g, q_is_prot, max_q, max_k, QG, ldqg, qcol, gcol, q_rows, KV, ldkv, kcol, vcol, k_rows, PB, O, ldo,
                          (cudaStream_t)stream);
}

int32_t fb_split_rows(const float* src, int32_t ld, int32_t M, int32_t K, void* dst, void* stream) {
  return split_rows(src, ld, K, nullptr, 0, 0, M, dst, (cudaStream_t)stream);
}

}  // extern "C"
