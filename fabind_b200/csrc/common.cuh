// Shared device helpers for the fabind_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define FB_OK 0
#define FB_ERR_BAD_ARG (-1)
#define FB_ERR_WORKSPACE (-2)
#define FB_ERR_CUDA (-3)
#define FB_ERR_UNSUPPORTED (-4)

#define FB_ACT_NONE 0
#define FB_ACT_SILU 1
#define FB_ACT_RELU 2

#define FB_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return FB_ERR_CUDA;             \
  } while (0)

typedef __nv_bfloat16 bf16;

namespace fb {

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerialization and starts with
// pdl_entry(): it lets the NEXT kernel in the stream begin scheduling its CTAs (launch latency and
// prologue overlap this kernel's tail) and then waits until the PREVIOUS grid has completed and flushed.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }

bool pdl_enabled();   // forward.cu (FB_PDL=0 disables the launch attribute)

// ---- dropout (FABind+ sampling mode: the reference runs the whole model in train() mode, P/test_sampling_fabind.py:118-124) ----
// Counter-based masks: keep(seed, site, row, col) is a pure function, so every kernel that touches an activation can evaluate
// it in place (GEMM epilogues, edge kernels) and a CPU restatement can reproduce it bit for bit (tests/emulate_packed.py).
//   site = which nn.Dropout of the reference (layer / sub-module, see forward.cu::Site), row = internal row id of the
//   activation (node / context edge / interface edge / pair row), col = feature column.  colonly = 1 ignores the row: a
//   mask that is invariant to row order, used to pin the PLACEMENT of every mask against the unmodified reference.
struct DropCfg {
  float p = 0.f, scale = 1.f;      // scale = 1 / (1 - p)
  uint32_t thresh = 0;             // keep iff hash >= thresh,  thresh = p * 2^32
  uint32_t seed = 0, site = 0;
  int row0 = 0;                    // added to the kernel-local row index
  int colonly = 0;
};
__host__ __device__ __forceinline__ uint32_t fb_drop_hash(uint32_t seed, uint32_t site, uint32_t row, uint32_t col) {
  uint32_t x = seed ^ (site * 0x9E3779B1u);
  x ^= row * 0x85EBCA77u;
  x = ((x << 13) | (x >> 19)) * 0xC2B2AE3Du;
  x ^= col * 0x27D4EB2Fu;
  x ^= x >> 15; x *= 0x2C1B3C6Du;
  x ^= x >> 12; x *= 0x297A2D39u;
  x ^= x >> 15;
  return x;
}
__device__ __forceinline__ float drop_apply(float x, const DropCfg& d, int row, int col) {
  const uint32_t h = fb_drop_hash(d.seed, d.site, d.colonly ? 0u : (uint32_t)(row + d.row0), (uint32_t)col);
  return h >= d.thresh ? x * d.scale : 0.f;
}
inline DropCfg make_drop(float p, uint32_t seed, uint32_t site, int row0, int colonly) {
  DropCfg d;
  if (p > 0.f) {
    d.p = p; d.scale = 1.0f / (1.0f - p);
    const double t = (double)p * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
    d.seed = seed; d.site = site; d.row0 = row0; d.colonly = colonly;
  }
  return d;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember which devices have it for a kernel
template <typename K>
inline bool ensure_smem_optin(K kernel, int bytes, unsigned long long& done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev < 64 && (done_mask >> dev) & 1ull) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
  if (dev < 64) done_mask |= 1ull << dev;
  return true;
}

template <typename... KArgs, typename... Args>
inline void fb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// host-side bookkeeping (forward.cu): number of kernels launched, optional per-category event timing
void count_launch(int n);
void prof_begin(int category, cudaStream_t st);
void prof_end(cudaStream_t st);
enum { CAT_GEMM_EDGE = 0, CAT_GEMM_NODE, CAT_GEMM_PAIR, CAT_GEMM_PAIR0, CAT_EDGE_ELEMWISE, CAT_ATTENTION, CAT_GRAPH_MISC, CAT_COUNT };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// SiLU as torch computes it: x / (1 + exp(-x))  (full-precision expf, no fast-math)
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }
// bf16-mode SiLU: x * sigmoid(x) with sigmoid(x) = 0.5 + 0.5 tanh(x/2) -- ONE MUFU op (tanh.approx, rel. error ~2^-11,
// below the bf16 rounding of the result) instead of ex2 + rcp; the SFU pipe (16 ops/clk/SM) is what bounds the
// SiLU epilogues and the edge pre-activation kernel
__device__ __forceinline__ float silu_fast(float x) {
  float t;
  const float hx = 0.5f * x;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
  return fmaf(hx, t, hx);
}

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == FB_ACT_SILU) return silu(x);
  if (ACT == FB_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}
__device__ __forceinline__ float apply_act_rt(float x, int act) {
  if (act == FB_ACT_SILU) return silu(x);
  if (act == FB_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

// element load/store with conversion (T = float or bf16)
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-wide vector access on 4-aligned element offsets
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb2 = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb2.x, fb2.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 8-wide vector access on 8-aligned element offsets (one 16-byte transaction in bf16)
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
  u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
  *reinterpret_cast<uint4*>(p) = u;
}


// per-complex radial norm from its RAD_SLICES partial sums of d^4 (layers.cu::radial_kernel)
constexpr int RAD_SLICES = 8;
__device__ __forceinline__ float radial_norm(const float* __restrict__ part, int b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < RAD_SLICES; ++k) s += part[b * RAD_SLICES + k];
  return sqrtf(s);
}

}  // namespace fb
