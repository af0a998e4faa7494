// Shared helpers for the dsfuse sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dsfuse.h"

namespace dsf {

// thread-local error string, set by fail(); read by dsf_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DSF_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ::dsf::set_error(__VA_ARGS__);        \
      return DSF_EINVAL;                    \
    }                                       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();
// channel tile of the NCHW token / upsample kernels: 8 for planes of >= 1024 pixels (stages 1-2: 4x the CTAs), else 32
inline int nchw_channel_tile(const dsf_geom* g) { return (g->H * g->W >= 1024 && g->C % 8 == 0) ? 8 : 32; }
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context) setting: a kernel's launcher keeps one
// "configured" flag per device, so a process that drives several GPUs (nn.DataParallel threads, cuda:1 without
// set_device(0)) configures the kernel on each of them.  Flags are only ever set (benign if two threads race).
inline bool& per_device_flag(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return flags[dev & 63];
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The hot kernels of the bf16 path are launched with cudaLaunchAttributeProgrammaticStreamSerialization: every kernel
// executes pdl_trigger() first (its successor may start occupying SMs as soon as ALL of this grid's CTAs are
// resident or done) and pdl_wait() before its first access to global memory (returns once the predecessor grid has
// completed and flushed).  The successor's launch latency, barrier/TMEM set-up and tensor-map prefetch thereby overlap
// the predecessor's tail.  Without the launch attribute both instructions are no-ops.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in check_launch()
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- scalar load/store by dtype
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store (16 B for float, 8 B for bf16); pointer must be aligned accordingly
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

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

// ---------------------------------------------------------------- counter-based dropout (Philox4x32, 7 rounds)
// nn.Dropout sites of the path (model2_seq.py:104,109,125,272).  The mask of element e of a site is a pure function of
// (seed, site, step, e): 16-bit lane e%8 of philox(key = seed, counter = (e/8, site, step)); keep iff lane >= t16 with
// t16 = round(p * 65536), kept values are scaled by 65536 / (65536 - t16) (= 1/(1-p) to 1.5e-5).  One Philox call
// decides 8 elements.  Forward and backward kernels recompute it, nothing is stored (attention excepted).
// Philox4x32 with 7 rounds (the Random123 "Crush-resistant" minimum; 10 is the library default with safety margin):
// the masks only have to be uncorrelated, and the generator runs inside GEMM epilogues and the softmax loop.
struct DropArgs {
  uint32_t thresh;  // 16-bit threshold: drop iff lane < thresh  (0 = dropout disabled)
  float scale;      // 65536 / (65536 - thresh)
  uint32_t seed_lo, seed_hi, site, step;
  const unsigned long long* seed_dev;  // nullable device word XOR-ed into the seed (graph-safe reseeding)
};
__host__ __device__ inline DropArgs make_drop(const dsf_dropout* d) {
  DropArgs a{0u, 1.0f, 0u, 0u, 0u, 0u, nullptr};
  if (d && d->p > 0.f) {
    long t = lrint((double)d->p * 65536.0);
    t = t < 1 ? 1 : (t > 65535 ? 65535 : t);
    a.thresh = (uint32_t)t;
    a.scale = 65536.0f / (65536.0f - (float)t);
    a.seed_lo = (uint32_t)(d->seed & 0xFFFFFFFFull);
    a.seed_hi = (uint32_t)(d->seed >> 32);
    a.site = d->site;
    a.step = d->step;
    a.seed_dev = reinterpret_cast<const unsigned long long*>(d->seed_dev);
  }
  return a;
}
// fold the device-resident seed word in (once per thread, after griddepcontrol.wait: an earlier kernel of the stream bumps it)
__device__ __forceinline__ DropArgs resolve_drop(DropArgs a) {
  if (a.thresh != 0u && a.seed_dev != nullptr) {
    const unsigned long long s = __ldg(a.seed_dev);
    a.seed_lo ^= (uint32_t)(s & 0xFFFFFFFFull);
    a.seed_hi ^= (uint32_t)(s >> 32);
  }
  return a;
}
constexpr int PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// scale factors (0 or scale) of the 8 consecutive elements e8*8 .. e8*8+7
__device__ __forceinline__ void drop_scale8(const DropArgs& a, uint64_t e8, float (&m)[8]) {
  const uint4 r = philox4x32(a.seed_lo, a.seed_hi, (uint32_t)e8, (uint32_t)(e8 >> 32), a.site, a.step);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    m[2 * k] = (w[k] & 0xFFFFu) >= a.thresh ? a.scale : 0.f;
    m[2 * k + 1] = (w[k] >> 16) >= a.thresh ? a.scale : 0.f;
  }
}
// the 4 consecutive elements e4*4 .. e4*4+3 (half of one 8-element group)
__device__ __forceinline__ void drop_scale4(const DropArgs& a, uint64_t e4, float (&m)[4]) {
  float m8[8];
  drop_scale8(a, e4 >> 1, m8);
  const int h = (int)(e4 & 1) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = m8[h + k];
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace dsf
