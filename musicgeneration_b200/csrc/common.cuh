// Shared device/host helpers for libmt_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mt_b200.h"

namespace mt {

// ---------------------------------------------------------------------------------------
// error plumbing: thread-local message, never throw across the ABI
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> code

#define MT_REQUIRE(cond, ...)                   \
  do {                                          \
    if (!(cond)) {                              \
      ::mt::set_error(__VA_ARGS__);             \
      return MT_E_ARG;                          \
    }                                           \
  } while (0)

// Programmatic dependent launch (decode step: ~46 dependent, few-microsecond kernels per event).
// A kernel launched through launch_chain() while chaining is on may start before its predecessor in
// the stream has finished; chain_prologue() at its very top lets ITS successor start launching and
// then blocks until the predecessor grid has completed and flushed.  Both are no-ops for a plain
// launch, so the same kernels serve the training path unchanged.
bool chain_enabled();
__device__ __forceinline__ void chain_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = chain_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline int dtype_size(int dt) { return dt == MT_F32 ? 4 : 2; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();
// cudaFuncSetAttribute is per DEVICE: launchers remember the devices they have configured as a bit mask
// (bit = device ordinal), not as one process-wide flag
inline unsigned long long attr_dev_bit() {
  int dev = 0;
  cudaGetDevice(&dev);
  return 1ull << (dev & 63);
}

// ---------------------------------------------------------------------------------------
// dtype conversion
// ---------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 4 consecutive elements <-> float4
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&r.x);
  __half2 b = *reinterpret_cast<__half2*>(&r.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (dropout masks are recomputed in backward from the same counters)
// counter = (elem4_lo, elem4_hi, site, 0), key = seed.  One call -> 4 uniforms for 4
// consecutive elements.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// keep-mask (x scale) for elements [4*e4, 4*e4+4): returns multipliers (0 or 1/(1-p))
__device__ __forceinline__ float4 dropout_mult4(uint64_t seed, uint64_t site, uint64_t e4, float p,
                                                float inv_keep) {
  uint4 r = philox4x32_10(make_uint4((uint32_t)e4, (uint32_t)(e4 >> 32), (uint32_t)site, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 2.3283064365386963e-10f;  // 2^-32
  float4 m;
  m.x = (r.x * s >= p) ? inv_keep : 0.f;
  m.y = (r.y * s >= p) ? inv_keep : 0.f;
  m.z = (r.z * s >= p) ? inv_keep : 0.f;
  m.w = (r.w * s >= p) ? inv_keep : 0.f;
  return m;
}

}  // namespace mt

// dtype dispatch helpers (host)
#define MT_DISPATCH_DTYPE(dt, T, ...)                         \
  switch (dt) {                                               \
    case MT_F32: { using T = float; __VA_ARGS__; } break;     \
    case MT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    case MT_F16: { using T = __half; __VA_ARGS__; } break;    \
    default: ::mt::set_error("bad dtype %d", dt); return MT_E_ARG; \
  }

#define MT_DISPATCH_F32_BF16(dt, T, ...)                      \
  switch (dt) {                                               \
    case MT_F32: { using T = float; __VA_ARGS__; } break;     \
    case MT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
    default: ::mt::set_error("dtype %d not supported here (f32/bf16 only)", dt); return MT_E_UNSUPPORTED; \
  }
