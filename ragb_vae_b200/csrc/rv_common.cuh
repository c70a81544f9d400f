// Shared helpers for librgbavae: error reporting, dtype load/store, warp reductions, launch
// accounting.  Internal to the library; the public surface is include/rgbavae.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/rgbavae.h"

namespace rv {

void set_error(const char* fmt, ...);

// Launch accounting (rv_launch_count) and the optional per-category CUDA-event profiler
// (rv_prof_begin / rv_prof_end).  Every kernel launch site wraps itself in a LaunchScope.
enum Category {
  CAT_CONV_TC = 0,
  CAT_CONV_DIRECT = 1,
  CAT_NORM = 2,
  CAT_SOFTMAX = 3,
  CAT_LAYOUT = 4,
  CAT_REPARAM = 5,
  CAT_LOSS = 6,
  CAT_PSNR = 7,
  CAT_ATTN = 8,
  CAT_CONV_UPS = 9,  // phase-folded up-sampling convs: booked at the algorithmic 3x3 cost, executing 4/9 of it
  CAT_COUNT = RV_PROF_CATEGORIES
};

struct LaunchScope {
  int cat;
  cudaStream_t stream;
  cudaEvent_t e0, e1;
  bool timed;
  LaunchScope(int cat, cudaStream_t stream, double work = 0.0);
  ~LaunchScope();
};

#define RV_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      rv::set_error(__VA_ARGS__);        \
      return 2;                          \
    }                                    \
  } while (0)

#define RV_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      rv::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define RV_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      rv::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));  \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// 16-byte vector of T: 4 floats or 8 bf16.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ float get(int i) const { return reinterpret_cast<const float*>(&v)[i]; }
  __device__ __forceinline__ void set(int i, float x) { reinterpret_cast<float*>(&v)[i] = x; }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ float get(int i) const {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&v)[i]);
  }
  __device__ __forceinline__ void set(int i, float x) {
    reinterpret_cast<__nv_bfloat16*>(&v)[i] = __float2bfloat16_rn(x);
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
};

// One 32-byte (whole DRAM sector) store of two 16-byte chunks; p must be 32-byte aligned.  A thread that owns a pixel row and
// writes it as separate 16-byte stores leaves half-written sectors behind each instruction, which measurably slows the
// write stream (conv epilogues: +15 % on the 96-channel layers from this change alone).
__device__ __forceinline__ void store32(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x * sigmoid(x).  The fp32 parity mode uses expf and an IEEE divide (within 1e-6 of torch's silu);
// bf16 outputs use the MUFU exp2 / reciprocal approximations (~2^-21 relative, far below bf16's 2^-9):
// the accurate form costs ~40 instructions per element, which makes a streaming kernel ALU-bound
// (measured 2.8 TB/s instead of HBM speed).
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
// silu(x) = x * sigmoid(x) with sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5, i.e. with h = x / 2:  silu = h * tanh(h) + h.
// ONE MUFU op (tanh.approx, 2^-11 relative) and two FMA-pipe ops per element (mul, fma) instead of ex2 + rcp and four -- the
// SFU runs 16 lanes per clock per SM, and at HBM speed a norm kernel has ~10 instruction slots per element in total.
// silu_from_half takes h directly: callers that scale by a constant anyway fold the 1/2 into it.
__device__ __forceinline__ float silu_from_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float silu_fast(float x) { return silu_from_half(0.5f * x); }
template <typename T> __device__ __forceinline__ float silu_t(float x);
template <> __device__ __forceinline__ float silu_t<float>(float x) { return silu(x); }
template <> __device__ __forceinline__ float silu_t<__nv_bfloat16>(float x) { return silu_fast(x); }

int num_sms();

}  // namespace rv
