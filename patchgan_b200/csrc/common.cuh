// Shared helpers for the patchgan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/patchgan_b200.h"

namespace pg {

typedef __nv_bfloat16 bf16;

// thread-local error text behind pg_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define PG_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      pg::set_error(__VA_ARGS__);             \
      return PG_ERR_INVALID;                  \
    }                                         \
  } while (0)

#define PG_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      cudaGetLastError(); /* clear the sticky-until-read error */                  \
      pg::set_error("%s failed: %s", #call, cudaGetErrorString(e__));              \
      return PG_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

__device__ __forceinline__ float act_apply(int act, float x) {
  switch (act) {
    case PG_ACT_RELU: return fmaxf(x, 0.f);
    case PG_ACT_LEAKYRELU: return x > 0.f ? x : 0.2f * x;
    case PG_ACT_TANH: return tanhf(x);
    case PG_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

// derivative expressed through the pre-activation x
__device__ __forceinline__ float act_grad_from_input(int act, float x) {
  switch (act) {
    case PG_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case PG_ACT_LEAKYRELU: return x > 0.f ? 1.f : 0.2f;
    case PG_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case PG_ACT_SIGMOID: { float s = 1.f / (1.f + __expf(-x)); return s * (1.f - s); }
    default: return 1.f;
  }
}

// derivative expressed through the output y = act(x)
__device__ __forceinline__ float act_grad_from_output(int act, float y) {
  switch (act) {
    case PG_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case PG_ACT_LEAKYRELU: return y > 0.f ? 1.f : 0.2f;
    case PG_ACT_TANH: return 1.f - y * y;
    case PG_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 x 16-bit (bf16 or fp16, selected by PgDType) <-> 8 float through one 16-byte vector
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// 8 bf16 <-> 8 float through one 16-byte vector
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

__device__ __forceinline__ void unpack8dt(const uint4& v, int dt, float* f) {
  if (dt == PG_F16) unpack8h(v, f); else unpack8(v, f);
}
__device__ __forceinline__ uint4 pack8dt(const float* f, int dt) { return dt == PG_F16 ? pack8h(f) : pack8(f); }
// i -> (x = i % W, y = (i / W) % H, b = i / (W * H)).  32-bit divisions when the index allows (a 64-bit division is ~100
// instructions, more than the rest of a pixel-per-thread kernel).
__device__ __forceinline__ void split_xyb(long long i, int W, int H, int& x, int& y, int& b) {
  if (i < 0x7fffffffLL) {
    const unsigned u = (unsigned)i, r = u / (unsigned)W;
    x = (int)(u - r * (unsigned)W);
    b = (int)(r / (unsigned)H);
    y = (int)(r - (unsigned)b * (unsigned)H);
  } else {
    const long long r = i / W;
    x = (int)(i - r * W);
    b = (int)(r / H);
    y = (int)(r - (long long)b * H);
  }
}

__device__ __forceinline__ unsigned short to16(float x, int dt) {
  if (dt == PG_F16) return __half_as_ushort(__float2half_rn(x));
  return __bfloat16_as_ushort(__float2bfloat16(x));
}
__device__ __forceinline__ float from16(unsigned short u, int dt) {
  if (dt == PG_F16) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}

// Counter-based RNG for dropout: one 32-bit draw per element index, reproducible in backward
// from (seed, index) alone -- no mask tensor is ever stored.  (splitmix64 finaliser)
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

// per-layer dropout stream: the device seed counter mixed with the layer's salt
__device__ __forceinline__ unsigned long long mix_seed(unsigned long long s, unsigned long long salt) {
  return s ^ (salt * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pg
