// InstanceNorm2d(affine=False, eps=1e-5) + activation + Dropout, forward and backward, NHWC.
// HBM-bound: every kernel streams 16-byte vectors, one (pixel, 8-channel group) per thread-iteration,
// consecutive threads on consecutive channel groups of the same pixel (fully coalesced).
// Reference: unet.py:20-28,55-66 and disc.py:32,42 (aten::instance_norm, activations, native_dropout).
#include "common.cuh"

namespace pg {

constexpr float IN_EPS = 1e-5f;
constexpr int NT = 256;

__device__ __forceinline__ unsigned long long mix_seed(unsigned long long s, unsigned long long salt) {
  return s ^ (salt * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull);
}

// is_f32 is a PgDType: PG_BF16 (0), PG_F32 (1) or PG_F16 (2)
__device__ __forceinline__ void load8(const void* base, int is_f32, long long off, float* f) {
  if (is_f32 == PG_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    float4 a = p[0], b = p[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + off);
    unpack8dt(v, is_f32, f);
  }
}
__device__ __forceinline__ void store8(void* base, int is_f32, long long off, const float* f) {
  if (is_f32 == PG_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
    p[0] = make_float4(f[0], f[1], f[2], f[3]);
    p[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + off) = pack8dt(f, is_f32);
  }
}

// mean / rstd of channel c of image b from the (sum, sumsq) pairs
__device__ __forceinline__ void mean_rstd(const float* sums, int b, int C, int c, long long HW, float& mean,
                                          float& rstd) {
  const double s = sums[((long long)b * C + c) * 2], ss = sums[((long long)b * C + c) * 2 + 1];
  const double m = s / (double)HW;
  double var = ss / (double)HW - m * m;
  if (var < 0) var = 0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)IN_EPS));
}

struct Span {  // how one block walks its image
  int cg, pl, my_cg, my_pl;
  bool active;
  long long pbeg, pend;
};
__device__ __forceinline__ Span make_span(int C, long long HW, long long ppb) {
  Span s;
  s.cg = C >> 3;
  s.pl = NT / s.cg;
  s.my_cg = threadIdx.x % s.cg;
  s.my_pl = threadIdx.x / s.cg;
  s.active = s.my_pl < s.pl;
  s.pbeg = (long long)blockIdx.x * ppb;
  s.pend = s.pbeg + ppb;
  if (s.pend > HW) s.pend = HW;
  return s;
}

// ------------------------------------------------------------------ statistics
__global__ void __launch_bounds__(NT) instnorm_stats_kernel(const void* x, int x_f32, long long HW, int C, int ld,
                                                            float* sums, long long ppb) {
  extern __shared__ float sh[];  // [2*C]
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += NT) sh[i] = 0.f;
  __syncthreads();
  Span s = make_span(C, HW, ppb);
  if (s.active) {
    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
    for (long long p = s.pbeg + s.my_pl; p < s.pend; p += s.pl) {
      float f[8];
      load8(x, x_f32, ((long long)b * HW + p) * ld + s.my_cg * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += f[j];
        a2[j] = fmaf(f[j], f[j], a2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[(s.my_cg * 8 + j) * 2], a1[j]);
      atomicAdd(&sh[(s.my_cg * 8 + j) * 2 + 1], a2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += NT) atomicAdd(&sums[(long long)b * C * 2 + i], sh[i]);
}

// ------------------------------------------------------------------ forward apply
__global__ void __launch_bounds__(NT) norm_act_fwd_kernel(const void* x, int x_f32, const float* sums, void* y,
                                                          int y_f32, void* y2, long long HW, int C, int ldx, int ldy, int act,
                                                          float drop_p, const unsigned long long* seed_ptr, unsigned long long salt,
                                                          long long ppb) {
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  Span s = make_span(C, HW, ppb);
  if (!s.active) return;
  float mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (sums != nullptr) mean_rstd(sums, b, C, s.my_cg * 8 + j, HW, mean[j], rstd[j]);
    else { mean[j] = 0.f; rstd[j] = 1.f; }
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long p = s.pbeg + s.my_pl; p < s.pend; p += s.pl) {
    float f[8];
    const long long pix = (long long)b * HW + p;
    load8(x, x_f32, pix * ldx + s.my_cg * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = act_apply(act, (f[j] - mean[j]) * rstd[j]);
      if (drop_p > 0.f) {
        const float u = uniform01(seed, (unsigned long long)(pix * C + s.my_cg * 8 + j));
        v = u >= drop_p ? v * keep_scale : 0.f;
      }
      f[j] = v;
    }
    store8(y, y_f32, pix * ldy + s.my_cg * 8, f);
    if (y2 != nullptr) store8(y2, PG_BF16, pix * ldy + s.my_cg * 8, f);
  }
}

// ------------------------------------------------------------------ backward
// dxhat for 8 channels of one pixel
__device__ __forceinline__ void dxhat8(const void* x, int x_f32, const void* dy1, int ld1, const void* dy2, int ld2,
                                       long long pix, int C, int c0, int ldx, int act, float drop_p,
                                       unsigned long long seed, const float* mean, const float* rstd, float* xhat,
                                       float* dxh) {
  float f[8], g[8];
  load8(x, x_f32, pix * ldx + c0, f);
  load8(dy1, 0, pix * ld1 + c0, g);
  if (dy2 != nullptr) {
    float g2[8];
    load8(dy2, 0, pix * ld2 + c0, g2);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += g2[j];
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xh = (f[j] - mean[j]) * rstd[j];
    float d = g[j] * act_grad_from_input(act, xh);
    if (drop_p > 0.f) {
      const float u = uniform01(seed, (unsigned long long)(pix * C + c0 + j));
      d = u >= drop_p ? d * keep_scale : 0.f;
    }
    xhat[j] = xh;
    dxh[j] = d;
  }
}

__global__ void __launch_bounds__(NT) norm_act_bwd_reduce_kernel(const void* x, int x_f32, const float* sums,
                                                                 const void* dy1, int ld1, const void* dy2, int ld2,
                                                                 float* bsums, long long HW, int C, int ldx, int act,
                                                                 float drop_p, const unsigned long long* seed_ptr,
                                                                 unsigned long long salt, long long ppb) {
  extern __shared__ float sh[];
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  for (int i = threadIdx.x; i < 2 * C; i += NT) sh[i] = 0.f;
  __syncthreads();
  Span s = make_span(C, HW, ppb);
  if (s.active) {
    float mean[8], rstd[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean_rstd(sums, b, C, s.my_cg * 8 + j, HW, mean[j], rstd[j]);
      a1[j] = a2[j] = 0.f;
    }
    for (long long p = s.pbeg + s.my_pl; p < s.pend; p += s.pl) {
      float xh[8], d[8];
      dxhat8(x, x_f32, dy1, ld1, dy2, ld2, (long long)b * HW + p, C, s.my_cg * 8, ldx, act, drop_p, seed, mean, rstd,
             xh, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += d[j];
        a2[j] = fmaf(d[j], xh[j], a2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[(s.my_cg * 8 + j) * 2], a1[j]);
      atomicAdd(&sh[(s.my_cg * 8 + j) * 2 + 1], a2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += NT) atomicAdd(&bsums[(long long)b * C * 2 + i], sh[i]);
}

__global__ void __launch_bounds__(NT) norm_act_bwd_apply_kernel(const void* x, int x_f32, const float* sums,
                                                                const void* dy1, int ld1, const void* dy2, int ld2,
                                                                const float* bsums, void* dx, int lddx, long long HW,
                                                                int C, int ldx, int act, float drop_p,
                                                                const unsigned long long* seed_ptr,
                                                                unsigned long long salt, long long ppb) {
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  Span s = make_span(C, HW, ppb);
  if (!s.active) return;
  float mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (sums != nullptr) {
      const int c = s.my_cg * 8 + j;
      mean_rstd(sums, b, C, c, HW, mean[j], rstd[j]);
      m1[j] = bsums[((long long)b * C + c) * 2] / (float)HW;
      m2[j] = bsums[((long long)b * C + c) * 2 + 1] / (float)HW;
    } else {
      mean[j] = 0.f; rstd[j] = 1.f; m1[j] = 0.f; m2[j] = 0.f;
    }
  }
  for (long long p = s.pbeg + s.my_pl; p < s.pend; p += s.pl) {
    float xh[8], d[8];
    const long long pix = (long long)b * HW + p;
    dxhat8(x, x_f32, dy1, ld1, dy2, ld2, pix, C, s.my_cg * 8, ldx, act, drop_p, seed, mean, rstd, xh, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = rstd[j] * (d[j] - m1[j] - xh[j] * m2[j]);
    store8(dx, 0, pix * lddx + s.my_cg * 8, d);
  }
}

__global__ void __launch_bounds__(NT) act_bwd_from_output_kernel(const void* y, int y_f32, int ldy, const void* dy,
                                                                 int lddy, void* dx, int lddx, long long npix, int C,
                                                                 int act) {
  const int cg = C >> 3;
  const long long total = npix * cg;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long pix = i / cg;
    const int c0 = (int)(i % cg) * 8;
    float f[8], g[8];
    load8(y, y_f32, pix * ldy + c0, f);
    load8(dy, 0, pix * lddy + c0, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= act_grad_from_output(act, f[j]);
    store8(dx, 0, pix * lddx + c0, g);
  }
}

__global__ void softmax_fwd_kernel(const float* x, float* y, long long npix, int C, int ld) {
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const float* xp = x + pix * ld;
    float* yp = y + pix * ld;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, xp[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(xp[c] - mx);
    const float inv = 1.f / s;
    for (int c = 0; c < C; ++c) yp[c] = expf(xp[c] - mx) * inv;
    for (int c = C; c < ld; ++c) yp[c] = 0.f;
  }
}

static void span_grid(int B, long long HW, int C, dim3& grid, long long& ppb) {
  const int pl = NT / (C >> 3);
  long long per_img = (4LL * num_sms() + B - 1) / B;
  if (per_img < 1) per_img = 1;
  ppb = (HW + per_img - 1) / per_img;
  if (ppb < 4LL * pl) ppb = 4LL * pl;
  const long long nb = (HW + ppb - 1) / ppb;
  grid = dim3((unsigned)nb, (unsigned)B, 1);
}

}  // namespace pg

using namespace pg;

static int check_c(const char* who, int C) {
  PG_REQUIRE(C >= 8 && (C % 8) == 0 && C <= 2048, "%s: C=%d must be a multiple of 8 in [8,2048]", who, C);
  return PG_OK;
}

extern "C" int pg_instnorm_stats(const void* x, int32_t x_f32, int32_t B, int64_t HW, int32_t C, int32_t ld,
                                 float* sums, void* stream) {
  if (int e = check_c("pg_instnorm_stats", C)) return e;
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb);
  instnorm_stats_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, x_f32, HW, C, ld, sums, ppb);
  return check_launch("instnorm_stats_kernel");
}

extern "C" int pg_norm_act_fwd(const void* x, int32_t x_f32, const float* sums, void* y, int32_t y_f32, void* y2, int32_t B,
                               int64_t HW, int32_t C, int32_t ldx, int32_t ldy, int32_t act, float drop_p,
                               const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_act_fwd", C)) return e;
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb);
  norm_act_fwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(x, x_f32, sums, y, y_f32, y2, HW, C, ldx, ldy, act, drop_p,
                                                            (const unsigned long long*)seed, salt, ppb);
  return check_launch("norm_act_fwd_kernel");
}

extern "C" int pg_norm_act_bwd_reduce(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                                      const void* dy2, int32_t ld2, float* bsums, int32_t B, int64_t HW, int32_t C,
                                      int32_t ldx, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt,
                                      void* stream) {
  if (int e = check_c("pg_norm_act_bwd_reduce", C)) return e;
  PG_REQUIRE(sums != nullptr, "pg_norm_act_bwd_reduce: sums is NULL");
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb);
  norm_act_bwd_reduce_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb);
  return check_launch("norm_act_bwd_reduce_kernel");
}

extern "C" int pg_norm_act_bwd_apply(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                                     const void* dy2, int32_t ld2, const float* bsums, void* dx, int32_t lddx,
                                     int32_t B, int64_t HW, int32_t C, int32_t ldx, int32_t act, float drop_p,
                                     const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_act_bwd_apply", C)) return e;
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb);
  norm_act_bwd_apply_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, dx,
                                                                  lddx, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb);
  return check_launch("norm_act_bwd_apply_kernel");
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }

extern "C" int pg_counter_add(uint64_t* ctr, uint64_t inc, void* stream) {
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)ctr, inc);
  return check_launch("counter_add_kernel");
}

extern "C" int pg_act_bwd_from_output(const void* y, int32_t y_f32, int32_t ldy, const void* dy, int32_t lddy,
                                      void* dx, int32_t lddx, int64_t npix, int32_t C, int32_t act, void* stream) {
  if (int e = check_c("pg_act_bwd_from_output", C)) return e;
  const long long total = npix * (C >> 3);
  long long blocks = (total + NT - 1) / NT;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  act_bwd_from_output_kernel<<<(unsigned)blocks, NT, 0, (cudaStream_t)stream>>>(y, y_f32, ldy, dy, lddy, dx, lddx,
                                                                               npix, C, act);
  return check_launch("act_bwd_from_output_kernel");
}

extern "C" int pg_softmax_fwd(const float* x, float* y, int64_t npix, int32_t C, int32_t ld, void* stream) {
  PG_REQUIRE(C >= 1 && C <= ld, "pg_softmax_fwd: bad C=%d ld=%d", C, ld);
  long long blocks = (npix + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  softmax_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, npix, C, ld);
  return check_launch("softmax_fwd_kernel");
}
