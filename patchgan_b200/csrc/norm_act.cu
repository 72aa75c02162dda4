// InstanceNorm2d(affine=False, eps=1e-5) + activation + Dropout, forward and backward, NHWC.
// HBM-bound: every kernel streams 16-byte vectors, one (pixel, 8-channel group) per thread-iteration,
// consecutive threads on consecutive channel groups of the same pixel (fully coalesced).
// Reference: unet.py:20-28,55-66 and disc.py:32,42 (aten::instance_norm, activations, native_dropout).
#include "common.cuh"

namespace pg {

constexpr float IN_EPS = 1e-5f;
constexpr int NT = 256;

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

// Per-block (mean, rstd) table in shared memory: C threads do the fp64 finish once instead of every thread 8 times.
__device__ __forceinline__ void load_stats(const float* sums, int b, int C, long long HW, float* sh_mean, float* sh_rstd) {
  for (int c = threadIdx.x; c < C; c += NT) {
    if (sums != nullptr) mean_rstd(sums, b, C, c, HW, sh_mean[c], sh_rstd[c]);
    else { sh_mean[c] = 0.f; sh_rstd[c] = 1.f; }
  }
  __syncthreads();
}

// Block-level reduction of per-thread (a1[8], a2[8]) channel partials into global pairs out[(b*C + c)*2 + {0,1}].
// Threads that share a channel group sit cg lanes apart, so for power-of-two cg <= 32 a few shuffles leave one partial
// per (warp, channel); the 8 warp partials are then summed by C threads.  (The first version let all 256 threads
// atomicAdd into 2*C shared words: a 64-way same-address conflict that cost more than the HBM traffic.)
__device__ __forceinline__ void reduce_channels(float* a1, float* a2, const Span& s, int C, float* sh, float* out, int b) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool pow2 = (s.cg & (s.cg - 1)) == 0 && s.cg <= 32;
  if (pow2) {
    if (!s.active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
    }
    for (int off = s.cg; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], off);
        a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], off);
      }
    }
    if (lane < s.cg) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sh[(warp * C + lane * 8 + j) * 2] = a1[j];
        sh[(warp * C + lane * 8 + j) * 2 + 1] = a2[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += NT) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) t += sh[w * C * 2 + i];
      atomicAdd(&out[(long long)b * C * 2 + i], t);
    }
  } else {
    for (int i = threadIdx.x; i < 2 * C; i += NT) sh[i] = 0.f;
    __syncthreads();
    if (s.active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sh[(s.my_cg * 8 + j) * 2], a1[j]);
        atomicAdd(&sh[(s.my_cg * 8 + j) * 2 + 1], a2[j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += NT) atomicAdd(&out[(long long)b * C * 2 + i], sh[i]);
  }
}
// shared floats needed by reduce_channels
static inline size_t reduce_smem(int C) {
  const int cg = C >> 3;
  const bool pow2 = (cg & (cg - 1)) == 0 && cg <= 32;
  return (size_t)(pow2 ? (NT / 32) * C * 2 : 2 * C) * sizeof(float);
}

// Optional per-channel affine transform between the normalisation and the activation (BatchNorm2d's weight / bias,
// unet.py:20,55 with norm_layer = nn.BatchNorm2d): z = gamma * xhat + beta.  gamma == nullptr: identity (InstanceNorm2d,
// affine=False).  Channels >= c_real (zero padding) use (1, 0).
struct Affine {
  const float* gamma;
  const float* beta;
  int c_real;
};
__device__ __forceinline__ void load_affine(const Affine& af, int c0, float* ga, float* be) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool on = af.gamma != nullptr && c0 + j < af.c_real;
    ga[j] = on ? af.gamma[c0 + j] : 1.f;
    be[j] = on ? af.beta[c0 + j] : 0.f;
  }
}

constexpr int UNR = 4;   // independent 16/32-byte loads in flight per thread

// ------------------------------------------------------------------ statistics
__global__ void __launch_bounds__(NT) instnorm_stats_kernel(const void* x, int x_f32, long long HW, int C, int ld,
                                                            float* sums, long long ppb) {
  extern __shared__ float sh[];
  const int b = blockIdx.y;
  Span s = make_span(C, HW, ppb);
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (s.active) {
    const long long base = (long long)b * HW;
    long long p = s.pbeg + s.my_pl;
    for (; p + (UNR - 1) * s.pl < s.pend; p += UNR * s.pl) {
      float f[UNR][8];
#pragma unroll
      for (int u = 0; u < UNR; ++u) load8(x, x_f32, (base + p + u * s.pl) * ld + s.my_cg * 8, f[u]);
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += f[u][j];
          a2[j] = fmaf(f[u][j], f[u][j], a2[j]);
        }
    }
    for (; p < s.pend; p += s.pl) {
      float f[8];
      load8(x, x_f32, (base + p) * ld + s.my_cg * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += f[j];
        a2[j] = fmaf(f[j], f[j], a2[j]);
      }
    }
  }
  reduce_channels(a1, a2, s, C, sh, sums, b);
}

// ------------------------------------------------------------------ forward apply
__global__ void __launch_bounds__(NT) norm_act_fwd_kernel(const void* x, int x_f32, const float* sums, void* y,
                                                          int y_f32, void* y2, long long HW, int C, int ldx, int ldy, int act,
                                                          float drop_p, const unsigned long long* seed_ptr, unsigned long long salt,
                                                          long long ppb, Affine af) {
  extern __shared__ float sh[];
  float* sh_mean = sh;
  float* sh_rstd = sh + C;
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  const Span s = make_span(C, HW, ppb);
  const long long base = (long long)b * HW;
  // the first batch of loads does not depend on the statistics: issue it before the (mean, rstd) prologue
  float f[UNR][8];
  long long p0 = s.pbeg + s.my_pl;
  if (s.active) {
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p0 + u * s.pl < s.pend) load8(x, x_f32, (base + p0 + u * s.pl) * ldx + s.my_cg * 8, f[u]);
  }
  load_stats(sums, b, C, HW, sh_mean, sh_rstd);
  if (!s.active) return;
  float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mean[j] = sh_mean[s.my_cg * 8 + j]; rstd[j] = sh_rstd[s.my_cg * 8 + j]; }
  load_affine(af, s.my_cg * 8, ga, be);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  while (p0 < s.pend) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long p = p0 + u * s.pl;
      if (p >= s.pend) break;
      const long long pix = base + p;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = act_apply(act, fmaf((f[u][j] - mean[j]) * rstd[j], ga[j], be[j]));
        if (drop_p > 0.f) {
          const float r = uniform01(seed, (unsigned long long)(pix * C + s.my_cg * 8 + j));
          v = r >= drop_p ? v * keep_scale : 0.f;
        }
        f[u][j] = v;
      }
      store8(y, y_f32, pix * ldy + s.my_cg * 8, f[u]);
      if (y2 != nullptr) store8(y2, PG_BF16, pix * ldy + s.my_cg * 8, f[u]);
    }
    p0 += UNR * s.pl;
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (p0 + u * s.pl < s.pend) load8(x, x_f32, (base + p0 + u * s.pl) * ldx + s.my_cg * 8, f[u]);
  }
}

// ------------------------------------------------------------------ backward
// dxhat for 8 channels of one pixel, from already loaded x / dy values
// xk (PG_X_* >> 8): what the saved tensor x holds -- 0: the pre-norm conv output, 1: xhat itself, 2: the block's OUTPUT
// y = act(xhat) with an invertible activation (LeakyReLU(0.2) / none; the fused forward kernel stores nothing else)
__device__ __forceinline__ void dxhat8(const float* f, float* g, long long pix, int C, int c0, int act, float drop_p,
                                       unsigned long long seed, const float* mean, const float* rstd, float* xhat, int xk,
                                       const float* ga = nullptr, const float* be = nullptr) {
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xh = xk == 0 ? (f[j] - mean[j]) * rstd[j]
                             : (xk == 2 && act == PG_ACT_LEAKYRELU && f[j] < 0.f ? 5.f * f[j] : f[j]);
    // (with an affine transform the activation sees z = gamma * xhat + beta; g stays d/dz, gamma is applied by the caller)
    float d = g[j] * act_grad_from_input(act, ga != nullptr ? fmaf(xh, ga[j], be[j]) : xh);
    if (drop_p > 0.f) {
      const float u = uniform01(seed, (unsigned long long)(pix * C + c0 + j));
      d = u >= drop_p ? d * keep_scale : 0.f;
    }
    xhat[j] = xh;
    g[j] = d;
  }
}
__device__ __forceinline__ void load_dy(const void* dy1, int ld1, const void* dy2, int ld2, long long pix, int c0, float* g) {
  load8(dy1, 0, pix * ld1 + c0, g);
  if (dy2 != nullptr) {
    float g2[8];
    load8(dy2, 0, pix * ld2 + c0, g2);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += g2[j];
  }
}

constexpr int UNB = 2;

__global__ void __launch_bounds__(NT) norm_act_bwd_reduce_kernel(const void* x, int x_f32, const float* sums,
                                                                 const void* dy1, int ld1, const void* dy2, int ld2,
                                                                 float* bsums, long long HW, int C, int ldx, int act,
                                                                 float drop_p, const unsigned long long* seed_ptr,
                                                                 unsigned long long salt, long long ppb, Affine af) {
  const int xk = x_f32 >> 8;
  x_f32 &= 0xff;
  extern __shared__ float sh[];
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  const Span s = make_span(C, HW, ppb);
  const long long base = (long long)b * HW;
  const int c0 = s.my_cg * 8;
  float ga[8], be[8];
  load_affine(af, c0, ga, be);
  // first batch of x / dy loads in flight across the (mean, rstd) prologue
  float f[UNB][8], g[UNB][8];
  long long p0 = s.pbeg + s.my_pl;
  if (s.active) {
#pragma unroll
    for (int u = 0; u < UNB; ++u)
      if (p0 + u * s.pl < s.pend) {
        load8(x, x_f32, (base + p0 + u * s.pl) * ldx + c0, f[u]);
        load_dy(dy1, ld1, dy2, ld2, base + p0 + u * s.pl, c0, g[u]);
      }
  }
  load_stats(sums, b, C, HW, sh, sh + C);
  float a1[8], a2[8], mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = a2[j] = 0.f;
    mean[j] = s.active ? sh[c0 + j] : 0.f;
    rstd[j] = s.active ? sh[C + c0 + j] : 1.f;
  }
  __syncthreads();       // sh is reused by reduce_channels
  if (s.active) {
    while (p0 < s.pend) {
#pragma unroll
      for (int u = 0; u < UNB; ++u) {
        if (p0 + u * s.pl >= s.pend) break;
        float xh[8];
        dxhat8(f[u], g[u], base + p0 + u * s.pl, C, c0, act, drop_p, seed, mean, rstd, xh, xk, ga, be);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += g[u][j];
          a2[j] = fmaf(g[u][j], xh[j], a2[j]);
        }
      }
      p0 += UNB * s.pl;
#pragma unroll
      for (int u = 0; u < UNB; ++u)
        if (p0 + u * s.pl < s.pend) {
          load8(x, x_f32, (base + p0 + u * s.pl) * ldx + c0, f[u]);
          load_dy(dy1, ld1, dy2, ld2, base + p0 + u * s.pl, c0, g[u]);
        }
    }
  }
  reduce_channels(a1, a2, s, C, sh, bsums, b);
}

__global__ void __launch_bounds__(NT) norm_act_bwd_apply_kernel(const void* x, int x_f32, const float* sums,
                                                                const void* dy1, int ld1, const void* dy2, int ld2,
                                                                const float* bsums, void* dx, int lddx, long long HW,
                                                                int C, int ldx, int act, float drop_p,
                                                                const unsigned long long* seed_ptr,
                                                                unsigned long long salt, long long ppb, Affine af) {
  const int xk = x_f32 >> 8;
  x_f32 &= 0xff;
  extern __shared__ float sh[];
  const int b = blockIdx.y;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  const Span s = make_span(C, HW, ppb);
  const long long base = (long long)b * HW;
  const int c0 = s.my_cg * 8;
  float ga[8], be[8];
  load_affine(af, c0, ga, be);
  float f[UNB][8], g[UNB][8];
  long long p0 = s.pbeg + s.my_pl;
  if (s.active) {
#pragma unroll
    for (int u = 0; u < UNB; ++u)
      if (p0 + u * s.pl < s.pend) {
        load8(x, x_f32, (base + p0 + u * s.pl) * ldx + c0, f[u]);
        load_dy(dy1, ld1, dy2, ld2, base + p0 + u * s.pl, c0, g[u]);
      }
  }
  load_stats(sums, b, C, HW, sh, sh + C);
  if (!s.active) return;
  float mean[8], rstd[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = sh[c0 + j];
    rstd[j] = sh[C + c0 + j];
    if (sums != nullptr) {
      m1[j] = bsums[((long long)b * C + c0 + j) * 2] / (float)HW;
      m2[j] = bsums[((long long)b * C + c0 + j) * 2 + 1] / (float)HW;
    } else {
      m1[j] = 0.f; m2[j] = 0.f;
    }
  }
  while (p0 < s.pend) {
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const long long pix = base + p0 + u * s.pl;
      if (p0 + u * s.pl >= s.pend) break;
      float xh[8];
      dxhat8(f[u], g[u], pix, C, c0, act, drop_p, seed, mean, rstd, xh, xk, ga, be);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[u][j] = ga[j] * rstd[j] * (g[u][j] - m1[j] - xh[j] * m2[j]);
      store8(dx, 0, pix * lddx + c0, g[u]);
    }
    p0 += UNB * s.pl;
#pragma unroll
    for (int u = 0; u < UNB; ++u)
      if (p0 + u * s.pl < s.pend) {
        load8(x, x_f32, (base + p0 + u * s.pl) * ldx + c0, f[u]);
        load_dy(dy1, ld1, dy2, ld2, base + p0 + u * s.pl, c0, g[u]);
      }
  }
}

// Small maps (HW <= 1024: the 2x2 .. 32x32 layers): reduce and apply in ONE launch.  A block owns (image, 8-channel group):
// every thread keeps the dxhat / xhat values of its <= 4 pixels in registers, the block reduces the two sums per channel
// (shuffles + 8 warp partials in shared memory), then the same registers are turned into dx.  No bsums round trip through
// global atomics, x and dy are read once, one launch instead of two on the generator's backward critical path.
constexpr int SMALL_PPT = 4;     // pixels per thread

__global__ void __launch_bounds__(NT) norm_act_bwd_small_kernel(const void* x, int x_f32, const float* sums, const void* dy1,
                                                                int ld1, const void* dy2, int ld2, void* dx, int lddx,
                                                                int HW, int C, int ldx, int act, float drop_p,
                                                                const unsigned long long* seed_ptr, unsigned long long salt) {
  const int xk = x_f32 >> 8;
  x_f32 &= 0xff;
  __shared__ float sh_stat[16];            // mean[8], rstd[8]
  __shared__ float sh_part[NT / 32][16];   // per-warp partial sums
  __shared__ float sh_tot[16];
  const int b = blockIdx.y, c0 = blockIdx.x * 8;
  const unsigned long long seed = drop_p > 0.f ? mix_seed(*seed_ptr, salt) : 0ull;
  const long long base = (long long)b * HW;
  float g[SMALL_PPT][8], xh[SMALL_PPT][8];      // xh holds the raw x until the statistics are known
#pragma unroll
  for (int u = 0; u < SMALL_PPT; ++u) {
    const int p = threadIdx.x + u * NT;
    if (p < HW) {
      load8(x, x_f32, (base + p) * ldx + c0, xh[u]);
      load_dy(dy1, ld1, dy2, ld2, base + p, c0, g[u]);
    }
  }
  if (threadIdx.x < 8) {
    float m = 0.f, r = 1.f;
    if (sums != nullptr) mean_rstd(sums, b, C, c0 + threadIdx.x, HW, m, r);
    sh_stat[threadIdx.x] = m;
    sh_stat[8 + threadIdx.x] = r;
  }
  __syncthreads();
  float mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mean[j] = sh_stat[j]; rstd[j] = sh_stat[8 + j]; }
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
#pragma unroll
  for (int u = 0; u < SMALL_PPT; ++u) {
    const int p = threadIdx.x + u * NT;
    if (p < HW) {
      dxhat8(xh[u], g[u], base + p, C, c0, act, drop_p, seed, mean, rstd, xh[u], xk);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += g[u][j];
        a2[j] = fmaf(g[u][j], xh[u][j], a2[j]);
      }
    }
  }
  // block reduction of the 16 sums
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = warp_sum(a1[j]);
    a2[j] = warp_sum(a2[j]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sh_part[warp][j] = a1[j]; sh_part[warp][8 + j] = a2[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += sh_part[w][threadIdx.x];
    sh_tot[threadIdx.x] = sums != nullptr ? t / (float)HW : 0.f;
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { m1[j] = sh_tot[j]; m2[j] = sh_tot[8 + j]; }
#pragma unroll
  for (int u = 0; u < SMALL_PPT; ++u) {
    const int p = threadIdx.x + u * NT;
    if (p < HW) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[u][j] = rstd[j] * (g[u][j] - m1[j] - xh[u][j] * m2[j]);
      store8(dx, 0, (base + p) * lddx + c0, g[u]);
    }
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

// bps: blocks of this kernel that fit one SM (registers).  One full wave of long-lived blocks: the (mean, rstd) prologue is
// paid once per block and there is no partial last wave.  PG_NORM_BPS=n overrides (8 = the earlier many-short-blocks grid).
// ------------------------------------------------------------------ BatchNorm2d on top of the per-image kernels
// The kernels above normalise image b with (sums[b][c] / HW).  BatchNorm2d (unet.py:20,55 with norm_layer = nn.BatchNorm2d)
// uses ONE mean / variance per channel over the whole batch: fold the per-image pairs into their batch mean and write it
// back into every image's slot, and the same kernels produce BatchNorm.  training != 0 also updates running_mean /
// running_var the way aten::batch_norm does (momentum, unbiased variance); training == 0 fills the slots from the running
// statistics instead (eval mode).  One thread per channel, fp64.
__global__ void bn_fold_fwd_kernel(float* sums, int B, int C, long long HW, float* running_mean, float* running_var, int c_real,
                                   float momentum, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  if (training) {
    for (int b = 0; b < B; ++b) {
      s += (double)sums[((long long)b * C + c) * 2];
      ss += (double)sums[((long long)b * C + c) * 2 + 1];
    }
    const double M = (double)B * (double)HW, mean = s / M;
    double var = ss / M - mean * mean;
    if (var < 0) var = 0;
    if (c < c_real && running_mean != nullptr) {
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * (M > 1 ? var * M / (M - 1) : var));
    }
    s /= B; ss /= B;
  } else {
    const double m = c < c_real ? (double)running_mean[c] : 0.0, v = c < c_real ? (double)running_var[c] : 1.0;
    s = m * (double)HW;
    ss = (v + m * m) * (double)HW;
  }
  for (int b = 0; b < B; ++b) {
    sums[((long long)b * C + c) * 2] = (float)s;
    sums[((long long)b * C + c) * 2 + 1] = (float)ss;
  }
}
// Backward: bsums[b][c] = (sum g, sum g * xhat) of image b.  dbeta[c] += sum_b sum g, dgamma[c] += sum_b sum g * xhat, and the
// slots are replaced by their batch mean (training) or by zero (eval mode: the statistics are constants, no mean terms).
__global__ void bn_fold_bwd_kernel(float* bsums, int B, int C, float* dgamma, float* dbeta, int c_real, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a1 = 0.0, a2 = 0.0;
  for (int b = 0; b < B; ++b) {
    a1 += (double)bsums[((long long)b * C + c) * 2];
    a2 += (double)bsums[((long long)b * C + c) * 2 + 1];
  }
  if (c < c_real && dgamma != nullptr) {
    dbeta[c] += (float)a1;
    dgamma[c] += (float)a2;
  }
  const float m1 = training ? (float)(a1 / B) : 0.f, m2 = training ? (float)(a2 / B) : 0.f;
  for (int b = 0; b < B; ++b) {
    bsums[((long long)b * C + c) * 2] = m1;
    bsums[((long long)b * C + c) * 2 + 1] = m2;
  }
}

static void span_grid(int B, long long HW, int C, dim3& grid, long long& ppb, int bps) {
  static const int bps_env = [] { const char* e = getenv("PG_NORM_BPS"); return e ? atoi(e) : 0; }();
  if (bps_env > 0) bps = bps_env;
  const int pl = NT / (C >> 3);
  long long per_img = bps_env > 0 ? ((long long)bps * num_sms() + B - 1) / B : ((long long)bps * num_sms()) / B;
  if (per_img < 1) per_img = 1;
  ppb = (HW + per_img - 1) / per_img;
  if (ppb < 1LL * pl) ppb = 1LL * pl;
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
  span_grid(B, HW, C, grid, ppb, 4);
  instnorm_stats_kernel<<<grid, NT, reduce_smem(C), (cudaStream_t)stream>>>(x, x_f32, HW, C, ld, sums, ppb);
  return check_launch("instnorm_stats_kernel");
}

extern "C" int pg_norm_act_fwd(const void* x, int32_t x_f32, const float* sums, void* y, int32_t y_f32, void* y2, int32_t B,
                               int64_t HW, int32_t C, int32_t ldx, int32_t ldy, int32_t act, float drop_p,
                               const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_act_fwd", C)) return e;
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 3);
  norm_act_fwd_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, x_f32, sums, y, y_f32, y2, HW, C, ldx, ldy, act, drop_p,
                                                            (const unsigned long long*)seed, salt, ppb, Affine{nullptr, nullptr, 0});
  return check_launch("norm_act_fwd_kernel");
}

extern "C" int pg_norm_affine_act_fwd(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                                      int32_t c_real, void* y, int32_t y_f32, void* y2, int32_t B, int64_t HW, int32_t C,
                                      int32_t ldx, int32_t ldy, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt,
                                      void* stream) {
  if (int e = check_c("pg_norm_affine_act_fwd", C)) return e;
  PG_REQUIRE(gamma != nullptr && beta != nullptr && c_real > 0 && c_real <= C, "pg_norm_affine_act_fwd: gamma / beta / c_real");
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 3);
  norm_act_fwd_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, x_f32, sums, y, y_f32, y2, HW, C, ldx, ldy, act, drop_p,
                                                            (const unsigned long long*)seed, salt, ppb, Affine{gamma, beta, c_real});
  return check_launch("norm_act_fwd_kernel");
}

extern "C" int pg_bn_fold_fwd(float* sums, int32_t B, int32_t C, int64_t HW, float* running_mean, float* running_var, int32_t c_real,
                              float momentum, int32_t training, void* stream) {
  PG_REQUIRE(sums != nullptr && B > 0 && C > 0 && HW > 0 && c_real >= 0 && c_real <= C, "pg_bn_fold_fwd: bad extents");
  PG_REQUIRE(training || (running_mean != nullptr && running_var != nullptr), "pg_bn_fold_fwd: eval mode needs the running statistics");
  PG_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "pg_bn_fold_fwd: running_mean / running_var go together");
  bn_fold_fwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, B, C, HW, running_mean, running_var, c_real, momentum,
                                                                       training);
  return check_launch("bn_fold_fwd_kernel");
}

extern "C" int pg_bn_fold_bwd(float* bsums, int32_t B, int32_t C, float* dgamma, float* dbeta, int32_t c_real, int32_t training,
                              void* stream) {
  PG_REQUIRE(bsums != nullptr && B > 0 && C > 0 && c_real >= 0 && c_real <= C, "pg_bn_fold_bwd: bad extents");
  PG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "pg_bn_fold_bwd: dgamma / dbeta go together");
  bn_fold_bwd_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bsums, B, C, dgamma, dbeta, c_real, training);
  return check_launch("bn_fold_bwd_kernel");
}

extern "C" int pg_norm_act_bwd_reduce(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                                      const void* dy2, int32_t ld2, float* bsums, int32_t B, int64_t HW, int32_t C,
                                      int32_t ldx, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt,
                                      void* stream) {
  if (int e = check_c("pg_norm_act_bwd_reduce", C)) return e;
  PG_REQUIRE(sums != nullptr, "pg_norm_act_bwd_reduce: sums is NULL");
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 2);
  norm_act_bwd_reduce_kernel<<<grid, NT, reduce_smem(C), (cudaStream_t)stream>>>(
      x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb,
      Affine{nullptr, nullptr, 0});
  return check_launch("norm_act_bwd_reduce_kernel");
}

extern "C" int pg_norm_affine_act_bwd_reduce(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                                             int32_t c_real, const void* dy1, int32_t ld1, const void* dy2, int32_t ld2,
                                             float* bsums, int32_t B, int64_t HW, int32_t C, int32_t ldx, int32_t act, float drop_p,
                                             const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_affine_act_bwd_reduce", C)) return e;
  PG_REQUIRE(sums != nullptr && gamma != nullptr && beta != nullptr && c_real > 0 && c_real <= C,
             "pg_norm_affine_act_bwd_reduce: sums / gamma / beta / c_real");
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 2);
  norm_act_bwd_reduce_kernel<<<grid, NT, reduce_smem(C), (cudaStream_t)stream>>>(
      x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb,
      Affine{gamma, beta, c_real});
  return check_launch("norm_act_bwd_reduce_kernel");
}

extern "C" int pg_norm_act_bwd_apply(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                                     const void* dy2, int32_t ld2, const float* bsums, void* dx, int32_t lddx,
                                     int32_t B, int64_t HW, int32_t C, int32_t ldx, int32_t act, float drop_p,
                                     const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_act_bwd_apply", C)) return e;
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 2);
  norm_act_bwd_apply_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, dx,
                                                                  lddx, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb,
                                                                  Affine{nullptr, nullptr, 0});
  return check_launch("norm_act_bwd_apply_kernel");
}

extern "C" int pg_norm_affine_act_bwd_apply(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                                            int32_t c_real, const void* dy1, int32_t ld1, const void* dy2, int32_t ld2,
                                            const float* bsums, void* dx, int32_t lddx, int32_t B, int64_t HW, int32_t C,
                                            int32_t ldx, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt,
                                            void* stream) {
  if (int e = check_c("pg_norm_affine_act_bwd_apply", C)) return e;
  PG_REQUIRE(sums != nullptr && gamma != nullptr && beta != nullptr && c_real > 0 && c_real <= C,
             "pg_norm_affine_act_bwd_apply: sums / gamma / beta / c_real");
  dim3 grid; long long ppb;
  span_grid(B, HW, C, grid, ppb, 2);
  norm_act_bwd_apply_kernel<<<grid, NT, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, dx,
                                                                  lddx, HW, C, ldx, act, drop_p, (const unsigned long long*)seed, salt, ppb,
                                                                  Affine{gamma, beta, c_real});
  return check_launch("norm_act_bwd_apply_kernel");
}

// Backward through dropout / activation / InstanceNorm in one call: small maps in one launch, larger ones as reduce + apply
// (bsums: zeroed float32 workspace [B][C][2], used by the two-pass path only).
extern "C" int pg_norm_act_bwd(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1, const void* dy2,
                               int32_t ld2, float* bsums, void* dx, int32_t lddx, int32_t B, int64_t HW, int32_t C, int32_t ldx,
                               int32_t act, float drop_p, const uint64_t* seed, uint64_t salt, void* stream) {
  if (int e = check_c("pg_norm_act_bwd", C)) return e;
  if (HW <= (int64_t)SMALL_PPT * NT) {
    dim3 grid((unsigned)(C / 8), (unsigned)B);
    norm_act_bwd_small_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(x, x_f32, sums, dy1, ld1, dy2, ld2, dx, lddx, (int)HW, C, ldx,
                                                                    act, drop_p, (const unsigned long long*)seed, salt);
    return check_launch("norm_act_bwd_small_kernel");
  }
  PG_REQUIRE(bsums != nullptr, "pg_norm_act_bwd: bsums workspace needed for HW > %d", SMALL_PPT * NT);
  if (int e = pg_norm_act_bwd_reduce(x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, B, HW, C, ldx, act, drop_p, seed, salt, stream))
    return e;
  return pg_norm_act_bwd_apply(x, x_f32, sums, dy1, ld1, dy2, ld2, bsums, dx, lddx, B, HW, C, ldx, act, drop_p, seed, salt, stream);
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
