// Single-pass segmentation and adversarial losses with their gradients.
// Reference: losses.py:18-39 (fc_tversky, MAE_loss, bce_loss) and trainer.py:71-85,101-103.
// HBM-bound; one thread per pixel, target read straight from the user's NCHW float tensor (coalesced
// per channel plane), prediction read from the generator's f32 NHWC output.
#include "common.cuh"

namespace pg {

constexpr int LT_TVERSKY = 0, LT_WBCE = 1, LT_MAE = 2, LT_NONE = 3;
constexpr int FINAL_SOFTMAX = 5;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (int)(blockDim.x >> 5) ? sh[l] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

__global__ void target_chsum_kernel(const float* __restrict__ t, float* chsum, long long HW) {
  __shared__ float sh[32];
  const int bc = blockIdx.y;
  float s = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x)
    s += t[(long long)bc * HW + p];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(chsum + bc, s);
}

__device__ __forceinline__ float bce_elem(float p, float t) {
  const float lp = fmaxf(logf(p), -100.f);
  const float l1p = fmaxf(log1pf(-p), -100.f);
  return -(t * lp + (1.f - t) * l1p);
}

__global__ void __launch_bounds__(256) seg_loss_partials_kernel(const float* __restrict__ p, int ld,
                                                                const float* __restrict__ t,
                                                                const float* __restrict__ chsum, float* part, int B,
                                                                int C, long long HW, int loss_type) {
  __shared__ float sh[32];
  __shared__ float wsh[64];
  const int b = blockIdx.y;
  if (loss_type == LT_WBCE) {
    // trainer.py:76-79: weight = 1 - sum_hw(t) / sum(t) if C > 1 else 1
    if (threadIdx.x < C) {
      float w = 1.f;
      if (C > 1) {
        float tot = 0.f;
        for (int i = 0; i < B * C; ++i) tot += chsum[i];
        w = 1.f - chsum[b * C + threadIdx.x] / tot;
      }
      wsh[threadIdx.x] = w;
    }
    __syncthreads();
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < HW;
       px += (long long)gridDim.x * blockDim.x) {
    const float* pp = p + ((long long)b * HW + px) * ld;
    for (int c = 0; c < C; ++c) {
      const float pv = pp[c];
      const float tv = t[((long long)b * C + c) * HW + px];
      a0 = fmaf(tv, pv, a0);
      a1 += tv;
      a2 += pv;
      a3 += fabsf(pv - tv);
      if (loss_type == LT_WBCE) a4 = fmaf(wsh[c], bce_elem(pv, tv), a4);
    }
  }
  a0 = block_sum(a0, sh); a1 = block_sum(a1, sh); a2 = block_sum(a2, sh); a3 = block_sum(a3, sh);
  a4 = block_sum(a4, sh);
  if (threadIdx.x == 0) {
    atomicAdd(part + b * 8 + 0, a0);
    atomicAdd(part + b * 8 + 1, a1);
    atomicAdd(part + b * 8 + 2, a2);
    atomicAdd(part + b * 8 + 3, a3);
    atomicAdd(part + b * 8 + 4, a4);
  }
}

// One warp; lane b owns batch element b (strided for B > 32): the per-sample terms are computed side by side and summed by
// shuffles (a single thread walking the batch with fp64 divisions took 13 us between the generator's forward and backward).
__global__ void seg_loss_finalize_kernel(const float* part, float* coef, float* losses, int slot, int B, int C,
                                         long long HW, int loss_type, float beta, float gamma, float seg_alpha) {
  if (blockIdx.x != 0) return;
  const int lane = threadIdx.x;
  const double cnt = (double)B * C * (double)HW;
  auto warp_sum_d = [](double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  };
  if (loss_type == LT_TVERSKY) {
    // losses.py:18-31 with smooth = 1; tp = sum t p, fn = sum t - tp, fp = sum p - tp
    double m = 0;
    for (int b = lane; b < B; b += 32) {
      const double tp = part[b * 8], st = part[b * 8 + 1], sp = part[b * 8 + 2];
      const double num = tp + 1.0, den = tp + beta * (st - tp) + (1.0 - beta) * (sp - tp) + 1.0;
      m += 1.0 - num / den;
    }
    m = warp_sum_d(m) / B;
    if (lane == 0) losses[slot] = (float)(seg_alpha * pow(m, (double)gamma));
    const double k0 = -seg_alpha * gamma * pow(m, (double)gamma - 1.0) / B;
    for (int b = lane; b < B; b += 32) {
      const double tp = part[b * 8], st = part[b * 8 + 1], sp = part[b * 8 + 2];
      const double num = tp + 1.0, den = tp + beta * (st - tp) + (1.0 - beta) * (sp - tp) + 1.0;
      // dTI/dp = (t*den - num*(1-beta)) / den^2
      coef[b * 4 + 0] = (float)(k0 / den);                          // multiplies t
      coef[b * 4 + 1] = (float)(-k0 * num * (1.0 - beta) / (den * den));  // constant term
    }
  } else {
    const int col = loss_type == LT_WBCE ? 4 : 3;
    double s = 0;
    for (int b = lane; b < B; b += 32) s += part[b * 8 + col];
    s = warp_sum_d(s);
    if (lane == 0) losses[slot] = (float)(seg_alpha * s / cnt);
    for (int b = lane; b < B; b += 32) coef[b * 4] = (float)(seg_alpha / cnt);
  }
}

__global__ void __launch_bounds__(256) gen_out_bwd_kernel(const float* __restrict__ p, int ld,
                                                          const float* __restrict__ t,
                                                          const float* __restrict__ chsum,
                                                          const float* __restrict__ coef, const bf16* __restrict__ dD,
                                                          int lddd, int dd_off, bf16* __restrict__ dx, int lddx,
                                                          int B, int C, long long HW, int loss_type, int final_act) {
  __shared__ float wsh[64];
  const int b = blockIdx.y;
  if (loss_type == LT_WBCE) {
    if (threadIdx.x < C) {
      float w = 1.f;
      if (C > 1) {
        float tot = 0.f;
        for (int i = 0; i < B * C; ++i) tot += chsum[i];
        w = 1.f - chsum[b * C + threadIdx.x] / tot;
      }
      wsh[threadIdx.x] = w;
    }
    __syncthreads();
  }
  const float c0 = loss_type == LT_NONE ? 0.f : coef[b * 4], c1 = loss_type == LT_NONE ? 0.f : coef[b * 4 + 1];
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < HW;
       px += (long long)gridDim.x * blockDim.x) {
    const long long pix = (long long)b * HW + px;
    const float* pp = p + pix * ld;
    float dp[16], pv[16];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { dp[c] = 0.f; pv[c] = 0.f; }
    for (int c = 0; c < C; ++c) {
      const float q = pp[c];
      const float tv = loss_type == LT_NONE ? 0.f : t[((long long)b * C + c) * HW + px];
      float g;
      if (loss_type == LT_TVERSKY) g = c0 * tv + c1;
      else if (loss_type == LT_WBCE) g = c0 * wsh[c] * (q - tv) / fmaxf(q * (1.f - q), 1e-12f);
      else if (loss_type == LT_MAE) { const float d = q - tv; g = c0 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)); }
      else g = 0.f;
      if (dD != nullptr) g += __bfloat162float(dD[pix * lddd + dd_off + c]);
      dp[c] = g;
      pv[c] = q;
      dot = fmaf(g, q, dot);
    }
    float outv[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float g = 0.f;
      if (c < C) {
        if (final_act == FINAL_SOFTMAX) g = pv[c] * (dp[c] - dot);
        else g = dp[c] * act_grad_from_output(final_act, pv[c]);
      }
      outv[c] = g;
    }
    if (lddx == 16) {          // the usual case: one 32-byte row
      *reinterpret_cast<uint4*>(dx + pix * 16) = pack8(outv);
      *reinterpret_cast<uint4*>(dx + pix * 16 + 8) = pack8(outv + 8);
    } else if (lddx == 4) {    // trimmed row (one real channel, consumed by the tap gather): one 8-byte store
      const uint4 v = pack8(outv);
      *reinterpret_cast<uint2*>(dx + pix * 4) = make_uint2(v.x, v.y);
    } else {
      for (int c = 0; c < lddx; ++c) dx[pix * lddx + c] = __float2bfloat16(c < 16 ? outv[c] : 0.f);
    }
  }
}

__global__ void __launch_bounds__(256) bce_const_kernel(const float* __restrict__ p, int ld, float label, float gscale,
                                                        float* losses, int slot, bf16* dz, int lddz, long long npix) {
  __shared__ float sh[32];
  float s = 0.f;
  const float inv = 1.f / (float)npix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const float q = p[i * ld];
    s += bce_elem(q, label);
    if (dz != nullptr) {
      // torch: grad_p = (p - t) / max(p (1 - p), 1e-12) / N ; sigmoid backward multiplies by p (1 - p)
      const float pq = q * (1.f - q);
      const float g = gscale * inv * (q - label) / fmaxf(pq, 1e-12f) * pq;
      dz[i * lddz] = __float2bfloat16(g);
      for (int c = 1; c < lddz; ++c) dz[i * lddz + c] = __float2bfloat16(0.f);
    }
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(losses + slot, s * inv);
}

// ---- the standalone loss API (losses.py:5-39 called on arbitrary tensors, outside the Trainer step)
// nn.BCELoss()(p, t) with a general target: losses[slot] += mean bce
__global__ void __launch_bounds__(256) bce_mean_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n,
                                                       float* losses, int slot) {
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += bce_elem(p[i], t[i]);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(losses + slot, s / (float)n);
}
// its gradient (aten::binary_cross_entropy_backward): dp = gout * (p - t) / max(p (1 - p), 1e-12) / n; gout: device scalar
__global__ void __launch_bounds__(256) bce_mean_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n,
                                                           const float* __restrict__ gout, float* __restrict__ dp) {
  const float g = *gout / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float q = p[i];
    dp[i] = g * (q - t[i]) / fmaxf(q * (1.f - q), 1e-12f);
  }
}
// gradient of the per-sample sums (tp, sum p) of tversky / fc_tversky(batch_mean=False) wrt the prediction (NCHW float):
//   dp[b, c, hw] = g_tp[b] * t[b, c, hw] + g_sp[b]
__global__ void __launch_bounds__(256) sample_sums_bwd_kernel(const float* __restrict__ t, const float* __restrict__ g_tp,
                                                              const float* __restrict__ g_sp, float* __restrict__ dp,
                                                              long long chw) {
  const int b = blockIdx.y;
  const float a = g_tp[b], c = g_sp[b];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chw; i += (long long)gridDim.x * blockDim.x)
    dp[(long long)b * chw + i] = fmaf(a, t[(long long)b * chw + i], c);
}

static dim3 img_grid(int B, long long HW) {
  long long per = (4LL * num_sms() + B - 1) / B;
  long long nb = (HW + 255) / 256;
  if (nb > per) nb = per;
  if (nb < 1) nb = 1;
  return dim3((unsigned)nb, (unsigned)B, 1);
}

}  // namespace pg
using namespace pg;

extern "C" int pg_target_chsum(const float* t, float* chsum, int32_t B, int32_t C, int64_t HW, void* stream) {
  long long nb = (HW + 1023) / 1024;
  if (nb > 64) nb = 64;
  dim3 grid((unsigned)nb, (unsigned)(B * C));
  target_chsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, chsum, HW);
  return check_launch("target_chsum_kernel");
}

extern "C" int pg_seg_loss_partials(const float* p, int32_t ld, const float* t, const float* chsum, float* part,
                                    int32_t B, int32_t C, int64_t HW, int32_t loss_type, void* stream) {
  PG_REQUIRE(C >= 1 && C <= 16 && C <= ld, "pg_seg_loss_partials: C=%d unsupported (1..16)", C);
  PG_REQUIRE(loss_type >= 0 && loss_type <= 2, "pg_seg_loss_partials: loss_type=%d", loss_type);
  PG_REQUIRE(loss_type != LT_WBCE || chsum != nullptr, "pg_seg_loss_partials: weighted_bce needs chsum");
  seg_loss_partials_kernel<<<img_grid(B, HW), 256, 0, (cudaStream_t)stream>>>(p, ld, t, chsum, part, B, C, HW,
                                                                             loss_type);
  return check_launch("seg_loss_partials_kernel");
}

extern "C" int pg_seg_loss_finalize(const float* part, float* coef, float* losses, int32_t slot, int32_t B, int32_t C,
                                    int64_t HW, int32_t loss_type, float beta, float gamma, float seg_alpha,
                                    void* stream) {
  seg_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(part, coef, losses, slot, B, C, HW, loss_type, beta,
                                                              gamma, seg_alpha);
  return check_launch("seg_loss_finalize_kernel");
}

extern "C" int pg_gen_out_bwd(const float* p, int32_t ld, const float* t, const float* chsum, const float* coef,
                              const void* dD, int32_t lddd, int32_t dd_off, void* dx, int32_t lddx, int32_t B,
                              int32_t C, int64_t HW, int32_t loss_type, int32_t final_act, float beta, void* stream) {
  (void)beta;
  PG_REQUIRE(C >= 1 && C <= 16 && C <= ld && C <= lddx, "pg_gen_out_bwd: C=%d unsupported (1..16)", C);
  gen_out_bwd_kernel<<<img_grid(B, HW), 256, 0, (cudaStream_t)stream>>>(p, ld, t, chsum, coef, (const bf16*)dD, lddd,
                                                                       dd_off, (bf16*)dx, lddx, B, C, HW, loss_type,
                                                                       final_act);
  return check_launch("gen_out_bwd_kernel");
}

extern "C" int pg_bce_const(const float* p, int32_t ld, float label, float gscale, float* losses, int32_t slot,
                            void* dz, int32_t lddz, int64_t npix, void* stream) {
  long long nb = (npix + 255) / 256;
  const long long cap = 2LL * num_sms();
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  bce_const_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(p, ld, label, gscale, losses, slot, (bf16*)dz, lddz,
                                                                  npix);
  return check_launch("bce_const_kernel");
}

static unsigned flat_blocks(long long n) {
  long long nb = (n + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (nb > cap) nb = cap;
  return (unsigned)(nb < 1 ? 1 : nb);
}

extern "C" int pg_bce_mean(const float* p, const float* t, int64_t n, float* losses, int32_t slot, void* stream) {
  PG_REQUIRE(p && t && losses && n > 0, "pg_bce_mean: bad arguments");
  bce_mean_kernel<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, t, n, losses, slot);
  return check_launch("bce_mean_kernel");
}

extern "C" int pg_bce_mean_bwd(const float* p, const float* t, int64_t n, const float* gout, float* dp, void* stream) {
  PG_REQUIRE(p && t && gout && dp && n > 0, "pg_bce_mean_bwd: bad arguments");
  bce_mean_bwd_kernel<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(p, t, n, gout, dp);
  return check_launch("bce_mean_bwd_kernel");
}

extern "C" int pg_sample_sums_bwd(const float* t, const float* g_tp, const float* g_sp, float* dp, int32_t B, int64_t chw,
                                  void* stream) {
  PG_REQUIRE(t && g_tp && g_sp && dp && B > 0 && chw > 0, "pg_sample_sums_bwd: bad arguments");
  dim3 grid(flat_blocks(chw), (unsigned)B);
  sample_sums_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, g_tp, g_sp, dp, chw);
  return check_launch("sample_sums_bwd_kernel");
}
