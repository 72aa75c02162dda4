// optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) over one flat fp32 parameter buffer.
// Reference: trainer.py:169-172 (construction), :90 and :107 (step).  HBM-bound: 4 reads + 3 writes of 4 bytes
// per parameter, float4-vectorised, one launch for all tensors of a network.
#include "common.cuh"

namespace pg {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   const float* __restrict__ hyper, const int* __restrict__ step,
                                                   float b1, float b2, float eps, float gscale) {
  const float lr = hyper[0];
  const int t = *step + 1;
  const float bc1 = 1.f - powf(b1, (float)t);
  const float bc2 = 1.f - powf(b2, (float)t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gscale;
      ma[j] = b1 * ma[j] + (1.f - b1) * gr;
      va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(va[j]) * inv_sqrt_bc2 + eps;
      pa[j] -= step_size * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
}

__global__ void bump_step_kernel(int* step) { *step += 1; }

}  // namespace pg
using namespace pg;

static int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                       float beta1, float beta2, float eps, float grad_scale, int bump, void* stream);

extern "C" int pg_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                            float beta1, float beta2, float eps, float grad_scale, void* stream) {
  return adam_launch(p, g, m, v, n, hyper, step, beta1, beta2, eps, grad_scale, 1, stream);
}

extern "C" int pg_adam_step_range(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                                  float beta1, float beta2, float eps, float grad_scale, int32_t bump, void* stream) {
  return adam_launch(p, g, m, v, n, hyper, step, beta1, beta2, eps, grad_scale, bump, stream);
}

static int adam_launch(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                       float beta1, float beta2, float eps, float grad_scale, int bump, void* stream) {
  PG_REQUIRE(n >= 0, "pg_adam_step: n < 0");
  PG_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
             "pg_adam_step: buffers must be 16-byte aligned");
  if (n == 0 && !bump) return PG_OK;
  long long nb = ((n >> 2) + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  adam_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, step, beta1, beta2, eps, grad_scale);
  if (int e = check_launch("adam_kernel")) return e;
  if (!bump) return PG_OK;
  bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step);
  return check_launch("bump_step_kernel");
}
