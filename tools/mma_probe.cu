// Microbenchmark: how fast can ONE thread of ONE CTA issue tcgen05.mma (cta_group::1, kind::f16, M=128) chains?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
// Prints ns per MMA for N in {16,64,128,256}, with 1/2/4 rotating accumulators, with/without a commit every 4 MMAs,
// and with 1 or 2 CTAs resident per SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit_generic(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((uint64_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * swz) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swz == 128 ? 2 : (swz == 64 ? 4 : 6)) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 2) probe(int N, int nacc, int commit_every, int nmma, long long* out, int swz, int form) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)(smem_raw + (base - smem_u32(smem_raw))))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    uint32_t cols = nacc * N < 32 ? 32 : nacc * N;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_desc(base, swz), bd = make_desc(base + 16384, swz);
    const int kk = swz / 32;
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      umma(tm + (uint32_t)((i % nacc) * N), ad + 2 * (i % kk), bd + 2 * (i % kk), idesc, i >= nacc);
      if (commit_every > 0 && (i % commit_every) == commit_every - 1) {
        if (form == 0) commit(smem_u32(&bar[1])); else commit_generic(&bar[1]);
      }
    }
    long long t1 = clock64();
    commit(smem_u32(&bar[0]));
    while (!mbar_try_wait(smem_u32(&bar[0]), 0)) {}
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t cols = nacc * N < 32 ? 32 : nacc * N;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(cols) : "memory");
  }
}


template <int NACC>
__global__ void __launch_bounds__(128, 2) probe_lean(int N, int commit_every8, int nmma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)(smem_raw + (base - smem_u32(smem_raw))))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t cols = NACC * N < 32 ? 32 : NACC * N;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = make_desc(base, 128), bd = make_desc(base + 16384, 128);
    const uint32_t b1 = smem_u32(&bar[1]);
    long long t0 = clock64();
    for (int i = 0; i < nmma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        umma(tm + (uint32_t)((j % NACC) * N), ad + 2 * (j & 3), bd + 2 * (j & 3), idesc, 1u);
      if (commit_every8) commit(b1);
    }
    long long t1 = clock64();
    commit(smem_u32(&bar[0]));
    while (!mbar_try_wait(smem_u32(&bar[0]), 0)) {}
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(cols) : "memory");
}

template <int NACC>
void run_lean(long long* out) {
  cudaFuncSetAttribute(probe_lean<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int ctas : {1, 2})
    for (int N : {16, 32, 64, 128, 256})
      for (int ce : {0, 1}) {
        if (NACC * N * ctas > 512) continue;
        probe_lean<NACC><<<148 * ctas, 128, ctas == 1 ? 100 * 1024 : 60 * 1024>>>(N, ce, 512, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[2];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("lean nacc %d ctas/SM %d N %3d commit/8 %d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (floor %d)\n", NACC, ctas, N, ce,
               (double)h[0] / 512, (double)h[1] / 512, N / 2);
      }
}

int main() {
  long long* out;
  cudaMalloc(&out, 1 << 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock %d kHz\n", clk);
  run_lean<1>(out); run_lean<2>(out); run_lean<4>(out);
  return 0;
  const int nmma = 256;
  for (int swz : {128, 32})
    for (int form : {0, 1})
      for (int ctas_per_sm : {1, 2})
        for (int N : {64, 256})
          for (int ce : {0, 1, 2, 4, 8, 16}) {
            const int nacc = 1;
            if (form == 1 && (ce == 0 || swz == 32)) continue;
            const size_t smem = ctas_per_sm == 1 ? 100 * 1024 : 60 * 1024;
            probe<<<148 * ctas_per_sm, 128, smem>>>(N, nacc, ce, nmma, out, swz, form);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[4];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("swz %3d form %d ctas/SM %d N %3d commit_every %2d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (floor %d)\n", swz, form,
                   ctas_per_sm, N, ce, (double)h[0] / nmma, (double)h[1] / nmma, N / 2);
          }
  return 0;
}
