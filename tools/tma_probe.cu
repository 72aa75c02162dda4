// Microbenchmark: L2 -> shared-memory fill rate per SM for (a) TMA 2D boxes of 128-byte rows (row pitch = ld bytes),
// (b) TMA 4D NHWC boxes with element stride 2, (c) 1-D cp.async.bulk copies of contiguous chunks.
// One CTA per SM (or --ctas-per-sm N), each CTA streams `iters` chunks of `chunk` bytes through a ring of mbarriers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ bool mb_try(uint32_t b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
  return ok;
}
__device__ __forceinline__ void mb_wait(uint32_t b, uint32_t ph) { while (!mb_try(b, ph)) {} }

#ifndef STAGES
#define STAGES 6
#endif

// mode 0: 2D box {64 elem (128B), rows} ; mode 1: 4D box NHWC es=2 ; mode 2: 1-D bulk
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap map, const uint8_t* base, int mode,
                                               int rows, int iters, long long span_rows, int chunk_bytes, int W, int H) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  const uint32_t sb = (s32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mb_init(s32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // issue-ahead ring: wait for stage s (issued STAGES iterations ago) before reusing it
    uint32_t phase_bits = 0;
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) { mb_wait(s32(&full[s]), (phase_bits >> s) & 1); phase_bits ^= 1u << s; }
      if (it < iters) {
        const uint32_t fb = s32(&full[s]);
        const uint32_t dst = sb + s * chunk_bytes;
        mb_expect(fb, chunk_bytes);
        const long long id = (long long)blockIdx.x * iters + it;
        if (mode == 0) {
          const int r0 = (int)((id * rows) % span_rows);
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(fb), "r"(0), "r"(r0) : "memory");
        } else if (mode == 3) {
          // 4D unit-stride box {64, tw, th, tb}: W = tw*8, H = th*8 tiles, walk them
          const int tw = W, th = H, tb = rows / (W * H);
          const long long t = id % 64;
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(fb), "r"(0), "r"((int)(t % 8) * tw), "r"((int)(t / 8) * th), "r"((int)((id / 64) % 4) * tb) : "memory");
        } else if (mode == 1) {
          // NHWC: box {64, TW*2 (es2), TH*2 (es2), 1}; walk tiles
          const int tw = 32, th = rows / 32;
          const int tiles_x = W / (2 * tw), tiles_y = H / (2 * th);
          const long long t = id % ((long long)tiles_x * tiles_y * 16);
          const int x0 = (int)(t % tiles_x) * 2 * tw, y0 = (int)((t / tiles_x) % tiles_y) * 2 * th, b = (int)(t / (tiles_x * tiles_y));
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(fb), "r"(0), "r"(x0 - 1), "r"(y0 - 1), "r"(b) : "memory");
        } else {
          const long long off = (id * chunk_bytes) % (span_rows * 128);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(base + off), "r"(chunk_bytes), "r"(fb) : "memory");
        }
      }
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void* sym; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)sym;
  const size_t bytes = 64ull << 20;   // 64 MB working set (fits the 126 MB L2)
  uint8_t* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  printf("SMs %d\n", sms);
  for (int ctas_per_sm = 1; ctas_per_sm <= 1; ++ctas_per_sm) {
    // (a) 2D boxes: row pitch ld bytes, 128B rows
    for (int ld : {128, 256, 512, 1024, 8192}) {
      for (int rows : {128, 256}) {
        CUtensorMap m;
        const long long span_rows = (long long)(bytes / ld);
        cuuint64_t dims[2] = {64, (cuuint64_t)span_rows};
        cuuint64_t str[1] = {(cuuint64_t)ld};
        cuuint32_t box[2] = {64, (cuuint32_t)rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int chunk = rows * 128;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          probe<<<sms * ctas_per_sm, 64, STAGES * chunk + 1024>>>(m, buf, 0, rows, iters, span_rows - rows, chunk, 0, 0);
          cudaEventRecord(e1); CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (rep == 1) {
            const double tot = (double)sms * ctas_per_sm * iters * chunk;
            printf("2D box  pitch %5d B rows %3d: %7.1f GB/s total, %6.1f GB/s per SM, %5.2f ns per 128B row per SM\n", ld, rows,
                   tot / ms / 1e6, tot / ms / 1e6 / sms, ms * 1e6 / ((double)iters * rows * ctas_per_sm));
          }
        }
      }
    }
    // (b) NHWC stride-2 boxes, C = 64 / 128 / 256 channels per pixel
    for (int C : {64, 128, 256}) {
      const int W = 128, H = 128, B = 16;
      CUtensorMap m;
      cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      const int rows = 128;
      cuuint32_t box[4] = {64, 64, (cuuint32_t)(rows / 32 * 2), 1};
      cuuint32_t es[4] = {1, 2, 2, 1};
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode4d failed %d\n", (int)r); continue; }
      const int chunk = rows * 128;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probe<<<sms * ctas_per_sm, 64, STAGES * chunk + 1024>>>(m, buf, 1, rows, iters, 0, chunk, W, H);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 1) {
          const double tot = (double)sms * ctas_per_sm * iters * chunk;
          printf("4D es=2 box C=%3d rows %3d: %7.1f GB/s total, %6.1f GB/s per SM, %5.2f ns per row per SM\n", C, rows,
                 tot / ms / 1e6, tot / ms / 1e6 / sms, ms * 1e6 / ((double)iters * rows));
        }
      }
    }
    // (b2) 4D unit-stride boxes of 128 rows with different shapes; pixel pitch = C*2 B (and 2x for the "phase view")
    for (int step : {1, 2}) {
      for (int shape = 0; shape < 5; ++shape) {
        const int tws[5] = {128, 32, 8, 2, 1}, ths[5] = {1, 4, 8, 2, 1}, tbs[5] = {1, 1, 2, 32, 128};
        const int tw = tws[shape], th = ths[shape], tb = tbs[shape];
        const int C = 256, W = tw * 8, H = th * 8, B = tb * 4;
        CUtensorMap m;
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)step * C * 2, (cuuint64_t)step * W * step * C * 2, (cuuint64_t)H * step * W * step * C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode4d(b2) failed %d\n", (int)r); continue; }
        const int rows = 128, chunk = rows * 128;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          probe<<<sms * ctas_per_sm, 64, STAGES * chunk + 1024>>>(m, buf, 3, rows, iters, 0, chunk, tw, th);
          cudaEventRecord(e1); CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (rep == 1) {
            const double tot = (double)sms * ctas_per_sm * iters * chunk;
            printf("4D unit box {64,%3d,%d,%3d} pitch x%d: %7.1f GB/s total, %6.1f GB/s per SM, %5.2f ns per row per SM\n", tw, th, tb,
                   step, tot / ms / 1e6, tot / ms / 1e6 / sms, ms * 1e6 / ((double)iters * rows));
          }
        }
      }
    }
    // (c) 1-D bulk copies
    for (int chunk : {4096, 16384, 32768}) {
      CUtensorMap m = {};
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        probe<<<sms * ctas_per_sm, 64, STAGES * chunk + 1024>>>(m, buf, 2, 0, iters, (long long)(bytes / 128) - chunk / 128, chunk, 0, 0);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 1) {
          const double tot = (double)sms * ctas_per_sm * iters * chunk;
          printf("1D bulk chunk %6d B      : %7.1f GB/s total, %6.1f GB/s per SM\n", chunk, tot / ms / 1e6, tot / ms / 1e6 / sms);
        }
      }
    }
  }
  return 0;
}
