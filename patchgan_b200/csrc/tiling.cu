// Inference tiling on the device: n_crop gather and build_mask overlap-average / threshold / argmax.
// Reference: infer.py:14-34 (n_crop) and infer.py:37-68 (build_mask), including the `j * ncropsy + i`
// crop-index quirk at infer.py:32,57.  double accumulation like numpy, so argmax is bit-exact.
#include "common.cuh"

namespace pg {

__device__ __forceinline__ int crop_start(int idx, int eff, int size, int extent) {
  int s = idx * eff;
  const int over = s + size - extent;
  if (over > 0) s -= over;
  return s;
}

__global__ void ncrop_kernel(const float* __restrict__ image, float* __restrict__ crops, int C, int H, int W, int size,
                             int eff, int ncy, int ncx) {
  const int k = blockIdx.y;  // destination crop index
  // the reference writes crops[j*ncy + i] for j in range(ncy), i in range(ncx): the LAST writer wins
  int sj = -1, si = -1;
  for (int j = ncy - 1; j >= 0 && sj < 0; --j) {
    const int i = k - j * ncy;
    if (i >= 0 && i < ncx) { sj = j; si = i; }
  }
  const long long per = (long long)C * size * size;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (sj >= 0) {
      const int x = (int)(e % size);
      const int y = (int)((e / size) % size);
      const int c = (int)(e / ((long long)size * size));
      const int sy = crop_start(sj, eff, size, H), sx = crop_start(si, eff, size, W);
      v = image[((long long)c * H + sy + y) * W + sx + x];
    }
    crops[(long long)k * per + e] = v;
  }
}

__global__ void build_mask_kernel(const float* __restrict__ masks, float* __restrict__ mask_out,
                                  int* __restrict__ argmax_out, int C, int H, int W, int size, int eff, int ncy,
                                  int ncx, float threshold) {
  const long long HW = (long long)H * W;
  for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < HW;
       px += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(px % W), y = (int)(px / W);
    double best = 0;
    int besti = 0;
    for (int c = 0; c < C; ++c) {
      double s = 0, cnt = 0;
      for (int j = 0; j < ncy; ++j) {
        const int sy = crop_start(j, eff, size, H);
        if (y < sy || y >= sy + size) continue;
        for (int i = 0; i < ncx; ++i) {
          const int sx = crop_start(i, eff, size, W);
          if (x < sx || x >= sx + size) continue;
          s += (double)masks[(((long long)(j * ncy + i) * C + c) * size + (y - sy)) * size + (x - sx)];
          cnt += 1;
        }
      }
      double v = s / cnt;
      if (threshold > 0.f) v = v >= (double)threshold ? 1.0 : 0.0;
      mask_out[(long long)c * HW + px] = (float)v;
      if (c == 0 || v > best) { best = v; besti = c; }
    }
    if (argmax_out != nullptr) argmax_out[px] = besti;
  }
}

}  // namespace pg
using namespace pg;

extern "C" int pg_ncrop(const float* image, float* crops, int32_t C, int32_t H, int32_t W, int32_t size, int32_t eff,
                        int32_t ncy, int32_t ncx, void* stream) {
  PG_REQUIRE(size <= H && size <= W && eff > 0, "pg_ncrop: crop %d larger than image %dx%d", size, H, W);
  const long long per = (long long)C * size * size;
  long long nb = (per + 255) / 256;
  if (nb > 1024) nb = 1024;
  dim3 grid((unsigned)nb, (unsigned)(ncx * ncy));
  ncrop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, crops, C, H, W, size, eff, ncy, ncx);
  return check_launch("ncrop_kernel");
}

extern "C" int pg_build_mask(const float* masks, float* mask_out, int32_t* argmax_out, int32_t C, int32_t H, int32_t W,
                             int32_t size, int32_t eff, int32_t ncy, int32_t ncx, float threshold, void* stream) {
  PG_REQUIRE(size <= H && size <= W && eff > 0, "pg_build_mask: crop %d larger than image %dx%d", size, H, W);
  const long long HW = (long long)H * W;
  long long nb = (HW + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (nb > cap) nb = cap;
  build_mask_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(masks, mask_out, argmax_out, C, H, W, size, eff, ncy,
                                                                   ncx, threshold);
  return check_launch("build_mask_kernel");
}
