// Device input pipeline: what COCOStuffDataset.__getitem__ does per sample on the CPU (io.py:38-58) -- /255, the uint8
// label shift +1, torchvision Resize (bilinear, align_corners=False, no antialias) of the stacked (image, labels) tensor,
// optional flips, one binary mask per requested label -- done on the GPU from the RAW uint8 image and label map, so that a
// step uploads 4 bytes per pixel instead of 4 * (3 + L) * 4.  HBM-bound: one thread per output pixel.
//
// Arithmetic mirrors ATen's upsample_bilinear2d (float): source index = max(scale * (dst + 0.5) - 0.5, 0), scale = in / out,
// lambda1 = src - floor(src), lambda0 = 1 - lambda1, value = l0y * (l0x * p00 + l1x * p01) + l1y * (l0x * p10 + l1x * p11),
// every product and sum rounded to fp32 on its own (no FMA contraction): the masks test `labels == label` on the
// interpolated float, so the last bit matters.  oracle/io_oracle.py restates the same and is pinned to the live reference.
#include "common.cuh"

namespace pg {

struct PrepP {
  const unsigned char* img;    // [B][3][Hs][Ws]
  const unsigned char* lab;    // [B][Hs][Ws]
  const unsigned char* flips;  // [B]: bit 0 horizontal, bit 1 vertical (applied after the resize), or null
  float* x;                    // [B][3][Ho][Wo]
  float* y;                    // [B][L][Ho][Wo]
  int labels[16];
  int L, B, Hs, Ws, Ho, Wo, resize;
  float sy, sx;                // Hs / Ho, Ws / Wo in fp32 (ATen: area_pixel_compute_scale)
};

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  real = fmaxf(real, 0.f);
  i0 = (int)floorf(real);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(__fsub_rn(real, (float)i0), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

__device__ __forceinline__ float lerp2(float p00, float p01, float p10, float p11, float lx0, float lx1, float ly0, float ly1) {
  const float t0 = __fadd_rn(__fmul_rn(p00, lx0), __fmul_rn(p01, lx1));
  const float t1 = __fadd_rn(__fmul_rn(p10, lx0), __fmul_rn(p11, lx1));
  return __fadd_rn(__fmul_rn(t0, ly0), __fmul_rn(t1, ly1));
}

__global__ void __launch_bounds__(256) prep_batch_u8_kernel(const PrepP p) {
  const long long hw = (long long)p.Ho * p.Wo;
  const long long total = (long long)p.B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int r = (int)(i - (long long)b * hw);
    int oy = r / p.Wo, ox = r - oy * p.Wo;
    // flips act on the resized tensor: output (oy, ox) shows resized (H-1-oy, W-1-ox)
    const unsigned char fl = p.flips != nullptr ? p.flips[b] : 0;
    const int ry = (fl & 2) ? p.Ho - 1 - oy : oy, rx = (fl & 1) ? p.Wo - 1 - ox : ox;
    const unsigned char* im = p.img + (long long)b * 3 * p.Hs * p.Ws;
    const unsigned char* lb = p.lab + (long long)b * p.Hs * p.Ws;
    float v[3], lv;
    if (!p.resize) {
      const long long s = (long long)ry * p.Ws + rx;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn((float)im[(long long)c * p.Hs * p.Ws + s], 255.f);
      lv = (float)(unsigned char)(lb[s] + 1);                         // uint8 arithmetic: 255 + 1 wraps to 0 (io.py:43)
    } else {
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      src_index(p.sy, ry, p.Hs, y0, y1, ly0, ly1);
      src_index(p.sx, rx, p.Ws, x0, x1, lx0, lx1);
      const long long a00 = (long long)y0 * p.Ws + x0, a01 = (long long)y0 * p.Ws + x1, a10 = (long long)y1 * p.Ws + x0,
                      a11 = (long long)y1 * p.Ws + x1;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned char* pl = im + (long long)c * p.Hs * p.Ws;
        v[c] = lerp2(__fdiv_rn((float)pl[a00], 255.f), __fdiv_rn((float)pl[a01], 255.f), __fdiv_rn((float)pl[a10], 255.f),
                     __fdiv_rn((float)pl[a11], 255.f), lx0, lx1, ly0, ly1);
      }
      lv = lerp2((float)(unsigned char)(lb[a00] + 1), (float)(unsigned char)(lb[a01] + 1), (float)(unsigned char)(lb[a10] + 1),
                 (float)(unsigned char)(lb[a11] + 1), lx0, lx1, ly0, ly1);
    }
    float* xo = p.x + (long long)b * 3 * hw + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) xo[(long long)c * hw] = v[c];
    float* yo = p.y + (long long)b * p.L * hw + r;
    for (int l = 0; l < p.L; ++l) yo[(long long)l * hw] = lv == (float)p.labels[l] ? 1.f : 0.f;
  }
}

}  // namespace pg
using namespace pg;

extern "C" int pg_prep_batch_u8(const uint8_t* img, const uint8_t* lab, const int32_t* labels, int32_t nlabels, int32_t B,
                                int32_t Hs, int32_t Ws, int32_t Ho, int32_t Wo, const uint8_t* flips, float* x, float* y,
                                void* stream) {
  PG_REQUIRE(img && lab && labels && x && y, "pg_prep_batch_u8: NULL pointer");
  PG_REQUIRE(nlabels >= 1 && nlabels <= 16, "pg_prep_batch_u8: 1..16 labels, got %d", nlabels);
  PG_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, "pg_prep_batch_u8: empty extent");
  PrepP p;
  p.img = img; p.lab = lab; p.flips = flips; p.x = x; p.y = y;
  for (int i = 0; i < 16; ++i) p.labels[i] = i < nlabels ? labels[i] : -1;      // `labels` is a HOST array
  p.L = nlabels; p.B = B; p.Hs = Hs; p.Ws = Ws; p.Ho = Ho; p.Wo = Wo;
  p.resize = (Hs != Ho || Ws != Wo) ? 1 : 0;
  p.sy = (float)Hs / (float)Ho;
  p.sx = (float)Ws / (float)Wo;
  const long long total = (long long)B * Ho * Wo;
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  prep_batch_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("prep_batch_u8_kernel");
}
