// Layout kernels: NCHW float <-> NHWC bf16 channel slices, weight packing.  HBM-bound byte movers.
// Reference: trainer.py:55-66,96-99 (.to(device) + torch.cat that feed the two nets).
#include "common.cuh"

namespace pg {

// one thread per pixel; reads are coalesced per channel plane, each thread writes its C channels
__global__ void pack_nchw_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, int C, long long HW,
                                 int ld, int c_off, int dt) {
  const int b = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    unsigned short* o = dst + ((long long)b * HW + p) * ld + c_off;
    const float* s = src + (long long)b * C * HW + p;
    for (int c = 0; c < C; ++c) o[c] = to16(s[(long long)c * HW], dt);
  }
}

__global__ void unpack_nhwc_kernel(const void* __restrict__ src, int src_f32, float* __restrict__ dst, int C,
                                   long long HW, int ld, int c_off) {
  const int b = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float* o = dst + (long long)b * C * HW + p;
    const long long base = ((long long)b * HW + p) * ld + c_off;
    for (int c = 0; c < C; ++c) {
      float v = src_f32 == PG_F32 ? reinterpret_cast<const float*>(src)[base + c]
                                  : from16(reinterpret_cast<const unsigned short*>(src)[base + c], src_f32);
      o[(long long)c * HW] = v;
    }
  }
}

__global__ void copy_f32_to_bf16_slice_kernel(const float* __restrict__ src, int lds, unsigned short* __restrict__ dst,
                                              int ldd, int c_off, int C, long long npix, int dt) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
       p += (long long)gridDim.x * blockDim.x) {
    for (int c = 0; c < C; ++c) dst[p * ldd + c_off + c] = to16(src[p * lds + c], dt);
  }
}

// dst[n][t][cp] over the padded extents; gathers from the fp32 reference layout
__global__ void pack_weight_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, int N, int Np, int C1,
                                   int C1p, int C2, int C2p, long long sn, long long sc, int flip, int dt) {
  const int Cp = C1p + C2p;
  const long long total = (long long)Np * 16 * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % Cp);
    const int t = (int)((i / Cp) % 16);
    const int n = (int)(i / ((long long)Cp * 16));
    int c = -1;
    if (cp < C1p) { if (cp < C1) c = cp; }
    else { if (cp - C1p < C2) c = C1 + (cp - C1p); }
    float v = 0.f;
    if (n < N && c >= 0) v = src[n * sn + c * sc + (flip ? 15 - t : t)];
    dst[i] = to16(v, dt);
  }
}

// Tiled version: one block moves an 8 (n) x 32 (c) x 16 (tap) brick.  Reads follow the source's contiguous axis
// (the 16 taps of one (n, c) are always contiguous; the next-fastest source axis is n when sn == 16, else c),
// writes are 32 consecutive packed channels (64 B) per (n, tap).
__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const float* __restrict__ src,
                                                               unsigned short* __restrict__ dst, int N, int Np, int C1,
                                                               int C1p, int C2, int C2p, long long sn, long long sc,
                                                               int flip, int dt) {
  __shared__ float tile[8][32][17];
  const int Cp = C1p + C2p;
  const int n0 = blockIdx.y * 8, cp0 = blockIdx.x * 32;
  // ---- load: thread -> (t fastest, then the source-contiguous one of (n, c))
  for (int e = threadIdx.x; e < 8 * 32 * 16; e += 256) {
    const int t = e & 15;
    int nl, cl;
    if (sn == 16) { nl = (e >> 4) & 7; cl = e >> 7; } else { cl = (e >> 4) & 31; nl = e >> 9; }
    const int n = n0 + nl, cp = cp0 + cl;
    int c = -1;
    if (cp < C1p) { if (cp < C1) c = cp; }
    else if (cp < Cp) { if (cp - C1p < C2) c = C1 + (cp - C1p); }
    float v = 0.f;
    if (n < N && c >= 0) v = src[n * sn + c * sc + t];
    tile[nl][cl][t] = v;
  }
  __syncthreads();
  // ---- store: 32 consecutive packed channels per (n, tap)
  for (int e = threadIdx.x; e < 8 * 32 * 16; e += 256) {
    const int cl = e & 31, t = (e >> 5) & 15, nl = e >> 9;
    const int n = n0 + nl, cp = cp0 + cl;
    if (n < Np && cp < Cp) dst[((long long)n * 16 + t) * Cp + cp] = to16(tile[nl][cl][flip ? 15 - t : t], dt);
  }
}

// Two NCHW float sources -> ONE full NHWC row per pixel: [src1 channels | src2 channels | zeros up to ld], written as
// 16-byte vectors, optionally also as a bf16 twin.  One thread per pixel (plane reads coalesced across the warp).
template <int LD>
__global__ void __launch_bounds__(256) pack2_rows_kernel(const float* __restrict__ s1, int C1, const float* __restrict__ s2,
                                                        int C2, unsigned short* __restrict__ dst,
                                                        unsigned short* __restrict__ dst2, long long HW, int dt) {
  const int b = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    float v[LD];
#pragma unroll
    for (int c = 0; c < LD; ++c) {
      float x = 0.f;
      if (c < C1) x = s1[((long long)b * C1 + c) * HW + p];
      else if (c - C1 < C2) x = s2[((long long)b * C2 + (c - C1)) * HW + p];
      v[c] = x;
    }
    const long long o = ((long long)b * HW + p) * LD;
#pragma unroll
    for (int c = 0; c < LD; c += 8) {
      *reinterpret_cast<uint4*>(dst + o + c) = pack8dt(v + c, dt);
      if (dst2 != nullptr) *reinterpret_cast<uint4*>(dst2 + o + c) = pack8(v + c);
    }
  }
}

// Channel-major stride-2 im2col of a first layer's input (Conv2d k4 s2 p1): for every output pixel o = (b, oy, ox) and
// source channel c:  dst[o*K + (k_off + c)*16 + kh*4 + kw] = src[b, c, 2oy-1+kh, 2ox-1+kw]   (0 outside the image).
// With k = c*16 + tap the row IS the layer's weight layout (Cout, Cin, 4, 4) flattened, so the first layers of both
// networks (3 / 4 real input channels: 16x channel padding as implicit GEMMs) run as one dense pointwise GEMM with
// K = 16*Cin and their weight-gradients land in the reference layout without a transpose.  The source is addressed
// with explicit element strides, so NCHW planes and a channel of an NHWC tensor both work.
// (src_b / C_b / dst_b_off: optional second source whose channels follow the first's and are written only to the rows
//  dst_b_off elements further on, while the first source's channels are written to BOTH places -- the discriminator batch
//  [cat(x, .) ; cat(x, y)] of trainer.py:65,96 in one launch.)
__global__ void __launch_bounds__(256) im2col_s2_kernel(const float* __restrict__ src, long long sb, long long sc, long long sy,
                                                       long long sx, int C, int H, int W, int Ho, int Wo,
                                                       unsigned short* __restrict__ dst, unsigned short* __restrict__ dst2,
                                                       int K, int k_off, int dt, long long per_c,
                                                       const float* __restrict__ src_b, long long sb_b, long long sc_b,
                                                       long long dst_b_off, int ctot) {
  // thread = (output pixel, channel), channel fastest: the threads of a warp write consecutive 32-byte segments of
  // consecutive rows (a contiguous kilobyte per store instruction when K = 16 * channels); the 4x-overlapping window reads
  // come out of L1 / L2 (the input is 1/8 of the bytes written)
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long i;
  int c;
  if (t < 0x7fffffffLL) {          // (32-bit division when it fits)
    const unsigned q = (unsigned)t / (unsigned)ctot;
    i = q; c = (int)((unsigned)t - q * (unsigned)ctot);
  } else {
    i = t / ctot; c = (int)(t - i * ctot);
  }
  if (i >= per_c) return;
  int ox, oy, b;
  split_xyb(i, Wo, Ho, ox, oy, b);
  const bool second = c >= C;
  const float* plane = second ? src_b + b * sb_b + (c - C) * sc_b : src + b * sb + c * sc;
  float v[16];
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int iy = 2 * oy - 1 + kh;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int ix = 2 * ox - 1 + kw;
      v[kh * 4 + kw] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(plane + iy * sy + ix * sx) : 0.f;
    }
  }
  const long long o = i * K + (long long)(k_off + c) * 16;
  const uint4 lo = pack8dt(v, dt), hi = pack8dt(v + 8, dt);
  if (!second) {
    *reinterpret_cast<uint4*>(dst + o) = lo;
    *reinterpret_cast<uint4*>(dst + o + 8) = hi;
  }
  if (second || dst_b_off != 0) {
    *reinterpret_cast<uint4*>(dst + dst_b_off + o) = lo;
    *reinterpret_cast<uint4*>(dst + dst_b_off + o + 8) = hi;
  }
  if (dst2 != nullptr) {
    const uint4 lo2 = pack8(v), hi2 = pack8(v + 8);
    if (!second) {
      *reinterpret_cast<uint4*>(dst2 + o) = lo2;
      *reinterpret_cast<uint4*>(dst2 + o + 8) = hi2;
    }
    if (second || dst_b_off != 0) {
      *reinterpret_cast<uint4*>(dst2 + dst_b_off + o) = lo2;
      *reinterpret_cast<uint4*>(dst2 + dst_b_off + o + 8) = hi2;
    }
  }
}

// Job lookup of the multi-tensor kernels: the job whose [tile_begin, next tile_begin) holds `tile`.  The first njobs threads
// each test one job (ONE global round trip; a per-block binary search by thread 0 was four dependent ones and made these
// short blocks latency-bound), then sizeof(Job)/4 threads copy the descriptor to shared memory.  Ends with __syncthreads().
template <typename Job>
__device__ __forceinline__ void find_job(const Job* __restrict__ jobs, int njobs, int tile, Job* sh_job, int* sh_idx) {
  for (int i = threadIdx.x; i < njobs; i += blockDim.x) {
    const int tb = jobs[i].tile_begin;
    const int te = i + 1 < njobs ? jobs[i + 1].tile_begin : 0x7fffffff;
    if (tb <= tile && tile < te) { sh_idx[0] = i; sh_idx[1] = te; }      // sh_idx[1]: first tile of the next job
  }
  __syncthreads();
  const int idx = sh_idx[0];
  if (threadIdx.x < sizeof(Job) / 4)
    reinterpret_cast<int*>(sh_job)[threadIdx.x] = reinterpret_cast<const int*>(jobs + idx)[threadIdx.x];
  __syncthreads();
}

// Tap-major weight-gradient scratch S[tap][Ns][Cs] -> reference layout dst[n*ld_n + c*16 + tap] for every layer of a
// network in ONE launch (jobs as in pack_weight_multi_kernel; one tile = one n, 32 channels, 16 taps; a block walks
// GF_TILES consecutive tiles with all their loads in flight before the first store).
struct GradJob {               // mirrors PgGradJob
  const float* S;
  float* dst;
  long long ld_n;
  int N, C, Ns, Cs;
  int tile_begin, ctiles;
};

constexpr int GF_TILES = 8;

__global__ void __launch_bounds__(256) grad_finalize_multi_kernel(const GradJob* __restrict__ jobs, int njobs, int total_tiles) {
  __shared__ float tile[GF_TILES][16][33];
  __shared__ GradJob job;
  __shared__ int job_idx[2];
  const int tile0 = blockIdx.x * GF_TILES;
  int ntiles = total_tiles - tile0;
  if (ntiles > GF_TILES) ntiles = GF_TILES;
  find_job(jobs, njobs, tile0, &job, job_idx);
  // (the first version decoded tile / tap / channel from a flat element index, with an integer division per 4-byte element:
  //  ncu showed 72 % issue-slot utilisation at 10 % of the DRAM bandwidth.  Now: lane = channel, warp = taps w and w + 8,
  //  one division per tile.)
  const int cl = threadIdx.x & 31, tq = threadIdx.x >> 5;
  int done = 0;
  while (done < ntiles) {
    // tiles [done, upto) of this block belong to the current job
    const int job_end = job_idx[1];
    int upto = job_end < total_tiles ? job_end - tile0 : ntiles;
    if (upto > ntiles) upto = ntiles;
    const long long tstride = (long long)job.Ns * job.Cs;
    const int first = tile0 + done - job.tile_begin;
    int n = first / job.ctiles, ct = first - n * job.ctiles;
    for (int k = done; k < upto; ++k) {
      const int c = ct * 32 + cl;
      const float* sp = job.S + (long long)n * job.Cs + c;
      const bool ok = c < job.C;
      tile[k][tq][cl] = ok ? sp[tq * tstride] : 0.f;
      tile[k][tq + 8][cl] = ok ? sp[(tq + 8) * tstride] : 0.f;
      if (++ct == job.ctiles) { ct = 0; ++n; }
    }
    __syncthreads();
    n = first / job.ctiles; ct = first - n * job.ctiles;
    for (int k = done; k < upto; ++k) {
      // dst[n][c0 .. c0+32)[16 taps] is 512 contiguous floats: two per thread
      float* dp = job.dst + (long long)n * job.ld_n + (long long)ct * 32 * 16;
      const int cmax = job.C - ct * 32;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = threadIdx.x + h * 256;
        if ((i >> 4) < cmax) dp[i] = tile[k][i & 15][i >> 4];
      }
      if (++ct == job.ctiles) { ct = 0; ++n; }
    }
    done = upto;
    if (done < ntiles) {           // the block straddles two tensors (block-uniform)
      __syncthreads();
      find_job(jobs, njobs, tile0 + done, &job, job_idx);
    }
  }
}

// All weight tensors of a network in ONE launch: blockIdx.x walks a concatenated list of 8x32x16 bricks.
struct PackJob {              // mirrors PgPackJob
  const float* src;
  unsigned short* dst;
  long long sn, sc;
  int N, Np, C1, C1p, C2, C2p, flip, dt;
  int tile_begin, ctiles;
};

__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[8][32][17];
  __shared__ PackJob job;
  __shared__ int job_idx[2];
  find_job(jobs, njobs, (int)blockIdx.x, &job, job_idx);
  const int local = blockIdx.x - job.tile_begin;
  if (job.flip == 2) {
    // flat job: dst[i] = convert(src[i]), i < sn (first-layer [N][Cin*16] and tap-product [Cin][16] operand copies,
    // which ARE the master layout); one block converts 4096 elements
    const long long base = (long long)local * 4096;
    for (int e = threadIdx.x; e < 4096; e += 256)
      if (base + e < job.sn) job.dst[base + e] = to16(job.src[base + e], job.dt);
    return;
  }
  const int n0 = (local / job.ctiles) * 8, cp0 = (local % job.ctiles) * 32;
  const int Cp = job.C1p + job.C2p;
  // Every thread owns ONE tap t and walks (n, c) with loop-invariant decoding (the flat-index version spent 65 % of the
  // issue slots on index arithmetic at 11 % of the DRAM bandwidth).
  const int t = threadIdx.x & 15, q = threadIdx.x >> 4;            // q = 0..15
  auto src_channel = [&](int cp) {                                 // padded channel -> real channel or -1
    if (cp < job.C1p) return cp < job.C1 ? cp : -1;
    if (cp < Cp) return cp - job.C1p < job.C2 ? job.C1 + (cp - job.C1p) : -1;
    return -1;
  };
  if (job.sn == 16) {
    // ConvTranspose2d master layout [Cin][Cout][16]: n fastest after the tap.  nl = q & 7, channels q >> 3 + 2 i
    const int nl = q & 7, n = n0 + nl;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int cl = (q >> 3) + 2 * i;
      const int c = src_channel(cp0 + cl);
      tile[nl][cl][t] = (n < job.N && c >= 0) ? job.src[(long long)n * 16 + (long long)c * job.sc + t] : 0.f;
    }
  } else {
    // Conv2d master layout [Cout][Cin][16]: c fastest after the tap.  channels q and q + 16, all 8 n
    const int ca = src_channel(cp0 + q), cb = src_channel(cp0 + q + 16);
#pragma unroll
    for (int nl = 0; nl < 8; ++nl) {
      const int n = n0 + nl;
      const float* sp = job.src + (long long)n * job.sn + t;
      tile[nl][q][t] = (n < job.N && ca >= 0) ? sp[(long long)ca * job.sc] : 0.f;
      tile[nl][q + 16][t] = (n < job.N && cb >= 0) ? sp[(long long)cb * job.sc] : 0.f;
    }
  }
  __syncthreads();
  // two channels (4 bytes) per store; Cp is a multiple of 16, so a pair never straddles the padded row end.
  // thread = (channel pair t', tap q'), loop over the 8 n
  {
    const int cl = (threadIdx.x & 15) * 2, tt = threadIdx.x >> 4;
    const int cp = cp0 + cl;
    const int ts = job.flip ? 15 - tt : tt;
    if (cp < Cp) {
#pragma unroll
      for (int nl = 0; nl < 8; ++nl) {
        const int n = n0 + nl;
        if (n < job.Np) {
          const unsigned lo = to16(tile[nl][cl][ts], job.dt), hi = to16(tile[nl][cl + 1][ts], job.dt);
          *reinterpret_cast<unsigned*>(job.dst + ((long long)n * 16 + tt) * Cp + cp) = lo | (hi << 16);
        }
      }
    }
  }
}

static unsigned grid1d(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = 16LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace pg
using namespace pg;

extern "C" int pg_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t B, int32_t C, int32_t H, int32_t W,
                                             int32_t ld, int32_t c_off, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(dst_dtype == PG_BF16 || dst_dtype == PG_F16, "pg_pack_nchw: dst_dtype must be 16-bit");
  PG_REQUIRE(B > 0 && C > 0 && c_off >= 0 && c_off + C <= ld, "pg_pack_nchw: bad C=%d c_off=%d ld=%d", C, c_off, ld);
  const long long HW = (long long)H * W;
  dim3 grid(grid1d(HW, 256), B);
  pack_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (unsigned short*)dst, C, HW, ld, c_off, dst_dtype);
  return check_launch("pack_nchw_kernel");
}

extern "C" int pg_unpack_nhwc_to_nchw_f32(const void* src, int32_t src_f32, float* dst, int32_t B, int32_t C,
                                          int32_t H, int32_t W, int32_t ld, int32_t c_off, void* stream) {
  PG_REQUIRE(B > 0 && C > 0 && c_off >= 0 && c_off + C <= ld, "pg_unpack_nhwc: bad C=%d c_off=%d ld=%d", C, c_off, ld);
  const long long HW = (long long)H * W;
  dim3 grid(grid1d(HW, 256), B);
  unpack_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_f32, dst, C, HW, ld, c_off);
  return check_launch("unpack_nhwc_kernel");
}

extern "C" int pg_copy_f32_to_bf16_slice(const float* src, int32_t lds, void* dst, int32_t ldd, int32_t c_off,
                                         int32_t C, int64_t npix, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(C > 0 && C <= lds && c_off + C <= ldd, "pg_copy_f32_to_bf16_slice: bad C=%d", C);
  copy_f32_to_bf16_slice_kernel<<<grid1d(npix, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, (unsigned short*)dst, ldd,
                                                                                    c_off, C, npix, dst_dtype);
  return check_launch("copy_f32_to_bf16_slice_kernel");
}

extern "C" int pg_pack_weight(const float* src, void* dst, int32_t N, int32_t Np, int32_t C1, int32_t C1p, int32_t C2,
                              int32_t C2p, int64_t sn, int64_t sc, int32_t flip, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(N <= Np && C1 <= C1p && C2 <= C2p, "pg_pack_weight: padded extents smaller than real ones");
  if (sn == 16 || sc == 16) {
    dim3 grid((unsigned)((C1p + C2p + 31) / 32), (unsigned)((Np + 7) / 8));
    pack_weight_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (unsigned short*)dst, N, Np, C1, C1p, C2, C2p,
                                                                     sn, sc, flip, dst_dtype);
    return check_launch("pack_weight_tiled_kernel");
  }
  const long long total = (long long)Np * 16 * (C1p + C2p);
  pack_weight_kernel<<<grid1d(total, 256), 256, 0, (cudaStream_t)stream>>>(src, (unsigned short*)dst, N, Np, C1, C1p, C2,
                                                                          C2p, sn, sc, flip, dst_dtype);
  return check_launch("pack_weight_kernel");
}

extern "C" int pg_pack2_nchw_rows(const float* src1, int32_t C1, const float* src2, int32_t C2, void* dst, void* dst2,
                                  int32_t B, int32_t H, int32_t W, int32_t ld, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(dst_dtype == PG_BF16 || dst_dtype == PG_F16, "pg_pack2_nchw_rows: dst_dtype must be 16-bit");
  PG_REQUIRE(C1 > 0 && C2 >= 0 && C1 + C2 <= ld && (ld == 16 || ld == 32), "pg_pack2_nchw_rows: C1=%d C2=%d ld=%d", C1, C2,
             ld);
  PG_REQUIRE(C2 == 0 || src2 != nullptr, "pg_pack2_nchw_rows: src2 is NULL");
  const long long HW = (long long)H * W;
  dim3 grid(grid1d(HW, 256), B);
  if (ld == 16)
    pack2_rows_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>(src1, C1, src2, C2, (unsigned short*)dst,
                                                                 (unsigned short*)dst2, HW, dst_dtype);
  else
    pack2_rows_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>(src1, C1, src2, C2, (unsigned short*)dst,
                                                                 (unsigned short*)dst2, HW, dst_dtype);
  return check_launch("pack2_rows_kernel");
}

extern "C" int pg_im2col_s2(const float* src, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int32_t C, int32_t B, int32_t H,
                            int32_t W, void* dst, void* dst2, int32_t K, int32_t k_off, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(src && dst && C > 0 && B > 0 && H > 1 && W > 1 && (H % 2) == 0 && (W % 2) == 0, "pg_im2col_s2: bad extents");
  PG_REQUIRE(K % 16 == 0 && k_off >= 0 && (k_off + C) * 16 <= K, "pg_im2col_s2: K=%d k_off=%d C=%d", K, k_off, C);
  PG_REQUIRE(dst_dtype == PG_BF16 || dst_dtype == PG_F16, "pg_im2col_s2: dst_dtype must be 16-bit");
  PG_REQUIRE((((uintptr_t)dst | (uintptr_t)dst2) & 15) == 0, "pg_im2col_s2: dst must be 16-byte aligned");
  const int Ho = H / 2, Wo = W / 2;
  const long long per_c = (long long)B * Ho * Wo;
  dim3 grid((unsigned)((per_c * C + 255) / 256));
  im2col_s2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sb, sc, sy, sx, C, H, W, Ho, Wo, (unsigned short*)dst,
                                                         (unsigned short*)dst2, K, k_off, dst_dtype, per_c, nullptr, 0, 0, 0, C);
  return check_launch("im2col_s2_kernel");
}

extern "C" int pg_im2col_s2_pair(const float* x, int32_t Cx, const float* y, int32_t Cy, int32_t B, int32_t H, int32_t W, void* dst,
                                 void* dst2, int32_t K, int32_t dst_dtype, void* stream) {
  PG_REQUIRE(x && y && dst && Cx > 0 && Cy > 0 && B > 0 && H > 1 && W > 1 && (H % 2) == 0 && (W % 2) == 0,
             "pg_im2col_s2_pair: bad extents");
  PG_REQUIRE(K % 16 == 0 && (Cx + Cy) * 16 <= K, "pg_im2col_s2_pair: K=%d Cx=%d Cy=%d", K, Cx, Cy);
  PG_REQUIRE(dst_dtype == PG_BF16 || dst_dtype == PG_F16, "pg_im2col_s2_pair: dst_dtype must be 16-bit");
  PG_REQUIRE((((uintptr_t)dst | (uintptr_t)dst2) & 15) == 0, "pg_im2col_s2_pair: dst must be 16-byte aligned");
  const int Ho = H / 2, Wo = W / 2;
  const long long per_c = (long long)B * Ho * Wo, HW = (long long)H * W;
  dim3 grid((unsigned)((per_c * (Cx + Cy) + 255) / 256));
  im2col_s2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, Cx * HW, HW, W, 1, Cx, H, W, Ho, Wo, (unsigned short*)dst,
                                                         (unsigned short*)dst2, K, 0, dst_dtype, per_c, y, Cy * HW, HW,
                                                         per_c * K, Cx + Cy);
  return check_launch("im2col_s2_kernel");
}

extern "C" int pg_grad_finalize_multi(const PgGradJob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream) {
  static_assert(sizeof(PgGradJob) == sizeof(GradJob), "PgGradJob layout");
  PG_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_tiles > 0, "pg_grad_finalize_multi: empty job list");
  grad_finalize_multi_kernel<<<(unsigned)((total_tiles + GF_TILES - 1) / GF_TILES), 256, 0, (cudaStream_t)stream>>>(
      (const GradJob*)jobs_dev, njobs, total_tiles);
  return check_launch("grad_finalize_multi_kernel");
}

extern "C" int pg_pack_weights_multi(const PgPackJob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream) {
  static_assert(sizeof(PgPackJob) == sizeof(PackJob), "PgPackJob layout");
  PG_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_tiles > 0, "pg_pack_weights_multi: empty job list");
  pack_weight_multi_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs_dev, njobs);
  return check_launch("pack_weight_multi_kernel");
}
