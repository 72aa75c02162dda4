// Layers with ONE real channel on one side
//   * generator output layer   ConvTranspose2d(2nf -> output_nc=1)      (unet.py:106-107)
//   * discriminator last layer  Conv2d(8ndf -> 1, stride 1)              (disc.py:45)
//   * the data-gradient of the discriminator's first layer w.r.t. the generated mask channel (trainer.py:84-89)
// and their data- / weight-gradients.  As 4x4 convolutions these are GEMMs with N = 1 (or K-per-tap = 1).
//
// This file holds (a) the cheap halves of the formulation the engine uses -- pointwise products over the 16 taps on the
// tensor cores (PG_CONV1X1 in conv_tc.cu) with taps_gather / taps_scatter around them (bottom of the file) -- and (b) a
// CUDA-core implementation of the same layers (fewout / fewin / wgrad1, PG_IMPL_SKINNY).  (b) was written first as a
// streaming kernel set; with ~15 FLOP per byte it turned out instruction-bound (~0.5 TB/s) and 1.5-3x slower than (a), and
// is kept as an independently written second implementation that the GPU tests compare with the oracle.
#include "common.cuh"

namespace pg {

// ------------------------------------------------------------------------------------------------------------------
// few-output forward:  out[pix][n_first + n] = act(bias + sum_{tap,c} in[q(pix,tap)][c] * W[n_first + n][tap][c]), n < NV
// A group of CS lanes owns a 4x4 tile of output pixels and splits the channels in 8-wide chunks (16-byte loads, CS
// consecutive lanes read CS*16 contiguous bytes of a pixel); every input pixel of the tile's halo window is loaded
// once and feeds all (output, tap) pairs it belongs to; partial sums are reduced over the group with shuffles.
//   PG_CONV stride 1:  window 7x7, input (oy0-pad+r, ox0-pad+q) feeds output (u,v) with tap (r-u, q-v)
//   PG_CONVT        :  window 4x4 around the 2x2 lattice tile, input (a0-1+r, b0-1+q) feeds output (U,V) with tap
//                      (U+3-2r, V+3-2q)
// ------------------------------------------------------------------------------------------------------------------
struct FewOutP {
  const void* src1;
  const void* src2;
  const void* w;        // packed [Np][16][Ctot], in_dt
  const float* bias;
  void* out;
  int mode, pad, B, Hin, Win, Hout, Wout, C1, C2, ld1, ld2, Ctot, n_first, ldo, act, out_dt, in_dt;
  int cs_log2;          // log2(lanes per tile)
  int tiles_y, tiles_x; // output tiles (4x4) per image
  long long ntiles;
};

template <int ACT>
__device__ __forceinline__ float act_sk(float x) {
  if (ACT == PG_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PG_ACT_LEAKYRELU) return fmaxf(x, 0.2f * x);
  if (ACT == PG_ACT_TANH) return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);
  if (ACT == PG_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
  return x;
}
__device__ __forceinline__ float act_rt(int act, float x) {
  switch (act) {
    case PG_ACT_RELU: return act_sk<PG_ACT_RELU>(x);
    case PG_ACT_LEAKYRELU: return act_sk<PG_ACT_LEAKYRELU>(x);
    case PG_ACT_TANH: return act_sk<PG_ACT_TANH>(x);
    case PG_ACT_SIGMOID: return act_sk<PG_ACT_SIGMOID>(x);
    default: return x;
  }
}

// Tile geometry: PG_CONVT 4x4 outputs <- 4x4 input window; PG_CONV (stride 1) 2x2 outputs <- 5x5 window.
// Shared-memory weights are laid out [tap][half][chunk][4 floats] so the CS lanes of a group read consecutive float4s.
template <int MODE>
__global__ void __launch_bounds__(256) fewout_kernel(const FewOutP p) {
  extern __shared__ float wsm[];
  constexpr int TO = MODE == PG_CONV ? 2 : 4;      // output tile edge
  constexpr int WIN = MODE == PG_CONV ? 5 : 4;     // input window edge
  const int nchunks = p.Ctot >> 3, nch1 = p.C1 >> 3;
  {
    const int total = 16 * p.Ctot;
    const long long base = (long long)p.n_first * 16 * p.Ctot;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int tap = i / p.Ctot, c = i - tap * p.Ctot;
      const float v = from16(reinterpret_cast<const unsigned short*>(p.w)[base + i], p.in_dt);
      wsm[((tap * 2 + ((c >> 2) & 1)) * nchunks + (c >> 3)) * 4 + (c & 3)] = v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CS = 1 << p.cs_log2;
  const int s = lane & (CS - 1), gl = lane >> p.cs_log2;
  const int groups_per_warp = 32 >> p.cs_log2;
  // persistent: the (expensive) weight preload above is paid once per block, warps walk the tiles
  const long long groups_total = (long long)gridDim.x * (blockDim.x >> 5) * groups_per_warp;
  for (long long t0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * groups_per_warp; t0 < p.ntiles; t0 += groups_total) {
  const long long tile = t0 + gl;
  const bool live = tile < p.ntiles;
  int b = 0, ty = 0, tx = 0;
  if (live) {
    tx = (int)(tile % p.tiles_x);
    const long long r = tile / p.tiles_x;
    ty = (int)(r % p.tiles_y);
    b = (int)(r / p.tiles_y);
  }
  const int iy0 = MODE == PG_CONV ? ty * TO - p.pad : ty * 2 - 1;
  const int ix0 = MODE == PG_CONV ? tx * TO - p.pad : tx * 2 - 1;
  float acc[TO * TO];
#pragma unroll
  for (int i = 0; i < TO * TO; ++i) acc[i] = 0.f;
  if (live) {
    for (int j = s; j < nchunks; j += CS) {
      const bool first = j < nch1;
      const char* src = reinterpret_cast<const char*>(first ? p.src1 : p.src2);
      const int ld = first ? p.ld1 : p.ld2;
      const int coff = first ? j * 8 : (j - nch1) * 8;
      uint4 xw[WIN][WIN];
#pragma unroll
      for (int r = 0; r < WIN; ++r) {
        const int iy = iy0 + r;
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          const int ix = ix0 + q;
          xw[r][q] = make_uint4(0, 0, 0, 0);
          if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
            xw[r][q] = __ldg(reinterpret_cast<const uint4*>(src + ((((long long)b * p.Hin + iy) * p.Win + ix) * ld + coff) * 2));
        }
      }
      const float4* wj = reinterpret_cast<const float4*>(wsm) + j;
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          const float4 w0 = wj[((kh * 4 + kw) * 2 + 0) * nchunks];
          const float4 w1 = wj[((kh * 4 + kw) * 2 + 1) * nchunks];
#pragma unroll
          for (int u = 0; u < TO; ++u) {
            // input row of the window that pairs with output row u through tap row kh
            const int r2 = MODE == PG_CONV ? 2 * (u + kh) : u + 3 - kh;
            if (r2 & 1) continue;
#pragma unroll
            for (int v = 0; v < TO; ++v) {
              const int q2 = MODE == PG_CONV ? 2 * (v + kw) : v + 3 - kw;
              if (q2 & 1) continue;
              float xf[8];
              if (p.in_dt == PG_F16) unpack8h(xw[r2 >> 1][q2 >> 1], xf); else unpack8(xw[r2 >> 1][q2 >> 1], xf);
              float a = acc[u * TO + v];
              a = fmaf(xf[0], w0.x, a); a = fmaf(xf[1], w0.y, a); a = fmaf(xf[2], w0.z, a); a = fmaf(xf[3], w0.w, a);
              a = fmaf(xf[4], w1.x, a); a = fmaf(xf[5], w1.y, a); a = fmaf(xf[6], w1.z, a); a = fmaf(xf[7], w1.w, a);
              acc[u * TO + v] = a;
            }
          }
        }
      }
    }
  }
  // reduce over the CS lanes of the group (all lanes end with the full sums)
  for (int o = 1; o < CS; o <<= 1) {
#pragma unroll
    for (int i = 0; i < TO * TO; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  }
  // lanes of the group store the outputs of the tile round-robin
#pragma unroll
  for (int i = 0; i < TO * TO; ++i) {
    if (!live || (i & (CS - 1)) != s) continue;
    const int oy = ty * TO + i / TO, ox = tx * TO + i % TO;
    if (oy >= p.Hout || ox >= p.Wout) continue;
    const long long opix = ((long long)b * p.Hout + oy) * p.Wout + ox;
    float y = acc[i];
    if (p.bias != nullptr) y += __ldg(p.bias + p.n_first);
    y = act_rt(p.act, y);
    const long long o = opix * p.ldo + p.n_first;
    if (p.out_dt == PG_F32) reinterpret_cast<float*>(p.out)[o] = y;
    else reinterpret_cast<unsigned short*>(p.out)[o] = to16(y, p.out_dt);
  }
  }   // tile loop
}

bool conv_fewout_supported(const PgConvDesc* d) {
  if (d->mode != PG_CONV && d->mode != PG_CONVT) return false;
  const int nv = d->n_valid - d->n_first;
  if (nv != 1) return false;
  if (d->mode == PG_CONV && d->stride != 1) return false;
  const int Ctot = d->C1 + d->C2;
  if ((Ctot & 7) || (d->C1 & 7) || Ctot < 8) return false;
  if ((size_t)nv * 16 * Ctot * 4 > 48 * 1024) return false;
  if (d->mode == PG_CONVT && ((d->Hout & 3) || (d->Wout & 3))) return false;
  return true;
}

int conv_fewout(const PgConvDesc* d, const void* src1, const void* src2, const void* w, const float* bias, void* out,
                cudaStream_t stream) {
  FewOutP p;
  p.src1 = src1; p.src2 = src2; p.w = w; p.bias = d->has_bias ? bias : nullptr; p.out = out;
  p.mode = d->mode; p.pad = d->pad; p.B = d->B; p.Hin = d->Hin; p.Win = d->Win; p.Hout = d->Hout; p.Wout = d->Wout;
  p.C1 = d->C1; p.C2 = d->C2; p.ld1 = d->ld1; p.ld2 = d->ld2; p.Ctot = d->C1 + d->C2; p.n_first = d->n_first;
  p.ldo = d->ldo; p.act = d->act; p.out_dt = d->out_f32; p.in_dt = d->in_dtype;
  const int nchunks = p.Ctot >> 3;
  int cs = 0;
  while ((1 << cs) < nchunks && cs < 5) ++cs;
  p.cs_log2 = cs;
  const int to = d->mode == PG_CONV ? 2 : 4;
  p.tiles_y = (d->Hout + to - 1) / to; p.tiles_x = (d->Wout + to - 1) / to;
  p.ntiles = (long long)d->B * p.tiles_y * p.tiles_x;
  const int tiles_per_block = 8 * (32 >> cs);
  long long blocks = (p.ntiles + tiles_per_block - 1) / tiles_per_block;
  const long long cap = (long long)num_sms() * (p.Ctot >= 256 ? 2 : 4);
  if (blocks > cap) blocks = cap;
  const size_t smem = (size_t)16 * p.Ctot * 4;
  if (d->mode == PG_CONV) fewout_kernel<PG_CONV><<<(unsigned)blocks, 256, smem, stream>>>(p);
  else fewout_kernel<PG_CONVT><<<(unsigned)blocks, 256, smem, stream>>>(p);
  return check_launch("fewout_kernel");
}

// ------------------------------------------------------------------------------------------------------------------
// few-input forward (the data-gradient of a 1-output-channel layer): PG_CONV stride 1 or 2
//   out[pix][n] = sum_tap in[q(pix,tap)][0] * W[n][tap][0]
// One thread = 8 output channels of 4 horizontally adjacent output pixels: the 16 tap weights of its channels come from
// shared memory once per 4 pixels, the 1-channel input from L1; the only real traffic is the coalesced 16-byte stores.
// ------------------------------------------------------------------------------------------------------------------
struct FewInP {
  const void* src;
  const void* w;      // packed [N][16][C], in_dt; only channel 0 is used
  void* out;          // 16-bit, out_dt
  void* out2;         // optional bf16 twin
  int stride, pad, B, Hin, Win, Hout, Wout, C, ld, N, ldo, out_dt, in_dt, act;
  int wq;             // pixel quads per output row
  long long nquads;
};

// Block = (range of 64 output channels) x (range of pixel quads): the block's 16 x 64 tap weights go to shared memory
// once (they are 2-byte reads at a 32-byte stride in the packed tensor, so every block must not fetch all N of them),
// then its threads walk the quads.  The 1-channel input row window (3*stride + 4 pixels) is read once per tap row.
template <int STRIDE>
__global__ void __launch_bounds__(256) fewin_kernel(const FewInP p) {
  __shared__ float wsm[16 * 64];   // [tap][64 channels of this block]
  const int nranges = (p.N + 63) >> 6;
  const int nr = blockIdx.x % nranges, qb = blockIdx.x / nranges, nqb = gridDim.x / nranges;
  const int n0 = nr * 64;
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int t = i >> 6, n = n0 + (i & 63);
    wsm[i] = n < p.N ? from16(reinterpret_cast<const unsigned short*>(p.w)[((long long)n * 16 + t) * p.C], p.in_dt) : 0.f;
  }
  __syncthreads();
  const int g = threadIdx.x & 7;                 // 8-channel group inside the 64-channel range
  if (n0 + g * 8 >= p.N) return;
  constexpr int WW = 3 * STRIDE + 4;
  const unsigned short* src = reinterpret_cast<const unsigned short*>(p.src);
  for (long long quad = (long long)qb * 32 + (threadIdx.x >> 3); quad < p.nquads; quad += (long long)nqb * 32) {
    const int xq = (int)(quad % p.wq);
    const long long r = quad / p.wq;
    const int oy = (int)(r % p.Hout), b = (int)(r / p.Hout);
    const int ox0 = xq * 4;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int iy = oy * STRIDE - p.pad + kh;
      if (iy < 0 || iy >= p.Hin) continue;
      const long long rowbase = ((long long)b * p.Hin + iy) * p.Win;
      float xr[WW];
#pragma unroll
      for (int q = 0; q < WW; ++q) {
        const int ix = ox0 * STRIDE - p.pad + q;
        xr[q] = 0.f;
        if (ix >= 0 && ix < p.Win) {
          const unsigned short u = __ldg(src + (rowbase + ix) * p.ld);
          xr[q] = p.in_dt == PG_F16 ? __half2float(__ushort_as_half(u)) : __uint_as_float((unsigned)u << 16);
        }
      }
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsm + (kh * 4 + kw) * 64 + g * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsm + (kh * 4 + kw) * 64 + g * 8 + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = xr[i * STRIDE + kw];
          acc[i][0] = fmaf(x, w0.x, acc[i][0]); acc[i][1] = fmaf(x, w0.y, acc[i][1]);
          acc[i][2] = fmaf(x, w0.z, acc[i][2]); acc[i][3] = fmaf(x, w0.w, acc[i][3]);
          acc[i][4] = fmaf(x, w1.x, acc[i][4]); acc[i][5] = fmaf(x, w1.y, acc[i][5]);
          acc[i][6] = fmaf(x, w1.z, acc[i][6]); acc[i][7] = fmaf(x, w1.w, acc[i][7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ox = ox0 + i;
      if (ox >= p.Wout) continue;
      const long long o = (((long long)b * p.Hout + oy) * p.Wout + ox) * p.ldo + n0 + g * 8;
      if (p.act != PG_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = act_rt(p.act, acc[i][j]);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out) + o) = pack8dt(acc[i], p.out_dt);
      if (p.out2 != nullptr) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out2) + o) = pack8(acc[i]);
    }
  }
}

bool conv_fewin_supported(const PgConvDesc* d) {
  if (d->mode != PG_CONV || d->C2 != 0 || d->c_valid != 1) return false;
  if (d->out_f32 == PG_F32 || d->has_bias) return false;
  if ((d->N & 7) || d->ldo < d->N || (d->ldo & 7)) return false;
  if (d->n_valid < d->N) return false;
  return true;
}

int conv_fewin(const PgConvDesc* d, const void* src, const void* w, void* out, void* out2, cudaStream_t stream) {
  FewInP p;
  p.src = src; p.w = w; p.out = out; p.out2 = out2;
  p.stride = d->stride; p.pad = d->pad; p.B = d->B; p.Hin = d->Hin; p.Win = d->Win; p.Hout = d->Hout; p.Wout = d->Wout;
  p.C = d->C1; p.ld = d->ld1; p.N = d->N; p.ldo = d->ldo; p.out_dt = d->out_f32; p.in_dt = d->in_dtype; p.act = d->act;
  p.wq = (d->Wout + 3) / 4;
  p.nquads = (long long)d->B * d->Hout * p.wq;
  const int nranges = (p.N + 63) / 64;
  long long qblocks = (p.nquads + 31) / 32;                 // 32 quads per block iteration
  const long long cap = (4LL * num_sms() + nranges - 1) / nranges;
  if (qblocks > cap) qblocks = cap;
  if (qblocks < 1) qblocks = 1;
  const unsigned blocks = (unsigned)(qblocks * nranges);
  if (p.stride == 2) fewin_kernel<2><<<blocks, 256, 0, stream>>>(p);
  else fewin_kernel<1><<<blocks, 256, 0, stream>>>(p);
  return check_launch("fewin_kernel");
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient with ONE real channel on one side:
//   dw[ch * chstride + tap] += sum_pix V[pix][ch] * S[q(pix, tap)][0]
//   s_is_input = 0 ("few-G", e.g. disc last layer): V = layer input a (pixel p = input position), S = dY;
//                 tap (kh,kw) pairs input (iy,ix) with output ((iy+pad-kh)/s, (ix+pad-kw)/s) when divisible / in range
//   s_is_input = 1 ("few-A", e.g. generator output layer): V = g (pixel p = output position), S = a;
//                 tap (kh,kw) pairs output (oy,ox) with input (oy*s-pad+kh, ox*s-pad+kw)
// A thread owns 8 channels and walks a strided slice of the pixels with 16x8 fp32 accumulators; pixel-lanes of a warp
// are reduced by shuffles, then one red.global.add.v4 per (channel, 4 taps).
// ------------------------------------------------------------------------------------------------------------------
struct Wg1P {
  const void* V;
  const void* S;
  float* dw;
  int s_is_input, stride, pad, B, Hv, Wv, Hs, Ws, C, ldv, lds, chstride, ch_real, dt;
  int pl_log2;          // log2(pixel lanes per warp)
  long long npix;
  int pix_per_block;
};

__global__ void __launch_bounds__(256) wgrad1_kernel(const Wg1P p) {
  // thread = 4 channels x a strided slice of the block's pixels, 16 taps x 4 fp32 accumulators
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ng = p.C >> 2;                     // 4-channel groups
  const int GW = 32 >> p.pl_log2;              // channel groups per warp
  const int PL = 1 << p.pl_log2;               // pixel lanes per warp
  const int gl = lane & (GW - 1), pl = lane >> (5 - p.pl_log2);
  const int gwarps = (ng + GW - 1) / GW;       // warps needed to cover all channel groups
  const int nwarps = blockDim.x >> 5;
  const int wrow = warp / gwarps;              // pixel sub-slice of this warp inside the block
  const int g = (warp % gwarps) * GW + gl;
  const int wrows = nwarps / gwarps;
  const int sl = p.stride == 2 ? 1 : 0;
  float acc[16][4];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const long long p0 = (long long)blockIdx.x * p.pix_per_block;
  long long p1 = p0 + p.pix_per_block;
  if (p1 > p.npix) p1 = p.npix;
  const unsigned short* S = reinterpret_cast<const unsigned short*>(p.S);
  if (g < ng && wrow < wrows) {
    for (long long px = p0 + wrow * PL + pl; px < p1; px += (long long)wrows * PL) {
      const int x = (int)(px % p.Wv);
      const long long r = px / p.Wv;
      const int y = (int)(r % p.Hv), b = (int)(r / p.Hv);
      const uint2 vv = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(p.V) + px * p.ldv + g * 4));
      float vf[4];
      if (p.dt == PG_F16) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&vv.x));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
        vf[0] = a.x; vf[1] = a.y; vf[2] = c.x; vf[3] = c.y;
      } else {
        vf[0] = __uint_as_float(vv.x << 16); vf[1] = __uint_as_float(vv.x & 0xffff0000u);
        vf[2] = __uint_as_float(vv.y << 16); vf[3] = __uint_as_float(vv.y & 0xffff0000u);
      }
      const long long sbase = (long long)b * p.Hs;
#pragma unroll
      for (int kh = 0; kh < 4; ++kh) {
        int sy;
        if (p.s_is_input) sy = (y << sl) - p.pad + kh;
        else {
          const int num = y + p.pad - kh;
          if (num < 0 || (num & sl)) continue;
          sy = num >> sl;
        }
        if (sy < 0 || sy >= p.Hs) continue;
#pragma unroll
        for (int kw = 0; kw < 4; ++kw) {
          int sx;
          if (p.s_is_input) sx = (x << sl) - p.pad + kw;
          else {
            const int num = x + p.pad - kw;
            if (num < 0 || (num & sl)) continue;
            sx = num >> sl;
          }
          if (sx < 0 || sx >= p.Ws) continue;
          const unsigned short u = __ldg(S + ((sbase + sy) * p.Ws + sx) * p.lds);
          const float sv = p.dt == PG_F16 ? __half2float(__ushort_as_half(u)) : __uint_as_float((unsigned)u << 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[kh * 4 + kw][j] = fmaf(vf[j], sv, acc[kh * 4 + kw][j]);
        }
      }
    }
  }
  // reduce over the pixel lanes of the warp
  for (int o = GW; o < 32; o <<= 1) {
#pragma unroll
    for (int t = 0; t < 16; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], o);
  }
  // reduce over the warps of the block that share channels (through shared memory), then one red.global.add.v4 per
  // (channel, 4 taps) per block
  __shared__ float red[256 * 32];                 // [wrow][group][8 taps][4 ch], wrows * ng <= 256; two rounds of 8 taps
#pragma unroll
  for (int th = 0; th < 16; th += 8) {
    if (pl == 0 && g < ng && wrow < wrows) {
      float* dst = red + ((size_t)wrow * ng + g) * 32;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<float4*>(dst + t * 4) = make_float4(acc[th + t][0], acc[th + t][1], acc[th + t][2], acc[th + t][3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C * 2; i += blockDim.x) {
      const int ch = i >> 1, tq = (i & 1) * 4;
      if (ch >= p.ch_real) continue;
      const int gg = ch >> 2, j = ch & 3;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int w = 0; w < wrows; ++w) {
        const float* src = red + ((size_t)w * ng + gg) * 32 + tq * 4 + j;
        s0 += src[0]; s1 += src[4]; s2 += src[8]; s3 += src[12];
      }
      float* dst = p.dw + (long long)ch * p.chstride + th + tq;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(s0), "f"(s1), "f"(s2), "f"(s3) : "memory");
    }
    __syncthreads();
  }
}

// few-G: n_real == 1 (S = g, V = a);  few-A: c_real == 1 (S = a, V = g)
bool conv_wgrad1_supported(const PgConvDesc* d, int ldg, const float* dw, int ld_n, int n_real, int c_real) {
  if (d->mode != PG_CONV || d->C2 != 0) return false;
  if (d->in_dtype != d->out_f32) return false;
  if ((((uintptr_t)dw) & 15) != 0) return false;
  if (n_real == 1) return (d->C1 & 7) == 0 && d->C1 <= 1024;
  if (c_real == 1) return (d->N & 7) == 0 && (ld_n & 3) == 0 && d->N <= 1024;
  (void)ldg;
  return false;
}

int conv_wgrad1(const PgConvDesc* d, const void* a, const void* g, int ldg, float* dw, int ld_n, int n_real, int c_real,
                cudaStream_t stream) {
  Wg1P p;
  p.stride = d->stride; p.pad = d->pad; p.B = d->B; p.dt = d->in_dtype; p.dw = dw;
  if (n_real == 1) {          // few-G: dw[0][c][tap]
    p.s_is_input = 0; p.V = a; p.S = g; p.Hv = d->Hin; p.Wv = d->Win; p.Hs = d->Hout; p.Ws = d->Wout;
    p.C = d->C1; p.ldv = d->ld1; p.lds = ldg; p.chstride = 16; p.ch_real = c_real;
  } else {                    // few-A: dw[n][0][tap]
    p.s_is_input = 1; p.V = g; p.S = a; p.Hv = d->Hout; p.Wv = d->Wout; p.Hs = d->Hin; p.Ws = d->Win;
    p.C = d->N; p.ldv = ldg; p.lds = d->ld1; p.chstride = ld_n; p.ch_real = n_real;
  }
  p.npix = (long long)d->B * p.Hv * p.Wv;
  const int ng = p.C >> 2;
  // lanes of a warp: GW channel groups x PL pixel lanes
  int gw = 1;
  while (gw < ng && gw < 32) gw <<= 1;
  int pl_log2 = 0;
  while ((gw << pl_log2) < 32) ++pl_log2;
  p.pl_log2 = pl_log2;
  const int gwarps = (ng + gw - 1) / gw;
  int nwarps = 8;
  if (gwarps > nwarps) nwarps = gwarps;           // C <= 1024 -> <= 8 warps of 32 groups
  nwarps = nwarps / gwarps * gwarps;
  // ~2 blocks per SM; every block ends with C*16 atomics
  long long blocks = 2LL * num_sms();
  const long long min_pix = 64;
  if (blocks * min_pix > p.npix) blocks = (p.npix + min_pix - 1) / min_pix;
  if (blocks < 1) blocks = 1;
  p.pix_per_block = (int)((p.npix + blocks - 1) / blocks);
  blocks = (p.npix + p.pix_per_block - 1) / p.pix_per_block;
  wgrad1_kernel<<<(unsigned)blocks, nwarps * 32, 0, stream>>>(p);
  return check_launch("wgrad1_kernel");
}

// ------------------------------------------------------------------------------------------------------------------
// Tap scatter / gather for layers with ONE real channel on one side.  Such a layer is run as a pointwise product on
// the tensor cores over the 16 taps (PG_CONV1X1):
//   forward (1 output channel):   P[q][tap] = sum_c in[q][c] * W[tap][c]        then  out[p] = act(bias + sum_tap P[q(p,tap)][tap])
//   data-gradient (1 channel in): G[q][tap] = dy[p(q,tap)]                       then  dx[q][c] = sum_tap G[q][tap] * W[c][tap]
//   weight-gradient:              dW[c][tap] = sum_q in[q][c] * G[q][tap]
// q runs over the pixels of the WIDE tensor, p over the pixels of the 1-channel tensor.  The two kernels below are the
// cheap halves: every element of P / G is touched once.
//   geometry PG_CONV (stride s, pad):  p = (q + pad - k) / s   (wide tensor = layer input;  q = p*s - pad + k)
//   geometry PG_CONVT               :  p = 2q - 1 + k          (wide tensor = layer input of a ConvTranspose2d)
// ------------------------------------------------------------------------------------------------------------------
struct TapsP {
  const void* src;
  void* dst;
  const float* bias;
  int mode, stride, pad, B, Hq, Wq, Hp, Wp, lds, ldd, ch, act, src_dt, dst_dt;
  long long n;
};

// wide-pixel coordinate q that pairs with narrow pixel p through tap k, or -1
__device__ __forceinline__ int tap_q_of_p(int mode, int stride, int pad, int p, int k) {
  if (mode == PG_CONVT) {
    const int num = p + 1 - k;
    return (num & 1) ? -1 : (num >> 1);
  }
  return p * stride - pad + k;
}
__device__ __forceinline__ int tap_p_of_q(int mode, int stride, int pad, int q, int k) {
  if (mode == PG_CONVT) return 2 * q - 1 + k;
  const int num = q + pad - k;
  if (num < 0) return -1;
  if (stride == 2) return (num & 1) ? -1 : (num >> 1);
  return num;
}

// out[p] = act(bias + sum_taps P[q(p,tap)][tap]);  P: f32 [B,Hq,Wq,lds];  out: element `ch` of [B,Hp,Wp,ldd]
__global__ void __launch_bounds__(256) taps_scatter_kernel(const TapsP t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.n) return;
  int px, py, b;
  split_xyb(i, t.Wp, t.Hp, px, py, b);
  const float* P = reinterpret_cast<const float*>(t.src);
  float acc = t.bias != nullptr ? __ldg(t.bias) : 0.f;
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int qy = tap_q_of_p(t.mode, t.stride, t.pad, py, kh);
    if (qy < 0 || qy >= t.Hq) continue;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int qx = tap_q_of_p(t.mode, t.stride, t.pad, px, kw);
      if (qx < 0 || qx >= t.Wq) continue;
      acc += __ldg(P + (((long long)b * t.Hq + qy) * t.Wq + qx) * t.lds + kh * 4 + kw);
    }
  }
  acc = act_rt(t.act, acc);
  const long long o = i * t.ldd + t.ch;
  if (t.dst_dt == PG_F32) reinterpret_cast<float*>(t.dst)[o] = acc;
  else reinterpret_cast<unsigned short*>(t.dst)[o] = to16(acc, t.dst_dt);
}

// G[q][tap] = src[p(q,tap)][ch] (0 outside);  src: 16-bit [B,Hp,Wp,lds];  G: 16-bit [B,Hq,Wq,16], same type
__global__ void __launch_bounds__(256) taps_gather_kernel(const TapsP t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.n) return;
  int qx, qy, b;
  split_xyb(i, t.Wq, t.Hq, qx, qy, b);
  const unsigned short* S = reinterpret_cast<const unsigned short*>(t.src);
  unsigned short v[16];
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int py = tap_p_of_q(t.mode, t.stride, t.pad, qy, kh);
    const bool yok = py >= 0 && py < t.Hp;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int px = tap_p_of_q(t.mode, t.stride, t.pad, qx, kw);
      unsigned short u = 0;
      if (yok && px >= 0 && px < t.Wp) u = __ldg(S + (((long long)b * t.Hp + py) * t.Wp + px) * t.lds + t.ch);
      v[kh * 4 + kw] = u;
    }
  }
  uint4 lo, hi;
  lo.x = v[0] | ((unsigned)v[1] << 16); lo.y = v[2] | ((unsigned)v[3] << 16);
  lo.z = v[4] | ((unsigned)v[5] << 16); lo.w = v[6] | ((unsigned)v[7] << 16);
  hi.x = v[8] | ((unsigned)v[9] << 16); hi.y = v[10] | ((unsigned)v[11] << 16);
  hi.z = v[12] | ((unsigned)v[13] << 16); hi.w = v[14] | ((unsigned)v[15] << 16);
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(t.dst) + i * t.ldd);
  dst[0] = lo;
  dst[1] = hi;
}

int taps_scatter(int mode, int stride, int pad, int B, int Hq, int Wq, int Hp, int Wp, const float* P, int ldp, const float* bias,
                 int act, void* out, int out_dt, int ldo, int ch, cudaStream_t stream) {
  TapsP t;
  t.src = P; t.dst = out; t.bias = bias; t.mode = mode; t.stride = stride; t.pad = pad; t.B = B; t.Hq = Hq; t.Wq = Wq;
  t.Hp = Hp; t.Wp = Wp; t.lds = ldp; t.ldd = ldo; t.ch = ch; t.act = act; t.src_dt = PG_F32; t.dst_dt = out_dt;
  t.n = (long long)B * Hp * Wp;
  taps_scatter_kernel<<<(unsigned)((t.n + 255) / 256), 256, 0, stream>>>(t);
  return check_launch("taps_scatter_kernel");
}

int taps_gather(int mode, int stride, int pad, int B, int Hq, int Wq, int Hp, int Wp, const void* src, int lds, int ch, void* G,
                cudaStream_t stream) {
  TapsP t;
  t.src = src; t.dst = G; t.bias = nullptr; t.mode = mode; t.stride = stride; t.pad = pad; t.B = B; t.Hq = Hq; t.Wq = Wq;
  t.Hp = Hp; t.Wp = Wp; t.lds = lds; t.ldd = 16; t.ch = ch; t.act = 0; t.src_dt = PG_BF16; t.dst_dt = PG_BF16;
  t.n = (long long)B * Hq * Wq;
  taps_gather_kernel<<<(unsigned)((t.n + 255) / 256), 256, 0, stream>>>(t);
  return check_launch("taps_gather_kernel");
}

// ------------------------------------------------------------------------------------------------------------
// Data-gradient of a tap-product layer (one real output channel; disc.py:45 backward), fused with the backward of the
// activation in front of it:   dx[q][c] = (sum_tap G[q][tap] * W[c][tap]) * act'(y[q][c])
// K = 16 and N = 8*ndf = 512: as a tcgen05 GEMM this is one MMA per tile in front of a 128 x 512 epilogue (60-70 us for
// cfg 3's 30 k pixels); it is a 16-term dot product per output, HBM-bound on the CUDA cores (read y, write dx).
// A thread owns TWO adjacent channels (weights in registers) and walks the pixels; the pixel's 16 taps come from shared
// memory as broadcasts; consecutive threads write consecutive channel pairs (coalesced 4-byte stores).
// ------------------------------------------------------------------------------------------------------------
constexpr int TD_PX = 32;     // pixels staged per block iteration

__global__ void __launch_bounds__(512) taps_dgrad_act_kernel(const bf16* __restrict__ G, const bf16* __restrict__ w16,
                                                             const unsigned short* __restrict__ y, int ldy, int y_dt, int act,
                                                             bf16* __restrict__ dx, int lddx, long long nq, int C) {
  __shared__ float Gs[TD_PX][16];
  const int c = 2 * threadIdx.x;
  const bool live = c < C;
  float w0[16], w1[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    w0[t] = live ? __bfloat162float(w16[(long long)c * 16 + t]) : 0.f;
    w1[t] = live ? __bfloat162float(w16[(long long)(c + 1) * 16 + t]) : 0.f;
  }
  for (long long q0 = (long long)blockIdx.x * TD_PX; q0 < nq; q0 += (long long)gridDim.x * TD_PX) {
    __syncthreads();
    for (int i = threadIdx.x; i < TD_PX * 2; i += blockDim.x) {      // 2 loads per pixel, 8 taps (16 bytes) each
      const int pxl = i >> 1, half = i & 1;
      float f[8];
      if (q0 + pxl < nq) {
        unpack8(*reinterpret_cast<const uint4*>(G + (q0 + pxl) * 16 + half * 8), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) Gs[pxl][half * 8 + j] = f[j];
    }
    __syncthreads();
    if (!live) continue;
    const int npx = nq - q0 < TD_PX ? (int)(nq - q0) : TD_PX;
    // eight pixels at a time: their y loads are issued together (a serial walk paid one DRAM latency per pixel)
    for (int p0 = 0; p0 < npx; p0 += 8) {
      unsigned yy[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        yy[u] = (y != nullptr && p0 + u < npx) ? __ldg(reinterpret_cast<const unsigned*>(y + (q0 + p0 + u) * ldy + c)) : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (p0 + u >= npx) break;
        const float4* g4 = reinterpret_cast<const float4*>(Gs[p0 + u]);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 gv = g4[j];
          a0 = fmaf(gv.x, w0[4 * j], a0); a0 = fmaf(gv.y, w0[4 * j + 1], a0); a0 = fmaf(gv.z, w0[4 * j + 2], a0); a0 = fmaf(gv.w, w0[4 * j + 3], a0);
          a1 = fmaf(gv.x, w1[4 * j], a1); a1 = fmaf(gv.y, w1[4 * j + 1], a1); a1 = fmaf(gv.z, w1[4 * j + 2], a1); a1 = fmaf(gv.w, w1[4 * j + 3], a1);
        }
        if (y != nullptr) {
          a0 *= act_grad_from_output(act, from16((unsigned short)(yy[u] & 0xffffu), y_dt));
          a1 *= act_grad_from_output(act, from16((unsigned short)(yy[u] >> 16), y_dt));
        }
        *reinterpret_cast<__nv_bfloat162*>(dx + (q0 + p0 + u) * lddx + c) = __floats2bfloat162_rn(a0, a1);
      }
    }
  }
}

int taps_dgrad_act(const void* G, const void* w16, void* dx, int lddx, const void* y, int ldy, int y_dt, int act, long long nq, int C,
                   cudaStream_t stream) {
  const int threads = ((C / 2 + 31) / 32) * 32;
  long long blocks = (nq + TD_PX - 1) / TD_PX;
  const long long cap = 6LL * num_sms();          // (the weights are loaded once per block: 64 bytes per thread, from L2)
  if (blocks > cap) blocks = cap;
  taps_dgrad_act_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const bf16*)G, (const bf16*)w16, (const unsigned short*)y, ldy,
                                                                y_dt, act, (bf16*)dx, lddx, nq, C);
  return check_launch("taps_dgrad_act_kernel");
}

}  // namespace pg
