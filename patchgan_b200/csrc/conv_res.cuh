// Convolution + InstanceNorm in ONE launch, accumulators resident in tensor memory (included by conv_tc.cu).
//
// DownSampleBlock / UpSampleBlock are conv -> InstanceNorm2d -> activation -> [Dropout] (unet.py:19-28, 53-66).  The
// statistics of a channel need the whole H x W map, so a conv epilogue cannot normalise its own tile -- unless the tile
// WAITS.  A B200 has 148 x 512 columns x 128 lanes of tensor memory = 9.7 M fp32 accumulators, more than every normalised
// layer of the 256 x 256 configurations produces for a whole batch.  conv_res_kernel therefore
//   phase 1: one CTA per SM computes up to 512 / BN output tiles, each into its OWN TMEM columns, and adds the per-(image,
//            channel) sums taken from the fp32 accumulators to global memory (red.global);
//   grid barrier (all CTAs are co-resident: grid <= #SMs, one CTA per SM);
//   phase 2: re-reads its accumulators from TMEM, normalises, applies activation (+ dropout) and stores the FINAL 16-bit
//            tensor (+ its bf16 twin) with TMA.
// The pre-norm tensor never exists in memory: no fp32 round trip, no separate norm_act_fwd launch.
//
// The same two-phase epilogue runs the BACKWARD of that block fused into the data-gradient convolution that produces
// dL/dy (kind = PG_FUSED_BWD): phase 1 forms g = (acc [+ skip gradient]) * act'(.) * dropout mask from the fp32
// accumulators and reduces sum(g), sum(g * xhat) per (image, channel); phase 2 writes
//   d(conv output) = rstd * (g - mean(g) - xhat * mean(g * xhat))                (autograd of aten::instance_norm)
// in bf16.  xhat is recovered from the saved activation output when the activation is invertible (LeakyReLU / none,
// no dropout), otherwise read from the xhat tensor the forward kernel saved.
#pragma once

struct ResParams {
  int kind;                     // PG_FUSED_FWD / PG_FUSED_BWD
  int n_norm;                   // bwd: output channels [0, n_norm) belong to the normalised layer; the rest is stored as is
  int cn;                       // channels per image in sums / bsums and in the dropout element index
  unsigned int* sync;           // two zeroed counters: grid barriers
  float* sums;                  // fwd: zeroed [B][N][2], receives (sum, sum of squares); bwd: the forward sums (read)
  float* bsums;                 // bwd: zeroed [B][n_norm][2], receives (sum g, sum g*xhat)
  float inv_hw;
  float drop_p, keep_scale;
  const unsigned long long* seed;
  unsigned long long salt;
  void* xhat;                   // fwd: optional output (16-bit, same dtype as the output); bwd: optional input
  int xhat_ld;
  const void* y;                // bwd: saved output of the layer (used when xhat == null)
  int y_ld, y_dt;
  const void* dskip;            // bwd: bf16 gradient of the skip connection, added to the accumulators (or null)
  int dskip_ld;
  uint32_t table_off;           // byte offset (from the 1024-aligned dynamic smem base) of the coefficient table
  int splits, kps;              // split-K mode (splits > 1): CTA = (tile, split) owns k-steps [split*kps, (split+1)*kps)
  float* ws;                    // split-K exchange scratch in global memory (L2-resident): [tile][split][128 rows][BN] fp32
};

constexpr int RES_MAX_SLOTS = 16;
constexpr float RES_IN_EPS = 1e-5f;

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_inc(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
}

// Sums of a[j], q[j] (16 channels per lane) over each group of 2^lgP consecutive lanes (lgP >= 5: the whole warp) are
// added to dst[j*2 + {0,1}] of the group's first lane.  dst_ok: this lane's image exists (b < B).
__device__ __forceinline__ void group_add16(float* a, float* q, int lgP, int lane, bool dst_ok, float* dst) {
  if (lgP >= 5) {
    // halving butterfly: 8+4+2+1+1 shuffles per quantity leave one channel per lane pair (see stats_add16)
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0;
    float a8[8], q8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a8[i] = (h16 ? a[i + 8] : a[i]) + __shfl_xor_sync(0xffffffffu, h16 ? a[i] : a[i + 8], 16);
      q8[i] = (h16 ? q[i + 8] : q[i]) + __shfl_xor_sync(0xffffffffu, h16 ? q[i] : q[i + 8], 16);
    }
    float a4[4], q4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a4[i] = (h8 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, h8 ? a8[i] : a8[i + 4], 8);
      q4[i] = (h8 ? q8[i + 4] : q8[i]) + __shfl_xor_sync(0xffffffffu, h8 ? q8[i] : q8[i + 4], 8);
    }
    float a2[2], q2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      a2[i] = (h4 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, h4 ? a4[i] : a4[i + 2], 4);
      q2[i] = (h4 ? q4[i + 2] : q4[i]) + __shfl_xor_sync(0xffffffffu, h4 ? q4[i] : q4[i + 2], 4);
    }
    float a1 = (h2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? a2[0] : a2[1], 2);
    float q1 = (h2 ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? q2[0] : q2[1], 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
    if ((lane & 1) == 0 && dst_ok) {
      const int col = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + col * 2), "f"(a1), "f"(q1) : "memory");
    }
  } else {
    // small maps (2x2 .. 4x4 lattice per image): several images per warp, plain xor tree inside each lane group
    for (int off = 1; off < (1 << lgP); off <<= 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        a[j] += __shfl_xor_sync(0xffffffffu, a[j], off);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
      }
    }
    if ((lane & ((1 << lgP) - 1)) == 0 && dst_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + j * 2), "f"(a[j]), "f"(q[j]) : "memory");
    }
  }
}

template <int ACT>
__device__ __forceinline__ float act_grad_in_fast(float x) {
  if (ACT == PG_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (ACT == PG_ACT_LEAKYRELU) return x > 0.f ? 1.f : 0.2f;
  if (ACT == PG_ACT_TANH) { const float t = act_fast<PG_ACT_TANH>(x); return 1.f - t * t; }
  if (ACT == PG_ACT_SIGMOID) { const float s = act_fast<PG_ACT_SIGMOID>(x); return s * (1.f - s); }
  return 1.f;
}

__device__ __forceinline__ void load16_16bit(const void* base, long long off, int dt, float* f) {
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(base) + off);
  const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
  if (dt == PG_F16) { unpack8h(v0, f); unpack8h(v1, f + 8); } else { unpack8(v0, f); unpack8(v1, f + 8); }
}

// Backward, one 16-channel chunk of one output pixel: v = raw accumulators -> g = dL/dxhat, xh = xhat
template <int ACT>
__device__ __forceinline__ void res_bwd_gx(const ResParams& r, const uint32_t* v, long long opix, int n, bool valid,
                                           unsigned long long seed, float* g, float* xh) {
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { g[j] = 0.f; xh[j] = 0.f; }
    return;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v[j]);
  if (r.dskip != nullptr) {
    float s[16];
    load16_16bit(r.dskip, opix * r.dskip_ld + n, PG_BF16, s);
#pragma unroll
    for (int j = 0; j < 16; ++j) g[j] += s[j];
  }
  if (r.xhat != nullptr) {
    load16_16bit(r.xhat, opix * r.xhat_ld + n, r.y_dt, xh);
#pragma unroll
    for (int j = 0; j < 16; ++j) g[j] *= act_grad_in_fast<ACT>(xh[j]);
  } else {
    float yv[16];
    load16_16bit(r.y, opix * r.y_ld + n, r.y_dt, yv);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // (only invertible activations take this route: LeakyReLU(0.2) and none)
      xh[j] = (ACT == PG_ACT_LEAKYRELU && yv[j] < 0.f) ? 5.f * yv[j] : yv[j];
      g[j] *= act_grad_out_fast<ACT>(yv[j]);
    }
  }
  if (r.drop_p > 0.f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float u = uniform01(seed, (unsigned long long)(opix * r.cn + n + j));
      g[j] = u >= r.drop_p ? g[j] * r.keep_scale : 0.f;
    }
  }
}

// ---- the two-phase epilogue.  One CTA owns an SM, so the epilogue gets SIXTEEN warps (four groups of four; warp w may
// touch TMEM lanes 32*(w % 4) .. +31, a group covers all 128 rows of a tile).  Work units are (tile, 16-column chunk),
// handed round-robin to the groups; a unit needs no synchronisation inside its group: statistics are warp shuffles +
// red.global, the final values are stored straight from registers (32 bytes per pixel and unit).
// Waiting warps must not spin: the TMA producer and the MMA issuer are single threads that share their schedulers with
// four epilogue warps each, and sixteen warps polling an mbarrier halved the main loop's k-step rate.  One lane per warp
// polls with a nanosleep back-off, the others wait in __syncwarp.
#ifndef RES_GROUPS_N
#define RES_GROUPS_N 3
#endif
constexpr int RES_GROUPS = RES_GROUPS_N;
constexpr int RES_EPI_THREADS = RES_GROUPS * 128;
constexpr int RES_THREADS = 128 + RES_EPI_THREADS;         // warps 0..3: TMA producer, MMA issuer, 2 idle; warps 4..19: epilogue
constexpr int RES_TABLE_MAX = 2048;                        // float4 coefficient entries in shared memory

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0 && !mbar_try_wait(bar, parity)) {
    const long long t0 = clock64();
    unsigned ns = 32;
    while (!mbar_try_wait(bar, parity)) {
      __nanosleep(ns);
      if (ns < 256) ns <<= 1;
      if (clock64() - t0 > 4000000000LL) {
        printf("conv_res: accumulator barrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
  __syncwarp();
}

// every CTA of the grid is resident (grid <= #SMs, one CTA per SM): a counter in global memory is a grid barrier
__device__ __forceinline__ void res_grid_barrier(unsigned int* ctr) {
  __threadfence();
  named_bar_sync(6, RES_EPI_THREADS);
  if (threadIdx.x == 128) {
    red_release_gpu_inc(ctr);
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(ctr) < gridDim.x) {
      __nanosleep(40);
      if (clock64() - t0 > 2000000000LL) {
        printf("conv_res: grid barrier timeout (block %d of %d, saw %u)\n", blockIdx.x, gridDim.x, ld_acquire_gpu_u32(ctr));
        __trap();
      }
    }
  }
  named_bar_sync(6, RES_EPI_THREADS);
  __threadfence();
}

struct RowPix {
  bool valid;        // this accumulator row is an output pixel
  int b;             // image
  long long opix;    // output pixel index (b * Hout + oy) * Wout + ox
};
__device__ __forceinline__ RowPix row_pixel(const TcParams& p, const TileXY& t, int row) {
  const int lgP = p.lgTW + p.lgTH;
  const int xl = row & (p.TW - 1), yl = (row >> p.lgTW) & (p.TH - 1), bl = row >> lgP;
  RowPix rp;
  rp.b = t.b0 + bl;
  const int a = t.y0 + yl, bb = t.x0 + xl;
  rp.valid = rp.b < p.B && a < p.Ha && bb < p.Wa;
  int oy = a, ox = bb;
  if (p.mode == PG_CONVT) { oy = 2 * a + t.py; ox = 2 * bb + t.px; }
  rp.opix = ((long long)rp.b * p.Hout + oy) * p.Wout + ox;
  return rp;
}

// coefficients of (image b, channel ch): fwd (mean, rstd), bwd (rstd, mean g, mean g*xhat)
template <int KIND>
__device__ __forceinline__ float4 res_coef(const TcParams& p, const ResParams& r, int b, int ch) {
  if (b >= p.B) return make_float4(0.f, 1.f, 0.f, 0.f);
  const double inv_hw = (double)r.inv_hw;
  const float* sp = r.sums + ((long long)b * r.cn + ch) * 2;
  const double m = (double)__ldcg(sp) * inv_hw;
  double var = (double)__ldcg(sp + 1) * inv_hw - m * m;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)RES_IN_EPS));
  if (KIND == PG_FUSED_FWD) return make_float4((float)m, rstd, 0.f, 0.f);
  const float* bs = r.bsums + ((long long)b * r.cn + ch) * 2;
  return make_float4(rstd, __ldcg(bs) * r.inv_hw, __ldcg(bs + 1) * r.inv_hw, 0.f);
}

// phase-1 quantities of one unit: v = 16 raw sums of this thread's row -> (a1, a2) whose per-(image, channel) sums are wanted
template <int KIND, int ACT>
__device__ __forceinline__ void res_stats16(const TcParams& p, const ResParams& r, const uint32_t* v, const RowPix& rp, int n,
                                            unsigned long long seed, int lane, int lgP) {
  float a1[16], a2[16];
  if (KIND == PG_FUSED_FWD) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { a1[j] = rp.valid ? __uint_as_float(v[j]) : 0.f; a2[j] = a1[j] * a1[j]; }
  } else {
    float xh[16];
    res_bwd_gx<ACT>(r, v, rp.opix, n, rp.valid, seed, a1, xh);
#pragma unroll
    for (int j = 0; j < 16; ++j) a2[j] = a1[j] * xh[j];
  }
  float* dst = (KIND == PG_FUSED_FWD ? r.sums : r.bsums) + ((long long)(rp.b < p.B ? rp.b : 0) * r.cn + n) * 2;
  group_add16(a1, a2, lgP, lane, rp.b < p.B, dst);
}

// phase-2 of one unit: v = 16 raw sums, cf = the 16 coefficient entries of this row's image -> final values, stored
template <int KIND, int ACT>
__device__ __forceinline__ void res_finish16(const TcParams& p, const ResParams& r, const uint32_t* v, const float4* cf,
                                             bool norm_tile, const RowPix& rp, int n, unsigned long long seed) {
  float f[16];
  if (!norm_tile) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  } else if (KIND == PG_FUSED_FWD) {
    float xh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 c4 = cf[j];
      xh[j] = (__uint_as_float(v[j]) - c4.x) * c4.y;
      f[j] = act_fast<ACT>(xh[j]);
    }
    if (r.drop_p > 0.f) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float u = uniform01(seed, (unsigned long long)(rp.opix * r.cn + n + j));
        f[j] = u >= r.drop_p ? f[j] * r.keep_scale : 0.f;
      }
    }
    if (n + 16 > p.n_valid) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (n + j >= p.n_valid) { f[j] = 0.f; xh[j] = 0.f; }
      }
    }
    if (r.xhat != nullptr && rp.valid) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(r.xhat) + rp.opix * r.xhat_ld + n);
      o[0] = pack8dt(xh, p.out_f32);
      o[1] = pack8dt(xh + 8, p.out_f32);
    }
  } else {
    float g[16], xh[16];
    res_bwd_gx<ACT>(r, v, rp.opix, n, rp.valid, seed, g, xh);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 c4 = cf[j];
      f[j] = c4.x * (g[j] - c4.y - xh[j] * c4.z);
    }
  }
  if (rp.valid && !(p.debug & 4)) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out) + rp.opix * p.ldo + n);
    o[0] = pack8dt(f, p.out_f32);
    o[1] = pack8dt(f + 8, p.out_f32);
    if (p.out2 != nullptr) {
      uint4* o2 = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(p.out2) + rp.opix * p.ldo + n);
      o2[0] = pack8(f);
      o2[1] = pack8(f + 8);
    }
  }
}

template <int KIND, int ACT>
__device__ __forceinline__ void res_epilogue(const TcParams& p, const ResParams& r, uint8_t* smem_gen, uint32_t tmem_base,
                                             uint64_t* tfull, int my_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, row = q * 32 + lane;
  const int grp = (warp - 4) >> 2;                       // 0 .. RES_GROUPS-1
  const int et = (int)threadIdx.x - 128;                 // 0 .. 511 among the epilogue threads
  const int lgP = p.lgTW + p.lgTH, lgTB = 7 - lgP, lgBN = 31 - __clz(p.BN);
  const int bl = row >> lgP;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const unsigned long long seed = r.drop_p > 0.f ? mix_seed(*r.seed, r.salt) : 0ull;
  const int nchunk = p.BN >> 4;
  const bool tracer = p.trace != nullptr && threadIdx.x == 128;
  float4* table = reinterpret_cast<float4*>(smem_gen + r.table_off);

  if (r.splits > 1) {
    // =========================== split-K mode: this CTA holds a PARTIAL tile ===========================
    // (small maps: a handful of output tiles, K = 16 * Cin in the thousands.  splits CTAs share a tile's k-range, dump
    //  their partial accumulators to an L2-resident scratch, and after a grid barrier every (tile, 16-column chunk) unit is
    //  summed by one epilogue group somewhere in the grid, which then owns it: statistics, second barrier, final values
    //  from registers.)
    const int tile = blockIdx.x / r.splits, split = blockIdx.x - tile * r.splits;
    mbar_wait_warp(smem_u32(&tfull[0]), 0, lane);
    tc_fence_after();
    if (tracer) { trace_raw(p.trace, 2, gtimer()); trace_raw(p.trace, 3, gtimer()); }
    {
      float* dst = r.ws + (((long long)tile * r.splits + split) * 128 + row) * p.BN;
      for (int c16 = grp; c16 < nchunk; c16 += RES_GROUPS) {
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_off + (uint32_t)(c16 << 4), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          __stcg(reinterpret_cast<float4*>(dst + (c16 << 4)) + j,
                 make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3])));
      }
    }
    if (tracer) trace_raw(p.trace, 4, gtimer());
    res_grid_barrier(r.sync);
    if (tracer) trace_raw(p.trace, 5, gtimer());
    // ---- this group's unit (at most one: the plan keeps units = tiles * nchunk <= RES_GROUPS * grid)
    const int unit = (int)blockIdx.x + grp * (int)gridDim.x;
    const bool have = unit < p.pers_total * nchunk;
    const int utile = have ? unit / nchunk : 0, c16 = have ? unit - utile * nchunk : 0;
    const TileXY t = pers_decode(p, utile);
    const RowPix rp = row_pixel(p, t, row);
    const int n = t.n0 + (c16 << 4);
    const bool norm_tile = KIND == PG_FUSED_FWD || t.n0 < r.n_norm;
    uint32_t v[16];
    if (have) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      const float* src = r.ws + (((long long)utile * r.splits) * 128 + row) * p.BN + (c16 << 4);
      const long long sstride = 128LL * p.BN;
      // (the partials of four splits are requested before any is added: one L2 round trip per four splits instead of one
      //  per split -- the loop used to cost ~0.5 us per split on the layer's critical path)
      int sp = 0;
      for (; sp + 4 <= r.splits; sp += 4) {
        float4 tq[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) tq[u][j] = __ldcg(reinterpret_cast<const float4*>(src + (sp + u) * sstride) + j);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc[4 * j] += tq[u][j].x; acc[4 * j + 1] += tq[u][j].y; acc[4 * j + 2] += tq[u][j].z; acc[4 * j + 3] += tq[u][j].w; }
      }
      for (; sp < r.splits; ++sp) {
        float4 tq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tq[j] = __ldcg(reinterpret_cast<const float4*>(src + sp * sstride) + j);
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[4 * j] += tq[j].x; acc[4 * j + 1] += tq[j].y; acc[4 * j + 2] += tq[j].z; acc[4 * j + 3] += tq[j].w; }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(acc[j]);
      if (norm_tile) res_stats16<KIND, ACT>(p, r, v, rp, n, seed, lane, lgP);
    }
    if (tracer) trace_raw(p.trace, 10, gtimer());
    // The statistics of a unit's images are complete when every unit holding pixels of those images has added its sums.
    // A conv-form tile that covers the whole map (maps up to 8 x 8: nx = ny = 1) holds ALL pixels of its images, and a unit
    // = (tile, 16 channels) belongs to this group alone: the group's own fence + barrier is enough, no grid-wide wait for
    // the slowest CTA.  (ConvTranspose2d-form tiles hold one output-parity class of an image: they need the grid.)
    if (p.nx == 1 && p.ny == 1 && p.mode != PG_CONVT) {
      __threadfence();
      named_bar_sync(1 + grp, 128);
    } else {
      res_grid_barrier(r.sync + 1);
    }
    if (tracer) trace_raw(p.trace, 11, gtimer());
    if (have) {
      // coefficient table of this group: [image of the tile][16 channels of the unit]
      float4* gtab = table + grp * (16 << lgTB);
      if (norm_tile)
        for (int e = (int)threadIdx.x - 128 - grp * 128; e < (16 << lgTB); e += 128)
          gtab[e] = res_coef<KIND>(p, r, t.b0 + (e >> 4), n + (e & 15));
    }
    named_bar_sync(1 + grp, 128);
    if (have) res_finish16<KIND, ACT>(p, r, v, table + grp * (16 << lgTB) + (bl << 4), norm_tile, rp, n, seed);
    if (tracer) { trace_raw(p.trace, 8, gtimer()); trace_raw(p.trace, 9, 1ull); }
    return;
  }

  // =========================== resident mode: whole tiles stay in TMEM across the barrier ===========================
  // ---- phase 1: per-(image, channel) sums
  int unit = 0;
  for (int i = 0; i < my_tiles; ++i) {
    const TileXY t = pers_decode(p, blockIdx.x + i * gridDim.x);
    mbar_wait_warp(smem_u32(&tfull[i]), 0, lane);
    tc_fence_after();
    if (tracer && i == 0) trace_raw(p.trace, 2, gtimer());
    if (tracer && i == my_tiles - 1) trace_raw(p.trace, 3, gtimer());
    if (KIND == PG_FUSED_BWD && t.n0 >= r.n_norm) { unit += nchunk; continue; }
    const RowPix rp = row_pixel(p, t, row);
    const uint32_t trow = tmem_base + (uint32_t)(i * p.BN) + lane_off;
    for (int c16 = 0; c16 < nchunk; ++c16, ++unit) {
      if (unit % RES_GROUPS != grp) continue;
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)(c16 << 4), v);
      tmem_ld_wait();
      res_stats16<KIND, ACT>(p, r, v, rp, t.n0 + (c16 << 4), seed, lane, lgP);
    }
  }
  if (tracer) trace_raw(p.trace, 4, gtimer());
  res_grid_barrier(r.sync);
  if (tracer) trace_raw(p.trace, 5, gtimer());

  // ---- coefficients of all my tiles: entry ((i * TB + image of the tile) * BN + channel)
  {
    const int per_tile = p.BN << lgTB;
    const int total = my_tiles * per_tile;
    for (int e = et; e < total; e += RES_EPI_THREADS) {
      const int i = e / per_tile, rem = e - i * per_tile;
      const TileXY t = pers_decode(p, blockIdx.x + i * gridDim.x);
      if (KIND == PG_FUSED_BWD && t.n0 >= r.n_norm) continue;
      table[e] = res_coef<KIND>(p, r, t.b0 + (rem >> lgBN), t.n0 + (rem & (p.BN - 1)));
    }
  }
  named_bar_sync(6, RES_EPI_THREADS);

  // ---- phase 2: normalise, activate, store
  unit = 0;
  for (int i = 0; i < my_tiles; ++i) {
    const TileXY t = pers_decode(p, blockIdx.x + i * gridDim.x);
    const bool norm_tile = KIND == PG_FUSED_FWD || t.n0 < r.n_norm;
    const RowPix rp = row_pixel(p, t, row);
    const uint32_t trow = tmem_base + (uint32_t)(i * p.BN) + lane_off;
    const float4* ctile = table + (((i << lgTB) + bl) << lgBN);
    for (int c16 = 0; c16 < nchunk; ++c16, ++unit) {
      if (unit % RES_GROUPS != grp) continue;
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)(c16 << 4), v);
      tmem_ld_wait();
      res_finish16<KIND, ACT>(p, r, v, ctile + (c16 << 4), norm_tile, rp, t.n0 + (c16 << 4), seed);
    }
  }
  if (p.trace != nullptr) {
    named_bar_sync(6, RES_EPI_THREADS);
    if (tracer) { trace_raw(p.trace, 8, gtimer()); trace_raw(p.trace, 9, (unsigned long long)my_tiles); }
  }
}

template <int KIND>
__global__ void __launch_bounds__(RES_THREADS, 1)
conv_res_kernel(const __grid_constant__ ActMaps mapsA, const __grid_constant__ CUtensorMap mapB, const TcParams p,
                const ResParams r) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull[RES_MAX_SLOTS];
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.stages * p.a_bytes;
  const int nk = p.nk1 + p.nk2;
  // resident mode: this CTA's tiles are blockIdx.x + i * gridDim.x, all k-steps; split-K mode: one (tile, split)
  const int my_tiles = r.splits > 1 ? 1 : (p.pers_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ks0 = r.splits > 1 ? ((int)blockIdx.x % r.splits) * r.kps : 0;
  const int ksteps = r.splits > 1 ? r.kps : p.ntaps * nk;
  if (p.trace != nullptr && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace_raw(p.trace, 0, gtimer());
    trace_raw(p.trace, 7, smid);
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapsA.m[0]);
    if (p.nk2 > 0) prefetch_tmap(&mapsA.m[4]);
    prefetch_tmap(&mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < RES_MAX_SLOTS; ++s) mbar_init(smem_u32(&tfull[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_sh), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  if (p.trace != nullptr && threadIdx.x == 0) trace_raw(p.trace, 1, gtimer());

  if (warp == 0) {
    // ===================== TMA producer: the ring runs ahead across this CTA's tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const TileXY t = pers_decode(p, r.splits > 1 ? (int)blockIdx.x / r.splits : (int)blockIdx.x + i * (int)gridDim.x);
        int tap = ks0 / nk, ck = ks0 - tap * nk, cx = 0, cy = 0, wtap = 0, ph = 0;
        bool newtap = true;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (newtap) {
            newtap = false;
            ph = 0;
            if (p.mode == PG_CONVT) {
              const int j = tap >> 1, ii = tap & 1;
              wtap = ((1 - t.py) + 2 * j) * 4 + (1 - t.px) + 2 * ii;
              cx = t.x0 + t.px - ii;
              cy = t.y0 + t.py - j;
            } else if (p.mode == PG_CONV1X1) {
              wtap = 0; cx = t.x0; cy = t.y0;
            } else {
              const int kh = tap >> 2, kw = tap & 3;
              wtap = tap;
              if (p.stride == 2) {
                const int u = kh - p.pad, v = kw - p.pad;
                ph = (u & 1) * 2 + (v & 1);
                cx = t.x0 + (v >> 1);
                cy = t.y0 + (u >> 1);
              } else {
                cx = t.x0 - p.pad + kw;
                cy = t.y0 - p.pad + kh;
              }
            }
          }
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, p.tx_bytes);
          if (ck < p.nk1) tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[ph], fb, ck * p.BK, cx, cy, t.b0);
          else tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[4 + ph], fb, (ck - p.nk1) * p.BK, cx, cy, t.b0);
          tma_load_2d(b_base + stage * p.b_bytes, &mapB, fb, wtap * p.Ctot + ck * p.BK, t.n0);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++ck == nk) { ck = 0; ++tap; newtap = true; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: tile i accumulates into TMEM columns [i*BN, (i+1)*BN) =====================
    if (lane == 0) {
      const uint32_t a_lo0 = (a_base & 0x3FFFF) >> 4, b_lo0 = (b_base & 0x3FFFF) >> 4;
      const uint64_t desc_hi = make_smem_desc(0, p.sbo, p.layout_type);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      MmaState st{a_lo0, b_lo0, 0u, 0u};
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t acc = tmem_base + (uint32_t)(i * p.BN);
        if (p.BK == 64) mma_issue_tile<4>(p, st, full0, empty0, smem_u32(&tfull[i]), a_lo0, b_lo0, desc_hi, acc, ksteps);
        else if (p.BK == 32) mma_issue_tile<2>(p, st, full0, empty0, smem_u32(&tfull[i]), a_lo0, b_lo0, desc_hi, acc, ksteps);
        else mma_issue_tile<1>(p, st, full0, empty0, smem_u32(&tfull[i]), a_lo0, b_lo0, desc_hi, acc, ksteps);
      }
    }
  } else if (warp >= 4) {
    // ===================== two-phase epilogue: 16 warps =====================
    switch (p.act) {
      case PG_ACT_RELU: res_epilogue<KIND, PG_ACT_RELU>(p, r, smem_gen, tmem_base, tfull, my_tiles); break;
      case PG_ACT_LEAKYRELU: res_epilogue<KIND, PG_ACT_LEAKYRELU>(p, r, smem_gen, tmem_base, tfull, my_tiles); break;
      case PG_ACT_TANH: res_epilogue<KIND, PG_ACT_TANH>(p, r, smem_gen, tmem_base, tfull, my_tiles); break;
      default: res_epilogue<KIND, PG_ACT_NONE>(p, r, smem_gen, tmem_base, tfull, my_tiles); break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (p.trace != nullptr && threadIdx.x == 0) trace_raw(p.trace, 6, gtimer());
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
// SMs the resident kernels may occupy (pg_set_sm_limit): every CTA spins on a grid barrier, so while another kernel that
// waits on PEER GPUs (an NCCL collective on a side stream) shares the device, some SMs must stay out of the grid's reach.
static int g_sm_limit = 0;
void set_sm_limit(int n) { g_sm_limit = n; }
static int res_sms() {
  const int n = num_sms();
  return g_sm_limit > 0 && g_sm_limit < n ? g_sm_limit : n;
}

struct ResPlan {
  TcPlan pl;
  int grid, slots, splits, kps;
  size_t smem, ws_bytes;
  uint32_t table_off;
};

// Tile width BN and K-split S are chosen by a small cost model (microseconds): a CTA moves (16 KB + BN * 128 B) per k-step
// from L2 at ~120 GB/s (TMA-latency-bound ring), the whole grid at most ~7 TB/s, and the epilogue costs per unit.
static bool make_res_plan(const PgConvDesc* d, const PgFusedNorm* fn, bool twin, ResPlan& rp) {
  (void)twin;
  TcParams& p = rp.pl.p;
  memset(&p, 0, sizeof(p));
  p.mode = d->mode; p.stride = d->stride; p.pad = d->pad; p.B = d->B;
  if (d->mode == PG_CONVT) { p.Ha = d->Hin; p.Wa = d->Win; p.ntaps = 4; }
  else if (d->mode == PG_CONV1X1) { p.Ha = d->Hout; p.Wa = d->Wout; p.ntaps = 1; }
  else { p.Ha = d->Hout; p.Wa = d->Wout; p.ntaps = 16; }
  p.Hout = d->Hout; p.Wout = d->Wout;
  p.TW = pow2_ceil(p.Wa); if (p.TW > 128) p.TW = 128;
  p.TH = pow2_ceil(p.Ha); if (p.TH > 128 / p.TW) p.TH = 128 / p.TW;
  p.TB = 128 / (p.TW * p.TH);
  p.lgTW = ilog2(p.TW); p.lgTH = ilog2(p.TH);
  p.nx = (p.Wa + p.TW - 1) / p.TW; p.ny = (p.Ha + p.TH - 1) / p.TH;
  if (p.TW * p.TH < 2) return false;                       // one lattice point per image: nothing to normalise over
  if (p.TW > 256 || p.TH > 256 || p.TB > 256) return false;
  if (d->out_f32 != PG_F16 && d->out_f32 != PG_BF16) return false;
  if (d->ldo < d->N) return false;
  const int nb = (p.B + p.TB - 1) / p.TB;
  int bk = 64;
  while (bk > 16 && ((d->C1 % bk) != 0 || (d->C2 % bk) != 0)) bk >>= 1;
  if ((d->C1 % bk) != 0 || (d->C2 % bk) != 0) return false;
  p.BK = bk; p.nk1 = d->C1 / bk; p.nk2 = d->C2 / bk; p.Ctot = d->C1 + d->C2;
  const int ncls = d->mode == PG_CONVT ? 4 : 1;
  const long long mt = (long long)p.nx * p.ny * nb * ncls;
  const int ksteps = p.ntaps * (p.nk1 + p.nk2);
  const int sms = res_sms();
  const int swz = bk * 2;
  const double a_bytes = 128.0 * swz;
  int best_bn = 0, best_s = 0;
  double best = 1e30;
  // (experiment knobs: PG_RES_BN / PG_RES_S restrict the candidates)
  static const int force_bn = [] { const char* e = getenv("PG_RES_BN"); return e ? atoi(e) : 0; }();
  static const int force_s = [] { const char* e = getenv("PG_RES_S"); return e ? atoi(e) : 0; }();
  for (int bn = 128; bn >= 16; bn >>= 1) {
    if (force_bn > 0 && bn != force_bn && (d->N % force_bn) == 0) continue;
    if ((d->N % bn) != 0 || (fn->kind == PG_FUSED_BWD && (fn->n_norm % bn) != 0)) continue;
    const long long tiles = mt * (d->N / bn);
    if (tiles >= (1LL << 30)) continue;
    const double step_bytes = a_bytes + (double)bn * swz;
    for (int S = 1; S <= 32; S <<= 1) {
      if (force_s > 0 && S != force_s) continue;
      double us;
      if (S == 1) {
        const int grid = (int)(tiles < sms ? tiles : sms);
        const int slots = (int)((tiles + grid - 1) / grid);
        if (slots > RES_MAX_SLOTS || slots * bn > 512) continue;
        if ((long long)slots * p.TB * bn > RES_TABLE_MAX) continue;
        const double cta = slots * ksteps * step_bytes, all = (double)tiles * ksteps * step_bytes;
        const double load = cta / 120e3 > all / 7e6 ? cta / 120e3 : all / 7e6;      // bytes / (bytes per us)
        us = load + 0.5 * slots * (bn / 16) / RES_GROUPS + 3.0;
      } else {
        if (tiles * S > sms || (ksteps % S) != 0 || ksteps / S < 2 || bn / 16 > RES_GROUPS * S) continue;
        if (fn->ws == nullptr || (size_t)tiles * S * 128 * bn * 4 > (size_t)fn->ws_bytes) continue;
        if (RES_GROUPS * 16 * p.TB > RES_TABLE_MAX) continue;
        const double cta = (double)(ksteps / S) * step_bytes, all = (double)tiles * ksteps * step_bytes;
        const double load = cta / 120e3 > all / 7e6 ? cta / 120e3 : all / 7e6;
        us = load + 0.3 * (bn / 16) / RES_GROUPS + 0.15 * S + 7.0;       // dump + S-plane sum + second barrier
      }
      if (us < best) { best = us; best_bn = bn; best_s = S; }
    }
  }
  if (best_bn == 0) return false;
  const int bn = best_bn;
  const long long tiles = mt * (d->N / bn);
  rp.splits = best_s;
  rp.kps = ksteps / best_s;
  int grid, slots;
  if (best_s == 1) {
    grid = (int)(tiles < sms ? tiles : sms);
    slots = (int)((tiles + grid - 1) / grid);
  } else {
    grid = (int)tiles * best_s;
    slots = 1;
  }
  rp.ws_bytes = best_s > 1 ? (size_t)tiles * best_s * 128 * bn * 4 : 0;
  p.BN = bn;
  p.N = d->N; p.ldo = d->ldo; p.n_valid = d->n_valid; p.act = fn->act; p.out_f32 = d->out_f32;
  p.splits = 1; p.nacc = 1;
  rp.pl.swz = swz;
  p.layout_type = swz == 128 ? 2u : (swz == 64 ? 4u : 6u);
  p.sbo = 8u * swz;
  const uint32_t fmt = d->in_dtype == PG_F16 ? 0u : 1u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.a_bytes = 128u * swz;
  p.b_bytes = ((uint32_t)bn * swz + 1023u) & ~1023u;
  p.tx_bytes = 128u * swz + (uint32_t)bn * swz;
  const uint32_t per_stage = p.a_bytes + p.b_bytes;
  const uint32_t table = (uint32_t)RES_TABLE_MAX * 16u;
  int stages = (int)((200u * 1024u - table) / per_stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > rp.kps * slots) stages = rp.kps * slots;
  if (stages < 1) return false;
  p.stages = stages;
  p.kps = rp.kps;
  uint32_t region = ((uint32_t)stages * per_stage + 1023u) & ~1023u;
  if (region < 120u * 1024u) region = 120u * 1024u;        // > half of the SM's shared memory: ONE CTA per SM, always
  rp.table_off = region;
  rp.smem = (size_t)region + table + 1024;
  int tcols = 32;
  while (tcols < slots * bn) tcols <<= 1;
  p.tmem_cols = (uint32_t)tcols;
  p.pers_total = (int)tiles;
  p.pers_mtiles = (int)(mt / ncls);
  p.pers_ntiles = d->N / bn;
  rp.grid = grid; rp.slots = slots;
  return true;
}

bool conv_res_supported(const PgConvDesc* d, const PgFusedNorm* fn, bool twin) {
  if (!tc_device_ok()) return false;
  ResPlan rp;
  return make_res_plan(d, fn, twin, rp);
}

int conv_res_launch(const PgConvDesc* d, const void* src1, const void* src2, const void* w, void* out, void* out2,
                    const PgFusedNorm* fn, cudaStream_t stream) {
  ResPlan rp;
  if (!make_res_plan(d, fn, out2 != nullptr, rp)) {
    set_error("conv_res: unsupported shape (does the layer's output fit tensor memory?)");
    return PG_ERR_UNSUPPORTED;
  }
  TcParams& p = rp.pl.p;
  p.out = out; p.out2 = out2;
  p.trace = g_trace;
  static const int skip = [] { const char* e = getenv("PG_TC_SKIP"); return e ? atoi(e) : 0; }();
  p.debug = skip;
  ResParams r;
  memset(&r, 0, sizeof(r));
  r.kind = fn->kind;
  r.n_norm = fn->kind == PG_FUSED_BWD ? fn->n_norm : d->N;
  r.cn = r.n_norm;
  r.sync = fn->sync; r.sums = fn->sums; r.bsums = fn->bsums;
  r.inv_hw = 1.0f / ((float)d->Hout * (float)d->Wout);
  r.drop_p = fn->drop_p;
  r.keep_scale = fn->drop_p > 0.f ? 1.f / (1.f - fn->drop_p) : 1.f;
  r.seed = (const unsigned long long*)fn->seed; r.salt = fn->salt;
  r.xhat = fn->xhat; r.xhat_ld = fn->xhat_ld;
  r.y = fn->y; r.y_ld = fn->y_ld; r.y_dt = fn->y_dtype;
  r.dskip = fn->dskip; r.dskip_ld = fn->dskip_ld;
  r.table_off = rp.table_off;
  r.splits = rp.splits; r.kps = rp.kps;
  r.ws = (float*)fn->ws;
  const bool phased = d->mode == PG_CONV && d->stride == 2;
  ActMaps mA;
  CUtensorMap mB;
  memset(&mA, 0, sizeof(mA));
  for (int src = 0; src < (d->C2 > 0 ? 2 : 1); ++src) {
    const void* base = src ? src2 : src1;
    const int C = src ? d->C2 : d->C1, ld = src ? d->ld2 : d->ld1;
    for (int ph = 0; ph < (phased ? 4 : 1); ++ph)
      if (int e = encode_act_map(&mA.m[src * 4 + ph], base, C, ld, d->B, d->Hin, d->Win, p.BK, p.TW, p.TH, p.TB,
                                 phased ? ph : -1, rp.pl.swz, d->in_dtype))
        return e;
  }
  {
    const int wtaps = d->mode == PG_CONV1X1 ? 1 : 16;
    cuuint64_t dims[2] = {(cuuint64_t)wtaps * p.Ctot, (cuuint64_t)d->N};
    cuuint64_t strides[1] = {(cuuint64_t)(d->mode == PG_CONV1X1 && d->ldw > 0 ? d->ldw : wtaps * p.Ctot) * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.BK, (cuuint32_t)p.BN};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = rp.pl.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (rp.pl.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult cr = get_encode()(&mB, d->in_dtype == PG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                               const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", wtaps * p.Ctot, d->N, (int)cr);
      return PG_ERR_CUDA;
    }
  }
  static bool smem_set = false;
  if (!smem_set) {
    PG_CUDA(cudaFuncSetAttribute(conv_res_kernel<PG_FUSED_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    PG_CUDA(cudaFuncSetAttribute(conv_res_kernel<PG_FUSED_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    smem_set = true;
  }
  static const bool dbg = getenv("PG_TC_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "conv_res: kind %d grid %d slots %d splits %d BN %d BK %d stages %d tmem %u smem %zu TW %d TH %d TB %d ksteps %d tiles %d\n",
            r.kind, rp.grid, rp.slots, rp.splits, p.BN, p.BK, p.stages, p.tmem_cols, rp.smem, p.TW, p.TH, p.TB,
            p.ntaps * (p.nk1 + p.nk2), p.pers_total);
  if (r.kind == PG_FUSED_FWD)
    conv_res_kernel<PG_FUSED_FWD><<<rp.grid, RES_THREADS, rp.smem, stream>>>(mA, mB, p, r);
  else
    conv_res_kernel<PG_FUSED_BWD><<<rp.grid, RES_THREADS, rp.smem, stream>>>(mA, mB, p, r);
  return check_launch("conv_res_kernel");
}
