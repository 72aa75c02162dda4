// CUDA-core implicit-GEMM 4x4 convolutions (any shape).  This is the validation / tail path: the
// tensor-core path lives in conv_tc.cu.  bf16 operands, fp32 accumulation, NHWC.
#include "common.cuh"

namespace pg {

struct ConvK {
  const bf16* src1;
  const bf16* src2;
  const bf16* w;
  const float* bias;
  void* out;
  void* out2;
  int mode, stride, pad, B, Hin, Win, Hout, Wout, C1, C2, ld1, ld2, N, ldo, n_valid, act, out_f32, in_dt;
  long long wrow;   // elements between the weight rows of consecutive n
  int Ha, Wa;  // lattice the tiles walk: PG_CONV -> (Hout, Wout); PG_CONVT -> (Hin, Win) per parity class
  long long M;
};

constexpr int TM = 64, TN = 64, TK = 32;

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvK p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int py = blockIdx.z >> 1, px = blockIdx.z & 1;
  const int ntaps = p.mode == PG_CONVT ? 4 : (p.mode == PG_CONV1X1 ? 1 : 16);
  const int Ctot = p.C1 + p.C2;

  const int lrow = tid >> 2, lkq = tid & 3;
  const long long lm = m0 + lrow;
  const bool mvalid = lm < p.M;
  int lb = 0, la = 0, lbb = 0;
  if (mvalid) {
    lbb = (int)(lm % p.Wa);
    long long r = lm / p.Wa;
    la = (int)(r % p.Ha);
    lb = (int)(r / p.Ha);
  }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < ntaps; ++t) {
    int iy, ix, wtap;
    if (p.mode == PG_CONVT) {
      const int j = t >> 1, i = t & 1;
      const int kh = (1 - py) + 2 * j, kw = (1 - px) + 2 * i;
      iy = la + py - j;
      ix = lbb + px - i;
      wtap = kh * 4 + kw;
    } else if (p.mode == PG_CONV1X1) {
      iy = la; ix = lbb; wtap = 0;
    } else {
      const int kh = t >> 2, kw = t & 3;
      iy = la * p.stride - p.pad + kh;
      ix = lbb * p.stride - p.pad + kw;
      wtap = t;
    }
    const bool inb = mvalid && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
    const long long pix = ((long long)lb * p.Hin + iy) * p.Win + ix;
    for (int c0 = 0; c0 < Ctot; c0 += TK) {
      const int c = c0 + lkq * 8;
      uint4 av = make_uint4(0, 0, 0, 0), wv = make_uint4(0, 0, 0, 0);
      if (inb && c < Ctot) {
        if (c < p.C1) av = *reinterpret_cast<const uint4*>(p.src1 + pix * p.ld1 + c);
        else av = *reinterpret_cast<const uint4*>(p.src2 + pix * p.ld2 + (c - p.C1));
      }
      if (n0 + lrow < p.N && c < Ctot)
        wv = *reinterpret_cast<const uint4*>(p.w + (long long)(n0 + lrow) * p.wrow + (long long)wtap * Ctot + c);
      float af[8], wf[8];
      unpack8dt(av, p.in_dt, af);
      unpack8dt(wv, p.in_dt, wf);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        As[lkq * 8 + j][lrow] = af[j];
        Bs[lkq * 8 + j][lrow] = wf[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int bb = (int)(m % p.Wa);
    const long long r = m / p.Wa;
    const int a = (int)(r % p.Ha);
    const int b = (int)(r / p.Ha);
    int oy = a, ox = bb;
    if (p.mode == PG_CONVT) {
      oy = 2 * a + py;
      ox = 2 * bb + px;
    }
    const long long opix = ((long long)b * p.Hout + oy) * p.Wout + ox;
    const int n = n0 + tx * 4;
    if (n >= p.N || n >= p.ldo) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j];
      if (p.bias != nullptr && n + j < p.n_valid) x += p.bias[n + j];
      x = act_apply(p.act, x);
      v[j] = (n + j < p.n_valid) ? x : 0.f;
    }
    if (p.out_f32 == PG_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * p.ldo + n) =
          make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint2 u;
      u.x = (uint32_t)to16(v[0], p.out_f32) | ((uint32_t)to16(v[1], p.out_f32) << 16);
      u.y = (uint32_t)to16(v[2], p.out_f32) | ((uint32_t)to16(v[3], p.out_f32) << 16);
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + opix * p.ldo + n) = u;
      if (p.out2 != nullptr) {
        u.x = (uint32_t)to16(v[0], PG_BF16) | ((uint32_t)to16(v[1], PG_BF16) << 16);
        u.y = (uint32_t)to16(v[2], PG_BF16) | ((uint32_t)to16(v[3], PG_BF16) << 16);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out2) + opix * p.ldo + n) = u;
      }
    }
  }
}

int conv_fwd_simt(const PgConvDesc* d, const void* src1, const void* src2, const void* w, const float* bias,
                  void* out, void* out2, cudaStream_t stream) {
  ConvK p;
  p.src1 = (const bf16*)src1;
  p.src2 = (const bf16*)src2;
  p.w = (const bf16*)w;
  p.bias = d->has_bias ? bias : nullptr;
  p.out = out;
  p.out2 = out2;
  p.mode = d->mode; p.stride = d->stride; p.pad = d->pad; p.B = d->B; p.Hin = d->Hin; p.Win = d->Win;
  p.Hout = d->Hout; p.Wout = d->Wout; p.C1 = d->C1; p.C2 = d->C2; p.ld1 = d->ld1; p.ld2 = d->ld2;
  p.N = d->N; p.ldo = d->ldo; p.n_valid = d->n_valid; p.act = d->act; p.out_f32 = d->out_f32;
  p.in_dt = d->in_dtype;
  p.wrow = d->mode == PG_CONV1X1 ? (d->ldw > 0 ? d->ldw : d->C1 + d->C2) : 16LL * (d->C1 + d->C2);
  if (d->mode == PG_CONVT) { p.Ha = d->Hin; p.Wa = d->Win; } else { p.Ha = d->Hout; p.Wa = d->Wout; }
  p.M = (long long)d->B * p.Ha * p.Wa;
  dim3 grid((unsigned)((p.M + TM - 1) / TM), (unsigned)((d->N + TN - 1) / TN), d->mode == PG_CONVT ? 4 : 1);
  conv_simt_kernel<<<grid, 256, 0, stream>>>(p);
  return check_launch("conv_simt_kernel");
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  dw[n*ld_n + c*16 + t] += sum_m g[m][n] * a[pix(m,t)][c]
// ---------------------------------------------------------------------------------------------
struct WgradK {
  const bf16* a;
  const bf16* g;
  float* dw;
  int stride, pad, B, Hin, Win, Hout, Wout, C, lda, N, ldg, ld_n, n_real, c_real, a_dt, g_dt, pointwise, ld_c;
  long long ld_t;
  long long M;
  int chunk;  // pixels per split
  int ctiles;
};

constexpr int WK = 32;

__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradK p) {
  __shared__ float Gs[WK][64 + 4];
  __shared__ float As[WK][64 + 4];
  const int tid = threadIdx.x;
  const int n0 = (blockIdx.x / p.ctiles) * 64;
  const int c0 = (blockIdx.x % p.ctiles) * 64;
  const int t = blockIdx.y;
  const int kh = t >> 2, kw = t & 3;
  const long long mbeg = (long long)blockIdx.z * p.chunk;
  long long mend = mbeg + p.chunk;
  if (mend > p.M) mend = p.M;
  const int lk = tid >> 3, lq = tid & 7;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long mb = mbeg; mb < mend; mb += WK) {
    const long long m = mb + lk;
    uint4 gv = make_uint4(0, 0, 0, 0), av = make_uint4(0, 0, 0, 0);
    if (m < mend) {
      const int ox = (int)(m % p.Wout);
      const long long r = m / p.Wout;
      const int oy = (int)(r % p.Hout);
      const int b = (int)(r / p.Hout);
      const int n = n0 + lq * 8;
      if (n < p.N) gv = *reinterpret_cast<const uint4*>(p.g + m * p.ldg + n);
      const int iy = p.pointwise ? oy : oy * p.stride - p.pad + kh, ix = p.pointwise ? ox : ox * p.stride - p.pad + kw;
      const int c = c0 + lq * 8;
      if (c < p.C && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win)
        av = *reinterpret_cast<const uint4*>(p.a + (((long long)b * p.Hin + iy) * p.Win + ix) * p.lda + c);
    }
    float gf[8], af[8];
    unpack8dt(gv, p.g_dt, gf);
    unpack8dt(av, p.a_dt, af);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      Gs[lk][lq * 8 + j] = gf[j];
      As[lk][lq * 8 + j] = af[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[k][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
      const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gg[i], aa[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.n_real) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= p.c_real) continue;
      atomicAdd(p.dw + (long long)n * p.ld_n + (long long)c * p.ld_c + (long long)t * p.ld_t, acc[i][j]);
    }
  }
}

int conv_wgrad_simt(const PgConvDesc* d, const void* a, const void* g, int ldg, float* dw, int ld_n, int n_real,
                    int c_real, int tap_major, int Cs, cudaStream_t stream) {
  WgradK p;
  p.a = (const bf16*)a; p.g = (const bf16*)g; p.dw = dw;
  p.stride = d->stride; p.pad = d->pad; p.B = d->B; p.Hin = d->Hin; p.Win = d->Win; p.Hout = d->Hout;
  p.Wout = d->Wout; p.C = d->C1; p.lda = d->ld1; p.N = d->N; p.ldg = ldg; p.ld_n = ld_n;
  p.n_real = n_real; p.c_real = c_real;
  p.a_dt = d->in_dtype; p.g_dt = d->out_f32;
  p.pointwise = d->mode == PG_CONV1X1 ? 1 : 0;
  p.ld_c = p.pointwise ? (d->ldw > 0 ? d->ldw : 16) : 16;
  const int taps = p.pointwise ? 1 : 16;
  p.ld_t = 1;
  if (tap_major) { p.ld_t = (long long)ld_n * Cs; p.ld_n = Cs; p.ld_c = 1; }   // S[tap][Ns = ld_n][Cs]
  p.M = (long long)d->B * d->Hout * d->Wout;
  const int ntiles = (d->N + 63) / 64;
  p.ctiles = (d->C1 + 63) / 64;
  const int base = ntiles * p.ctiles * taps;
  long long splits = (4LL * num_sms() + base - 1) / base;
  const long long maxsplits = (p.M + 255) / 256;
  if (splits > maxsplits) splits = maxsplits;
  if (splits < 1) splits = 1;
  long long chunk = (p.M + splits - 1) / splits;
  chunk = (chunk + WK - 1) / WK * WK;
  splits = (p.M + chunk - 1) / chunk;
  p.chunk = (int)chunk;
  dim3 grid(ntiles * p.ctiles, taps, (unsigned)splits);
  wgrad_simt_kernel<<<grid, 256, 0, stream>>>(p);
  return check_launch("wgrad_simt_kernel");
}

// ---------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ---------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const bf16* g, long long M, int ldg, int n_real, float* db, long long chunk) {
  // block: 256 threads = 32 rows x 8 column-groups of 8... keep it simple: thread = column, loop rows
  const int n = threadIdx.x;
  const long long mbeg = (long long)blockIdx.x * chunk;
  long long mend = mbeg + chunk;
  if (mend > M) mend = M;
  if (n >= n_real) return;
  float s = 0.f;
  for (long long m = mbeg; m < mend; ++m) s += __bfloat162float(g[m * ldg + n]);
  atomicAdd(db + n, s);
}

}  // namespace pg

__global__ void __launch_bounds__(256) colsum_vec_kernel(const pg::bf16* __restrict__ g, long long M, int ldg, int n_real,
                                                        float* db, long long rows_per_block) {
  // 8 channels per thread (16 B loads), consecutive threads on consecutive channel groups of a row
  __shared__ float sh[1024];
  const int cg = (n_real + 7) >> 3;
  const int pl = 256 / cg;
  const int my_cg = threadIdx.x % cg, my_pl = threadIdx.x / cg;
  for (int i = threadIdx.x; i < cg * 8; i += 256) sh[i] = 0.f;
  __syncthreads();
  const long long beg = (long long)blockIdx.x * rows_per_block;
  long long end = beg + rows_per_block;
  if (end > M) end = M;
  if (my_pl < pl) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (long long m = beg + my_pl; m < end; m += pl) {
      float f[8];
      pg::unpack8(*reinterpret_cast<const uint4*>(g + m * ldg + my_cg * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
    if ((cg & (cg - 1)) == 0 && cg <= 32) {
      // lanes that share a channel group sit cg apart: shuffle them together first (256 threads adding to the same 8
      // shared words, as with cg = 1 for the one-channel bias, serialised for ~20 us)
      for (int off = cg; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += __shfl_xor_sync(0xffffffffu, a[j], off);
      }
      if ((threadIdx.x & 31) < cg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sh[my_cg * 8 + j], a[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sh[my_cg * 8 + j], a[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_real; i += 256) atomicAdd(db + i, sh[i]);
}

extern "C" int pg_colsum(const void* g, int64_t M, int32_t ldg, int32_t n_real, float* db, void* stream) {
  PG_REQUIRE(n_real >= 1 && n_real <= 1024, "pg_colsum: n_real=%d out of range", n_real);
  if (M <= 0) return PG_OK;
  if (ldg % 8 == 0 && ((n_real + 7) / 8) * 8 <= ldg && (((uintptr_t)g) & 15) == 0) {
    long long blocks = 4LL * pg::num_sms();
    long long rpb = (M + blocks - 1) / blocks;
    if (rpb < 64) rpb = 64;
    blocks = (M + rpb - 1) / rpb;
    colsum_vec_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const pg::bf16*)g, M, ldg, n_real, db, rpb);
    return pg::check_launch("colsum_vec_kernel");
  }
  long long chunk = 512;
  long long blocks = (M + chunk - 1) / chunk;
  int threads = ((n_real + 31) / 32) * 32;
  pg::colsum_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>((const pg::bf16*)g, M, ldg, n_real, db,
                                                                          chunk);
  return pg::check_launch("colsum_kernel");
}
