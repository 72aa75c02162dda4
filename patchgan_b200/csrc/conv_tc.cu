// Implicit-GEMM 4x4 convolutions on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, tiled mode, zero OOB fill; stride-2 layers through four phase-split views) -> swizzled smem
//   -> tcgen05.mma (cta_group::1, kind::f16, f16 x f16 or bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
// No im2col of an NHWC activation is ever materialised: for every (tap, channel-chunk) k-step the A operand is ONE TMA box
// of the activation tensor.  (The only im2col in the package is layout.cu's pg_im2col_s2 over the 3 / 4-channel NCHW float
// IMAGES in front of the two first layers: it doubles as the NCHW -> operand-layout conversion that is needed anyway, see
// DESIGN.md section 4.4 for the byte count.)
//   PG_CONV  : stride 1: box {BK ch, TW, TH, TB} at (c, ox0-pad+kw, oy0-pad+kh, b0).  Stride 2: the input is viewed as
//              four phase tensors X[b][2y'+ry][2x'+rx][c] (one tensor map each); tap (kh,kw) with u = kh-pad, v = kw-pad
//              is the STRIDE-1 box at (c, ox0+floor(v/2), oy0+floor(u/2), b0) of phase (u&1, v&1).  (Element-strided
//              boxes do the same job but the TMA engine walks them at ~3.7 ns per row, 35 GB/s per SM -- measured with
//              tools/tma_probe.cu -- which made every stride-2 layer TMA-bound.)
//   PG_CONVT : per output-parity class (py,px) a stride-1 box {BK, TW, TH, TB} at (c, x0+px-i, y0+py-j, b0)
//   PG_CONV1X1: pointwise, one "tap" (first-layer im2col GEMMs and the tap products of the one-real-channel layers)
// padding = TMA out-of-bounds zero fill; the skip concat = a second tensor map (K loop walks src1 then src2).
// The B operand is a 2D box {BK, BN} of the packed weights [N][16*Ctot].
//
// Kernels in this file:
//   conv_tc_kernel       one output tile per CTA, up to 4 CTAs per SM; optional split-K over a CTA cluster
//   conv_tc_pers_kernel  persistent: two CTAs per SM walk the tiles, double-buffered TMEM accumulators; <.., PAIR = true>:
//                        clusters of two CTAs, tcgen05.mma.cta_group::2 on 256-pixel tiles, 8 epilogue warps
//   wgrad_tc_kernel      weight gradient (MN-major operands, split-K over pixel tiles, TMA bulk-reduce epilogue)
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (each owns the TMEM lane quarter warp_id % 4): bias / activation / activation-gradient /
// InstanceNorm statistics / TMA store.
// Reference ops replaced: aten::convolution under nn.Conv2d / nn.ConvTranspose2d (unet.py:19,53; disc.py:19-45)
// and their dgrad / wgrad.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace pg {

// --------------------------------------------------------------------------------------------
// PTX wrappers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_tc: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// CTA-pair (cta_group::2) forms: the destination is this CTA's shared memory, the mbarrier may live in the peer CTA
// (the pair's leader collects the bytes of both halves of a stage)
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// shared::cluster address of `addr` (an address in this CTA's shared memory) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync_grp(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA pair: D[256 x N] over the two SMs' tensor cores; each CTA's shared memory holds its 128 rows of A and N/2 rows of B
// at the SAME offsets, each CTA's TMEM its 128 rows of D.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ... and the completion of the pair's MMAs arrives on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// arrives on the mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address [0,14), LBO [16,30),
// SBO [32,46), version=1 [46,48), layout type [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                      // LBO (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;       // stride between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

// --------------------------------------------------------------------------------------------
// kernel
// --------------------------------------------------------------------------------------------
struct TcParams {
  int mode, stride, pad;
  int B, Ha, Wa, Hout, Wout;
  int TW, TH, lgTW, lgTH, TB;
  int nx, ny;
  int BN, BK, nk1, nk2, ntaps, Ctot;
  int N, ldo, n_valid, act, out_f32;
  const float* bias;
  void* out;
  void* out2;
  uint32_t idesc, sbo, layout_type;
  int stages;
  uint32_t a_bytes, b_bytes, tx_bytes;
  uint32_t tmem_cols;
  int nacc;   // independent TMEM accumulators the MMAs rotate over (summed in the epilogue)
  int debug;  // timing experiments only: 1 = skip the MMAs, 2 = skip the TMA loads, 4 = skip the epilogue stores
  // TMA-store epilogue: the tile is staged in (reused) ring smem as 128 rows x st_rowbytes, swizzled, st_cw columns at a time
  int tma_store, st_rowbytes, st_cw, st_nbuf, st_twin;
  unsigned long long* trace;   // debug: 8 timestamps per CTA (pg_debug_set_trace), else null
  int pers_total, pers_mtiles, pers_ntiles;   // persistent variant: work items = (m-tile fastest, n-tile, class)
  int pers_defer, pers_nch;                   // deferred store drain across tiles; store chunks per tile
  uint32_t pers_stage_off, pers_slot_cols;    // staging buffers behind the ring; TMEM columns per accumulator slot
  int splits, kps;             // K-split cluster: `splits` CTAs (cluster dims (1,1,splits)) share one tile, kps k-steps each
  float* ws;                   // split-K exchange buffer in global memory (L2-resident): [tile][rank][128 rows][BN] fp32
  const void* mul_y;           // data-gradient fusion: out = acc * act'(y) with y = saved activation OUTPUT at the same pixel /
  int mul_ld, mul_dt;          //   channel (pixel stride mul_ld, PgDType mul_dt); the activation is p.act (not applied forward)
  float* stats;                // fused InstanceNorm statistics: sums[(b*N + n)*2 + {0,1}] += {x, x^2} over the tile (or null)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_raw(unsigned long long* tr, int slot, unsigned long long v) {
  if (tr != nullptr) {
    const unsigned long long cta = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
    tr[cta * 16 + slot] = v;
  }
}
__device__ __forceinline__ void trace_val(const TcParams& p, int slot, unsigned long long v) {
  if (p.trace != nullptr) {
    const unsigned long long cta = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
    p.trace[cta * 16 + slot] = v;
  }
}
__device__ __forceinline__ void trace_put(const TcParams& p, int slot) {
  if (p.trace != nullptr) {
    const unsigned long long cta = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
    p.trace[cta * 16 + slot] = gtimer();
  }
}

constexpr int TC_THREADS = 192;
#ifndef EPI_WIDE
#define EPI_WIDE 0
#endif
constexpr int MAX_STAGES = 12;
constexpr int TC_MAX_DYN_SMEM = 224 * 1024;

struct ActMaps {
  CUtensorMap m[8];   // [source (0/1)][phase ry*2+rx]; stride-1 layers use phase 0 only
};

// --------------------------------------------------------------------------------------------
// epilogue (warps 2..5): TMEM -> registers -> bias / activation / padding mask -> global
// --------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_fast(float x) {
  if (ACT == PG_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PG_ACT_LEAKYRELU) return fmaxf(x, 0.2f * x);
  if (ACT == PG_ACT_TANH) {
    // 1 - 2 / (e^{2x} + 1): branch-free (ex2 + rcp), absolute error ~1e-7, saturates correctly at +-inf
    return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);
  }
  if (ACT == PG_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
  return x;
}

struct EpiCtx {
  uint32_t smem_base, tmem_acc;
  int x0, y0, b0, n0, py, px, cls;
  int seq0 = 0;      // store chunks this CTA has issued before this tile (persistent kernel with deferred drain), else 0
  int defer = 0;     // 1: do not wait for the tile's bulk stores here; the next use of a staging buffer (or the CTA's exit) does
  int grp = 0, ngrp = 1;   // epilogue warp groups (4 warps each; CTA-pair kernel: 2): group g takes the store chunks
                           // ch = g (mod ngrp) of every tile, owns staging buffer g and named barrier 1 + g
};

template <int ACT>
__device__ __forceinline__ float act_grad_out_fast(float y) {
  if (ACT == PG_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (ACT == PG_ACT_LEAKYRELU) return y > 0.f ? 1.f : 0.2f;
  if (ACT == PG_ACT_TANH) return 1.f - y * y;
  if (ACT == PG_ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

// f[0..16) *= act'(y[opix][n .. n+16))   (data-gradient through an activation, from its saved output)
template <int ACT>
__device__ __forceinline__ void epi_mul16(const TcParams& p, long long opix, int n, bool row_valid, float* f) {
  if (!row_valid) return;
  const unsigned short* yp = reinterpret_cast<const unsigned short*>(p.mul_y) + opix * p.mul_ld + n;
  const uint4 y0 = __ldg(reinterpret_cast<const uint4*>(yp));
  const uint4 y1 = __ldg(reinterpret_cast<const uint4*>(yp) + 1);
  float yf[16];
  if (p.mul_dt == PG_F16) { unpack8h(y0, yf); unpack8h(y1, yf + 8); } else { unpack8(y0, yf); unpack8(y1, yf + 8); }
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] *= act_grad_out_fast<ACT>(yf[j]);
}

// bias / activation / padding mask on U raw sums v[] of channels [n, n+U) -> f[]
template <int ACT, int U>
__device__ __forceinline__ void epi_finish(const TcParams& p, int n, uint32_t* v, float* f) {
  if (p.bias != nullptr) {
    if (n + U <= p.n_valid) {
#pragma unroll
      for (int j = 0; j < U; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        v[j] = __float_as_uint(__uint_as_float(v[j]) + b4.x);
        v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b4.y);
        v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b4.z);
        v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b4.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (n + j < p.n_valid) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + n + j));
    }
  }
  if (p.mul_y != nullptr) {      // (p.act names the activation whose derivative multiplies the result: epi_mul16)
#pragma unroll
    for (int j = 0; j < U; ++j) f[j] = __uint_as_float(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < U; ++j) f[j] = act_fast<ACT>(__uint_as_float(v[j]));
  }
  if (n + U > p.n_valid) {       // zero the padded output channels
#pragma unroll
    for (int j = 0; j < U; ++j) f[j] = (n + j < p.n_valid) ? f[j] : 0.f;
  }
}

// U accumulator columns [c, c+U) of this thread's row, summed over the rotating accumulators -> raw sums v[]
template <int U>
__device__ __forceinline__ void epi_tmem(const TcParams& p, uint32_t trow, int c, uint32_t* v) {
#pragma unroll
  for (int i = 0; i < U; i += 16) tmem_ld16(trow + (uint32_t)(c + i), v + i);
  if (U == 16 && p.nacc == 2) {           // the loads of all rotating accumulators in flight under ONE wait
    uint32_t w[16];
    tmem_ld16(trow + (uint32_t)(p.BN + c), w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    return;
  }
  if (U == 16 && p.nacc == 4) {           // two loads per wait (three would not fit the 80-register budget of occupancy 4)
    uint32_t w[16];
    tmem_ld16(trow + (uint32_t)(p.BN + c), w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    uint32_t w2[16];
    tmem_ld16(trow + (uint32_t)(2 * p.BN + c), w);
    tmem_ld16(trow + (uint32_t)(3 * p.BN + c), w2);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + (__uint_as_float(w[j]) + __uint_as_float(w2[j])));
    return;
  }
  tmem_ld_wait();
  for (int a = 1; a < p.nacc; ++a) {
#pragma unroll
    for (int i = 0; i < U; i += 16) {
      uint32_t w[16];
      tmem_ld16(trow + (uint32_t)(a * p.BN + c + i), w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) v[i + j] = __float_as_uint(__uint_as_float(v[i + j]) + __uint_as_float(w[j]));
    }
  }
}

template <int ACT, int U>
__device__ __forceinline__ void epi_load(const TcParams& p, uint32_t trow, int c, int n, float* f) {
  uint32_t v[U];
  epi_tmem<U>(p, trow, c, v);
  epi_finish<ACT, U>(p, n, v, f);
}

// Fused InstanceNorm statistics (unet.py:20,55): every epilogue thread holds 16 channels of ONE output pixel, the 32
// lanes of a warp hold 32 pixels of the same image.  A halving butterfly (8+4+2+1+1 shuffles per quantity instead of
// 16 x 5) leaves each even lane with the warp's sum of one channel, added to sums[(b*N + n)*2 + {0,1}] with red.add.
__device__ __forceinline__ void stats_add16(const float* f, bool row_valid, float* sums_bn, int lane) {
  float a[16], q[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { a[j] = row_valid ? f[j] : 0.f; q[j] = a[j] * a[j]; }
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
  if ((lane & 1) == 0) {
    const int col = (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0);
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(sums_bn + col * 2), "f"(a1), "f"(q1) : "memory");
  }
}

template <int ACT, int U>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const ActMaps& mapsO, const EpiCtx& e) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int r = q * 32 + lane;            // accumulator row == lattice point inside the tile
  const uint32_t trow = e.tmem_acc + ((uint32_t)(q * 32) << 16);
  // fused statistics (U == 16 only): image of this warp's 32 rows, validity of this thread's row
  const int s_b = e.b0 + (r >> (p.lgTW + p.lgTH));
  const bool s_valid = s_b < p.B && e.y0 + ((r >> p.lgTW) & (p.TH - 1)) < p.Ha && e.x0 + (r & (p.TW - 1)) < p.Wa;
  long long m_opix = 0;           // output pixel of this thread's row (activation-gradient fusion)
  if (p.mul_y != nullptr) {
    int oy = e.y0 + ((r >> p.lgTW) & (p.TH - 1)), ox = e.x0 + (r & (p.TW - 1));
    if (p.mode == PG_CONVT) { oy = 2 * oy + e.py; ox = 2 * ox + e.px; }
    m_opix = ((long long)s_b * p.Hout + oy) * p.Wout + ox;
  }
  if (p.tma_store) {
    // ---- stage the tile in the (now idle) ring smem with the TMA swizzle, store it with cp.async.bulk.tensor:
    //      full 128-byte rows instead of 32 scattered 16-byte stores per warp instruction; tails are clipped by TMA
    const int et = (threadIdx.x - 64) & 127;             // 0..127 among the epilogue threads of this group
    const uint32_t bufbytes = 128u * (uint32_t)p.st_rowbytes;
    const int sh = p.st_rowbytes == 128 ? 0 : (p.st_rowbytes == 64 ? 1 : 2);
    const uint32_t xr = (uint32_t)(r >> sh) & (uint32_t)((p.st_rowbytes >> 4) - 1);   // swizzle XOR of this row
    const int cls = e.cls;
    const int nch = p.BN / p.st_cw;
    for (int ch = e.grp; ch < nch; ch += e.ngrp) {
      int buf;
      if (e.ngrp > 1) {
        // one staging buffer per group (st_nbuf == ngrp): its previous store (this group's previous chunk, possibly of the
        // previous tile: seq0 = tiles done) must have been read; the other group keeps working meanwhile
        buf = e.grp;
        if (e.seq0 > 0 || ch >= e.ngrp) {
          if (et == 0) bulk_wait_read<0>();
          epi_bar_sync_grp(e.grp);
        }
      } else {
        const int g = e.seq0 + ch;                        // running chunk number: bulk groups complete in this order
        buf = g % p.st_nbuf;
        if (g >= p.st_nbuf) {                             // the buffer's previous store must have been read
          if (et == 0) { if (p.st_nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
          epi_bar_sync();
        }
      }
      const uint32_t prim = e.smem_base + (uint32_t)buf * bufbytes + (uint32_t)r * p.st_rowbytes;
      const uint32_t twin = e.smem_base + (uint32_t)(p.st_nbuf + buf) * bufbytes + (uint32_t)r * p.st_rowbytes;
      for (int sub = 0; sub < p.st_cw; sub += U) {
        float f[U];
        epi_load<ACT, U>(p, trow, ch * p.st_cw + sub, e.n0 + ch * p.st_cw + sub, f);
        if (U == 16 && p.mul_y != nullptr) epi_mul16<ACT>(p, m_opix, e.n0 + ch * p.st_cw + sub, s_valid, f);
        if (U == 16 && p.stats != nullptr && s_b < p.B)
          stats_add16(f, s_valid, p.stats + ((long long)s_b * p.N + e.n0 + ch * p.st_cw + sub) * 2, lane);
        if (p.out_f32 == PG_F32) {
          const uint32_t u0 = (uint32_t)(sub * 4) >> 4;
#pragma unroll
          for (int j = 0; j < U / 4; ++j) {
            uint4 v4 = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                                  __float_as_uint(f[4 * j + 3]));
            st_shared_v4(prim + (((u0 + j) ^ xr) << 4), v4);
          }
        } else {
          const uint32_t u0 = (uint32_t)(sub * 2) >> 4;
          if (p.out_f32 == PG_F16) {
#pragma unroll
            for (int j = 0; j < U / 8; ++j) st_shared_v4(prim + (((u0 + j) ^ xr) << 4), pack8h(f + 8 * j));
          } else {
#pragma unroll
            for (int j = 0; j < U / 8; ++j) st_shared_v4(prim + (((u0 + j) ^ xr) << 4), pack8(f + 8 * j));
          }
          if (p.st_twin) {
#pragma unroll
            for (int j = 0; j < U / 8; ++j) st_shared_v4(twin + (((u0 + j) ^ xr) << 4), pack8(f + 8 * j));
          }
        }
      }
      fence_proxy_async();
      if (e.ngrp > 1) epi_bar_sync_grp(e.grp); else epi_bar_sync();
      if (et == 0 && !(p.debug & 4)) {
        tma_store_4d(&mapsO.m[cls], e.smem_base + (uint32_t)buf * bufbytes, e.n0 + ch * p.st_cw, e.x0, e.y0, e.b0);
        if (p.st_twin)
          tma_store_4d(&mapsO.m[4 + cls], e.smem_base + (uint32_t)(p.st_nbuf + buf) * bufbytes, e.n0 + ch * p.st_cw, e.x0,
                       e.y0, e.b0);
        bulk_commit();
      }
    }
    if (et == 0 && !e.defer) bulk_wait_read<0>();
  } else {
    const int xl = r & (p.TW - 1);
    const int yl = (r >> p.lgTW) & (p.TH - 1);
    const int bl = r >> (p.lgTW + p.lgTH);
    const int b = e.b0 + bl, a = e.y0 + yl, bb = e.x0 + xl;
    const bool valid = b < p.B && a < p.Ha && bb < p.Wa;
    int oy = a, ox = bb;
    if (p.mode == PG_CONVT) { oy = 2 * a + e.py; ox = 2 * bb + e.px; }
    const long long opix = ((long long)b * p.Hout + oy) * p.Wout + ox;
    for (int c = e.grp * U; c < p.BN; c += U * e.ngrp) {
      float f[U];
      const int n = e.n0 + c;
      epi_load<ACT, U>(p, trow, c, n, f);
      if (U == 16 && p.mul_y != nullptr) epi_mul16<ACT>(p, m_opix, n, s_valid, f);
      if (U == 16 && p.stats != nullptr && s_b < p.B) stats_add16(f, s_valid, p.stats + ((long long)s_b * p.N + n) * 2, lane);
      const int keep = p.ldo - n;     // channels of this chunk that exist in the (possibly trimmed) output row
      if (valid && keep > 0 && !(p.debug & 4)) {
        if (p.out_f32 == PG_F32) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * p.ldo + n);
#pragma unroll
          for (int j = 0; j < U / 4; ++j)
            if (4 * j < keep) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + opix * p.ldo + n);
#pragma unroll
          for (int j = 0; j < U / 8; ++j)
            if (8 * j < keep) o[j] = pack8dt(f + 8 * j, p.out_f32);
          if (p.out2 != nullptr) {
            uint4* o2 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out2) + opix * p.ldo + n);
#pragma unroll
            for (int j = 0; j < U / 8; ++j)
              if (8 * j < keep) o2[j] = pack8(f + 8 * j);
          }
        }
      }
    }
  }
}

// ---- direct (non-TMA) store of 16 finished channels [n, n+16) of output pixel opix
__device__ __forceinline__ void store16(const TcParams& p, long long opix, int n, const float* f) {
  const int keep = p.ldo - n;     // channels of this chunk that exist in the (possibly trimmed) output row
  if (keep <= 0) return;
  if (p.out_f32 == PG_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix * p.ldo + n);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * j < keep) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  } else {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + opix * p.ldo + n);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (8 * j < keep) o[j] = pack8dt(f + 8 * j, p.out_f32);
    if (p.out2 != nullptr) {
      uint4* o2 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out2) + opix * p.ldo + n);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (8 * j < keep) o2[j] = pack8(f + 8 * j);
    }
  }
}

// --------------------------------------------------------------------------------------------
// Split-K over a thread-block cluster (the 2x2 .. 16x16 bottleneck layers: a handful of M tiles, K = 4096).
// The `splits` CTAs of a cluster (1,1,splits) own disjoint k-ranges of ONE output tile.  Each dumps its fp32 partial
// accumulators to its own shared memory, the cluster synchronises, and CTA r reduces the column slice
// [r*BN/splits, (r+1)*BN/splits) over all peers through distributed shared memory (ld.shared::cluster) and runs the
// normal epilogue (bias / activation / statistics / store) on it.  No atomics, no zero-fill, no extra pass.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// Exchange through L2: every CTA stores its partial tile (64-byte runs per thread) into its slot of a global scratch
// buffer, the cluster barrier (release / acquire at cluster scope) orders the stores before the owners' loads.
// (Distributed shared memory was tried first: ld.shared::cluster reached ~8 GB/s per SM and st.shared::cluster ~15 GB/s,
// 4-8x slower than the same bytes through L2.)
__device__ __forceinline__ void split_push(const TcParams& p, const EpiCtx& e, int rank, long long tile_id) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, r = q * 32 + lane;
  const uint32_t trow = e.tmem_acc + ((uint32_t)(q * 32) << 16);
  float* dst = p.ws + ((tile_id * p.splits + rank) * 128 + r) * p.BN;
  for (int c = 0; c < p.BN; c += 16) {
    uint32_t v[16];
    epi_tmem<16>(p, trow, c, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      __stcg(reinterpret_cast<float4*>(dst + c) + j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
  }
}

template <int ACT>
__device__ __forceinline__ void split_reduce_store(const TcParams& p, const EpiCtx& e, int rank, long long tile_id) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, r = q * 32 + lane;
  const int xl = r & (p.TW - 1);
  const int yl = (r >> p.lgTW) & (p.TH - 1);
  const int bl = r >> (p.lgTW + p.lgTH);
  const int b = e.b0 + bl, a = e.y0 + yl, bb = e.x0 + xl;
  const bool valid = b < p.B && a < p.Ha && bb < p.Wa;
  int oy = a, ox = bb;
  if (p.mode == PG_CONVT) { oy = 2 * a + e.py; ox = 2 * bb + e.px; }
  const long long opix = ((long long)b * p.Hout + oy) * p.Wout + ox;
  const int s_b = e.b0 + (r >> (p.lgTW + p.lgTH));
  const int w = p.BN / p.splits;
  const float* src0 = p.ws + ((tile_id * p.splits) * 128 + r) * p.BN;
  for (int c = rank * w; c < (rank + 1) * w; c += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int src = 0; src < p.splits; src += 2) {
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[j] = __ldcg(reinterpret_cast<const float4*>(src0 + (long long)src * 128 * p.BN + c) + j);
        t[4 + j] = __ldcg(reinterpret_cast<const float4*>(src0 + (long long)(src + 1) * 128 * p.BN + c) + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[4 * j] += t[j].x + t[4 + j].x; acc[4 * j + 1] += t[j].y + t[4 + j].y;
        acc[4 * j + 2] += t[j].z + t[4 + j].z; acc[4 * j + 3] += t[j].w + t[4 + j].w;
      }
    }
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(acc[j]);
    float f[16];
    const int n = e.n0 + c;
    epi_finish<ACT, 16>(p, n, v, f);
    if (p.mul_y != nullptr) epi_mul16<ACT>(p, opix, n, valid, f);
    if (p.stats != nullptr && s_b < p.B) stats_add16(f, valid, p.stats + ((long long)s_b * p.N + n) * 2, lane);
    if (valid && !(p.debug & 4)) store16(p, opix, n, f);
  }
}

// --------------------------------------------------------------------------------------------
// MMA issue loop (one thread).  A single thread retires roughly one dependent instruction every 4-6 cycles, so the
// instruction count per tcgen05.mma IS the issue rate (tools/mma_probe.cu: 150-190 cycles per MMA with a runtime modulo
// and descriptor rebuild in the loop, whatever N is).  Everything loop-invariant lives in registers, descriptors advance
// by adds, the accumulator rotation is a mask, KK = MMAs per k-step is a template parameter.
// --------------------------------------------------------------------------------------------
template <int KK>
__device__ __forceinline__ void mma_issue(const TcParams& p, uint64_t* full_bar, uint64_t* empty_bar, uint64_t* acc_bar,
                                          uint32_t a_base, uint32_t b_base, uint32_t tmem_acc, int ksteps) {
  const uint32_t idesc = p.idesc;
  const uint32_t stages = (uint32_t)p.stages;
  const uint32_t a_step = p.a_bytes >> 4, b_step = p.b_bytes >> 4;     // descriptor start-address units (16 B)
  const uint64_t desc_hi = make_smem_desc(0, p.sbo, p.layout_type);     // everything but the start address
  const uint32_t a_lo0 = (a_base & 0x3FFFF) >> 4, b_lo0 = (b_base & 0x3FFFF) >> 4;
  const uint32_t acc_wrap = (uint32_t)(p.nacc * p.BN);                  // nacc and BN are powers of two
  const uint32_t bn = (uint32_t)p.BN;
  const bool skip = (p.debug & 1) != 0;
  uint32_t a_lo = a_lo0, b_lo = b_lo0, stage = 0, phase = 0, acc_off = 0;
  uint32_t fresh = (uint32_t)p.nacc;          // MMAs that still start their accumulator (accumulate = 0)
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const bool tracing = p.trace != nullptr;
  long long waited = 0;
  for (int ks = 0; ks < ksteps; ++ks) {
    if (tracing) {
      const long long w0 = clock64();
      mbar_wait(full0 + stage * 8, phase);
      if (ks > 0) waited += clock64() - w0;
    } else {
      mbar_wait(full0 + stage * 8, phase);
    }
    tc_fence_after();
    if (ks == 0) trace_put(p, 2);
    if (!skip) {
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        // +32 bytes (16 elements) along K inside the swizzle row: start-address field += 2
        const uint32_t accum = fresh == 0 ? 1u : 0u;
        umma_bf16(tmem_acc + acc_off, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, accum);
        fresh -= accum ^ 1u;
        acc_off += bn;
        if (acc_off == acc_wrap) acc_off = 0;
      }
    }
    umma_commit(empty0 + stage * 8);
    a_lo += a_step; b_lo += b_step;
    if (++stage == stages) { stage = 0; phase ^= 1; a_lo = a_lo0; b_lo = b_lo0; }
  }
  umma_commit(smem_u32(acc_bar));
  if (tracing) trace_val(p, 8, (unsigned long long)waited);   // cycles the issuer waited for operands (after the first)
}

__global__ void __launch_bounds__(TC_THREADS, 4)
conv_tc_kernel(const __grid_constant__ ActMaps mapsA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ ActMaps mapsO, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dynamic smem base rounded up to 1024 B (SWIZZLE_128B atoms)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.stages * p.a_bytes;

  // tile coordinates
  const int tile = blockIdx.x;
  const int tx_i = tile % p.nx;
  const int ty_i = (tile / p.nx) % p.ny;
  const int tb_i = tile / (p.nx * p.ny);
  const int x0 = tx_i * p.TW, y0 = ty_i * p.TH, b0 = tb_i * p.TB;
  const int n0 = blockIdx.y * p.BN;
  const int split = p.splits > 1 ? (int)cluster_rank() : 0;        // cluster dims (1,1,splits): rank = blockIdx.z % splits
  const int cls = p.splits > 1 ? blockIdx.z / p.splits : blockIdx.z;
  const int py = cls >> 1, px = cls & 1;
  const int nk = p.nk1 + p.nk2;
  const int ks0 = split * p.kps, ks1 = ks0 + p.kps;                 // this CTA's k-steps (all of them without split-K)
  const long long tile_id = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * cls);
  if (threadIdx.x == 0) {
    trace_put(p, 0);
    if (p.trace != nullptr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const unsigned long long cta = blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
      p.trace[cta * 16 + 7] = smid;
    }
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapsA.m[0]);
    if (p.nk2 > 0) prefetch_tmap(&mapsA.m[4]);
    prefetch_tmap(&mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&acc_bar), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_sh), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;
  if (threadIdx.x == 0) trace_put(p, 1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long pwait = 0;
      const bool tracing = p.trace != nullptr;
      int t = ks0 / nk, ck = ks0 - t * nk;
      int cx = 0, cy = 0, wtap = 0, ph = 0;
      bool newtap = true;
      for (int ks = ks0; ks < ks1; ++ks) {
        if (newtap) {
          newtap = false;
          ph = 0;
          if (p.mode == PG_CONVT) {
            const int j = t >> 1, i = t & 1;
            wtap = ((1 - py) + 2 * j) * 4 + (1 - px) + 2 * i;
            cx = x0 + px - i;
            cy = y0 + py - j;
          } else if (p.mode == PG_CONV1X1) {
            wtap = 0; cx = x0; cy = y0;
          } else {
            const int kh = t >> 2, kw = t & 3;
            wtap = t;
            if (p.stride == 2) {
              const int u = kh - p.pad, v = kw - p.pad;       // input row = 2*oy + u
              ph = (u & 1) * 2 + (v & 1);
              cx = x0 + (v >> 1);                              // arithmetic shift = floor
              cy = y0 + (u >> 1);
            } else {
              cx = x0 - p.pad + kw;
              cy = y0 - p.pad + kh;
            }
          }
        }
        if (tracing) {
          const long long w0 = clock64();
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          pwait += clock64() - w0;
        } else {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        }
        const uint32_t fb = smem_u32(&full_bar[stage]);
        if (p.debug & 2) {
          mbar_expect_tx(fb, 0);
        } else {
          mbar_expect_tx(fb, p.tx_bytes);
          if (ck < p.nk1) tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[ph], fb, ck * p.BK, cx, cy, b0);
          else tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[4 + ph], fb, (ck - p.nk1) * p.BK, cx, cy, b0);
          tma_load_2d(b_base + stage * p.b_bytes, &mapB, fb, wtap * p.Ctot + ck * p.BK, n0);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++ck == nk) { ck = 0; ++t; newtap = true; }
      }
      if (tracing) { trace_val(p, 9, (unsigned long long)pwait); trace_put(p, 10); }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (p.BK == 64) mma_issue<4>(p, full_bar, empty_bar, &acc_bar, a_base, b_base, tmem_acc, ks1 - ks0);
      else if (p.BK == 32) mma_issue<2>(p, full_bar, empty_bar, &acc_bar, a_base, b_base, tmem_acc, ks1 - ks0);
      else mma_issue<1>(p, full_bar, empty_bar, &acc_bar, a_base, b_base, tmem_acc, ks1 - ks0);
      trace_put(p, 3);
    }
  } else {
    // ===================== epilogue =====================
    mbar_wait(smem_u32(&acc_bar), 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_put(p, 4);
    const EpiCtx e{smem_base, tmem_acc, x0, y0, b0, n0, py, px, cls};
    if (p.splits > 1) {
      split_push(p, e, split, tile_id);   // partial accumulators -> this CTA's slot of the L2 exchange buffer
      if (threadIdx.x == 64) trace_put(p, 11);
    } else
    // one uniform dispatch per CTA: the per-element code below is straight-line (4 epilogue warps = one warp per
    // scheduler, so every branch / dependent-issue bubble of a per-element `switch` was fully exposed: 0.1 us per
    // accumulator column before this was templated)
    if (EPI_WIDE && p.BN >= 32) {
      switch (p.act) {
        case PG_ACT_RELU: tc_epilogue<PG_ACT_RELU, 32>(p, mapsO, e); break;
        case PG_ACT_LEAKYRELU: tc_epilogue<PG_ACT_LEAKYRELU, 32>(p, mapsO, e); break;
        case PG_ACT_TANH: tc_epilogue<PG_ACT_TANH, 32>(p, mapsO, e); break;
        case PG_ACT_SIGMOID: tc_epilogue<PG_ACT_SIGMOID, 32>(p, mapsO, e); break;
        default: tc_epilogue<PG_ACT_NONE, 32>(p, mapsO, e); break;
      }
    } else {
      switch (p.act) {
        case PG_ACT_RELU: tc_epilogue<PG_ACT_RELU, 16>(p, mapsO, e); break;
        case PG_ACT_LEAKYRELU: tc_epilogue<PG_ACT_LEAKYRELU, 16>(p, mapsO, e); break;
        case PG_ACT_TANH: tc_epilogue<PG_ACT_TANH, 16>(p, mapsO, e); break;
        case PG_ACT_SIGMOID: tc_epilogue<PG_ACT_SIGMOID, 16>(p, mapsO, e); break;
        default: tc_epilogue<PG_ACT_NONE, 16>(p, mapsO, e); break;
      }
    }
  }
  if (p.splits > 1) {
    cluster_sync_all();                                   // every peer's partial tile is visible (release / acquire)
    if (threadIdx.x == 64) trace_put(p, 12);
    if (warp >= 2) {
      const EpiCtx e{smem_base, tmem_acc, x0, y0, b0, n0, py, px, cls};
      switch (p.act) {
        case PG_ACT_RELU: split_reduce_store<PG_ACT_RELU>(p, e, split, tile_id); break;
        case PG_ACT_LEAKYRELU: split_reduce_store<PG_ACT_LEAKYRELU>(p, e, split, tile_id); break;
        case PG_ACT_TANH: split_reduce_store<PG_ACT_TANH>(p, e, split, tile_id); break;
        case PG_ACT_SIGMOID: split_reduce_store<PG_ACT_SIGMOID>(p, e, split, tile_id); break;
        default: split_reduce_store<PG_ACT_NONE>(p, e, split, tile_id); break;
      }
      if (threadIdx.x == 64) trace_put(p, 13);
    }
  }
  if (threadIdx.x == 64) trace_put(p, 5);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, p.tmem_cols);
  }
  if (threadIdx.x == 0) trace_put(p, 6);
}

// --------------------------------------------------------------------------------------------
// Persistent variant for grids of many short tiles (first layers, 64-channel decoder / data-gradient layers: thousands of
// CTAs whose life is setup + one TMA round trip + a few MMAs + epilogue + store drain).  One CTA per SM slot walks the
// tiles; the operand ring runs ahead across tile boundaries, the accumulator is double-buffered in TMEM (tfull / tempty
// mbarriers), so the epilogue and the store drain of tile i overlap the loads and MMAs of tile i+1, and barrier init /
// TMEM allocation / descriptor prefetch are paid once.
// --------------------------------------------------------------------------------------------
struct MmaState {
  uint32_t a_lo, b_lo, stage, phase;
};

template <int KK, bool TR = false, bool PAIR = false>
__device__ __forceinline__ void mma_issue_tile(const TcParams& p, MmaState& st, uint32_t full0, uint32_t empty0, uint32_t done_bar,
                                               uint32_t a_lo0, uint32_t b_lo0, uint64_t desc_hi, uint32_t tmem_acc, int ksteps,
                                               long long* waited = nullptr) {
  const uint32_t idesc = p.idesc, stages = (uint32_t)p.stages;
  const uint32_t a_step = p.a_bytes >> 4, b_step = p.b_bytes >> 4;
  const uint32_t acc_wrap = (uint32_t)(p.nacc * p.BN), bn = (uint32_t)p.BN;
  uint32_t acc_off = 0, fresh = (uint32_t)p.nacc;
  for (int ks = 0; ks < ksteps; ++ks) {
    if (TR) {
      const long long w0 = clock64();
      mbar_wait(full0 + st.stage * 8, st.phase);
      *waited += clock64() - w0;
    } else {
      mbar_wait(full0 + st.stage * 8, st.phase);
    }
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const uint32_t accum = fresh == 0 ? 1u : 0u;
      if (PAIR) umma_bf16_pair(tmem_acc + acc_off, desc_hi | (uint64_t)(st.a_lo + 2 * k), desc_hi | (uint64_t)(st.b_lo + 2 * k), idesc, accum);
      else umma_bf16(tmem_acc + acc_off, desc_hi | (uint64_t)(st.a_lo + 2 * k), desc_hi | (uint64_t)(st.b_lo + 2 * k), idesc, accum);
      fresh -= accum ^ 1u;
      acc_off += bn;
      if (acc_off == acc_wrap) acc_off = 0;
    }
    if (PAIR) umma_commit_pair(empty0 + st.stage * 8); else umma_commit(empty0 + st.stage * 8);
    st.a_lo += a_step; st.b_lo += b_step;
    if (++st.stage == stages) { st.stage = 0; st.phase ^= 1; st.a_lo = a_lo0; st.b_lo = b_lo0; }
  }
  if (PAIR) umma_commit_pair(done_bar); else umma_commit(done_bar);
}

struct TileXY {
  int x0, y0, b0, n0, py, px, cls;
};
// (CTA pair: pers_mtiles counts PAIRS of adjacent m-tiles, the CTA of rank r owns the r-th of its pair)
__device__ __forceinline__ TileXY pers_decode(const TcParams& p, int w, int pair = 0, int rank = 0) {
  const int mt = pair ? 2 * (w % p.pers_mtiles) + rank : w % p.pers_mtiles;
  const int r = w / p.pers_mtiles;
  const int nt = r % p.pers_ntiles;
  TileXY t;
  t.cls = r / p.pers_ntiles;
  t.x0 = (mt % p.nx) * p.TW;
  t.y0 = ((mt / p.nx) % p.ny) * p.TH;
  t.b0 = (mt / (p.nx * p.ny)) * p.TB;
  t.n0 = nt * p.BN;
  t.py = t.cls >> 1; t.px = t.cls & 1;
  return t;
}

// TR = true: the per-CTA trace build (tools/conv_trace.py): slots 0 start, 1 setup done, 2 first accumulator ready, 6 exit
// (%globaltimer), 7 SM id, 8 producer cycles waiting for ring slots, 9 issuer cycles waiting for operands, 10 issuer cycles
// waiting for a drained accumulator, 11 epilogue cycles waiting for an accumulator, 12 epilogue busy cycles, 13 tiles, 15 = 1
// PAIR = true: launched as clusters of two CTAs (one TPC); the pair accumulates a 256-row x BN tile with
// tcgen05.mma.cta_group::2 -- each CTA loads its own 128 rows of A and HALF of the B tile (the weights), which is what
// lowers the L2 -> SM bytes per flop of these L2-throughput-bound loops by 25..33 % -- the leader's full barriers collect the
// bytes of both CTAs, the leader issues the MMAs, and its commits release the ring slots / publish the accumulators in
// both CTAs.  Each CTA runs the ordinary epilogue on its own 128 TMEM lanes.
constexpr int PAIR_EPI_GROUPS = 2;                      // one CTA per SM: 8 epilogue warps, like two co-resident single CTAs
constexpr int PAIR_THREADS = 64 + 128 * PAIR_EPI_GROUPS;
template <bool TR, bool PAIR = false>
__global__ void __launch_bounds__(PAIR ? PAIR_THREADS : TC_THREADS, PAIR ? 1 : 2)
conv_tc_pers_kernel(const __grid_constant__ ActMaps mapsA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ ActMaps mapsO, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull[2];
  __shared__ __align__(8) uint64_t tempty[2];
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + p.stages * p.a_bytes;
  const int nk = p.nk1 + p.nk2;
  const int ksteps = p.ntaps * nk;
  const int rank = PAIR ? (int)cluster_rank() : 0;
  const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapsA.m[0]);
    if (p.nk2 > 0) prefetch_tmap(&mapsA.m[4]);
    prefetch_tmap(&mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull[s]), 1);
      mbar_init(smem_u32(&tempty[s]), PAIR ? 256 * PAIR_EPI_GROUPS : 128);   // pair: the epilogue threads of both CTAs release the leader's slot
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (TR && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace_raw(p.trace, 0, gtimer());
    trace_raw(p.trace, 7, smid);
    trace_raw(p.trace, 15, 1);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(smem_u32(&tmem_base_sh), p.tmem_cols);
    else tmem_alloc(smem_u32(&tmem_base_sh), p.tmem_cols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();      // (pair: the peer's barriers exist before anything signals them)
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;
  if (TR && threadIdx.x == 0) trace_raw(p.trace, 1, gtimer());

  if (warp == 0) {
    // ===================== TMA producer: the ring runs ahead across tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long pwait = 0;
      const int b_row0 = PAIR ? rank * (p.BN >> 1) : 0;      // pair: this CTA's half of the weight tile
      for (int w = w_first; w < p.pers_total; w += w_step) {
        const TileXY t = pers_decode(p, w, PAIR, rank);
        int tap = 0, ck = 0, cx = 0, cy = 0, wtap = 0, ph = 0;
        bool newtap = true;
        for (int ks = 0; ks < ksteps; ++ks) {
          if (newtap) {
            newtap = false;
            ph = 0;
            if (p.mode == PG_CONVT) {
              const int j = tap >> 1, i = tap & 1;
              wtap = ((1 - t.py) + 2 * j) * 4 + (1 - t.px) + 2 * i;
              cx = t.x0 + t.px - i;
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
          if (TR) {
            const long long w0 = clock64();
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
            pwait += clock64() - w0;
          } else {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          }
          if (PAIR) {
            // both CTAs' bytes complete on the LEADER's barrier (tx_bytes counts both); only the leader arrives on it
            if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), p.tx_bytes);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (ck < p.nk1) tma_load_4d_pair(a_base + stage * p.a_bytes, &mapsA.m[ph], fb, ck * p.BK, cx, cy, t.b0);
            else tma_load_4d_pair(a_base + stage * p.a_bytes, &mapsA.m[4 + ph], fb, (ck - p.nk1) * p.BK, cx, cy, t.b0);
            tma_load_2d_pair(b_base + stage * p.b_bytes, &mapB, fb, wtap * p.Ctot + ck * p.BK, t.n0 + b_row0);
          } else {
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, p.tx_bytes);
            if (ck < p.nk1) tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[ph], fb, ck * p.BK, cx, cy, t.b0);
            else tma_load_4d(a_base + stage * p.a_bytes, &mapsA.m[4 + ph], fb, (ck - p.nk1) * p.BK, cx, cy, t.b0);
            tma_load_2d(b_base + stage * p.b_bytes, &mapB, fb, wtap * p.Ctot + ck * p.BK, t.n0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++ck == nk) { ck = 0; ++tap; newtap = true; }
        }
      }
      if (PAIR) {
        // the leader's last commits still arrive on this CTA's ring barriers: see them land before the CTA may exit
        for (int s = 0; s < p.stages; ++s) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (TR) trace_raw(p.trace, 8, (unsigned long long)pwait);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: accumulator slot it & 1 =====================
    if (lane == 0 && rank == 0) {
      const uint32_t a_lo0 = (a_base & 0x3FFFF) >> 4, b_lo0 = (b_base & 0x3FFFF) >> 4;
      const uint64_t desc_hi = make_smem_desc(0, p.sbo, p.layout_type);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      MmaState st{a_lo0, b_lo0, 0u, 0u};
      uint32_t it = 0;
      long long wfull = 0, wempty = 0;
      for (int w = w_first; w < p.pers_total; w += w_step, ++it) {
        const uint32_t slot = it & 1u;
        if (TR) {
          const long long w0 = clock64();
          mbar_wait(smem_u32(&tempty[slot]), ((it >> 1) & 1u) ^ 1u);
          wempty += clock64() - w0;
        } else {
          mbar_wait(smem_u32(&tempty[slot]), ((it >> 1) & 1u) ^ 1u);      // epilogue has drained this slot
        }
        tc_fence_after();
        const uint32_t acc = tmem_base + slot * p.pers_slot_cols;
        if (p.BK == 64) mma_issue_tile<4, TR, PAIR>(p, st, full0, empty0, smem_u32(&tfull[slot]), a_lo0, b_lo0, desc_hi, acc, ksteps, &wfull);
        else if (p.BK == 32) mma_issue_tile<2, TR, PAIR>(p, st, full0, empty0, smem_u32(&tfull[slot]), a_lo0, b_lo0, desc_hi, acc, ksteps, &wfull);
        else mma_issue_tile<1, TR, PAIR>(p, st, full0, empty0, smem_u32(&tfull[slot]), a_lo0, b_lo0, desc_hi, acc, ksteps, &wfull);
      }
      if (TR) { trace_raw(p.trace, 9, (unsigned long long)wfull); trace_raw(p.trace, 10, (unsigned long long)wempty); }
    }
  } else {
    // ===================== epilogue =====================
    uint32_t it = 0;
    long long ewait = 0, ebusy = 0, e0 = 0;
    for (int w = w_first; w < p.pers_total; w += w_step, ++it) {
      const TileXY t = pers_decode(p, w, PAIR, rank);
      const uint32_t slot = it & 1u;
      if (TR) {
        const long long w0 = clock64();
        mbar_wait(smem_u32(&tfull[slot]), (it >> 1) & 1u);
        e0 = clock64();
        ewait += e0 - w0;
        if (it == 0 && threadIdx.x == 64) trace_raw(p.trace, 2, gtimer());
      } else {
        mbar_wait(smem_u32(&tfull[slot]), (it >> 1) & 1u);
      }
      tc_fence_after();
      // (deferred drain: the bulk stores of this tile are still reading the staging buffer while the next tile's
      //  accumulators are read and converted; a buffer is waited for right before it is written again)
      const EpiCtx e{smem_base + p.pers_stage_off, tmem_base + slot * p.pers_slot_cols, t.x0, t.y0, t.b0, t.n0, t.py, t.px, t.cls,
                     PAIR ? (int)it : (p.pers_defer ? (int)it * p.pers_nch : 0), PAIR ? 1 : p.pers_defer,
                     PAIR ? (warp - 2) >> 2 : 0, PAIR ? PAIR_EPI_GROUPS : 1};
      switch (p.act) {
        case PG_ACT_RELU: tc_epilogue<PG_ACT_RELU, 16>(p, mapsO, e); break;
        case PG_ACT_LEAKYRELU: tc_epilogue<PG_ACT_LEAKYRELU, 16>(p, mapsO, e); break;
        case PG_ACT_TANH: tc_epilogue<PG_ACT_TANH, 16>(p, mapsO, e); break;
        case PG_ACT_SIGMOID: tc_epilogue<PG_ACT_SIGMOID, 16>(p, mapsO, e); break;
        default: tc_epilogue<PG_ACT_NONE, 16>(p, mapsO, e); break;
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[slot]), 0));
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[slot])) : "memory");
      if (!PAIR && !p.pers_defer) epi_bar_sync();      // the staging buffers are free again (thread 64 has waited for the bulk stores)
      if (TR) ebusy += clock64() - e0;
    }
    if ((PAIR || p.pers_defer) && p.tma_store && ((threadIdx.x - 64) & 127) == 0) bulk_wait_read<0>();   // smem must outlive the last stores
    if (TR && threadIdx.x == 64) {
      trace_raw(p.trace, 11, (unsigned long long)ewait);
      trace_raw(p.trace, 12, (unsigned long long)ebusy);
      trace_raw(p.trace, 13, it);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (TR && threadIdx.x == 0) trace_raw(p.trace, 6, gtimer());
}

static unsigned long long* g_trace = nullptr;
void set_tc_trace(void* buf) { g_trace = (unsigned long long*)buf; }
static unsigned long long g_pair_launches = 0;
static int g_pair_mode = -1;     // 0 never, 1 where it measured faster, 2 every eligible shape; default from PG_TC_PAIR (1)
static int pair_mode() {
  if (g_pair_mode < 0) { const char* e = getenv("PG_TC_PAIR"); g_pair_mode = e ? atoi(e) : 1; }
  return g_pair_mode;
}
void set_pair_mode(int m) { g_pair_mode = m < 0 ? -1 : (m > 2 ? 2 : m); }
unsigned long long pair_launch_count() { return g_pair_launches; }

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
    else
      cudaGetLastError();
  }
  return fn;
}

bool tc_device_ok() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); ok = 0; return false; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10 && get_encode() != nullptr) ? 1 : 0;
  }
  return ok == 1;
}

static int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Split-K exchange scratch (caller-owned, pg_conv_set_workspace), one per device: consecutive split launches take
// consecutive slices so that launches in flight on different streams never share one.  The cursor never wraps: the caller
// re-registers the buffer (which rewinds the cursor) at a point where no convolution is in flight -- the engine does so at
// the beginning of every step -- and a launch that does not fit any more fails loudly instead of reusing a slice that
// another stream may still be reducing through.
constexpr int MAX_DEVICES = 16;
struct WsState { float* ws; size_t bytes, next; };
static WsState g_wss[MAX_DEVICES];
static WsState& ws_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return g_wss[dev >= 0 && dev < MAX_DEVICES ? dev : 0];
}
void set_conv_workspace(void* ws, size_t bytes) {
  WsState& w = ws_state();
  w.ws = (float*)ws; w.bytes = bytes; w.next = 0;
}

struct TcPlan {
  TcParams p;
  int swz;  // bytes
  dim3 grid;
  size_t smem;
};

static bool make_plan(const PgConvDesc* d, TcPlan& pl) {
  TcParams& p = pl.p;
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
  const int nb = (p.B + p.TB - 1) / p.TB;
  // channel chunk: largest of 64/32/16 dividing both sources
  int bk = 64;
  while (bk > 16 && ((d->C1 % bk) != 0 || (d->C2 % bk) != 0)) bk >>= 1;
  if ((d->C1 % bk) != 0 || (d->C2 % bk) != 0) return false;
  p.BK = bk; p.nk1 = d->C1 / bk; p.nk2 = d->C2 / bk; p.Ctot = d->C1 + d->C2;
  int bn = 256;
  while (bn > 16 && (d->N % bn) != 0) bn >>= 1;
  if ((d->N % bn) != 0) return false;
  // Small M grids (the 2x2 .. 16x16 bottleneck layers) would otherwise run on a handful of SMs, each pulling the
  // whole weight matrix through its own ~80 GB/s L2 port.  First choice: split K over a cluster of CTAs that reduce
  // through distributed shared memory (tiles stay 128 x 128, every operand byte is fetched once per tile row/column);
  // otherwise split N further so the weights stream on more SMs.
  p.splits = 1;
  {
    const long long mtiles = (long long)p.nx * p.ny * nb * (d->mode == PG_CONVT ? 4 : 1);
    const int ksteps0 = p.ntaps * (p.nk1 + p.nk2);
    static const int splitk_env = [] { const char* e = getenv("PG_TC_SPLITK"); return e ? atoi(e) : 1; }();
    // 128 x 64 tiles, clusters of <= 4: the exchanged partial tile is 32 KB, 4-SM clusters pack into every GPC, and
    // two such CTAs fit one SM
    int sbn = bn > 64 ? 64 : bn;
    const WsState& wst = ws_state();
    if (splitk_env && wst.ws != nullptr && d->mode != PG_CONV1X1 && mtiles * (d->N / sbn) * 2 <= num_sms() && ksteps0 >= 8) {
      int S = 4;
      while (S > 1 && (sbn / S < 16 || (ksteps0 % S) != 0 || ksteps0 / S < 2)) S >>= 1;
      while (S > 2 && mtiles * (d->N / sbn) * S > 2LL * num_sms()) S >>= 1;
      if (S > 1 && (size_t)mtiles * (d->N / sbn) * S * 128 * sbn * 4 <= wst.bytes) { p.splits = S; bn = sbn; }
    }
    if (p.splits == 1)
      while (bn > 16 && mtiles * (d->N / bn) < num_sms()) bn >>= 1;
  }
  p.BN = bn;
  p.N = d->N; p.ldo = d->ldo; p.n_valid = d->n_valid; p.act = d->act; p.out_f32 = d->out_f32;
  pl.swz = bk * 2;
  p.layout_type = pl.swz == 128 ? 2u : (pl.swz == 64 ? 4u : 6u);
  p.sbo = 8u * pl.swz;
  // instruction descriptor: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major both, N>>3 at 17, M>>4 at 24
  const uint32_t fmt = d->in_dtype == PG_F16 ? 0u : 1u;   // F16F32Format: F16 = 0, BF16 = 1
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  p.a_bytes = 128u * pl.swz;
  p.b_bytes = ((uint32_t)bn * pl.swz + 1023u) & ~1023u;
  p.tx_bytes = 128u * pl.swz + (uint32_t)bn * pl.swz;
  const uint32_t per_stage = p.a_bytes + p.b_bytes;
  // Two CTAs share an SM (each with half of the smem ring and <= 256 TMEM columns) so that one CTA's prologue /
  // epilogue overlaps the other's main loop.  PG_TC_OCC=1 restores one CTA per SM with the full ring.
  static const int occ_env = [] { const char* e = getenv("PG_TC_OCC"); return e ? atoi(e) : 4; }();
  int occ = occ_env < 1 ? 1 : (occ_env > 4 ? 4 : occ_env);
  {
    // small grids cannot fill more than one CTA per SM anyway: give them one CTA with a deep ring (latency-bound
    // k-loops need bytes in flight), large grids get co-resident CTAs with shallow rings
    const long long ctas = (long long)p.nx * p.ny * nb * (d->N / bn) * (d->mode == PG_CONVT ? 4 : 1) * p.splits;
    const int want = (int)((ctas + num_sms() - 1) / num_sms());
    if (occ > want) occ = want < 1 ? 1 : want;
    if (p.splits > 1 && occ < 2) occ = 2;      // split-K CTAs are short: half a ring each, so two clusters can share SMs
  }
  static const int nacc_env = [] { const char* e = getenv("PG_TC_NACC"); return e ? atoi(e) : 4; }();
  {
    const int total_mma = p.ntaps * (p.nk1 + p.nk2) * (bk / 16) / p.splits;
    int nacc = bn <= 64 ? nacc_env : (bn == 128 ? (nacc_env >= 2 ? 2 : 1) : 1);
    while (nacc > 1 && (nacc * bn > 256 || total_mma < 2 * nacc)) nacc >>= 1;
    p.nacc = nacc < 1 ? 1 : nacc;
  }
  const int tmem_need = p.nacc * bn < 32 ? 32 : p.nacc * bn;        // (power of two: nacc and bn both are)
  if (occ * tmem_need > 512) occ = 512 / tmem_need;                 // co-resident CTAs share the 512 TMEM columns
  while (occ > 1 && (220u * 1024u / occ) / per_stage < 2) --occ;    // keep at least a 2-deep ring per CTA
  const uint32_t budget = occ >= 2 ? 220u * 1024u / occ - 1024u : (uint32_t)TC_MAX_DYN_SMEM - 2048u;
  int stages = (int)(budget / per_stage);
  if (stages < 1) stages = 1;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  {
    static const int st_env = [] { const char* e = getenv("PG_TC_STAGES"); return e ? atoi(e) : 0; }();
    if (st_env > 0 && st_env < stages) stages = st_env;      // experiment knob
  }
  const int ksteps = p.ntaps * (p.nk1 + p.nk2);
  p.kps = ksteps / p.splits;
  if (stages > p.kps) stages = p.kps;
  if (stages < 1) return false;
  p.stages = stages;
  p.tmem_cols = (uint32_t)tmem_need;
  pl.smem = (size_t)stages * per_stage + 1024;
  pl.grid = dim3((unsigned)(p.nx * p.ny * nb), (unsigned)(d->N / bn), (unsigned)((d->mode == PG_CONVT ? 4 : 1) * p.splits));
  // TMA box limits
  if (p.TW > 256 || p.TH > 256 || p.TB > 256) return false;
  return true;
}

bool conv_fwd_tc_supported(const PgConvDesc* d, const void* src1, const void* src2, const void* w, const void* out) {
  if (!tc_device_ok()) return false;
  if ((((uintptr_t)src1 | (uintptr_t)src2 | (uintptr_t)w | (uintptr_t)out) & 15) != 0) return false;
  TcPlan pl;
  return make_plan(d, pl);
}

// Tensor map over an NHWC activation.  phase < 0: the plain tensor {C, W, H, B}.  phase = ry*2+rx: the stride-2 phase
// view X[b][2y'+ry][2x'+rx][c] with extents {C, ceil((W-rx)/2), ceil((H-ry)/2), B}.  Boxes are always unit-stride.
static int encode_act_map(CUtensorMap* m, const void* base, int C, int ld, int B, int H, int W, int bk, int tw, int th,
                          int tb, int phase, int swz, int dt) {
  const size_t es_ = dt == PG_F32 ? 4 : 2;
  const int ry = phase < 0 ? 0 : (phase >> 1), rx = phase < 0 ? 0 : (phase & 1);
  const int step = phase < 0 ? 1 : 2;
  const int Wv = phase < 0 ? W : (W - rx + 1) / 2, Hv = phase < 0 ? H : (H - ry + 1) / 2;
  if (Wv <= 0 || Hv <= 0) {            // degenerate phase (1-pixel-wide input): never addressed in bounds; map phase 0
    return encode_act_map(m, base, C, ld, B, H, W, bk, tw, th, tb, phase < 0 ? -1 : 0, swz, dt);
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)step * ld * es_, (cuuint64_t)step * W * ld * es_, (cuuint64_t)H * W * ld * es_};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const char* origin = (const char*)base + ((size_t)ry * W + rx) * ld * es_;
  CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                     : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUtensorMapDataType cdt = dt == PG_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                               : (dt == PG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = get_encode()(m, cdt, 4, const_cast<char*>(origin), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation C=%d ld=%d B=%d H=%d W=%d box=%d,%d,%d,%d phase=%d) failed: %d", C, ld, B,
              H, W, bk, tw, th, tb, phase, (int)r);
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

// stats fusion needs every warp's 32 rows inside one image and the 16-column epilogue
bool conv_fwd_tc_stats_ok(const PgConvDesc* d) {
  TcPlan pl;
  if (!make_plan(d, pl)) return false;
  return pl.p.TW * pl.p.TH >= 32 && !(EPI_WIDE && pl.p.BN >= 32);
}

int conv_fwd_tc(const PgConvDesc* d, const void* src1, const void* src2, const void* w, const float* bias, void* out,
                void* out2, float* stats, const void* mul_y, int mul_ld, int mul_dt, cudaStream_t stream) {
  TcPlan pl;
  if (!make_plan(d, pl)) {
    set_error("conv_fwd_tc: unsupported shape");
    return PG_ERR_UNSUPPORTED;
  }
  TcParams& p = pl.p;
  p.bias = d->has_bias ? bias : nullptr;
  p.out = out;
  p.out2 = out2;
  p.stats = stats;
  p.mul_y = mul_y; p.mul_ld = mul_ld; p.mul_dt = mul_dt;
  const bool phased = d->mode == PG_CONV && d->stride == 2;
  // ---- CTA-pair variant (cta_group::2) of the persistent kernel: wide tiles of the big layers, whose main loops run at
  //      the chip's L2 -> SM throughput; the pair shares one weight tile, so each CTA fetches half of it
  bool pair = false;
  {
    const int pair_env = pair_mode();
    static const int pair_min_bn = [] { const char* e = getenv("PG_TC_PAIR_MINBN"); return e ? atoi(e) : 128; }();
    const int ncls = d->mode == PG_CONVT ? 4 : 1;
    const long long mtiles = (long long)pl.grid.x, ntn = d->N / p.BN;
    const int sms = num_sms() & ~1;
    const long long npairs = (mtiles / 2) * ntn * ncls;
    pair = pair_env && p.splits == 1 && (mtiles & 1) == 0 && p.BN >= pair_min_bn && p.BN <= 256 && stats == nullptr &&
           npairs >= sms / 2 && npairs < (1LL << 30);
    // Measured on B200 (tools/conv_trace.py d, PG_TC_PAIR=2 vs 0): the pair wins 5..8 % on 128-wide tiles walked >= 4 per
    // cluster (d1 forward, d2 data-gradient) and loses to two co-resident single CTAs on the 256-wide layers, whose main
    // loops already run at 78..98 % of the sustained tensor rate and whose tile counts quantise worse over 74 clusters
    // than over 296 CTA slots.  PG_TC_PAIR=2 takes every eligible shape (A/B runs, the parity tests), 0 none.
    if (pair && pair_env == 1 && !(p.BN == 128 && npairs >= 4LL * (sms / 2))) pair = false;
    if (pair) {
      p.b_bytes = ((uint32_t)(p.BN / 2) * pl.swz + 1023u) & ~1023u;
      p.tx_bytes = 2u * (128u * pl.swz + (uint32_t)(p.BN / 2) * pl.swz);
      p.idesc = (p.idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
    }
  }
  ActMaps mA;
  CUtensorMap mB;
  memset(&mA, 0, sizeof(mA));
  for (int src = 0; src < (d->C2 > 0 ? 2 : 1); ++src) {
    const void* base = src ? src2 : src1;
    const int C = src ? d->C2 : d->C1, ld = src ? d->ld2 : d->ld1;
    for (int ph = 0; ph < (phased ? 4 : 1); ++ph)
      if (int e = encode_act_map(&mA.m[src * 4 + ph], base, C, ld, d->B, d->Hin, d->Win, p.BK, p.TW, p.TH, p.TB,
                                 phased ? ph : -1, pl.swz, d->in_dtype))
        return e;
  }
  {
    const int wtaps = d->mode == PG_CONV1X1 ? 1 : 16;
    cuuint64_t dims[2] = {(cuuint64_t)wtaps * p.Ctot, (cuuint64_t)d->N};
    cuuint64_t strides[1] = {(cuuint64_t)(d->mode == PG_CONV1X1 && d->ldw > 0 ? d->ldw : wtaps * p.Ctot) * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.BK, (cuuint32_t)(pair ? p.BN / 2 : p.BN)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = pl.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (pl.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = get_encode()(&mB, d->in_dtype == PG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", 16 * p.Ctot, d->N, (int)r);
      return PG_ERR_CUDA;
    }
  }
  static bool smem_set = false;
  if (!smem_set) {
    // 227 KB opt-in limit per block minus this kernel's static shared memory (barriers, 1 KB with alignment)
    PG_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    smem_set = true;
  }
  // ---- persistent variant: many short tiles (decided first: it has dedicated store-staging buffers behind the ring)
  bool pers = false;
  size_t pers_smem = 0;
  int pers_nbuf = 2;
  dim3 pers_grid;
  {
    static const int pers_env = [] { const char* e = getenv("PG_TC_PERSIST"); return e ? atoi(e) : 1; }();
    static const int tma_st_env2 = [] { const char* e = getenv("PG_TC_TMA_STORE"); return e ? atoi(e) : 1; }();
    const int ncls = d->mode == PG_CONVT ? 4 : 1;
    const long long mtiles = (long long)pl.grid.x, ntn = d->N / p.BN;
    const long long total = mtiles * ntn * ncls;
    static const int pers_min = [] { const char* e = getenv("PG_TC_PERSIST_MIN"); return e ? atoi(e) : 1; }();
    if (pair || (pers_env && p.splits == 1 && total >= (long long)pers_min * num_sms() && total < (1LL << 30))) {
      int nacc = p.nacc;
      while (nacc > 1 && 2 * nacc * p.BN > (pair ? 512 : 256)) nacc >>= 1;
      const int cols = 2 * nacc * p.BN;
      int tcols = 32;
      while (tcols < cols) tcols <<= 1;
      const int occ = (tcols <= 256 && !pair) ? 2 : 1;      // a CTA pair owns its two SMs
      const int esz = d->out_f32 == PG_F32 ? 4 : 2;
      const int rowbytes = p.BN * esz > 128 ? 128 : p.BN * esz;
      const int twin = out2 != nullptr ? 1 : 0;
      const bool tst = tma_st_env2 && d->ldo >= d->N && rowbytes >= 32 && ((uintptr_t)out & 15) == 0 &&
                       (twin == 0 || ((uintptr_t)out2 & 15) == 0);
      // staging: double-buffered unless the tile is a single store chunk anyway
      static const int defer_env = [] { const char* e = getenv("PG_TC_DEFER"); return e ? atoi(e) : 1; }();
      const int nch = p.BN / (rowbytes / esz);
      const uint32_t per_stage = p.a_bytes + p.b_bytes;
      // (one persistent CTA per SM loses to four short-lived ones when the tile is epilogue-bound: keep two per SM)
      const uint32_t budget = 220u * 1024u / occ - 2048u;
      // with the deferred drain one-chunk tiles alternate between two staging buffers as well (if the ring keeps 2 stages)
      int nbuf = (nch >= 2 || (defer_env && tst)) ? 2 : 1;
      if (nch < 2 && nbuf == 2 && budget < 2u * (1 + twin) * 128u * rowbytes + 2u * per_stage) nbuf = 1;
      const uint32_t staging = tst ? (uint32_t)nbuf * (1 + twin) * 128u * rowbytes : 0u;
      int stages = budget > staging ? (int)((budget - staging) / per_stage) : 0;
      p.pers_defer = defer_env && tst ? 1 : 0;
      p.pers_nch = nch;
      pers_nbuf = nbuf;
      if (stages > MAX_STAGES) stages = MAX_STAGES;
      // (BN = 256 needs all 512 TMEM columns for two slots -> one CTA per SM, which measured slower than two co-resident
      //  non-persistent CTAs: 124 vs 109 us on the largest discriminator layer)
      if ((tcols <= 256 || pair) && stages >= 2) {
        pers = true;
        p.nacc = nacc;
        p.tmem_cols = (uint32_t)tcols;
        p.pers_slot_cols = (uint32_t)(nacc * p.BN);
        p.stages = stages;
        p.pers_stage_off = (uint32_t)stages * per_stage;
        p.pers_total = (int)total; p.pers_mtiles = (int)mtiles; p.pers_ntiles = (int)ntn;
        pers_smem = (size_t)stages * per_stage + staging + 1024;
        const long long slots = (long long)num_sms() * occ;
        pers_grid = dim3((unsigned)(total < slots ? total : slots));
        if (pair) {          // work items are pairs of m-tiles; one cluster of two CTAs per TPC
          p.pers_mtiles = (int)(mtiles / 2);
          p.pers_total = (int)(total / 2);
          const long long cl = (long long)(num_sms() / 2);
          pers_grid = dim3((unsigned)(2 * (p.pers_total < cl ? p.pers_total : cl)));
        }
      } else if (pair) {
        set_error("conv_tc: CTA-pair plan does not fit (BN %d stages %d)", p.BN, stages);
        return PG_ERR_UNSUPPORTED;
      }
    }
  }
  // ---- output path: TMA store of a smem-staged tile unless the output row is trimmed / tiny
  ActMaps mO;
  memset(&mO, 0, sizeof(mO));
  {
    static const int tma_st_env = [] { const char* e = getenv("PG_TC_TMA_STORE"); return e ? atoi(e) : 1; }();
    const int esz = d->out_f32 == PG_F32 ? 4 : 2;
    int rowbytes = p.BN * esz > 128 ? 128 : p.BN * esz;
    const int twin = out2 != nullptr ? 1 : 0;
    const uint32_t ring = (uint32_t)p.stages * (p.a_bytes + p.b_bytes);
    const uint32_t need1 = 128u * rowbytes * (1 + twin);
    p.tma_store = tma_st_env && p.splits == 1 && d->ldo >= d->N && rowbytes >= 32 && (pers || ring >= need1) &&
                  ((uintptr_t)out & 15) == 0 && (twin == 0 || ((uintptr_t)out2 & 15) == 0);
    if (p.tma_store) {
      p.st_rowbytes = rowbytes;
      p.st_cw = rowbytes / esz;
      p.st_nbuf = pers ? pers_nbuf : (ring >= 2 * need1 ? 2 : 1);
      p.st_twin = twin;
      const bool cls = d->mode == PG_CONVT;
      for (int ph = 0; ph < (cls ? 4 : 1); ++ph) {
        if (int e = encode_act_map(&mO.m[ph], out, d->N, d->ldo, d->B, d->Hout, d->Wout, p.st_cw, p.TW, p.TH, p.TB,
                                   cls ? ph : -1, rowbytes, d->out_f32))
          return e;
        if (twin)
          if (int e = encode_act_map(&mO.m[4 + ph], out2, d->N, d->ldo, d->B, d->Hout, d->Wout, p.st_cw, p.TW, p.TH, p.TB,
                                     cls ? ph : -1, rowbytes, PG_BF16))
            return e;
      }
    }
  }
  static const int skip = [] { const char* e = getenv("PG_TC_SKIP"); return e ? atoi(e) : 0; }();
  p.debug = skip;
  p.trace = g_trace;
  static const bool dbg = getenv("PG_TC_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "conv_tc: grid (%u,%u,%u) BN %d BK %d stages %d nacc %d tmem %u smem %zu TW %d TH %d TB %d ksteps %d splits %d\n",
            pl.grid.x, pl.grid.y, pl.grid.z, p.BN, p.BK, p.stages, p.nacc, p.tmem_cols, pl.smem, p.TW, p.TH, p.TB,
            p.ntaps * (p.nk1 + p.nk2), p.splits);
  if (pers) {
    static bool pers_set = false;
    if (!pers_set) {
      PG_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
      PG_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
      pers_set = true;
    }
    if (dbg) fprintf(stderr, "conv_tc: persistent grid %u stages %d nacc %d tmem %u smem %zu work %d\n", pers_grid.x, p.stages, p.nacc,
                     p.tmem_cols, pers_smem, p.pers_total);
    if (pair) {
      static bool pair_set = false;
      if (!pair_set) {
        PG_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
        PG_CUDA(cudaFuncSetAttribute(conv_tc_pers_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
        pair_set = true;
      }
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = pers_grid;
      cfg.blockDim = dim3(PAIR_THREADS);
      cfg.dynamicSmemBytes = pers_smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (p.trace != nullptr) PG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_pers_kernel<true, true>, mA, mB, mO, p));
      else PG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_pers_kernel<false, true>, mA, mB, mO, p));
      ++g_pair_launches;
      return check_launch("conv_tc_pers_kernel<pair>");
    }
    if (p.trace != nullptr) conv_tc_pers_kernel<true><<<pers_grid, TC_THREADS, pers_smem, stream>>>(mA, mB, mO, p);
    else conv_tc_pers_kernel<false><<<pers_grid, TC_THREADS, pers_smem, stream>>>(mA, mB, mO, p);
    return check_launch("conv_tc_pers_kernel");
  }
  if (p.splits > 1) {
    const size_t need = (size_t)pl.grid.x * pl.grid.y * pl.grid.z * 128 * p.BN * 4;
    WsState& wsm = ws_state();
    if (wsm.next + need > wsm.bytes) {
      set_error("conv_tc: split-K workspace exhausted (%zu of %zu bytes handed out since the last pg_conv_set_workspace): "
                "re-register it when no convolution is in flight, or make it larger", wsm.next, wsm.bytes);
      return PG_ERR_INVALID;
    }
    p.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(wsm.ws) + wsm.next);
    wsm.next += (need + 255) & ~(size_t)255;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = pl.grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)p.splits;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PG_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, mA, mB, mO, p));
  } else {
    conv_tc_kernel<<<pl.grid, TC_THREADS, pl.smem, stream>>>(mA, mB, mO, p);
  }
  return check_launch("conv_tc_kernel");
}

// --------------------------------------------------------------------------------------------
// weight gradient on tcgen05:  dW[n][c][tap] += sum_pix G[pix][n] * A[pix(tap)][c]
//
// Per tap this is a GEMM D[n][c] with the reduction over PIXELS.  Both operands are stored pixel-major with
// channels contiguous (NHWC), i.e. MN-major for the tensor core: a TMA box {channels, TW, TH, TB} lands in smem
// as KP=64 rows (pixels) of one swizzle-width of channels, which is exactly the canonical MN-major layout
// ((atom, n), (8, k)) with SBO = 8 rows and LBO = one box.  The G tile (dY, or the layer input for
// ConvTranspose2d) is loaded once per pixel tile and reused by the T taps whose accumulators [128 x ct] share
// the 512 TMEM columns; the A tile of every tap is the same strided/zero-filled box the forward kernel loads.
// Split-K over pixel tiles; the epilogue adds fp32 partials into dW with red.global.add.v4.f32 (4 taps of one
// (n, c) are contiguous in the reference weight layout).
// --------------------------------------------------------------------------------------------
struct WgParams {
  int stride, pad;
  int B, Hout, Wout;
  int TW, TH, TB;
  int nx, ny, total_tiles, tiles_per_split;
  int g_box, n_atoms, a_box, c_atoms, ct, T, ctiles;
  uint32_t idesc;
  uint32_t g_rowbytes, a_rowbytes, g_layout, a_layout, g_boxbytes, a_boxbytes;
  uint32_t g_stage_bytes, a_stage_bytes;
  int g_stages, a_stages;
  uint32_t tmem_cols;
  float* dw;
  int ld_n, n_real, c_real;
  int pointwise, ld_c;     // PG_CONV1X1: one tap, no shift, dw[n*ld_n + c*ld_c]
  int tapmajor;            // epilogue = TMA reduce-add of [128 n][ct c] tiles into S[tap][n][c] (mapS)
  int st_rowbytes;         // bytes per staged row (128, or 64 when ct == 16)
  int n_atoms_load;        // G boxes that exist (the accumulator rows of the others are never stored)
  int n_split;             // > 0: the G operand is the virtual concat of two tensors: channels [0, n_split) | [n_split, N)
  int pair;                // CTA pair (cta_group::2): two n-tiles share the activation tiles, each CTA loads half of them
  unsigned long long* trace;
};

constexpr int WG_KP = 64;
constexpr int WG_MAX_G = 4, WG_MAX_A = 12;

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                      uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between MN atoms (one TMA box)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride between groups of 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// The kernel body: one CTA = (n-tile bx, (channel tile, tap group) by, pixel-tile split bz) of ONE weight gradient.
// Called by wgrad_tc_kernel (one weight gradient per launch) and wgrad_group_kernel (many per launch).
// PAIR = true (wgrad_tc_pair_kernel, clusters of two CTAs along bx): the two n-tiles of a pair accumulate [256 n][ct c] per
// tap with tcgen05.mma.cta_group::2.  Each CTA loads its own G tile and HALF of every activation tile (the MMA's N
// operand is split over the pair), so a pixel tile costs a CTA 32 KB instead of 48 KB through its L2 port; barriers as in
// conv_tc_pers_kernel<.., PAIR>: the leader's full barriers collect both CTAs' bytes, its commits are multicast.
template <bool PAIR>
__device__ __forceinline__ void wgrad_body(const CUtensorMap* mapG, const CUtensorMap* mapG2, const CUtensorMap* mapsA,
                                           const CUtensorMap* mapS, const WgParams& p, int bx, int by, int bz) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t gfull[WG_MAX_G], gempty[WG_MAX_G], afull[WG_MAX_A], aempty[WG_MAX_A];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_sh;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t g_base = smem_base;
  const uint32_t a_base = smem_base + p.g_stages * p.g_stage_bytes;

  const int n0 = bx * 128;
  const int ctile = by % p.ctiles;
  const int tgrp = by / p.ctiles;
  const int c0 = ctile * p.ct;
  const int t0 = tgrp * p.T;
  const int tile_beg = bz * p.tiles_per_split;
  int tile_end = tile_beg + p.tiles_per_split;
  if (tile_end > p.total_tiles) tile_end = p.total_tiles;
  const int ntiles = tile_end > tile_beg ? tile_end - tile_beg : 0;
  const int rank = PAIR ? (int)cluster_rank() : 0;
  const int a_first = PAIR ? rank * (p.c_atoms >> 1) : 0, a_count = PAIR ? (p.c_atoms >> 1) : p.c_atoms;
  if (threadIdx.x == 0 && p.trace != nullptr) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace_raw(p.trace, 0, gtimer());
    trace_raw(p.trace, 7, smid);
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(mapG);
    prefetch_tmap(&mapsA[0]);
    for (int s = 0; s < p.g_stages; ++s) { mbar_init(smem_u32(&gfull[s]), 1); mbar_init(smem_u32(&gempty[s]), 1); }
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(smem_u32(&afull[s]), 1); mbar_init(smem_u32(&aempty[s]), 1); }
    mbar_init(smem_u32(&acc_bar), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(smem_u32(&tmem_base_sh), p.tmem_cols);
    else tmem_alloc(smem_u32(&tmem_base_sh), p.tmem_cols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_sh;
  if (threadIdx.x == 0 && p.trace != nullptr) trace_raw(p.trace, 1, gtimer());

  if (warp == 0) {
    if (lane == 0 && ntiles > 0) {
      int gs = 0, as = 0;
      uint32_t gph = 0, aph = 0;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        const int x0 = (tile % p.nx) * p.TW;
        const int y0 = ((tile / p.nx) % p.ny) * p.TH;
        const int b0 = (tile / (p.nx * p.ny)) * p.TB;
        mbar_wait(smem_u32(&gempty[gs]), gph ^ 1);
        if (PAIR) {
          if (rank == 0) mbar_expect_tx(smem_u32(&gfull[gs]), 2u * p.n_atoms_load * p.g_boxbytes);
          const uint32_t gb = mapa_u32(smem_u32(&gfull[gs]), 0);
          for (int a = 0; a < p.n_atoms_load; ++a)
            tma_load_4d_pair(g_base + gs * p.g_stage_bytes + a * p.g_boxbytes, mapG, gb, n0 + a * p.g_box, x0, y0, b0);
        } else {
          const uint32_t gb = smem_u32(&gfull[gs]);
          mbar_expect_tx(gb, p.n_atoms_load * p.g_boxbytes);
          for (int a = 0; a < p.n_atoms_load; ++a) {
            const int ch = n0 + a * p.g_box;          // (a box never straddles the two sources: n_split % g_box == 0)
            if (p.n_split > 0 && ch >= p.n_split)
              tma_load_4d(g_base + gs * p.g_stage_bytes + a * p.g_boxbytes, mapG2, gb, ch - p.n_split, x0, y0, b0);
            else
              tma_load_4d(g_base + gs * p.g_stage_bytes + a * p.g_boxbytes, mapG, gb, ch, x0, y0, b0);
          }
        }
        if (++gs == p.g_stages) { gs = 0; gph ^= 1; }
        for (int tl = 0; tl < p.T; ++tl) {
          const int t = t0 + tl, kh = t >> 2, kw = t & 3;
          mbar_wait(smem_u32(&aempty[as]), aph ^ 1);
          uint32_t ab = smem_u32(&afull[as]);
          if (PAIR) {
            if (rank == 0) mbar_expect_tx(ab, p.c_atoms * p.a_boxbytes);      // both halves complete on the leader's barrier
            ab = mapa_u32(ab, 0);
          } else {
            mbar_expect_tx(ab, p.c_atoms * p.a_boxbytes);
          }
          int ph = 0, cx = x0 - p.pad + kw, cy = y0 - p.pad + kh;
          if (p.pointwise) { cx = x0; cy = y0; }
          else if (p.stride == 2) {
            const int u = kh - p.pad, v = kw - p.pad;
            ph = (u & 1) * 2 + (v & 1);
            cx = x0 + (v >> 1);
            cy = y0 + (u >> 1);
          }
          for (int a = 0; a < a_count; ++a) {
            if (PAIR) tma_load_4d_pair(a_base + as * p.a_stage_bytes + a * p.a_boxbytes, &mapsA[ph], ab,
                                       c0 + (a_first + a) * p.a_box, cx, cy, b0);
            else tma_load_4d(a_base + as * p.a_stage_bytes + a * p.a_boxbytes, &mapsA[ph], ab, c0 + a * p.a_box, cx, cy, b0);
          }
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
        }
      }
      if (PAIR) {      // the leader's last commits still arrive on this CTA's ring barriers
        for (int s = 0; s < p.g_stages; ++s) { mbar_wait(smem_u32(&gempty[gs]), gph ^ 1); if (++gs == p.g_stages) { gs = 0; gph ^= 1; } }
        for (int s = 0; s < p.a_stages; ++s) { mbar_wait(smem_u32(&aempty[as]), aph ^ 1); if (++as == p.a_stages) { as = 0; aph ^= 1; } }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0 && rank == 0) {
      // lean single-thread issue loop (see mma_issue): loop invariants in registers, descriptors advance by adds
      const uint32_t idesc = p.idesc, T = (uint32_t)p.T, ct = (uint32_t)p.ct;
      const uint32_t g_stages = (uint32_t)p.g_stages, a_stages = (uint32_t)p.a_stages;
      const uint64_t g_hi = make_smem_desc_mn(0, p.g_boxbytes, 8 * p.g_rowbytes, p.g_layout);
      const uint64_t a_hi = make_smem_desc_mn(0, p.a_boxbytes, 8 * p.a_rowbytes, p.a_layout);
      const uint32_t g_lo0 = (g_base & 0x3FFFF) >> 4, a_lo0 = (a_base & 0x3FFFF) >> 4;
      const uint32_t g_step = p.g_stage_bytes >> 4, a_step = p.a_stage_bytes >> 4;
      const uint32_t gk = (16 * p.g_rowbytes) >> 4, ak = (16 * p.a_rowbytes) >> 4;   // 16 pixels (rows) further along K
      const uint32_t gfull0 = smem_u32(gfull), gempty0 = smem_u32(gempty), afull0 = smem_u32(afull), aempty0 = smem_u32(aempty);
      uint32_t gs = 0, as = 0, gph = 0, aph = 0, g_lo = g_lo0, a_lo = a_lo0;
      uint32_t accum = 0;
      for (int it = 0; it < ntiles; ++it) {
        mbar_wait(gfull0 + gs * 8, gph);
        tc_fence_after();
        if (it == 0 && p.trace != nullptr) trace_raw(p.trace, 2, gtimer());
        uint32_t tcol = tmem_acc;
        for (uint32_t tl = 0; tl < T; ++tl) {
          mbar_wait(afull0 + as * 8, aph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < WG_KP / 16; ++k) {
            if (PAIR) umma_bf16_pair(tcol, g_hi | (uint64_t)(g_lo + k * gk), a_hi | (uint64_t)(a_lo + k * ak), idesc, (accum | k) ? 1u : 0u);
            else umma_bf16(tcol, g_hi | (uint64_t)(g_lo + k * gk), a_hi | (uint64_t)(a_lo + k * ak), idesc, (accum | k) ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(aempty0 + as * 8); else umma_commit(aempty0 + as * 8);
          tcol += ct;
          a_lo += a_step;
          if (++as == a_stages) { as = 0; aph ^= 1; a_lo = a_lo0; }
        }
        accum = 1;
        if (PAIR) umma_commit_pair(gempty0 + gs * 8); else umma_commit(gempty0 + gs * 8);
        g_lo += g_step;
        if (++gs == g_stages) { gs = 0; gph ^= 1; g_lo = g_lo0; }
      }
      if (PAIR) umma_commit_pair(smem_u32(&acc_bar)); else umma_commit(smem_u32(&acc_bar));
      if (p.trace != nullptr) trace_raw(p.trace, 3, gtimer());
    }
  } else if (ntiles > 0) {
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    mbar_wait(smem_u32(&acc_bar), 0);
    tc_fence_after();
    if (threadIdx.x == 64 && p.trace != nullptr) trace_raw(p.trace, 4, gtimer());
    if (p.tapmajor) {
      // Tap-major accumulation: the [128 n][ct c] accumulator of tap t is staged in the (idle) operand ring with the
      // TMA swizzle and added into S[t][n][c] by ONE bulk reduce per 32 columns (full 128-byte lines, tails clipped by
      // the tensor map) -- instead of 128 x ct scattered 8/16-byte red.global per tap pair, which cost ~9 us per CTA
      // and serialised on the few addresses of the small layers.
      const int et = threadIdx.x - 64;
      const int r = q * 32 + lane;
      const int cw = p.st_rowbytes >> 2;                       // fp32 columns per staged row: 32 (or 16)
      const uint32_t bufbytes = 128u * (uint32_t)p.st_rowbytes;
      const int sh = p.st_rowbytes == 128 ? 0 : 1;
      const uint32_t xr = (uint32_t)(r >> sh) & (uint32_t)((p.st_rowbytes >> 4) - 1);
      int idx = 0;
      for (int tl = 0; tl < p.T; ++tl) {
        for (int cc = 0; cc < p.ct; cc += cw, ++idx) {
          const int buf = idx & 1;
          if (idx >= 2) {
            if (et == 0) bulk_wait_read<1>();
            epi_bar_sync();
          }
          const uint32_t row = smem_base + (uint32_t)buf * bufbytes + (uint32_t)r * p.st_rowbytes;
          for (int sub = 0; sub < cw; sub += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * p.ct + cc + sub), v);
            tmem_ld_wait();
            const uint32_t u0 = (uint32_t)(sub * 4) >> 4;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(row + (((u0 + j) ^ xr) << 4), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
          fence_proxy_async();
          epi_bar_sync();
          if (et == 0) {
            tma_reduce_add_3d(mapS, smem_base + (uint32_t)buf * bufbytes, c0 + cc, n0, t0 + tl);
            bulk_commit();
          }
        }
      }
      if (et == 0) bulk_wait_read<0>();
    } else if (p.pointwise) {
      for (int c16 = 0; c16 < p.ct; c16 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c16, v);
        tmem_ld_wait();
        if (n < p.n_real) {
          if (p.ld_c == 1 && c0 + c16 + 16 <= p.c_real && (p.ld_n & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              red_add_v4(p.dw + (long long)n * p.ld_n + c0 + c16 + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                         __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = c0 + c16 + i;
              if (c < p.c_real) atomicAdd(p.dw + (long long)n * p.ld_n + (long long)c * p.ld_c, __uint_as_float(v[i]));
            }
          }
        }
      }
    } else if (p.T == 2) {
      for (int c16 = 0; c16 < p.ct; c16 += 16) {
        uint32_t v[2][16];
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c16, v[0]);
        tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.ct + c16), v[1]);
        tmem_ld_wait();
        if (n < p.n_real) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + c16 + i;
            if (c < p.c_real)
              red_add_v2(p.dw + (long long)n * p.ld_n + (long long)c * 16 + t0, __uint_as_float(v[0][i]),
                         __uint_as_float(v[1][i]));
          }
        }
      }
    } else
    for (int c16 = 0; c16 < p.ct; c16 += 16) {
      for (int tq = 0; tq < p.T; tq += 4) {
        uint32_t v[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)((tq + j) * p.ct + c16), v[j]);
        tmem_ld_wait();
        if (n < p.n_real) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + c16 + i;
            if (c < p.c_real)
              red_add_v4(p.dw + (long long)n * p.ld_n + (long long)c * 16 + t0 + tq, __uint_as_float(v[0][i]),
                         __uint_as_float(v[1][i]), __uint_as_float(v[2][i]), __uint_as_float(v[3][i]));
          }
        }
      }
    }
  }
  if (threadIdx.x == 64 && p.trace != nullptr) trace_raw(p.trace, 5, gtimer());
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_acc, p.tmem_cols);
    else tmem_dealloc(tmem_acc, p.tmem_cols);
  }
  if (threadIdx.x == 0 && p.trace != nullptr) trace_raw(p.trace, 6, gtimer());
}

__global__ void __launch_bounds__(TC_THREADS, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ ActMaps mapsA,
                const __grid_constant__ CUtensorMap mapS, const WgParams p) {
  wgrad_body<false>(&mapG, &mapG, mapsA.m, &mapS, p, blockIdx.x, blockIdx.y, blockIdx.z);
}
// clusters of two CTAs along x (n-tiles 2i, 2i+1)
__global__ void __launch_bounds__(TC_THREADS, 2)
wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ ActMaps mapsA,
                     const __grid_constant__ CUtensorMap mapS, const WgParams p) {
  wgrad_body<true>(&mapG, &mapG, mapsA.m, &mapS, p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Many weight gradients in ONE launch.  The generator's backward has ~20 of them (14 layers, two sources per decoder
// layer), each a 10-20 us launch of a few hundred short CTAs whose phases (setup, one TMA round trip, a few MMAs, the
// reduce-add epilogue) barely overlap -- and while the one-CTA-per-SM convolution kernels run nothing else does, so
// they only took turns with the data-gradient chain.  Here every job keeps its own tensor maps and parameters (the whole
// table is a kernel parameter, 1 KB per job); CTAs of different jobs share the SMs, so one job's epilogue runs under
// another's loads.
constexpr int WG_GROUP_MAX = 24;
struct __align__(64) WgJobDev {
  CUtensorMap mapG;
  CUtensorMap mapG2;       // second source of the G operand (virtual concat), else a copy of mapG
  CUtensorMap mapA[4];
  CUtensorMap mapS;
  WgParams p;
  int gx, gy, gz, pad0;
};
struct __align__(64) WgGroupArgs {
  int njobs;
  int cta_begin[WG_GROUP_MAX + 1];
  WgJobDev jobs[WG_GROUP_MAX];
};

__global__ void __launch_bounds__(TC_THREADS, 2) wgrad_group_kernel(const __grid_constant__ WgGroupArgs g) {
  int j = 0;
  while (j + 1 < g.njobs && (int)blockIdx.x >= g.cta_begin[j + 1]) ++j;
  const WgJobDev& jb = g.jobs[j];
  const int local = (int)blockIdx.x - g.cta_begin[j];
  const int bx = local % jb.gx, r = local / jb.gx;
  wgrad_body<false>(&jb.mapG, &jb.mapG2, jb.mapA, &jb.mapS, jb.p, bx, r % jb.gy, r / jb.gy);
}

static int box_of(int ch) { return ch >= 64 ? 64 : (ch >= 32 ? 32 : 16); }
static uint32_t layout_of(int rowbytes) { return rowbytes == 128 ? 2u : (rowbytes == 64 ? 4u : 6u); }

static bool make_wg_plan(const PgConvDesc* d, WgParams& p, dim3& grid, size_t& smem, int split_cap = 0, bool want_pair = false) {
  memset(&p, 0, sizeof(p));
  if ((d->mode != PG_CONV && d->mode != PG_CONV1X1) || d->C2 != 0) return false;
  p.pointwise = d->mode == PG_CONV1X1 ? 1 : 0;
  p.stride = p.pointwise ? 1 : d->stride; p.pad = d->pad; p.B = d->B; p.Hout = d->Hout; p.Wout = d->Wout;
  p.TW = pow2_ceil(d->Wout); if (p.TW > WG_KP) p.TW = WG_KP;
  p.TH = pow2_ceil(d->Hout); if (p.TH > WG_KP / p.TW) p.TH = WG_KP / p.TW;
  p.TB = WG_KP / (p.TW * p.TH);
  p.nx = (d->Wout + p.TW - 1) / p.TW; p.ny = (d->Hout + p.TH - 1) / p.TH;
  const int nb = (d->B + p.TB - 1) / p.TB;
  p.total_tiles = p.nx * p.ny * nb;
  const int N = d->N, C = d->C1;
  p.g_box = box_of(N); p.n_atoms = 128 / p.g_box;
  p.a_box = box_of(C);
  // T taps x ct channels = 256 accumulator columns, so two CTAs fit the 512 TMEM columns of an SM
  if (p.pointwise) { p.T = 1; p.ct = C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : (C >= 32 ? 32 : 16))); }
  else if (C >= 128) { p.ct = 128; p.T = 2; }
  else if (C >= 64) { p.ct = 64; p.T = 4; }
  else if (C >= 32) { p.ct = 32; p.T = 8; }
  else { p.ct = 16; p.T = 16; }
  p.c_atoms = p.ct / p.a_box;
  p.ctiles = (C + p.ct - 1) / p.ct;
  p.g_rowbytes = p.g_box * 2; p.a_rowbytes = p.a_box * 2;
  p.g_layout = layout_of(p.g_rowbytes); p.a_layout = layout_of(p.a_rowbytes);
  p.g_boxbytes = WG_KP * p.g_rowbytes; p.a_boxbytes = WG_KP * p.a_rowbytes;
  // CTA pair: two full n-tiles, an even number of activation boxes per tap to split between the CTAs
  p.pair = want_pair && !p.pointwise && (N % 256) == 0 && p.c_atoms >= 2 && (p.c_atoms & 1) == 0 ? 1 : 0;
  p.g_stage_bytes = (p.n_atoms * p.g_boxbytes + 1023u) & ~1023u;
  p.a_stage_bytes = ((p.pair ? p.c_atoms / 2 : p.c_atoms) * p.a_boxbytes + 1023u) & ~1023u;
  // Both rings are TMA-latency-bound (a box takes ~2 us to land): the pixel-tile (G) ring is as deep as shared memory allows
  // while the tap ring keeps at least T + 2 slots (one tile's taps plus a head start on the next), at most 4 / 12 slots.
  static const int gst_env = [] { const char* e = getenv("PG_WG_GSTAGES"); return e ? atoi(e) : WG_MAX_G; }();
  int gst = gst_env < 2 ? 2 : (gst_env > WG_MAX_G ? WG_MAX_G : gst_env);
  int as = 0;
  for (;; --gst) {
    as = (int)((110u * 1024u - (uint32_t)gst * p.g_stage_bytes) / p.a_stage_bytes);
    if (as >= p.T + 2 || gst == 2) break;
  }
  p.g_stages = gst;
  if (as > WG_MAX_A) as = WG_MAX_A;
  if (as < 2) return false;
  p.a_stages = as;
  const int cols = p.T * p.ct;
  p.tmem_cols = cols <= 32 ? 32u : (cols <= 64 ? 64u : (cols <= 128 ? 128u : (cols <= 256 ? 256u : 512u)));
  // A operand = G (grad side, d->out_f32 holds its dtype), B operand = activations (d->in_dtype); both MN-major
  const uint32_t afmt = d->out_f32 == PG_F16 ? 0u : 1u, bfmt = d->in_dtype == PG_F16 ? 0u : 1u;
  p.idesc = (1u << 4) | (afmt << 7) | (bfmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.ct >> 3) << 17) |
            ((uint32_t)((p.pair ? 256 : 128) >> 4) << 24);
  {
    const int rows = N < 128 ? N : 128;              // (the last n-tile of a larger N may be partial: load all its boxes)
    p.n_atoms_load = N <= 128 ? (rows + p.g_box - 1) / p.g_box : p.n_atoms;
  }
  const int gx = (N + 127) / 128, gy = p.ctiles * (p.pointwise ? 1 : 16 / p.T);
  // Split-K over pixel tiles: every CTA ends with 128 x 256 fp32 atomics into dW, so a CTA should own enough pixel
  // tiles (>= 8, ~1 us of MMA) to amortise them; beyond that, split until ~2 CTAs per SM exist.
  // (rounded DOWN: 2 CTAs fit an SM, and a few CTAs beyond 2 x SMs would run as a second wave on an otherwise idle GPU --
  //  the largest discriminator layer had 320 CTAs for 296 slots)
  static const int split_ceil = [] { const char* e = getenv("PG_WG_SPLIT_CEIL"); return e ? atoi(e) : 0; }();
  int splits = split_ceil ? (2 * num_sms() + gx * gy - 1) / (gx * gy) : (2 * num_sms()) / (gx * gy);
  const int max_splits = p.total_tiles / 8 > 0 ? p.total_tiles / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  if (split_cap > 0 && splits > split_cap) splits = split_cap;
  if (splits > p.total_tiles) splits = p.total_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.total_tiles + splits - 1) / splits;
  splits = (p.total_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)splits);
  smem = (size_t)p.g_stages * p.g_stage_bytes + (size_t)p.a_stages * p.a_stage_bytes + 1024;
  return true;
}

bool conv_wgrad_tc_supported(const PgConvDesc* d, const void* a, const void* g, int ldg) {
  if (!tc_device_ok()) return false;
  if ((((uintptr_t)a | (uintptr_t)g) & 15) != 0 || (ldg % 8) != 0) return false;
  if (d->out_f32 != PG_BF16 && d->out_f32 != PG_F16) return false;
  if (d->out_f32 != d->in_dtype) return false;   // kind::f16 needs one operand format (mixed = illegal instruction)
  WgParams p; dim3 grid; size_t smem;
  return make_wg_plan(d, p, grid, smem);
}

// Everything a launch of one weight gradient needs: plan, parameters, tensor maps.
struct WgPrepared {
  WgParams p;
  dim3 grid;
  size_t smem;
  CUtensorMap mG, mG2, mS;
  ActMaps mA;
};

// tap_major != 0: dw is S[16][Ns = ld_n][Cs = c_stride] (fp32, zeroed by the caller); else the reference layout
static int wg_prepare(const PgConvDesc* d, const void* a, const void* g, int ldg, float* dw, int ld_n, int n_real, int c_real,
                      int tap_major, int Cs, int split_cap, WgPrepared& w, const void* g2 = nullptr, int ldg2 = 0,
                      int n_split = 0, bool want_pair = false) {
  WgParams& p = w.p;
  if (!make_wg_plan(d, p, w.grid, w.smem, split_cap, want_pair && g2 == nullptr)) {
    set_error("conv_wgrad_tc: unsupported shape");
    return PG_ERR_UNSUPPORTED;
  }
  if (!p.pointwise && !tap_major && ((((uintptr_t)dw) & 15) != 0 || (ld_n % 4) != 0)) {
    set_error("conv_wgrad_tc: dw must be 16-byte aligned with ld_n %% 4 == 0");
    return PG_ERR_UNSUPPORTED;
  }
  memset(&w.mS, 0, sizeof(w.mS));
  p.tapmajor = 0;
  if (tap_major) {
    if (p.pointwise || (((uintptr_t)dw) & 15) != 0 || (Cs % 4) != 0 || Cs < c_real || ld_n < n_real) {
      set_error("conv_wgrad_tc: bad tap-major scratch (Ns=%d Cs=%d)", ld_n, Cs);
      return PG_ERR_UNSUPPORTED;
    }
    p.tapmajor = 1;
    p.st_rowbytes = p.ct >= 32 ? 128 : 64;
    cuuint64_t dims[3] = {(cuuint64_t)Cs, (cuuint64_t)ld_n, 16};
    cuuint64_t strides[2] = {(cuuint64_t)Cs * 4, (cuuint64_t)Cs * ld_n * 4};
    cuuint32_t box[3] = {(cuuint32_t)(p.st_rowbytes / 4), 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(&w.mS, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              p.st_rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(tap-major dW Ns=%d Cs=%d) failed: %d", ld_n, Cs, (int)r);
      return PG_ERR_CUDA;
    }
  }
  p.dw = dw; p.ld_n = ld_n; p.n_real = n_real; p.c_real = c_real; p.ld_c = d->ldw > 0 ? d->ldw : 16;
  p.trace = g_trace;
  static const bool wdbg = getenv("PG_TC_DEBUG") != nullptr;
  if (wdbg)
    fprintf(stderr, "wgrad_tc: grid (%u,%u,%u) T %d ct %d tiles %d per-split %d g_stages %d a_stages %d smem %zu tmem %u\n", w.grid.x,
            w.grid.y, w.grid.z, p.T, p.ct, p.total_tiles, p.tiles_per_split, p.g_stages, p.a_stages, w.smem, p.tmem_cols);
  memset(&w.mA, 0, sizeof(w.mA));
  p.n_split = 0;
  if (g2 != nullptr) {
    // G = [g (n_split channels) | g2 (N - n_split channels)]: two tensor maps, every box inside one of them
    if (n_split <= 0 || n_split >= d->N || (n_split % p.g_box) != 0 || ((d->N - n_split) % p.g_box) != 0 || (ldg2 % 8) != 0 ||
        (((uintptr_t)g2) & 15) != 0) {
      set_error("conv_wgrad_tc: bad two-source G operand (n_split=%d of N=%d, box %d)", n_split, d->N, p.g_box);
      return PG_ERR_UNSUPPORTED;
    }
    p.n_split = n_split;
    if (int e = encode_act_map(&w.mG, g, n_split, ldg, d->B, d->Hout, d->Wout, p.g_box, p.TW, p.TH, p.TB, -1, p.g_rowbytes,
                               d->out_f32))
      return e;
    if (int e = encode_act_map(&w.mG2, g2, d->N - n_split, ldg2, d->B, d->Hout, d->Wout, p.g_box, p.TW, p.TH, p.TB, -1,
                               p.g_rowbytes, d->out_f32))
      return e;
  } else {
    if (int e = encode_act_map(&w.mG, g, d->N, ldg, d->B, d->Hout, d->Wout, p.g_box, p.TW, p.TH, p.TB, -1, p.g_rowbytes,
                               d->out_f32))
      return e;
    w.mG2 = w.mG;
  }
  const bool phased = !p.pointwise && d->stride == 2;
  for (int ph = 0; ph < (phased ? 4 : 1); ++ph)
    if (int e = encode_act_map(&w.mA.m[ph], a, d->C1, d->ld1, d->B, d->Hin, d->Win, p.a_box, p.TW, p.TH, p.TB,
                               phased ? ph : -1, p.a_rowbytes, d->in_dtype))
      return e;
  return PG_OK;
}

int conv_wgrad_tc(const PgConvDesc* d, const void* a, const void* g, int ldg, float* dw, int ld_n, int n_real,
                  int c_real, int tap_major, int Cs, cudaStream_t stream) {
  WgPrepared w;
  // (CTA pairs: default on for every eligible weight gradient -- PG_WG_PAIR=0 or pg_set_pair_mode(0) turn them off)
  static const int wg_pair_env = [] { const char* e = getenv("PG_WG_PAIR"); return e ? atoi(e) : 1; }();
  if (int e = wg_prepare(d, a, g, ldg, dw, ld_n, n_real, c_real, tap_major, Cs, 0, w, nullptr, 0, 0,
                         wg_pair_env != 0 && pair_mode() != 0))
    return e;
  static bool smem_set = false;
  if (!smem_set) {
    PG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    PG_CUDA(cudaFuncSetAttribute(wgrad_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    smem_set = true;
  }
  if (w.p.pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = w.grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = w.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PG_CUDA(cudaLaunchKernelEx(&cfg, wgrad_tc_pair_kernel, w.mG, w.mA, w.mS, w.p));
    ++g_pair_launches;
    return check_launch("wgrad_tc_pair_kernel");
  }
  wgrad_tc_kernel<<<w.grid, TC_THREADS, w.smem, stream>>>(w.mG, w.mA, w.mS, w.p);
  return check_launch("wgrad_tc_kernel");
}

// One launch for a list of weight gradients (see wgrad_group_kernel).  The pixel-tile splits of every job are capped so
// that the whole group is a few waves of the 2 x #SM resident CTAs: a job no longer has to fill the GPU by itself.
int conv_wgrad_group_tc(const PgWgradJob* jobs, int njobs, cudaStream_t stream) {
  if (njobs < 1 || njobs > WG_GROUP_MAX) {
    set_error("conv_wgrad_group_tc: 1..%d jobs per launch, got %d", WG_GROUP_MAX, njobs);
    return PG_ERR_INVALID;
  }
  static WgGroupArgs args;         // (32 KB: not on the stack; launches are serialised by the caller's thread)
  static WgPrepared w;
  const int cap = (3 * 2 * num_sms() + njobs - 1) / njobs;       // ~3 waves in total
  size_t smem = 0;
  int total = 0;
  for (int j = 0; j < njobs; ++j) {
    const PgWgradJob& jb = jobs[j];
    // splits such that this job contributes at most `cap` CTAs
    WgParams probe; dim3 g0; size_t s0;
    if (!make_wg_plan(&jb.desc, probe, g0, s0, 0)) {
      set_error("conv_wgrad_group_tc: job %d: unsupported shape", j);
      return PG_ERR_UNSUPPORTED;
    }
    int split_cap = cap / (int)(g0.x * g0.y);
    if (split_cap < 1) split_cap = 1;
    if (int e = wg_prepare(&jb.desc, jb.a, jb.g, jb.ldg, jb.dw, jb.ld_n, jb.n_real, jb.c_real, jb.tap_major, jb.Cs, split_cap, w,
                           jb.g2, jb.ldg2, jb.n_split))
      return e;
    WgJobDev& dst = args.jobs[j];
    dst.mapG = w.mG;
    dst.mapG2 = w.mG2;
    for (int ph = 0; ph < 4; ++ph) dst.mapA[ph] = w.mA.m[ph];
    dst.mapS = w.mS;
    dst.p = w.p;
    dst.gx = (int)w.grid.x; dst.gy = (int)w.grid.y; dst.gz = (int)w.grid.z;
    args.cta_begin[j] = total;
    total += (int)(w.grid.x * w.grid.y * w.grid.z);
    if (w.smem > smem) smem = w.smem;
  }
  args.njobs = njobs;
  args.cta_begin[njobs] = total;
  static bool smem_set = false;
  if (!smem_set) {
    PG_CUDA(cudaFuncSetAttribute(wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_DYN_SMEM));
    smem_set = true;
  }
  static const bool wdbg = getenv("PG_TC_DEBUG") != nullptr;
  if (wdbg) fprintf(stderr, "wgrad_group: %d jobs, %d CTAs, smem %zu\n", njobs, total, smem);
  wgrad_group_kernel<<<total, TC_THREADS, smem, stream>>>(args);
  return check_launch("wgrad_group_kernel");
}

#include "conv_res.cuh"

}  // namespace pg
