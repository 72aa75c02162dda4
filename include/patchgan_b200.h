/*
 * patchgan_b200 -- C-ABI of the B200 (sm_100a) kernels under the patchGAN hot path.
 *
 * The reference (ramanakumars/patchGAN) has no FFI of its own: every arithmetic call is a
 * PyTorch operator.  Each entry point below replaces the PyTorch operator call(s) at the cited
 * reference line(s); the Python host side (patchgan_b200/*.py) binds them with ctypes and keeps
 * the reference's module API (UNet / Discriminator / losses / Trainer).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns all memory (including workspaces); nothing is allocated or freed here;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     and is CUDA-graph capturable;
 *   - return value: 0 on success, non-zero PgStatus otherwise; pg_last_error() gives the text
 *     (thread-local);
 *   - activations are NHWC 16-bit (bf16 or f16, see PgDType) with an explicit pixel stride `ld` (elements) so that a tensor can be a
 *     channel slice of a wider buffer (virtual concat); channel counts seen by the conv kernels are
 *     multiples of 16 (the host zero-pads 3/4/1/7-channel tensors and the packed weights);
 *   - packed weights are bf16 [N][16 taps][C] (tap = kh*4+kw, C = C1+C2 in concat order).
 */
#ifndef PATCHGAN_B200_H
#define PATCHGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum PgStatus {
  PG_OK = 0,
  PG_ERR_INVALID = 1,     /* bad argument / unsupported shape            */
  PG_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed         */
  PG_ERR_UNSUPPORTED = 3  /* requested implementation cannot run this op */
} PgStatus;

typedef enum PgAct {
  PG_ACT_NONE = 0,
  PG_ACT_RELU = 1,
  PG_ACT_LEAKYRELU = 2,   /* slope 0.2  (unet.py:17, disc.py:20) */
  PG_ACT_TANH = 3,
  PG_ACT_SIGMOID = 4
} PgAct;

/* element types of activation / gradient tensors.  Forward operands may be bf16 or f16 (same tensor-core
 * rate; f16 carries 3 more mantissa bits and post-InstanceNorm activations are O(1)); gradients are bf16. */
typedef enum PgDType {
  PG_BF16 = 0,
  PG_F32 = 1,
  PG_F16 = 2
} PgDType;

typedef enum PgConvMode {
  PG_CONV = 0,            /* nn.Conv2d(k=4, stride, pad)                          */
  PG_CONVT = 1,           /* nn.ConvTranspose2d(k=4, stride=2, pad=1)             */
  PG_CONV1X1 = 2          /* pointwise: out[pix][n] = sum_c in[pix][c] * W[n][c]  (W rows at stride ldw).  Used to run the
                             one-real-channel layers as [pixels x C] x [C x 16 taps] products (see pg_taps_scatter /
                             pg_taps_gather) */
} PgConvMode;

typedef enum PgImpl {
  PG_IMPL_AUTO = 0,
  PG_IMPL_SIMT = 1,       /* CUDA-core implicit GEMM (any shape; validation path)  */
  PG_IMPL_TCGEN05 = 2,    /* TMA + tcgen05.mma + TMEM implicit GEMM               */
  PG_IMPL_SKINNY = 3      /* CUDA-core streaming kernels for layers with one real channel on one side (HBM-bound);
                             PG_IMPL_AUTO picks them when the shape qualifies */
} PgImpl;

/* Geometry of one 4x4 convolution-shaped contraction.
 *   PG_CONV : out[b,oy,ox,n] = sum_{kh,kw,c} in[b, oy*stride-pad+kh, ox*stride-pad+kw, c] * W[n][kh*4+kw][c]
 *   PG_CONVT: out[b,oy,ox,n] = sum_{kh,kw,c : oy=2*iy-1+kh, ox=2*ix-1+kw} in[b,iy,ix,c] * W[n][kh*4+kw][c]
 * `in` is the virtual concat of src1 (C1 channels, stride ld1) and src2 (C2 channels, stride ld2).
 * The same two forms express every data-gradient:
 *   dgrad(Conv2d s=2)        = PG_CONVT over dY with W'[ci][tap][co]
 *   dgrad(Conv2d s=1,p=1)    = PG_CONV stride 1 pad 2 over dY with taps flipped
 *   dgrad(ConvTranspose2d)   = PG_CONV stride 2 pad 1 over dY with W'[ci][tap][co]
 */
typedef struct PgConvDesc {
  int32_t mode;      /* PgConvMode */
  int32_t stride;    /* PG_CONV: 1 or 2; PG_CONVT: 2 */
  int32_t pad;       /* PG_CONV: 1 or 2; PG_CONVT: 1 */
  int32_t B, Hin, Win;
  int32_t Hout, Wout;
  int32_t C1, C2;    /* multiples of 16; C2 = 0 without concat */
  int32_t ld1, ld2;
  int32_t N;         /* output channels computed, multiple of 16 */
  int32_t ldo;       /* output pixel stride */
  int32_t n_valid;   /* channels >= n_valid are stored as 0 (padding must stay 0 after sigmoid) */
  int32_t act;       /* PgAct fused into the epilogue */
  int32_t out_f32;   /* PgDType of the output: PG_BF16, PG_F32 or PG_F16 */
  int32_t has_bias;
  int32_t in_dtype;  /* PgDType of src1/src2 and of the packed weights: PG_BF16 or PG_F16 */
  int32_t n_first;   /* only output channels [n_first, n_valid) are needed by the caller; channels below n_first may be
                        left unwritten (0 = all).  Lets the 1-real-channel layers run as streaming kernels. */
  int32_t c_valid;   /* real (non-padding) input channels of src1 when C2 == 0; 0 = unknown / all C1 */
  int32_t ldw;       /* PG_CONV1X1: row stride (elements) of the weight matrix W[n][c]; 0 = C1 + C2 */
} PgConvDesc;

const char* pg_last_error(void);
int pg_version(void);
/* number of kernels this library has launched in this process (bench.py reports the per-step delta) */
int64_t pg_launch_count(void);
/* PgImpl the last pg_conv_fwd / pg_conv_wgrad of this thread dispatched to (profiling aid) */
int pg_last_conv_impl(void);
/* number of PG_IMPL_AUTO calls of this process that found no tensor-core plan for their shape and ran on the CUDA-core
 * kernel (the first one is also reported on stderr).  bench.py and the step tests require it to stay 0. */
int64_t pg_fallback_count(void);
/* CTA-pair variant of the persistent convolution kernel (tcgen05.mma.cta_group::2 on clusters of two CTAs: a 256-pixel x BN
 * tile per pair, each CTA fetching half of the weight tile).  Eligible: tile width 128 or 256 output channels, an even number
 * of >= 148 pixel tiles, no split-K.  mode 0 = never, 1 = only where it measured faster than two co-resident single CTAs
 * (128-wide tiles, >= 4 tile pairs per cluster), 2 = every eligible shape (parity tests, A/B runs), -1 = back to the default
 * (environment variable PG_TC_PAIR, else 1).  pg_pair_launch_count: convolutions of this process that took the variant. */
int pg_set_pair_mode(int32_t mode);
int64_t pg_pair_launch_count(void);
/* 1 if the library was built with the tcgen05 path and the current device is sm_100. */
int pg_tcgen05_available(void);
/* Debug hook (kernel tuning only): when buf != NULL every conv_tc CTA writes 16 uint64 (globaltimer ns at entry, after
 * setup, first operands landed, last MMA issued, accumulator ready, epilogue done, exit; SM id) at buf[cta*16 ...]. */
int pg_debug_set_trace(void* buf);
/* Diagnostics: a one-thread kernel on `stream` that writes %globaltimer (ns) to *slot when the stream gets there; can be
 * captured into a CUDA graph (tools/timeline.py builds the in-graph timeline of a step from it).  Not a reference op. */
int pg_debug_stamp(uint64_t* slot, void* stream);

/* Caller-owned device scratch for the split-K convolutions (the 2x2 .. 16x16 bottleneck layers run their K = 16*Cin
 * reduction on a cluster of CTAs that exchange fp32 partial tiles through this L2-resident buffer).  256-byte aligned;
 * 64 MB covers every layer of the reference configurations; NULL / 0 disables split-K (N is split instead).  The buffer must
 * stay alive while convolutions are in flight.  State is per device (the current one).  Successive launches take
 * successive slices, so that launches in flight on different streams never share one, and the slices are never reused
 * behind the caller's back: calling this again rewinds the cursor -- do it where no convolution is in flight (the Python
 * engine does at the start of every step) -- and a launch that no longer fits returns PG_ERR_INVALID. */
int pg_conv_set_workspace(void* ws, int64_t bytes);

/* ---- convolutions: replaces aten::convolution behind nn.Conv2d / nn.ConvTranspose2d
 *      (unet.py:19, unet.py:53, disc.py:19,27,37,45) and their autograd dgrad ---- */
/* out2 (nullable): a second, bf16 copy of the output with the same pixel stride.  tcgen05 kind::f16 needs both
 * MMA operands in ONE 16-bit format, so when forward tensors are f16 the weight-gradient GEMMs (whose other
 * operand is a bf16 gradient) read this bf16 twin. */
int pg_conv_fwd(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed,
                const float* bias, void* out, void* out2, int impl, void* stream);

/* pg_conv_fwd + the InstanceNorm statistics of its output (unet.py:19-20, 53-55) in one call:
 *   sums[(b*N + n)*2 + {0,1}] += {sum, sum of squares} of out[b,:,:,n]   (caller zeroes sums; out: all N channels, ldo >= N)
 * On the tcgen05 path the sums are taken from the fp32 accumulators in the epilogue (no second pass over the output);
 * tile shapes that cannot fuse run pg_instnorm_stats after the convolution. */
int pg_conv_fwd_stats(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed, const float* bias,
                      void* out, float* sums, int impl, void* stream);

/* Data-gradient through a convolution AND the activation in front of it, in one call (autograd of disc.py:19-42):
 *   dx = conv_dgrad(dy) * act'(y),  act = d->act, y = the activation's saved OUTPUT (16-bit NHWC, pixel stride ldy)
 * `d` is the data-gradient geometry as for pg_conv_fwd (d->act is NOT applied forward); dx: bf16, all N channels.
 * On the tcgen05 path the product is taken on the fp32 accumulators in the epilogue; otherwise the plain data-gradient
 * runs first and pg_act_bwd_from_output follows in place. */
int pg_conv_dgrad_act(const PgConvDesc* d, const void* dy, const void* w_packed, void* dx, const void* y, int32_t ldy,
                      int32_t y_dtype, int impl, void* stream);

/* ---- convolution + InstanceNorm2d + activation (+ Dropout) in ONE launch: DownSampleBlock / UpSampleBlock
 *      (unet.py:19-28, 53-66) forward, and their autograd backward fused into the data-gradient convolution that
 *      produces dL/d(block output).  The accumulators of the whole layer stay resident in tensor memory (148 SMs x 512
 *      columns x 128 lanes = 9.7 M fp32 values) across a grid barrier: phase 1 reduces the per-(image, channel) sums from
 *      the fp32 accumulators, phase 2 normalises / activates and stores the final 16-bit tensor.  The pre-norm tensor is
 *      never written.  Layers whose output does not fit tensor memory (or with a 1 x 1 map) are refused
 *      (pg_conv_norm_supported == 0): run pg_conv_fwd_stats + pg_norm_act_fwd / pg_norm_act_bwd instead.
 *      One such launch may be in flight per device at a time (it occupies every SM until its grid barrier). ---- */
/* SMs the one-launch kernels below may occupy (0 = all).  Their CTAs wait for each other on a grid barrier, so they must
 * not compete for the last SMs with a kernel that itself waits for other GPUs: the data-parallel Trainer overlaps NCCL
 * all-reduces with the backward pass and keeps 16 SMs out of these kernels' reach. */
int pg_set_sm_limit(int32_t n);
typedef enum PgFusedKind { PG_FUSED_FWD = 0, PG_FUSED_BWD = 1 } PgFusedKind;
typedef struct PgFusedNorm {
  int32_t kind;        /* PgFusedKind */
  int32_t act;         /* PgAct of the normalised block (none / relu / leakyrelu / tanh) */
  int32_t n_norm;      /* BWD: output channels [0, n_norm) are gradients of the normalised block's output; channels
                          [n_norm, N) (the skip half of a concat input) are stored unchanged.  FWD: ignored (= N) */
  float drop_p;        /* Dropout probability (0 = none); the mask is uniform(mix(*seed, salt), pixel*C + channel) >= p */
  const uint64_t* seed;
  uint64_t salt;
  float* sums;         /* FWD: zeroed [B][N][2], receives (sum, sum of squares) of the conv output per (image, channel).
                          BWD: the sums the forward call of the block left, [B][n_norm][2] */
  float* bsums;        /* BWD: zeroed workspace [B][n_norm][2] */
  uint32_t* sync;      /* TWO zeroed 32-bit counters (grid barriers) */
  void* xhat;          /* FWD: optional extra output: the normalised pre-activation, same dtype / layout as out with pixel
                          stride xhat_ld (needed by BWD for relu / tanh / dropout blocks).  BWD: that tensor, or NULL */
  int32_t xhat_ld;
  const void* y;       /* BWD, when xhat == NULL: the block's saved output (invertible activation: leakyrelu / none) */
  int32_t y_ld;
  int32_t y_dtype;     /* PgDType of y / xhat */
  const void* dskip;   /* BWD: bf16 gradient arriving over the skip connection, added before the activation backward */
  int32_t dskip_ld;
  void* ws;            /* optional scratch (256-byte aligned, need not be zeroed, must not be shared by launches in flight):
                          with it, layers of a few output tiles and K in the thousands (the 2x2 .. 16x16 maps) split K over
                          the SMs and exchange fp32 partial tiles through it; 8 MB covers every layer of the reference
                          configurations.  NULL: such layers split N instead (slower). */
  int64_t ws_bytes;
} PgFusedNorm;
/* 1 if the fused kernel can run this geometry (d as for pg_conv_fwd), else 0 */
int pg_conv_norm_supported(const PgConvDesc* d, const PgFusedNorm* fn, int32_t has_twin);
/* out (+ bf16 twin out2, nullable) = dropout(act(instance_norm(conv(src1 | src2)))); d->out_f32 = PG_F16 / PG_BF16,
 * no bias (the blocks have none), d->act is ignored (fn->act) */
int pg_conv_norm_fwd(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed, void* out, void* out2,
                     const PgFusedNorm* fn, void* stream);
/* dx[.., 0:n_norm] = instance_norm_backward(act_backward(dropout_backward(dgrad(dy) [+ dskip]))),
 * dx[.., n_norm:N] = dgrad(dy);  d = the data-gradient geometry as for pg_conv_fwd, dx bf16 */
int pg_conv_dgrad_norm_bwd(const PgConvDesc* d, const void* dy, const void* w_packed, void* dx, const PgFusedNorm* fn,
                           void* stream);

/* weight gradient of PG_CONV geometry `d` (autograd wgrad of unet.py:19,53 / disc.py:19-45):
 *   dw[n*ld_n + c*16 + tap] += sum_{b,oy,ox} g[b,oy,ox,n] * a[b, oy*s-p+kh, ox*s-p+kw, c]
 * g: [B,Hout,Wout] x N (stride ldg), a: [B,Hin,Win] x C1 (stride ld1).  Only n < n_real, c < c_real are
 * written.  Accumulates atomically into dw (caller zeroes).  For ConvTranspose2d swap the roles:
 * a = dY (2H x 2W), g = layer input.
 * PG_CONV1X1 (pointwise): dw[n*ld_n + c*d->ldw] += sum_pix g[pix,n] * a[pix,c]. */
int pg_conv_wgrad(const PgConvDesc* d, const void* a, const void* g, int32_t ldg, float* dw,
                  int32_t ld_n, int32_t n_real, int32_t c_real, int impl, void* stream);
/* (d->in_dtype is the type of `a`; d->out_f32 is reused as the PgDType of `g`: PG_BF16 or PG_F16; the tcgen05
 * implementation needs both to be the same type) */

/* Same contraction accumulated TAP-MAJOR: S[(tap*Ns + n)*Cs + c] += sum g[..,n] * a[..(tap)..,c], n < Ns, c < Cs (fp32, the
 * caller zeroes S; Cs % 4 == 0).  On the tcgen05 path every [128 n][32 c] accumulator tile of a tap is added with one TMA
 * bulk reduce (cp.reduce.async.bulk.tensor .add) instead of per-element atomics; pg_grad_finalize_multi then writes the
 * reference layout (Cout, Cin, 4, 4) of every layer of a network in one launch. */
int pg_conv_wgrad_tapmajor(const PgConvDesc* d, const void* a, const void* g, int32_t ldg, float* S, int32_t Ns, int32_t Cs,
                           int impl, void* stream);
/* A list of weight gradients in ONE launch (tcgen05 only): the generator's backward has ~20 of them, each too small to
 * fill the GPU and too short to overlap its own phases.  Every job is one pg_conv_wgrad (tap_major = 0: dw / ld_n as
 * there) or pg_conv_wgrad_tapmajor (tap_major = 1: dw = S, ld_n = Ns, Cs) call; at most 24 jobs per launch.  All operands
 * must be ready on `stream`. */
typedef struct PgWgradJob {
  PgConvDesc desc;
  const void* a;
  const void* g;
  int32_t ldg;
  int32_t tap_major;
  float* dw;
  int32_t ld_n, n_real, c_real, Cs;
  const void* g2;      /* optional: g is the virtual concat [g (n_split channels) | g2 (desc.N - n_split)] -- the two sources of
                          a decoder layer's input (unet.py:127) in one job, so that the tap-shifted operand `a` is read once
                          for both; n_split and desc.N - n_split must be multiples of min(64, ...) channel boxes */
  int32_t ldg2, n_split;
} PgWgradJob;
int pg_conv_wgrad_group(const PgWgradJob* jobs, int32_t njobs, void* stream);

typedef struct PgGradJob {
  const float* S;      /* [16][Ns][Cs] */
  float* dst;          /* dst[n*ld_n + c*16 + tap] = S[tap][n][c], n < N, c < C (overwrites) */
  int64_t ld_n;
  int32_t N, C, Ns, Cs;
  int32_t tile_begin, ctiles;   /* first block of the job in the launch; ctiles = ceil(C/32); the job has N*ctiles blocks */
} PgGradJob;
int pg_grad_finalize_multi(const PgGradJob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream);

/* ---- layers with ONE real channel on one side (generator output ConvTranspose2d(2nf -> 1), unet.py:106-107;
 *      discriminator last Conv2d(8ndf -> 1), disc.py:45; the mask-channel data-gradient of the discriminator's first
 *      layer, trainer.py:84-89).  They run as pointwise (PG_CONV1X1) products over the 16 taps on the tensor cores:
 *        forward        P[q][tap] = sum_c in[q][c] W[tap][c];   out[p] = act(bias + sum_tap P[q(p,tap)][tap])   (scatter)
 *        data-gradient  G[q][tap] = dy[p(q,tap)]  (gather);     dx[q][c] = sum_tap G[q][tap] W[c][tap]
 *        weight-grad.   dW[c][tap] = sum_q in[q][c] G[q][tap]    (pg_conv_wgrad with PG_CONV1X1, ld_n = 1, ldw = 16)
 *      q = pixel of the wide tensor (Hq x Wq), p = pixel of the 1-channel tensor (Hp x Wp).  mode / stride / pad are
 *      the layer's: PG_CONV: q = p*stride - pad + k;  PG_CONVT: p = 2q - 1 + k. ---- */
/* P: f32 [B,Hq,Wq,ldp] (taps in channels 0..15) -> element `ch` of out [B,Hp,Wp,ldo] (PgDType out_dtype); bias: 1 float or NULL */
int pg_taps_scatter(int32_t mode, int32_t stride, int32_t pad, int32_t B, int32_t Hq, int32_t Wq, int32_t Hp, int32_t Wp,
                    const float* P, int32_t ldp, const float* bias, int32_t act, void* out, int32_t out_dtype, int32_t ldo,
                    int32_t ch, void* stream);
/* element `ch` of the 16-bit src [B,Hp,Wp,lds] -> G [B,Hq,Wq,16] of the same 16-bit type (zeros where the tap falls outside) */
int pg_taps_gather(int32_t mode, int32_t stride, int32_t pad, int32_t B, int32_t Hq, int32_t Wq, int32_t Hp, int32_t Wp,
                   const void* src, int32_t lds, int32_t ch, void* G, void* stream);

/* data-gradient of a tap-product layer fused with the backward of the activation in front of it (disc.py:39-46 backward):
 *   dx[q][c] = (sum_tap G[q][tap] * w16[c][tap]) * act'(y[q][c]),  q < nq, c < C
 * G: bf16 [nq][16] (pg_taps_gather), w16: bf16 [C][16] (the layer's master weight), y: the activation's saved OUTPUT
 * (16-bit, pixel stride ldy; NULL: no activation), dx: bf16, pixel stride lddx.  K = 16 per output: HBM-bound on the CUDA
 * cores (as a tensor-core GEMM it is one MMA per tile in front of a 128 x C epilogue). */
int pg_taps_dgrad_act(const void* G, const void* w16, void* dx, int32_t lddx, const void* y, int32_t ldy, int32_t y_dtype,
                      int32_t act, int64_t nq, int32_t C, void* stream);

/* bias gradient: db[n] += sum_m g[m*ldg + n], n < n_real (disc.py:19,45 biases) */
int pg_colsum(const void* g, int64_t M, int32_t ldg, int32_t n_real, float* db, void* stream);

/* ---- layout ---- */
/* NCHW float -> NHWC bf16 channel slice [c_off, c_off+C) of a buffer with pixel stride ld
 * (trainer.py:55-60,65,96: .to(device) + torch.cat feeding the nets) */
int pg_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int32_t B, int32_t C, int32_t H, int32_t W,
                                  int32_t ld, int32_t c_off, int32_t dst_dtype, void* stream);
/* NHWC (src_f32 = PgDType of src, pixel stride ld, channel offset c_off) -> NCHW float */
int pg_unpack_nhwc_to_nchw_f32(const void* src, int32_t src_f32, float* dst, int32_t B, int32_t C, int32_t H,
                               int32_t W, int32_t ld, int32_t c_off, void* stream);
/* f32 NHWC (stride lds) channels [0,C) -> bf16 NHWC (stride ldd) channels [c_off, c_off+C) */
int pg_copy_f32_to_bf16_slice(const float* src, int32_t lds, void* dst, int32_t ldd, int32_t c_off, int32_t C,
                              int64_t npix, int32_t dst_dtype, void* stream);
/* fp32 reference-layout weight -> packed bf16 [Np][16][C1p+C2p].
 * src element (n, c, tap) is at src[n*sn + c*sc + tap]; c < C1 maps to packed channel c, C1 <= c < C1+C2 maps to
 * C1p + (c-C1); flip != 0 reverses the taps (15 - tap); everything else is zero. */
int pg_pack_weight(const float* src, void* dst, int32_t N, int32_t Np, int32_t C1, int32_t C1p, int32_t C2,
                   int32_t C2p, int64_t sn, int64_t sc, int32_t flip, int32_t dst_dtype, void* stream);

/* Two NCHW float sources -> one full NHWC row per pixel [src1 | src2 | zeros up to ld] (ld = 16 or 32), plus an optional
 * bf16 twin dst2: the whole torch.cat((input, mask), 1) + cast + pad of trainer.py:65,96 in one pass. */
int pg_pack2_nchw_rows(const float* src1, int32_t C1, const float* src2, int32_t C2, void* dst, void* dst2, int32_t B,
                       int32_t H, int32_t W, int32_t ld, int32_t dst_dtype, void* stream);
/* Channel-major stride-2 im2col of a FIRST layer's input (Conv2d k4 s2 p1 with 3 / 4 real input channels; unet.py:84,
 * disc.py:19): dst[o*K + (k_off + c)*16 + kh*4 + kw] = src[b, c, 2oy-1+kh, 2ox-1+kw] for output pixel o = (b,oy,ox), 0 outside.
 * src element (b,c,y,x) is at src[b*sb + c*sc + y*sy + x*sx] (NCHW planes or channels of an f32 NHWC tensor).  With
 * k = c*16 + tap the layer is a PG_CONV1X1 product with its (Cout, Cin*16) weight matrix, and pg_conv_wgrad (PG_CONV1X1,
 * ld_n = Cin*16, ldw = 1) writes the weight-gradient straight into the reference layout.  dst2: optional bf16 twin. */
int pg_im2col_s2(const float* src, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int32_t C, int32_t B, int32_t H, int32_t W,
                 void* dst, void* dst2, int32_t K, int32_t k_off, int32_t dst_dtype, void* stream);
/* The discriminator batch of trainer.py:65,96 in one launch: dst = im2col rows of [cat(x, .) ; cat(x, y)] for 2B images
 * (x: NCHW (B,Cx,H,W), y: NCHW (B,Cy,H,W)); the mask columns of the first B images are filled later from the generator's
 * output with pg_im2col_s2 (k_off = Cx). */
int pg_im2col_s2_pair(const float* x, int32_t Cx, const float* y, int32_t Cy, int32_t B, int32_t H, int32_t W, void* dst,
                      void* dst2, int32_t K, int32_t dst_dtype, void* stream);
/* Every weight tensor of a network in one launch.  jobs_dev: DEVICE array; each job is one pg_pack_weight call;
 * tile_begin = first 8x32x16 brick of the job in the launch, ctiles = ceil((C1p+C2p)/32).
 * flip == 2: flat job, dst[i] = convert(src[i]) for i < sn, 4096 elements per brick (operand copies that keep the
 * master layout: first-layer [N][Cin*16], tap-product [Cin][16]). */
typedef struct PgPackJob {
  const float* src;
  void* dst;
  int64_t sn, sc;
  int32_t N, Np, C1, C1p, C2, C2p, flip, dst_dtype;
  int32_t tile_begin, ctiles;
} PgPackJob;
int pg_pack_weights_multi(const PgPackJob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream);

/* ---- InstanceNorm2d(affine=False, eps=1e-5) + activation + Dropout(0.2)
 *      (unet.py:20-28,55-66; disc.py:32,42) ---- */
/* In this group x_f32 / y_f32 are PgDType values (PG_BF16, PG_F32, PG_F16); dy / dx gradients are bf16.
 * The backward calls (pg_norm_act_bwd*) accept one of two flags OR-ed into x_f32, telling what the saved tensor x holds
 * when the forward was the one-launch kernel (pg_conv_norm_fwd), which never writes the pre-norm tensor: */
#define PG_X_IS_XHAT 0x100    /* x is the normalised pre-activation xhat itself */
#define PG_X_IS_OUTPUT 0x200  /* x is the block's output y = act(xhat), act = leakyrelu / none, no dropout (xhat is recovered) */
/* sums[(b*C + c)*2 + {0,1}] += {sum, sum of squares} over the HW pixels of image b (caller zeroes sums) */
int pg_instnorm_stats(const void* x, int32_t x_f32, int32_t B, int64_t HW, int32_t C, int32_t ld, float* sums,
                      void* stream);
/* y = dropout(act((x - mean) * rstd)); mean/rstd from sums (sums == NULL: no normalisation).
 * drop_p == 0: no dropout; else keep = uniform(mix(*seed, salt), element index) >= drop_p, scaled 1/(1-p).
 * seed is a DEVICE counter (so a captured CUDA graph draws a fresh mask every replay), salt is per layer;
 * the mask is regenerated in backward from (seed, salt, index) -- no mask tensor is stored. */
/* y2 (nullable): bf16 twin of y, same stride (see pg_conv_fwd). */
int pg_norm_act_fwd(const void* x, int32_t x_f32, const float* sums, void* y, int32_t y_f32, void* y2, int32_t B, int64_t HW,
                    int32_t C, int32_t ldx, int32_t ldy, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt,
                    void* stream);
/* backward, pass 1: with dxhat = (dy1 [+ dy2]) * mask * act'(xhat):
 *   bsums[(b*C+c)*2 + {0,1}] += {sum dxhat, sum dxhat*xhat}   (caller zeroes) */
int pg_norm_act_bwd_reduce(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                           const void* dy2, int32_t ld2, float* bsums, int32_t B, int64_t HW, int32_t C, int32_t ldx,
                           int32_t act, float drop_p, const uint64_t* seed, uint64_t salt, void* stream);
/* backward, pass 2: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))  -> bf16 (stride lddx).
 * sums == NULL (no norm): dx = dxhat. */
int pg_norm_act_bwd_apply(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1,
                          const void* dy2, int32_t ld2, const float* bsums, void* dx, int32_t lddx, int32_t B,
                          int64_t HW, int32_t C, int32_t ldx, int32_t act, float drop_p, const uint64_t* seed,
                          uint64_t salt, void* stream);
/* Both passes in one call.  Maps of up to 1024 pixels (the 2x2 .. 32x32 layers) run as ONE launch (a block owns an image and
 * 8 channels, keeps dxhat in registers between the reduction and the apply); larger maps run the two kernels above.
 * bsums: zeroed workspace [B][C][2] (only touched by the two-pass path). */
int pg_norm_act_bwd(const void* x, int32_t x_f32, const float* sums, const void* dy1, int32_t ld1, const void* dy2,
                    int32_t ld2, float* bsums, void* dx, int32_t lddx, int32_t B, int64_t HW, int32_t C, int32_t ldx,
                    int32_t act, float drop_p, const uint64_t* seed, uint64_t salt, void* stream);
/* ---- norm_layer = nn.BatchNorm2d (unet.py:77: the blocks call norm_layer(output_filt) -> BatchNorm2d(C): affine weight /
 *      bias, running statistics, momentum 0.1, eps 1e-5).  The kernels above normalise image b with sums[b][c]; BatchNorm2d
 *      uses one mean / variance per channel over the whole batch:
 *      pg_bn_fold_fwd replaces every image's (sum, sum of squares) pair by the batch mean of the pairs, so the same kernels
 *      then normalise with batch statistics.  training != 0: also running_mean / running_var <- (1 - momentum) * running +
 *      momentum * (batch mean / unbiased batch variance), as aten::batch_norm; training == 0 (eval): the pairs are filled
 *      from the running statistics instead.  Channels >= c_real (zero padding) are left out of the running buffers.
 *      The *_affine_* variants insert z = gamma * xhat + beta (BatchNorm2d.weight / .bias, c_real entries; padded channels
 *      use (1, 0)) between the normalisation and the activation, forward and backward.
 *      pg_bn_fold_bwd: after pg_norm_affine_act_bwd_reduce left bsums[b][c] = (sum g, sum g*xhat) per image (g = dL/dz):
 *      dbeta[c] += sum_b sum g, dgamma[c] += sum_b sum g*xhat, and the pairs are replaced by their batch mean (training) or
 *      zero (eval: constant statistics), so pg_norm_affine_act_bwd_apply writes
 *      dx = gamma * rstd * (g - mean(g) - xhat * mean(g*xhat)) with batch-wide means. ---- */
int pg_bn_fold_fwd(float* sums, int32_t B, int32_t C, int64_t HW, float* running_mean, float* running_var, int32_t c_real,
                   float momentum, int32_t training, void* stream);
int pg_bn_fold_bwd(float* bsums, int32_t B, int32_t C, float* dgamma, float* dbeta, int32_t c_real, int32_t training,
                   void* stream);
int pg_norm_affine_act_fwd(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                           int32_t c_real, void* y, int32_t y_f32, void* y2, int32_t B, int64_t HW, int32_t C, int32_t ldx,
                           int32_t ldy, int32_t act, float drop_p, const uint64_t* seed, uint64_t salt, void* stream);
int pg_norm_affine_act_bwd_reduce(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                                  int32_t c_real, const void* dy1, int32_t ld1, const void* dy2, int32_t ld2, float* bsums,
                                  int32_t B, int64_t HW, int32_t C, int32_t ldx, int32_t act, float drop_p,
                                  const uint64_t* seed, uint64_t salt, void* stream);
int pg_norm_affine_act_bwd_apply(const void* x, int32_t x_f32, const float* sums, const float* gamma, const float* beta,
                                 int32_t c_real, const void* dy1, int32_t ld1, const void* dy2, int32_t ld2,
                                 const float* bsums, void* dx, int32_t lddx, int32_t B, int64_t HW, int32_t C, int32_t ldx,
                                 int32_t act, float drop_p, const uint64_t* seed, uint64_t salt, void* stream);
/* *ctr += inc on the device (advances the dropout seed once per step, graph-capturable) */
int pg_counter_add(uint64_t* ctr, uint64_t inc, void* stream);
/* activation backward from the saved OUTPUT y (layers without norm: unet.py:97-99,106-107; disc.py):
 *   dx = dy * act'(y)   (relu/leakyrelu/tanh/sigmoid are all recoverable from y) */
int pg_act_bwd_from_output(const void* y, int32_t y_f32, int32_t ldy, const void* dy, int32_t lddy, void* dx,
                           int32_t lddx, int64_t npix, int32_t C, int32_t act, void* stream);
/* nn.Softmax(dim=1) over the first C channels of each pixel (unet.py:47), f32 NHWC in/out */
int pg_softmax_fwd(const float* x, float* y, int64_t npix, int32_t C, int32_t ld, void* stream);

/* ---- losses (losses.py:18-39, trainer.py:71-85,101-103) ---- */
/* Per-sample sums for fc_tversky / weighted BCE / MAE over p (f32 NHWC, stride ld, C channels) and the
 * target t (NCHW float, as given by the user):
 *   part[b*8 + 0..4] += { sum t*p, sum t, sum p, sum |p-t|, sum_c w_c * bce(p,t) }   (caller zeroes)
 *   chsum[b*C + c]  : input (sum_hw t) when wbce_w != NULL semantics are needed -- see pg_target_chsum. */
int pg_target_chsum(const float* t, float* chsum, int32_t B, int32_t C, int64_t HW, void* stream);
int pg_seg_loss_partials(const float* p, int32_t ld, const float* t, const float* chsum, float* part, int32_t B,
                         int32_t C, int64_t HW, int32_t loss_type, void* stream);
/* Finalise: writes losses[slot] = seg_alpha * loss and coef[b*4..] used by pg_gen_out_bwd.
 * loss_type: 0 fc_tversky(beta,gamma), 1 weighted_bce, 2 MAE. */
int pg_seg_loss_finalize(const float* part, float* coef, float* losses, int32_t slot, int32_t B, int32_t C,
                         int64_t HW, int32_t loss_type, float beta, float gamma, float seg_alpha, void* stream);
/* d(raw) of the generator's last layer: (dseg + dD) * final_act'(p)  -> bf16 NHWC (stride lddx).
 * dD: bf16 NHWC gradient from the discriminator's first layer at channel offset dd_off (NULL: none).
 * final_act: PgAct or 5 = softmax.  up = upstream scale of the seg loss (1.0). */
int pg_gen_out_bwd(const float* p, int32_t ld, const float* t, const float* chsum, const float* coef,
                   const void* dD, int32_t lddd, int32_t dd_off, void* dx, int32_t lddx, int32_t B, int32_t C,
                   int64_t HW, int32_t loss_type, int32_t final_act, float beta, void* stream);
/* nn.BCELoss(p, const label) on the discriminator map p (f32, one value per pixel, stride ld):
 *   losses[slot] += mean bce;  if dz != NULL: dz[pix*lddz + 0] = gscale * dBCE/dp * p(1-p) (bf16), other
 *   channels of the pixel (1..lddz-1) = 0.   (trainer.py:84,101,102 + disc.py:46 sigmoid backward) */
int pg_bce_const(const float* p, int32_t ld, float label, float gscale, float* losses, int32_t slot, void* dz,
                 int32_t lddz, int64_t npix, void* stream);

/* The standalone loss API (losses.py:5-39 called by user code on arbitrary tensors; the Trainer step uses the fused
 * calls above).  nn.BCELoss()(p, t) with a general target (losses.py:39): losses[slot] += mean bce(p, t), log clamped
 * at -100 like aten::binary_cross_entropy; its gradient dp = *gout * (p - t) / max(p (1 - p), 1e-12) / n (gout: DEVICE
 * scalar, the upstream gradient autograd hands over -- no host read). */
int pg_bce_mean(const float* p, const float* t, int64_t n, float* losses, int32_t slot, void* stream);
int pg_bce_mean_bwd(const float* p, const float* t, int64_t n, const float* gout, float* dp, void* stream);
/* gradient of the per-sample sums (sum t*p, sum p) behind tversky / fc_tversky(batch_mean=False) (losses.py:5-31) wrt the
 * prediction, NCHW float: dp[b, i] = g_tp[b] * t[b, i] + g_sp[b], i < chw */
int pg_sample_sums_bwd(const float* t, const float* g_tp, const float* g_sp, float* dp, int32_t B, int64_t chw,
                       void* stream);

/* ---- optim.Adam (trainer.py:169-172, 90, 107), one launch for a whole flat parameter buffer.
 * hyper (device): [0] = lr.  step (device int32): number of steps taken so far; incremented by the call. ---- */
int pg_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                 float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The same update for a sub-range of the flat buffers (pointers already offset, 16-byte aligned): the parameters of a
 * network may be updated in several launches of one optimizer step -- e.g. the layers whose gradients are final first,
 * on a side stream, while the backward pass of the remaining layers still runs.  Every launch of the step uses the step
 * count as it was before the step; exactly one of them (the last in stream order) passes bump = 1. */
int pg_adam_step_range(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, int32_t* step,
                       float beta1, float beta2, float eps, float grad_scale, int32_t bump, void* stream);

/* ---- device input pipeline: COCOStuffDataset.__getitem__ (io.py:38-58) from RAW uint8 data, "next" row f3 ----
 * img: uint8 [B][3][Hs][Ws] (decoded RGB), lab: uint8 [B][Hs][Ws] (decoded label map), both on the device;
 * x[b] = resize(img / 255), labels' = resize(uint8(lab + 1)) (the reference adds 1 in uint8, 255 wraps to 0),
 * y[b][l] = (labels' == labels[l]) as 0 / 1 float; resize = torchvision Resize((Ho, Wo)) on a float tensor (bilinear,
 * align_corners = False, no antialias; identity when the sizes match); flips[b] (device, nullable): bit 0 horizontal, bit 1
 * vertical flip after the resize (RandomHorizontalFlip / RandomVerticalFlip; the caller draws them).  `labels` is a HOST
 * array of nlabels <= 16 values.  x: float [B][3][Ho][Wo], y: float [B][nlabels][Ho][Wo]. */
int pg_prep_batch_u8(const uint8_t* img, const uint8_t* lab, const int32_t* labels, int32_t nlabels, int32_t B, int32_t Hs,
                     int32_t Ws, int32_t Ho, int32_t Wo, const uint8_t* flips, float* x, float* y, void* stream);

/* ---- inference tiling (infer.py:14-68), "next" row ---- */
int pg_ncrop(const float* image, float* crops, int32_t C, int32_t H, int32_t W, int32_t size, int32_t eff,
             int32_t ncy, int32_t ncx, void* stream);
int pg_build_mask(const float* masks, float* mask_out, int32_t* argmax_out, int32_t C, int32_t H, int32_t W,
                  int32_t size, int32_t eff, int32_t ncy, int32_t ncx, float threshold, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PATCHGAN_B200_H */
