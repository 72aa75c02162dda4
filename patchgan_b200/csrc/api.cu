// C-ABI glue: error text, argument validation, implementation dispatch.
#include <stdarg.h>

#include "common.cuh"

namespace pg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
static unsigned long long g_simt_fallbacks = 0;
static thread_local int g_last_impl = 0;

// PG_IMPL_AUTO found no tensor-core plan for a shape and ran the CUDA-core kernel instead: counted (pg_fallback_count,
// asserted to stay 0 by bench.py and the step tests) and reported once per process -- a 50x cliff must not be silent.
static void note_simt_fallback(const char* what, const PgConvDesc* d) {
  if (g_simt_fallbacks++ == 0)
    fprintf(stderr, "patchgan_b200: %s mode %d stride %d B %d %dx%d -> %dx%d C %d+%d N %d runs on the CUDA-core kernel (no tcgen05 plan "
            "for this shape / alignment); further fallbacks are only counted (pg_fallback_count)\n", what, d->mode, d->stride,
            d->B, d->Hin, d->Win, d->Hout, d->Wout, d->C1, d->C2, d->N);
}

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return PG_ERR_CUDA;
  }
  return PG_OK;
}

int conv_fwd_simt(const PgConvDesc*, const void*, const void*, const void*, const float*, void*, void*, cudaStream_t);
int conv_wgrad_simt(const PgConvDesc*, const void*, const void*, int, float*, int, int, int, int, int, cudaStream_t);
// conv_tc.cu
int conv_fwd_tc(const PgConvDesc*, const void*, const void*, const void*, const float*, void*, void*, float*, const void*, int, int,
                cudaStream_t);
bool conv_fwd_tc_stats_ok(const PgConvDesc*);
bool conv_fwd_tc_supported(const PgConvDesc*, const void*, const void*, const void*, const void*);
int conv_wgrad_tc(const PgConvDesc*, const void*, const void*, int, float*, int, int, int, int, int, cudaStream_t);
bool conv_wgrad_tc_supported(const PgConvDesc*, const void*, const void*, int);
int conv_wgrad_group_tc(const PgWgradJob*, int, cudaStream_t);
bool tc_device_ok();
bool conv_res_supported(const PgConvDesc*, const PgFusedNorm*, bool);
int conv_res_launch(const PgConvDesc*, const void*, const void*, const void*, void*, void*, const PgFusedNorm*, cudaStream_t);
void set_tc_trace(void*);
unsigned long long pair_launch_count();
void set_pair_mode(int);
void set_sm_limit(int);
void set_conv_workspace(void*, size_t);
// conv_skinny.cu
bool conv_fewout_supported(const PgConvDesc*);
int conv_fewout(const PgConvDesc*, const void*, const void*, const void*, const float*, void*, cudaStream_t);
bool conv_fewin_supported(const PgConvDesc*);
int conv_fewin(const PgConvDesc*, const void*, const void*, void*, void*, cudaStream_t);
int taps_scatter(int, int, int, int, int, int, int, int, const float*, int, const float*, int, void*, int, int, int, cudaStream_t);
int taps_gather(int, int, int, int, int, int, int, int, const void*, int, int, void*, cudaStream_t);
int taps_dgrad_act(const void*, const void*, void*, int, const void*, int, int, int, long long, int, cudaStream_t);
bool conv_wgrad1_supported(const PgConvDesc*, int, const float*, int, int, int);
int conv_wgrad1(const PgConvDesc*, const void*, const void*, int, float*, int, int, int, cudaStream_t);

static int validate(const PgConvDesc* d, const char* who) {
  PG_REQUIRE(d != nullptr, "%s: desc is NULL", who);
  PG_REQUIRE(d->mode == PG_CONV || d->mode == PG_CONVT || d->mode == PG_CONV1X1, "%s: bad mode %d", who, d->mode);
  PG_REQUIRE(d->B > 0 && d->Hin > 0 && d->Win > 0 && d->Hout > 0 && d->Wout > 0, "%s: empty extent", who);
  PG_REQUIRE(d->C1 > 0 && d->C1 % 16 == 0 && d->C2 >= 0 && d->C2 % 16 == 0, "%s: C1=%d C2=%d must be multiples of 16",
             who, d->C1, d->C2);
  PG_REQUIRE(d->N > 0 && d->N % 16 == 0, "%s: N=%d must be a multiple of 16", who, d->N);
  PG_REQUIRE(d->ld1 >= d->C1 && d->ld1 % 8 == 0 && (d->C2 == 0 || (d->ld2 >= d->C2 && d->ld2 % 8 == 0)),
             "%s: bad pixel strides", who);
  // ldo may be smaller than N when only the first n_valid (<= ldo) channels are wanted: the rest is not stored
  PG_REQUIRE((d->ldo >= d->N || d->ldo >= d->n_valid) && d->ldo % (d->out_f32 == PG_F32 ? 4 : 8) == 0,
             "%s: bad output stride %d", who, d->ldo);
  PG_REQUIRE(d->in_dtype == PG_BF16 || d->in_dtype == PG_F16, "%s: in_dtype must be PG_BF16 or PG_F16", who);
  PG_REQUIRE(d->out_f32 >= PG_BF16 && d->out_f32 <= PG_F16, "%s: bad output dtype %d", who, d->out_f32);
  PG_REQUIRE(d->n_first >= 0 && d->n_first < d->n_valid && d->c_valid >= 0 && d->c_valid <= d->C1,
             "%s: bad n_first=%d / c_valid=%d", who, d->n_first, d->c_valid);
  if (d->mode == PG_CONV1X1) {
    PG_REQUIRE(d->Hout == d->Hin && d->Wout == d->Win && d->ldw >= 0, "%s: pointwise needs Hout = Hin, Wout = Win", who);
  } else if (d->mode == PG_CONV) {
    PG_REQUIRE((d->stride == 1 || d->stride == 2) && (d->pad == 1 || d->pad == 2), "%s: stride/pad unsupported", who);
    PG_REQUIRE(d->Hout == (d->Hin + 2 * d->pad - 4) / d->stride + 1 && d->Wout == (d->Win + 2 * d->pad - 4) / d->stride + 1,
               "%s: Hout/Wout inconsistent with Hin/Win", who);
  } else {
    PG_REQUIRE(d->stride == 2 && d->pad == 1 && d->Hout == 2 * d->Hin && d->Wout == 2 * d->Win,
               "%s: convT needs stride 2 pad 1 and Hout = 2 Hin", who);
  }
  return PG_OK;
}

}  // namespace pg
using namespace pg;

extern "C" const char* pg_last_error(void) { return g_err; }
extern "C" int pg_version(void) { return 100; }
extern "C" int64_t pg_launch_count(void) { return (int64_t)g_launches; }
extern "C" int pg_last_conv_impl(void) { return g_last_impl; }
extern "C" int64_t pg_fallback_count(void) { return (int64_t)g_simt_fallbacks; }
extern "C" int64_t pg_pair_launch_count(void) { return (int64_t)pair_launch_count(); }
extern "C" int pg_set_pair_mode(int32_t mode) { set_pair_mode(mode); return PG_OK; }
extern "C" int pg_tcgen05_available(void) { return tc_device_ok() ? 1 : 0; }
extern "C" int pg_debug_set_trace(void* buf) { set_tc_trace(buf); return PG_OK; }
extern "C" int pg_set_sm_limit(int32_t n) {
  PG_REQUIRE(n >= 0, "pg_set_sm_limit: negative limit");
  set_sm_limit(n);
  return PG_OK;
}

// In-stream time stamp (tools/timeline.py): *slot = %globaltimer when the stream reaches this point.  Capturable.
__global__ void debug_stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
extern "C" int pg_debug_stamp(uint64_t* slot, void* stream) {
  PG_REQUIRE(slot != nullptr, "pg_debug_stamp: slot is NULL");
  debug_stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)slot);
  return check_launch("debug_stamp_kernel");
}
extern "C" int pg_conv_set_workspace(void* ws, int64_t bytes) {
  PG_REQUIRE((ws == nullptr) == (bytes == 0) && bytes >= 0 && (((uintptr_t)ws) & 255) == 0, "pg_conv_set_workspace: bad buffer");
  set_conv_workspace(ws, (size_t)bytes);
  return PG_OK;
}

extern "C" int pg_instnorm_stats(const void* x, int32_t x_f32, int32_t B, int64_t HW, int32_t C, int32_t ld, float* sums,
                                 void* stream);

static int conv_fwd_any(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed, const float* bias,
                        void* out, void* out2, float* stats, const void* mul_y, int mul_ld, int mul_dt, int impl, void* stream);
extern "C" int pg_act_bwd_from_output(const void* y, int32_t y_f32, int32_t ldy, const void* dy, int32_t lddy, void* dx,
                                      int32_t lddx, int64_t npix, int32_t C, int32_t act, void* stream);

extern "C" int pg_conv_fwd(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed,
                           const float* bias, void* out, void* out2, int impl, void* stream) {
  return conv_fwd_any(d, src1, src2, w_packed, bias, out, out2, nullptr, nullptr, 0, 0, impl, stream);
}

extern "C" int pg_conv_fwd_stats(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed,
                                 const float* bias, void* out, float* sums, int impl, void* stream) {
  PG_REQUIRE(d != nullptr && sums != nullptr, "pg_conv_fwd_stats: NULL argument");
  PG_REQUIRE(d->ldo >= d->N, "pg_conv_fwd_stats: the output row must hold all N channels");
  return conv_fwd_any(d, src1, src2, w_packed, bias, out, nullptr, sums, nullptr, 0, 0, impl, stream);
}

extern "C" int pg_conv_dgrad_act(const PgConvDesc* d, const void* dy, const void* w_packed, void* dx, const void* y, int32_t ldy,
                                 int32_t y_dtype, int impl, void* stream) {
  PG_REQUIRE(d != nullptr && y != nullptr && dx != nullptr, "pg_conv_dgrad_act: NULL argument");
  PG_REQUIRE(d->out_f32 == PG_BF16 && !d->has_bias && d->C2 == 0 && d->ldo >= d->N && d->n_valid == d->N,
             "pg_conv_dgrad_act: needs a bf16 output holding all N channels, no bias, one source");
  PG_REQUIRE((y_dtype == PG_BF16 || y_dtype == PG_F16) && ldy >= d->N && ldy % 8 == 0 && (((uintptr_t)y) & 15) == 0,
             "pg_conv_dgrad_act: bad y (16-bit, ldy >= N, 16-byte aligned)");
  return conv_fwd_any(d, dy, nullptr, w_packed, nullptr, dx, nullptr, nullptr, y, ldy, y_dtype, impl, stream);
}

static int conv_fwd_any(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed, const float* bias,
                        void* out, void* out2, float* stats, const void* mul_y, int mul_ld, int mul_dt, int impl, void* stream) {
  if (int e = validate(d, "pg_conv_fwd")) return e;
  if (mul_y != nullptr && (impl == PG_IMPL_SIMT || impl == PG_IMPL_SKINNY ||
                           !conv_fwd_tc_supported(d, src1, src2, w_packed, out))) {
    // no fused epilogue on this path: the plain data-gradient, then the activation backward in place
    PgConvDesc plain = *d;
    plain.act = PG_ACT_NONE;
    if (int e = conv_fwd_any(&plain, src1, src2, w_packed, bias, out, out2, stats, nullptr, 0, 0, impl, stream)) return e;
    return pg_act_bwd_from_output(mul_y, mul_dt, mul_ld, out, d->ldo, out, d->ldo, (int64_t)d->B * d->Hout * d->Wout, d->N,
                                  d->act, stream);
  }
  PG_REQUIRE(src1 && w_packed && out && (d->C2 == 0 || src2), "pg_conv_fwd: NULL pointer");
  PG_REQUIRE(!d->has_bias || bias, "pg_conv_fwd: has_bias but bias is NULL");
  PG_REQUIRE(d->mode != PG_CONV1X1 || d->ldw % 8 == 0, "pg_conv_fwd: pointwise weight rows need ldw %% 8 == 0");
  cudaStream_t s = (cudaStream_t)stream;
  g_last_impl = PG_IMPL_SIMT;
  PG_REQUIRE(out2 == nullptr || d->out_f32 != PG_F32, "pg_conv_fwd: out2 (bf16 twin) needs a 16-bit primary output");
  if (impl == PG_IMPL_SIMT) return conv_fwd_simt(d, src1, src2, w_packed, bias, out, out2, s);
  if (impl == PG_IMPL_SKINNY) {
    // CUDA-core kernels for one real channel on one side (kept as a second implementation for validation; the
    // engine runs these layers as PG_CONV1X1 tap products on the tensor cores, which is faster)
    if (out2 == nullptr && stats == nullptr && conv_fewout_supported(d)) {
      g_last_impl = PG_IMPL_SKINNY;
      return conv_fewout(d, src1, src2, w_packed, bias, out, s);
    }
    if (stats == nullptr && conv_fewin_supported(d)) {
      g_last_impl = PG_IMPL_SKINNY;
      return conv_fewin(d, src1, w_packed, out, out2, s);
    }
    if (impl == PG_IMPL_SKINNY) {
      set_error("pg_conv_fwd: shape does not qualify for the skinny kernels");
      return PG_ERR_UNSUPPORTED;
    }
  }
  const bool ok = conv_fwd_tc_supported(d, src1, src2, w_packed, out);
  if (ok) g_last_impl = PG_IMPL_TCGEN05;
  if (impl == PG_IMPL_TCGEN05 && !ok) {
    set_error("pg_conv_fwd: tcgen05 path does not support this shape / alignment");
    return PG_ERR_UNSUPPORTED;
  }
  // InstanceNorm statistics: fused into the tcgen05 epilogue when the tile geometry allows, else a second launch
  if (!ok) note_simt_fallback("pg_conv_fwd", d);
  const bool fuse = stats != nullptr && ok && conv_fwd_tc_stats_ok(d);
  int e = ok ? conv_fwd_tc(d, src1, src2, w_packed, bias, out, out2, fuse ? stats : nullptr, mul_y, mul_ld, mul_dt, s)
             : conv_fwd_simt(d, src1, src2, w_packed, bias, out, out2, s);
  if (e == PG_OK && stats != nullptr && !fuse)
    e = pg_instnorm_stats(out, d->out_f32, d->B, (int64_t)d->Hout * d->Wout, d->N, d->ldo, stats, stream);
  return e;
}

static int validate_fused(const PgConvDesc* d, const PgFusedNorm* fn, const char* who) {
  if (int e = validate(d, who)) return e;
  PG_REQUIRE(fn != nullptr, "%s: PgFusedNorm is NULL", who);
  PG_REQUIRE(fn->kind == PG_FUSED_FWD || fn->kind == PG_FUSED_BWD, "%s: bad kind %d", who, fn->kind);
  PG_REQUIRE(fn->act == PG_ACT_NONE || fn->act == PG_ACT_RELU || fn->act == PG_ACT_LEAKYRELU || fn->act == PG_ACT_TANH,
             "%s: activation %d cannot follow an InstanceNorm block", who, fn->act);
  PG_REQUIRE(fn->drop_p >= 0.f && fn->drop_p < 1.f && (fn->drop_p == 0.f || fn->seed != nullptr), "%s: bad dropout arguments", who);
  PG_REQUIRE(!d->has_bias && d->ldo >= d->N && (d->out_f32 == PG_BF16 || d->out_f32 == PG_F16),
             "%s: needs a 16-bit output holding all N channels and no bias", who);
  if (fn->kind == PG_FUSED_BWD) {
    PG_REQUIRE(fn->n_norm > 0 && fn->n_norm <= d->N && fn->n_norm % 16 == 0, "%s: bad n_norm %d", who, fn->n_norm);
    PG_REQUIRE(d->out_f32 == PG_BF16 && d->C2 == 0, "%s: gradients are bf16, one source", who);
    PG_REQUIRE((fn->xhat != nullptr) != (fn->y != nullptr), "%s: exactly one of xhat / y", who);
    PG_REQUIRE(fn->y == nullptr || ((fn->act == PG_ACT_NONE || fn->act == PG_ACT_LEAKYRELU) && fn->drop_p == 0.f),
               "%s: recovering xhat from y needs an invertible activation and no dropout (pass xhat)", who);
    PG_REQUIRE(fn->y_dtype == PG_BF16 || fn->y_dtype == PG_F16, "%s: bad y_dtype", who);
    const void* t = fn->xhat != nullptr ? fn->xhat : fn->y;
    const int ld = fn->xhat != nullptr ? fn->xhat_ld : fn->y_ld;
    PG_REQUIRE((((uintptr_t)t) & 15) == 0 && ld >= fn->n_norm && ld % 8 == 0, "%s: xhat / y must be 16-byte aligned, ld >= n_norm", who);
    PG_REQUIRE(fn->dskip == nullptr || ((((uintptr_t)fn->dskip) & 15) == 0 && fn->dskip_ld >= fn->n_norm && fn->dskip_ld % 8 == 0),
               "%s: bad dskip", who);
  } else {
    PG_REQUIRE(fn->xhat == nullptr || ((((uintptr_t)fn->xhat) & 15) == 0 && fn->xhat_ld >= d->N && fn->xhat_ld % 8 == 0),
               "%s: bad xhat output", who);
  }
  return PG_OK;
}

extern "C" int pg_conv_norm_supported(const PgConvDesc* d, const PgFusedNorm* fn, int32_t has_twin) {
  if (validate_fused(d, fn, "pg_conv_norm_supported") != PG_OK) return 0;
  return conv_res_supported(d, fn, has_twin != 0) ? 1 : 0;
}

extern "C" int pg_conv_norm_fwd(const PgConvDesc* d, const void* src1, const void* src2, const void* w_packed, void* out,
                                void* out2, const PgFusedNorm* fn, void* stream) {
  if (int e = validate_fused(d, fn, "pg_conv_norm_fwd")) return e;
  PG_REQUIRE(fn->kind == PG_FUSED_FWD, "pg_conv_norm_fwd: kind must be PG_FUSED_FWD");
  PG_REQUIRE(src1 && w_packed && out && (d->C2 == 0 || src2) && fn->sums && fn->sync, "pg_conv_norm_fwd: NULL pointer");
  PG_REQUIRE(((((uintptr_t)src1) | ((uintptr_t)src2) | ((uintptr_t)w_packed) | ((uintptr_t)out) | ((uintptr_t)out2)) & 15) == 0,
             "pg_conv_norm_fwd: pointers must be 16-byte aligned");
  g_last_impl = PG_IMPL_TCGEN05;
  return conv_res_launch(d, src1, src2, w_packed, out, out2, fn, (cudaStream_t)stream);
}

extern "C" int pg_conv_dgrad_norm_bwd(const PgConvDesc* d, const void* dy, const void* w_packed, void* dx, const PgFusedNorm* fn,
                                      void* stream) {
  if (int e = validate_fused(d, fn, "pg_conv_dgrad_norm_bwd")) return e;
  PG_REQUIRE(fn->kind == PG_FUSED_BWD, "pg_conv_dgrad_norm_bwd: kind must be PG_FUSED_BWD");
  PG_REQUIRE(dy && w_packed && dx && fn->sums && fn->bsums && fn->sync, "pg_conv_dgrad_norm_bwd: NULL pointer");
  PG_REQUIRE(((((uintptr_t)dy) | ((uintptr_t)w_packed) | ((uintptr_t)dx)) & 15) == 0,
             "pg_conv_dgrad_norm_bwd: pointers must be 16-byte aligned");
  g_last_impl = PG_IMPL_TCGEN05;
  return conv_res_launch(d, dy, nullptr, w_packed, dx, nullptr, fn, (cudaStream_t)stream);
}

extern "C" int pg_taps_scatter(int32_t mode, int32_t stride, int32_t pad, int32_t B, int32_t Hq, int32_t Wq, int32_t Hp, int32_t Wp,
                               const float* P, int32_t ldp, const float* bias, int32_t act, void* out, int32_t out_dtype,
                               int32_t ldo, int32_t ch, void* stream) {
  PG_REQUIRE(P && out && B > 0 && Hq > 0 && Wq > 0 && Hp > 0 && Wp > 0 && ldp >= 16, "pg_taps_scatter: bad arguments");
  PG_REQUIRE(mode == PG_CONV || mode == PG_CONVT, "pg_taps_scatter: bad mode %d", mode);
  return taps_scatter(mode, stride, pad, B, Hq, Wq, Hp, Wp, P, ldp, bias, act, out, out_dtype, ldo, ch, (cudaStream_t)stream);
}

extern "C" int pg_taps_gather(int32_t mode, int32_t stride, int32_t pad, int32_t B, int32_t Hq, int32_t Wq, int32_t Hp, int32_t Wp,
                              const void* src, int32_t lds, int32_t ch, void* G, void* stream) {
  PG_REQUIRE(src && G && B > 0 && Hq > 0 && Wq > 0 && Hp > 0 && Wp > 0, "pg_taps_gather: bad arguments");
  PG_REQUIRE(mode == PG_CONV || mode == PG_CONVT, "pg_taps_gather: bad mode %d", mode);
  PG_REQUIRE((((uintptr_t)G) & 15) == 0, "pg_taps_gather: G must be 16-byte aligned");
  return taps_gather(mode, stride, pad, B, Hq, Wq, Hp, Wp, src, lds, ch, G, (cudaStream_t)stream);
}

extern "C" int pg_taps_dgrad_act(const void* G, const void* w16, void* dx, int32_t lddx, const void* y, int32_t ldy, int32_t y_dtype,
                                 int32_t act, int64_t nq, int32_t C, void* stream) {
  PG_REQUIRE(G && w16 && dx && nq > 0, "pg_taps_dgrad_act: NULL pointer / empty");
  PG_REQUIRE(C >= 2 && C % 2 == 0 && C <= 1024 && lddx >= C && lddx % 2 == 0, "pg_taps_dgrad_act: C=%d must be even, <= 1024, lddx >= C", C);
  PG_REQUIRE((((uintptr_t)G) & 15) == 0 && (((uintptr_t)dx) & 3) == 0, "pg_taps_dgrad_act: G must be 16-byte, dx 4-byte aligned");
  PG_REQUIRE(y == nullptr || ((y_dtype == PG_BF16 || y_dtype == PG_F16) && ldy >= C && ldy % 2 == 0 && (((uintptr_t)y) & 3) == 0),
             "pg_taps_dgrad_act: bad y");
  return taps_dgrad_act(G, w16, dx, lddx, y, ldy, y_dtype, act, nq, C, (cudaStream_t)stream);
}

static int conv_wgrad_any(const PgConvDesc* d, const void* a, const void* g, int32_t ldg, float* dw, int32_t ld_n,
                          int32_t n_real, int32_t c_real, int tap_major, int Cs, int impl, void* stream) {
  if (int e = validate(d, "pg_conv_wgrad")) return e;
  PG_REQUIRE((d->mode == PG_CONV || d->mode == PG_CONV1X1) && d->C2 == 0,
             "pg_conv_wgrad: geometry must be PG_CONV / PG_CONV1X1 with one source");
  PG_REQUIRE(a && g && dw && ldg >= d->N && ldg % 8 == 0, "pg_conv_wgrad: bad pointers / ldg");
  PG_REQUIRE(n_real <= d->N && c_real <= d->C1, "pg_conv_wgrad: n_real / c_real exceed padded extents");
  cudaStream_t s = (cudaStream_t)stream;
  g_last_impl = PG_IMPL_SIMT;
  if (impl == PG_IMPL_SIMT) return conv_wgrad_simt(d, a, g, ldg, dw, ld_n, n_real, c_real, tap_major, Cs, s);
  if (impl == PG_IMPL_SKINNY && !tap_major) {
    if (conv_wgrad1_supported(d, ldg, dw, ld_n, n_real, c_real)) {
      g_last_impl = PG_IMPL_SKINNY;
      return conv_wgrad1(d, a, g, ldg, dw, ld_n, n_real, c_real, s);
    }
    set_error("pg_conv_wgrad: shape does not qualify for the skinny kernels");
    return PG_ERR_UNSUPPORTED;
  }
  const bool ok = conv_wgrad_tc_supported(d, a, g, ldg) && !(tap_major && (Cs % 4) != 0);
  if (ok) g_last_impl = PG_IMPL_TCGEN05;
  if (impl == PG_IMPL_TCGEN05 && !ok) {
    set_error("pg_conv_wgrad: tcgen05 path does not support this shape / alignment");
    return PG_ERR_UNSUPPORTED;
  }
  if (ok) return conv_wgrad_tc(d, a, g, ldg, dw, ld_n, n_real, c_real, tap_major, Cs, s);
  note_simt_fallback("pg_conv_wgrad", d);
  return conv_wgrad_simt(d, a, g, ldg, dw, ld_n, n_real, c_real, tap_major, Cs, s);
}

extern "C" int pg_conv_wgrad(const PgConvDesc* d, const void* a, const void* g, int32_t ldg, float* dw, int32_t ld_n,
                             int32_t n_real, int32_t c_real, int impl, void* stream) {
  return conv_wgrad_any(d, a, g, ldg, dw, ld_n, n_real, c_real, 0, 0, impl, stream);
}

extern "C" int pg_conv_wgrad_group(const PgWgradJob* jobs, int32_t njobs, void* stream) {
  PG_REQUIRE(jobs != nullptr && njobs > 0, "pg_conv_wgrad_group: no jobs");
  for (int j = 0; j < njobs; ++j) {
    const PgWgradJob& jb = jobs[j];
    if (int e = validate(&jb.desc, "pg_conv_wgrad_group")) return e;
    PG_REQUIRE((jb.desc.mode == PG_CONV || jb.desc.mode == PG_CONV1X1) && jb.desc.C2 == 0,
               "pg_conv_wgrad_group: job %d: geometry must be PG_CONV / PG_CONV1X1 with one source", j);
    PG_REQUIRE(jb.a && jb.g && jb.dw && jb.ldg % 8 == 0 && jb.ldg >= (jb.g2 ? jb.n_split : jb.desc.N),
               "pg_conv_wgrad_group: job %d: bad pointers / ldg", j);
    PG_REQUIRE(jb.g2 == nullptr || (jb.n_split > 0 && jb.n_split < jb.desc.N && jb.ldg2 >= jb.desc.N - jb.n_split),
               "pg_conv_wgrad_group: job %d: bad second source", j);
    PG_REQUIRE(jb.n_real <= jb.desc.N && jb.c_real <= jb.desc.C1, "pg_conv_wgrad_group: job %d: n_real / c_real exceed padded extents", j);
    PG_REQUIRE(!jb.tap_major || (jb.desc.mode == PG_CONV && jb.Cs % 4 == 0), "pg_conv_wgrad_group: job %d: bad tap-major job", j);
    PG_REQUIRE(conv_wgrad_tc_supported(&jb.desc, jb.a, jb.g, jb.ldg), "pg_conv_wgrad_group: job %d has no tcgen05 plan", j);
  }
  g_last_impl = PG_IMPL_TCGEN05;
  return conv_wgrad_group_tc(jobs, njobs, (cudaStream_t)stream);
}

extern "C" int pg_conv_wgrad_tapmajor(const PgConvDesc* d, const void* a, const void* g, int32_t ldg, float* S, int32_t Ns,
                                      int32_t Cs, int impl, void* stream) {
  PG_REQUIRE(d != nullptr && d->mode == PG_CONV, "pg_conv_wgrad_tapmajor: geometry must be PG_CONV");
  PG_REQUIRE(Ns > 0 && Ns <= d->N && Cs > 0 && Cs <= d->C1, "pg_conv_wgrad_tapmajor: Ns / Cs exceed the padded extents");
  return conv_wgrad_any(d, a, g, ldg, S, Ns, Ns, Cs, 1, Cs, impl, stream);
}
