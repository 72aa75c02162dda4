"""ctypes binding of libpatchgan_b200.so (the C-ABI declared in include/patchgan_b200.h).

There is NO fallback: if the shared library is missing or cannot be loaded, importing the kernels
raises, and every op raises when handed a non-CUDA tensor.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libpatchgan_b200.so')

PG_CONV, PG_CONVT, PG_CONV1X1 = 0, 1, 2
ACT = {'none': 0, None: 0, 'relu': 1, 'leakyrelu': 2, 'tanh': 3, 'sigmoid': 4, 'softmax': 5}
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, IMPL_SKINNY = 0, 1, 2, 3
DT_BF16, DT_F32, DT_F16 = 0, 1, 2
LOSS = {'tversky': 0, 'weighted_bce': 1, 'MAE': 2, 'none': 3}


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'mode', 'stride', 'pad', 'B', 'Hin', 'Win', 'Hout', 'Wout', 'C1', 'C2', 'ld1', 'ld2', 'N', 'ldo',
        'n_valid', 'act', 'out_f32', 'has_bias', 'in_dtype', 'n_first', 'c_valid', 'ldw')]


vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
DP = C.POINTER(ConvDesc)

FUSED_FWD, FUSED_BWD = 0, 1


class FusedNorm(C.Structure):
    """PgFusedNorm (include/patchgan_b200.h): arguments of the conv + InstanceNorm one-launch kernels."""
    _fields_ = [('kind', i32), ('act', i32), ('n_norm', i32), ('drop_p', f32), ('seed', vp), ('salt', u64),
                ('sums', vp), ('bsums', vp), ('sync', vp), ('xhat', vp), ('xhat_ld', i32), ('y', vp), ('y_ld', i32),
                ('y_dtype', i32), ('dskip', vp), ('dskip_ld', i32), ('ws', vp), ('ws_bytes', i64)]


FP = C.POINTER(FusedNorm)


class WgradJob(C.Structure):
    """PgWgradJob: one weight gradient of a grouped launch (pg_conv_wgrad_group)."""
    _fields_ = [('desc', ConvDesc), ('a', vp), ('g', vp), ('ldg', i32), ('tap_major', i32), ('dw', vp), ('ld_n', i32),
                ('n_real', i32), ('c_real', i32), ('Cs', i32), ('g2', vp), ('ldg2', i32), ('n_split', i32)]

_SIGS = {
    'pg_version': ([], C.c_int),
    'pg_tcgen05_available': ([], C.c_int),
    'pg_debug_set_trace': ([vp], C.c_int),
    'pg_set_sm_limit': ([i32], C.c_int),
    'pg_debug_stamp': ([vp, vp], C.c_int),
    'pg_conv_set_workspace': ([vp, i64], C.c_int),
    'pg_launch_count': ([], C.c_int64),
    'pg_last_conv_impl': ([], C.c_int),
    'pg_fallback_count': ([], C.c_int64),
    'pg_pair_launch_count': ([], C.c_int64),
    'pg_set_pair_mode': ([i32], C.c_int),
    'pg_conv_fwd': ([DP, vp, vp, vp, vp, vp, vp, C.c_int, vp], C.c_int),
    'pg_conv_fwd_stats': ([DP, vp, vp, vp, vp, vp, vp, C.c_int, vp], C.c_int),
    'pg_conv_dgrad_act': ([DP, vp, vp, vp, vp, i32, i32, C.c_int, vp], C.c_int),
    'pg_conv_norm_supported': ([DP, FP, i32], C.c_int),
    'pg_conv_norm_fwd': ([DP, vp, vp, vp, vp, vp, FP, vp], C.c_int),
    'pg_conv_dgrad_norm_bwd': ([DP, vp, vp, vp, FP, vp], C.c_int),
    'pg_conv_wgrad': ([DP, vp, vp, i32, vp, i32, i32, i32, C.c_int, vp], C.c_int),
    'pg_conv_wgrad_group': ([C.POINTER(WgradJob), i32, vp], C.c_int),
    'pg_conv_wgrad_tapmajor': ([DP, vp, vp, i32, vp, i32, i32, C.c_int, vp], C.c_int),
    'pg_grad_finalize_multi': ([vp, i32, i32, vp], C.c_int),
    'pg_colsum': ([vp, i64, i32, i32, vp, vp], C.c_int),
    'pg_taps_scatter': ([i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, i32, vp], C.c_int),
    'pg_taps_dgrad_act': ([vp, vp, vp, i32, vp, i32, i32, i32, i64, i32, vp], C.c_int),
    'pg_taps_gather': ([i32, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, vp, vp], C.c_int),
    'pg_pack_nchw_f32_to_nhwc_bf16': ([vp, vp, i32, i32, i32, i32, i32, i32, i32, vp], C.c_int),
    'pg_unpack_nhwc_to_nchw_f32': ([vp, i32, vp, i32, i32, i32, i32, i32, i32, vp], C.c_int),
    'pg_copy_f32_to_bf16_slice': ([vp, i32, vp, i32, i32, i32, i64, i32, vp], C.c_int),
    'pg_pack_weight': ([vp, vp, i32, i32, i32, i32, i32, i32, i64, i64, i32, i32, vp], C.c_int),
    'pg_pack2_nchw_rows': ([vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp], C.c_int),
    'pg_pack_weights_multi': ([vp, i32, i32, vp], C.c_int),
    'pg_im2col_s2_pair': ([vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp], C.c_int),
    'pg_im2col_s2': ([vp, i64, i64, i64, i64, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp], C.c_int),
    'pg_instnorm_stats': ([vp, i32, i32, i64, i32, i32, vp, vp], C.c_int),
    'pg_norm_act_fwd': ([vp, i32, vp, vp, i32, vp, i32, i64, i32, i32, i32, i32, f32, vp, u64, vp], C.c_int),
    'pg_norm_act_bwd_reduce': ([vp, i32, vp, vp, i32, vp, i32, vp, i32, i64, i32, i32, i32, f32, vp, u64, vp], C.c_int),
    'pg_norm_act_bwd_apply': ([vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, i32, i64, i32, i32, i32, f32, vp, u64, vp],
                              C.c_int),
    'pg_norm_act_bwd': ([vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, i32, i64, i32, i32, i32, f32, vp, u64, vp], C.c_int),
    'pg_bn_fold_fwd': ([vp, i32, i32, i64, vp, vp, i32, f32, i32, vp], C.c_int),
    'pg_bn_fold_bwd': ([vp, i32, i32, vp, vp, i32, i32, vp], C.c_int),
    'pg_norm_affine_act_fwd': ([vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, i64, i32, i32, i32, i32, f32, vp, u64, vp], C.c_int),
    'pg_norm_affine_act_bwd_reduce': ([vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, vp, i32, i64, i32, i32, i32, f32, vp, u64, vp],
                                      C.c_int),
    'pg_norm_affine_act_bwd_apply': ([vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, vp, vp, i32, i32, i64, i32, i32, i32, f32, vp,
                                      u64, vp], C.c_int),
    'pg_act_bwd_from_output': ([vp, i32, i32, vp, i32, vp, i32, i64, i32, i32, vp], C.c_int),
    'pg_softmax_fwd': ([vp, vp, i64, i32, i32, vp], C.c_int),
    'pg_target_chsum': ([vp, vp, i32, i32, i64, vp], C.c_int),
    'pg_seg_loss_partials': ([vp, i32, vp, vp, vp, i32, i32, i64, i32, vp], C.c_int),
    'pg_seg_loss_finalize': ([vp, vp, vp, i32, i32, i32, i64, i32, f32, f32, f32, vp], C.c_int),
    'pg_gen_out_bwd': ([vp, i32, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i64, i32, i32, f32, vp], C.c_int),
    'pg_bce_const': ([vp, i32, f32, f32, vp, i32, vp, i32, i64, vp], C.c_int),
    'pg_bce_mean': ([vp, vp, i64, vp, i32, vp], C.c_int),
    'pg_bce_mean_bwd': ([vp, vp, i64, vp, vp, vp], C.c_int),
    'pg_sample_sums_bwd': ([vp, vp, vp, vp, i32, i64, vp], C.c_int),
    'pg_adam_step': ([vp, vp, vp, vp, i64, vp, vp, f32, f32, f32, f32, vp], C.c_int),
    'pg_adam_step_range': ([vp, vp, vp, vp, i64, vp, vp, f32, f32, f32, f32, i32, vp], C.c_int),
    'pg_counter_add': ([vp, u64, vp], C.c_int),
    'pg_prep_batch_u8': ([vp, vp, C.POINTER(i32), i32, i32, i32, i32, i32, i32, vp, vp, vp, vp], C.c_int),
    'pg_ncrop': ([vp, vp, i32, i32, i32, i32, i32, i32, i32, vp], C.c_int),
    'pg_build_mask': ([vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp], C.c_int),
}

_lib = None


class KernelLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises KernelLibraryError if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelLibraryError(
                f'{LIB_PATH} is missing: build it with `python -m patchgan_b200.build` '
                '(nvcc, sm_100a).  patchgan_b200 has no CPU or library fallback.')
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise KernelLibraryError(f'cannot load {LIB_PATH}: {e}') from e
        for name, (args, res) in _SIGS.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = res
        handle.pg_last_error.restype = C.c_char_p
        handle.pg_last_error.argtypes = []
        _lib = handle
    return _lib


def exported_symbols():
    return ['pg_last_error'] + list(_SIGS)


def check(status, what=''):
    if status != 0:
        msg = lib().pg_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'patchgan_b200 kernel call failed ({what}, status {status}): {msg}')


PROFILER = None      # set to an object with begin(name)/end(token) to time every C-ABI call (bench.py)


STAMPER = None       # set to a callable (name, args) invoked after every C-ABI call (tools/timeline.py)


def call(name, *args):
    if PROFILER is None:
        check(getattr(lib(), name)(*args), name)
        if STAMPER is not None:
            STAMPER(name, args)
    else:
        tok = PROFILER.begin(name)
        check(getattr(lib(), name)(*args), name)
        PROFILER.end(tok)
