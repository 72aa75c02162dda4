"""Host-side schedule of the sm_100a kernels for the U-Net generator and the PatchGAN discriminator.

This is the layer between the reference-shaped ``nn.Module`` API (unet.py / disc.py / trainer.py in this
package) and the C-ABI library: it owns the packed bf16 weights, allocates NHWC activations with torch,
and issues forward / data-gradient / weight-gradient launches on the current CUDA stream.  torch is used
for memory and streams only; no torch operator does arithmetic on this path.

Layer semantics follow the reference: ``DownSampleBlock`` (/root/reference/patchgan/unet.py:8-35),
``UpSampleBlock`` (unet.py:38-72), ``UNet.forward`` (unet.py:112-134), ``Discriminator`` (disc.py:8-51).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib as L

DROP_P = 0.2  # nn.Dropout(0.2), unet.py:28,65


def rup16(c):
    return (c + 15) // 16 * 16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f'patchgan_b200: {what} must be a CUDA tensor -- this package runs on sm_100a '
                           'kernels only and has no CPU path')


BF16, F32, F16 = L.DT_BF16, L.DT_F32, L.DT_F16
TORCH_DT = {BF16: torch.bfloat16, F32: torch.float32, F16: torch.float16}


class Act:
    """A (possibly channel-sliced) NHWC activation: keeps its storage alive, carries ptr / C / pixel stride / dtype."""
    __slots__ = ('t', 'ptr', 'C', 'ld', 'B', 'H', 'W', 'dt', 'tw', 'im2col')

    def __init__(self, t, B, H, W, C, ld=None, off=0, dt=BF16):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        self.ld = ld if ld is not None else C
        self.dt = dt
        self.ptr = t.data_ptr() + off * self.esize
        self.tw = None      # bf16 twin of an f16 activation (same shape / stride), read by the wgrad GEMMs
        self.im2col = 0     # > 0: this is the stride-2 im2col A[b, oy, ox, c*16 + tap] of an image with `im2col` channels

    @property
    def b16(self):
        """The bf16 version of this activation (itself if it already is bf16)."""
        if self.dt == BF16:
            return self
        if self.tw is None:
            raise RuntimeError('activation has no bf16 twin (forward ran without save=True)')
        return self.tw

    @property
    def twptr(self):
        return self.tw.ptr if self.tw is not None else None

    @property
    def esize(self):
        return 4 if self.dt == F32 else 2

    def slice(self, c0, C):
        a = Act(self.t, self.B, self.H, self.W, C, self.ld, (self.ptr - self.t.data_ptr()) // self.esize + c0, self.dt)
        if self.tw is not None:
            a.tw = self.tw.slice(c0, C)
        return a

    def first(self, nb):
        """The first nb images of the batch (same storage)."""
        return self.images(0, nb)

    def images(self, b0, nb):
        """Images b0 .. b0+nb-1 of the batch (same storage)."""
        a = Act(self.t, nb, self.H, self.W, self.C, self.ld, 0, self.dt)
        a.ptr = self.ptr + b0 * self.H * self.W * self.ld * self.esize
        a.im2col = self.im2col
        if self.tw is not None:
            a.tw = self.tw.images(b0, nb)
        return a


# ------------------------------------------------------------------------------------------------
# Streams.  The step is ~140 short launches, most of them far too small to fill 148 SMs, and it contains long
# independent chains (weight-gradients vs. the data-gradient chain; the discriminator update vs. the generator
# backward; D(real) vs. the generator forward).  Those chains are issued on side streams that fork from / join
# into the caller's stream (captured as parallel branches of the step's CUDA graph).  Every buffer allocated
# during a step is kept alive until the next step begins, so the caching allocator can never hand a block
# that a side stream is still reading to a later allocation on another stream.
# ------------------------------------------------------------------------------------------------
_KEEP = []
_KEEPING = False
_SIDE = {}


_ZPOOL = {'buf': None, 'off': 0}
ZPOOL_FLOATS = 1 << 20          # 4 MB of zeros per step for the small accumulators (statistics, loss partials, ...)


def begin_step(device=None):
    """Drop the previous step's buffers (all streams were joined at its end) and start keeping this step's.
    With a device: also zero ONE pool that `zeros()` carves the step's small accumulators from (one fill kernel at the
    start of the step instead of ~40 tiny ones on the critical path)."""
    global _KEEPING
    del _KEEP[:]
    _FORKED.clear()
    _KEEPING = True
    _ZPOOL['buf'] = None
    if device is not None:
        _ZPOOL['buf'] = keep(torch.zeros(ZPOOL_FLOATS, device=device, dtype=torch.float32))
        _ZPOOL['off'] = 0
        ensure_workspace(torch.device(device), rewind=True)


def end_step():
    """Stop collecting (the collected buffers stay alive until the next begin_step)."""
    global _KEEPING
    _KEEPING = False
    _ZPOOL['buf'] = None


def keep(t):
    if _KEEPING:
        _KEEP.append(t)
    return t


def side_streams(device, n=7):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return _SIDE[key]


_FORKED = set()


def fork(side):
    """`side` continues after everything issued so far on the current stream."""
    side.wait_stream(torch.cuda.current_stream())
    _FORKED.add(side)


def join(side):
    """The current stream continues after everything issued so far on `side` (no-op if `side` was never forked in this
    step: joining a stream that is not part of the graph being captured would be an error).  `side` may be a list."""
    if isinstance(side, (list, tuple)):
        for s_ in side:
            join(s_)
        return
    if side in _FORKED:
        torch.cuda.current_stream().wait_stream(side)


def zeros(shape, device, dtype=torch.float32):
    """Zero-initialised accumulator: a 256-byte aligned slice of the step's zero pool when one is open, else torch.zeros."""
    pool = _ZPOOL['buf']
    if _KEEPING and pool is not None and dtype == torch.float32 and pool.device == torch.device(device):
        n = 1
        for d in (shape if isinstance(shape, (tuple, list)) else (shape,)):
            n *= int(d)
        off = _ZPOOL['off']
        if off + n <= pool.numel():
            _ZPOOL['off'] = (off + n + 63) // 64 * 64
            return pool[off:off + n].view(shape)
    return keep(torch.zeros(shape, device=device, dtype=dtype))


_POISON = os.environ.get('PATCHGAN_B200_POISON', '0') != '0'    # debug: fill fresh buffers with NaN instead of leaving them


def _empty_nan(shape, device, dtype):
    return torch.full(shape, float('nan'), device=device, dtype=dtype)


def new_act(B, H, W, C, device, dt=BF16, zero=False, twin=False):
    fn = torch.zeros if zero else (_empty_nan if _POISON else torch.empty)
    t = keep(fn((B, H, W, C), device=device, dtype=TORCH_DT[dt]))
    a = Act(t, B, H, W, C, dt=dt)
    if twin and dt == F16:
        a.tw = Act(keep(fn((B, H, W, C), device=device, dtype=torch.bfloat16)), B, H, W, C, dt=BF16)
    return a


def conv_desc(mode, stride, pad, B, Hin, Win, Hout, Wout, C1, C2, ld1, ld2, N, ldo, n_valid=None, act=0, out_dt=BF16,
              has_bias=0, in_dt=BF16, n_first=0, c_valid=0, ldw=0):
    return L.ConvDesc(mode, stride, pad, B, Hin, Win, Hout, Wout, C1, C2, ld1, ld2, N, ldo,
                      N if n_valid is None else n_valid, act, out_dt, has_bias, in_dt, n_first, c_valid, ldw)


class Config:
    """Process-wide switches (tests flip `impl` to compare the tcgen05 path with the SIMT path).
    PATCHGAN_B200_IMPL = auto | simt | tcgen05 selects the convolution implementation."""
    impl = {'auto': L.IMPL_AUTO, 'simt': L.IMPL_SIMT, 'tcgen05': L.IMPL_TCGEN05}[
        os.environ.get('PATCHGAN_B200_IMPL', 'auto')]
    # 16-bit type of the FORWARD tensor-core operands (activations + weights).  fp16 and bf16 run at the same
    # tcgen05 rate; fp16's 3 extra mantissa bits keep every layer within the 1e-2 activation tolerance (bf16
    # reaches 1.3-2e-2 at the 2x2 bottleneck).  Gradient tensors are always bf16 (range).
    fwd_dt = {'fp16': L.DT_F16, 'bf16': L.DT_BF16}[os.environ.get('PATCHGAN_B200_FWD_DTYPE', 'fp16')]
    # independent chains of the step on side streams (0 = everything on the caller's stream)
    streams = os.environ.get('PATCHGAN_B200_STREAMS', '1') != '0'
    # one-real-channel layers as pointwise tap products (0 = run them as 16x padded 4x4 implicit GEMMs)
    taps = os.environ.get('PATCHGAN_B200_TAPS', '1') != '0'
    # conv + InstanceNorm + activation as ONE launch with TMEM-resident accumulators (forward), and the block's backward
    # fused into the data-gradient convolution that feeds it (0 = separate statistics / apply / backward kernels)
    fused_fwd = os.environ.get('PATCHGAN_B200_FUSED_FWD', '1') != '0'
    fused_bwd = os.environ.get('PATCHGAN_B200_FUSED_BWD', '1') != '0'
    # the generator's ~20 weight-gradients as grouped launches (pg_conv_wgrad_group) instead of one launch each
    group_wgrad = os.environ.get('PATCHGAN_B200_GROUP_WGRAD', '1') != '0'


def conv_flops(desc):
    """2*MACs of the contraction as launched (padded channel counts)."""
    c = desc.C1 + desc.C2
    if desc.mode == L.PG_CONV1X1:
        return 2.0 * desc.B * desc.Hout * desc.Wout * c * desc.N
    if desc.mode == L.PG_CONVT:
        return 2.0 * desc.B * desc.Hin * desc.Win * c * desc.N * 16
    return 2.0 * desc.B * desc.Hout * desc.Wout * c * desc.N * 16


def desc_tag(d):
    return (f"{('conv', 'convT', 'conv1x1')[d.mode]} s{d.stride}p{d.pad} B{d.B} {d.Hin}x{d.Win}->{d.Hout}x{d.Wout} "
            f"C{d.C1}+{d.C2} N{d.N}")


_WS = {}


def ensure_workspace(device, rewind=False):
    """Register the split-K exchange scratch of the library (one 64 MB buffer per device, allocated once).
    rewind=True (start of a step: nothing is in flight) hands the slices out from the beginning again."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _WS:
        _WS[key] = torch.empty(64 << 20, device=device, dtype=torch.uint8)
        _WS[('fused', key)] = torch.empty(8 << 20, device=device, dtype=torch.uint8)
        rewind = True
    if rewind:
        with torch.cuda.device(key):
            # (not a stream call: bypass the profiler / stamper hooks of L.call)
            L.check(L.lib().pg_conv_set_workspace(_WS[key].data_ptr(), _WS[key].numel()), 'pg_conv_set_workspace')


def fused_workspace(device):
    """Scratch of the one-launch conv + InstanceNorm kernels' split-K mode (one such launch is in flight at a time)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ensure_workspace(device)
    return _WS[('fused', key)]


def run_conv(desc, src1, src2, w, bias, out, stats=None):
    """stats: zeroed float32 (B, N, 2) tensor -> also accumulates the InstanceNorm sums of the output."""
    if L.PROFILER is not None:
        L.PROFILER.note(conv_flops(desc), desc_tag(desc))
    if stats is not None:
        L.call('pg_conv_fwd_stats', ctypes.byref(desc), src1.ptr, src2.ptr if src2 is not None else None,
               w if isinstance(w, int) else w.data_ptr(), bias.data_ptr() if bias is not None else None, out.ptr,
               stats.data_ptr(), Config.impl, _stream())
        return
    L.call('pg_conv_fwd', ctypes.byref(desc), src1.ptr, src2.ptr if src2 is not None else None,
           w if isinstance(w, int) else w.data_ptr(), bias.data_ptr() if bias is not None else None, out.ptr, out.twptr,
           Config.impl, _stream())


SKIP = set(filter(None, os.environ.get('PATCHGAN_B200_SKIP', '').split(',')))   # timing experiments only


def run_conv_dgrad_act(desc, dy, w, out, y):
    """Data-gradient + backward of the activation in front of it (desc.act, from its saved output y) in one launch."""
    if L.PROFILER is not None:
        L.PROFILER.note(conv_flops(desc), desc_tag(desc))
    L.call('pg_conv_dgrad_act', ctypes.byref(desc), dy.ptr, w if isinstance(w, int) else w.data_ptr(), out.ptr, y.ptr, y.ld,
           y.dt, Config.impl, _stream())


X_IS_XHAT, X_IS_OUTPUT = 0x100, 0x200      # flags of pg_norm_act_bwd's x dtype argument (include/patchgan_b200.h)
INVERTIBLE = ('leakyrelu', 'none', None)    # activations whose input is recovered from the saved output


def fused_conv_norm(desc, src1, src2, w, out, act, drop_p, seed, salt, want_xhat):
    """conv -> InstanceNorm -> act -> dropout in one launch (pg_conv_norm_fwd).  Returns (sums, xhat Act or None), or None
    if the geometry does not fit tensor memory (the caller then runs the separate kernels)."""
    if not (taps_enabled() and Config.fused_fwd):
        return None
    fn = L.FusedNorm()
    fn.kind, fn.act, fn.drop_p, fn.salt = L.FUSED_FWD, act, float(drop_p), salt
    fn.seed = seed.data_ptr() if drop_p > 0 else None
    ws = fused_workspace(out.t.device)
    fn.ws, fn.ws_bytes = ws.data_ptr(), ws.numel()
    if not L.lib().pg_conv_norm_supported(ctypes.byref(desc), ctypes.byref(fn), int(out.tw is not None)):
        return None
    dev = out.t.device
    sums = zeros((desc.B, desc.N, 2), dev)
    sync = zeros(16, dev)            # (held until the call is issued: a dropped temporary's block would be handed out again)
    fn.sums, fn.sync = sums.data_ptr(), sync.data_ptr()
    xh = None
    if want_xhat:
        xh = new_act(out.B, out.H, out.W, out.C, dev, dt=out.dt)
        fn.xhat, fn.xhat_ld = xh.ptr, xh.ld
    if L.PROFILER is not None:
        L.PROFILER.note(conv_flops(desc), desc_tag(desc))
    L.call('pg_conv_norm_fwd', ctypes.byref(desc), src1.ptr, src2.ptr if src2 is not None else None,
           w if isinstance(w, int) else w.data_ptr(), out.ptr, out.twptr, ctypes.byref(fn), _stream())
    return sums, xh


class Block:
    """What the backward of one conv -> [norm] -> act -> [dropout] block needs from its forward."""
    __slots__ = ('spec', 'raw', 'sums', 'out', 'dp', 'xh', 'salt', 'bn')

    def __init__(self, spec, raw, sums, out, dp, xh, salt, bn=None):
        self.spec, self.raw, self.sums, self.out, self.dp, self.xh, self.salt = spec, raw, sums, out, dp, xh, salt
        self.bn = bn      # BatchNorm2d blocks: (gamma, beta, dgamma, dbeta, training) tensors / flag, else None


def _shift_desc(desc, n0, n):
    """desc restricted to output channels [n0, n0 + n) (the pointers are offset by the caller)."""
    d = L.ConvDesc()
    ctypes.memmove(ctypes.byref(d), ctypes.byref(desc), ctypes.sizeof(d))
    d.N = n
    d.n_valid = n
    return d


def dgrad_block_bwd(desc, dy, w, wrow_bytes, din, n_norm, blk, dskip, seed):
    """din = data-gradient of a layer (geometry desc: dy -> din, all desc.N channels); its first n_norm channels are the
    gradient wrt the output of block `blk` and are taken through that block's backward (dropout, activation, InstanceNorm;
    `dskip` = gradient arriving over the skip connection, added first).  Returns d(conv output of blk) as a bf16 Act.
    One launch when the fused kernel applies (pg_conv_dgrad_norm_bwd), else data-gradient + pg_norm_act_bwd."""
    s = blk.spec
    act = L.ACT[s.act]
    wp = w if isinstance(w, int) else w.data_ptr()
    can = blk.xh is not None or (s.act in INVERTIBLE and blk.dp == 0)
    if s.norm and can and taps_enabled() and Config.fused_bwd and blk.bn is None:
        dev = din.t.device
        fn = L.FusedNorm()
        fn.kind, fn.act, fn.n_norm, fn.drop_p, fn.salt = L.FUSED_BWD, act, n_norm, float(blk.dp), blk.salt
        fn.seed = seed.data_ptr() if blk.dp > 0 else None
        fn.y_dtype = blk.out.dt
        if blk.xh is not None:
            fn.xhat, fn.xhat_ld = blk.xh.ptr, blk.xh.ld
        else:
            fn.y, fn.y_ld = blk.out.ptr, blk.out.ld
        if dskip is not None:
            fn.dskip, fn.dskip_ld = dskip.ptr, dskip.ld
        fn.sums = blk.sums.data_ptr()
        fws = fused_workspace(dev)
        fn.ws, fn.ws_bytes = fws.data_ptr(), fws.numel()
        plans = [(desc, None)]
        if n_norm < desc.N:       # the normalised half alone may fit tensor memory when the whole concat gradient does not
            plans.append((_shift_desc(desc, 0, n_norm), _shift_desc(desc, n_norm, desc.N - n_norm)))
        for d1, d2 in plans:
            if not L.lib().pg_conv_norm_supported(ctypes.byref(d1), ctypes.byref(fn), 0):
                continue
            # one zeroed workspace: [bsums | barrier counter], alive until the call is issued
            ws = zeros(desc.B * n_norm * 2 + 16, dev)
            fn.bsums = ws.data_ptr()
            fn.sync = ws.data_ptr() + desc.B * n_norm * 2 * 4
            if L.PROFILER is not None:
                L.PROFILER.note(conv_flops(d1), desc_tag(d1))
            L.call('pg_conv_dgrad_norm_bwd', ctypes.byref(d1), dy.ptr, wp, din.ptr, ctypes.byref(fn), _stream())
            if d2 is not None:
                rest = din.slice(n_norm, desc.N - n_norm)
                run_conv(d2, dy, None, wp + n_norm * wrow_bytes, None, rest)
            return din.slice(0, n_norm)
    run_conv(desc, dy, None, wp, None, din)
    d_prev = din.slice(0, n_norm)
    if not s.norm:
        return act_bwd_out(blk.out, d_prev, act)
    if blk.bn is not None:
        return batchnorm_bwd(blk.raw, blk.sums, blk.bn, s.cout, d_prev, dskip, act, blk.dp, seed, blk.salt)
    if blk.raw is not None:
        x, kind = blk.raw, 0
    elif blk.xh is not None:
        x, kind = blk.xh, X_IS_XHAT
    else:
        x, kind = blk.out, X_IS_OUTPUT
    return norm_bwd(x, blk.sums, d_prev, dskip, act, blk.dp, seed, blk.salt, kind)


_WRR = [0]


def pick_wstream(wstream):
    """wstream may be one side stream or a list of them (independent weight-gradients then go round-robin, so that one
    layer's launch does not wait behind another's for no reason)."""
    if isinstance(wstream, (list, tuple)):
        _WRR[0] += 1
        return wstream[_WRR[0] % len(wstream)]
    return wstream


def run_wgrad(desc, a, g, dw_ptr, ld_n, n_real, c_real, wstream=None):
    """wstream: issue the weight-gradient on that side stream (it forks here, after its operands were produced on
    the current stream; the caller joins it before the optimizer step)."""
    if 'wgrad' in SKIP:
        return
    wstream = pick_wstream(wstream)
    if wstream is not None:
        fork(wstream)
        with torch.cuda.stream(wstream):
            return run_wgrad(desc, a, g, dw_ptr, ld_n, n_real, c_real)
    if L.PROFILER is not None:
        L.PROFILER.note(conv_flops(desc), desc_tag(desc))
    L.call('pg_conv_wgrad', ctypes.byref(desc), a.ptr, g.ptr, g.ld, dw_ptr, ld_n, n_real, c_real, Config.impl, _stream())


# ------------------------------------------------------------------------------------------------
# Layers with ONE real channel on one side (generator output ConvTranspose2d(2nf -> 1), discriminator last
# Conv2d(8ndf -> 1), the mask-channel data-gradient of the discriminator's first layer).  As 4x4 convolutions they are
# GEMMs with N = 1 (16x padding on the tensor cores) and too many FLOPs per byte for the CUDA cores; as pointwise
# products over the 16 taps they are GEMMs with N = 16 (or K = 16) that read / write the wide tensor exactly once:
#   forward        P[q][tap] = sum_c in[q][c] W[tap][c]          then scatter: out[p] = act(b + sum_tap P[q(p,tap)][tap])
#   data-gradient  gather: G[q][tap] = dy[p(q,tap)]              then dx[q][c] = sum_tap G[q][tap] W[c][tap]
#   weight-grad.   dW[c][tap] = sum_q in[q][c] G[q][tap]
# (include/patchgan_b200.h: PG_CONV1X1, pg_taps_scatter, pg_taps_gather)
# ------------------------------------------------------------------------------------------------

def taps_enabled():
    return Config.impl == L.IMPL_AUTO and Config.taps


def taps_forward(mode, stride, pad, src1, src2, w_ptr, bias, act, out, ch):
    """out[..., ch] = act(bias + 4x4 conv / convT of (src1 | src2) with the 16 x Ctot tap matrix at w_ptr)."""
    B, Hq, Wq = src1.B, src1.H, src1.W
    P = new_act(B, Hq, Wq, 16, src1.t.device, dt=F32)
    c2, ld2 = (src2.C, src2.ld) if src2 is not None else (0, 0)
    run_conv(conv_desc(L.PG_CONV1X1, 1, 0, B, Hq, Wq, Hq, Wq, src1.C, c2, src1.ld, ld2, 16, 16, out_dt=F32,
                       in_dt=src1.dt), src1, src2, w_ptr, None, P)
    L.call('pg_taps_scatter', mode, stride, pad, B, Hq, Wq, out.H, out.W, P.ptr, 16,
           bias.data_ptr() if bias is not None else None, act, out.ptr, out.dt, out.ld, ch, _stream())


def taps_gather(mode, stride, pad, dy, ch, B, Hq, Wq):
    """G[q][tap] = dy[p(q, tap)][ch] as a bf16 Act (B, Hq, Wq, 16)."""
    G = new_act(B, Hq, Wq, 16, dy.t.device)
    L.call('pg_taps_gather', mode, stride, pad, B, Hq, Wq, dy.H, dy.W, dy.ptr, dy.ld, ch, G.ptr, _stream())
    return G


def taps_dgrad(G, w16, out, y=None, act=0):
    """out[q][c] = sum_tap G[q][tap] * w16[c][tap]   (w16: bf16 [C][16]);  with y: times act'(y) (fused activation backward)"""
    if out.C % 2 == 0 and out.C <= 1024:
        # 16 FMAs per output: a CUDA-core streaming kernel (as a GEMM it would be K = 16 in front of a C-wide epilogue)
        L.call('pg_taps_dgrad_act', G.ptr, w16.data_ptr(), out.ptr, out.ld, y.ptr if y is not None else None,
               y.ld if y is not None else 0, y.dt if y is not None else 0, act if y is not None else 0,
               G.B * G.H * G.W, out.C, _stream())
        return
    d = conv_desc(L.PG_CONV1X1, 1, 0, G.B, G.H, G.W, G.H, G.W, 16, 0, 16, 0, out.C, out.ld, act=act if y is not None else 0)
    if y is not None:
        run_conv_dgrad_act(d, G, w16, out, y)
    else:
        run_conv(d, G, None, w16, None, out)


def taps_wgrad(G, x, dw_ptr, c_real, wstream=None, eng=None):
    """dw[c*16 + tap] += sum_q x[q][c] * G[q][tap]   (eng: the network engine, when its weight-gradients are grouped)"""
    d = conv_desc(L.PG_CONV1X1, 1, 0, G.B, G.H, G.W, G.H, G.W, x.C, 0, x.ld, 0, 16, 16, out_dt=BF16, in_dt=BF16, ldw=16)
    (eng.wgrad_direct if eng is not None else run_wgrad)(d, x, G, dw_ptr, 1, 16, c_real, wstream)



def first_im2col(B, H, W, cin, device, twin):
    """Empty stride-2 im2col matrix of a first layer's input: Act (B, H/2, W/2, cin*16), k = c*16 + tap."""
    a = new_act(B, H // 2, W // 2, cin * 16, device, dt=Config.fwd_dt, twin=twin)
    a.im2col = cin
    if a.tw is not None:
        a.tw.im2col = cin
    return a


def im2col_fill(a, b0, src_ptr, strides, C, k_off, B, H, W):
    """Write channels [k_off, k_off + C) of images [b0, b0 + B) of the im2col matrix `a` from an f32 source with element
    strides (batch, channel, row, column)."""
    off = b0 * a.H * a.W * a.ld * 2
    L.call('pg_im2col_s2', src_ptr, strides[0], strides[1], strides[2], strides[3], C, B, H, W, a.ptr + off,
           a.tw.ptr + off if a.tw is not None else None, a.ld, k_off, a.dt, _stream())


def nchw_strides(t):
    B, C, H, W = t.shape
    return (C * H * W, H * W, W, 1)


def first_conv(a, w_first, bias, act, out, n_valid, stats=None):
    """First Conv2d(k4, s2, p1) on the im2col matrix: a pointwise product with the (Np, cin*16) weight matrix.
    n_valid = real output channels (the bias has only that many entries; padded channels are stored as 0)."""
    run_conv(conv_desc(L.PG_CONV1X1, 1, 0, a.B, a.H, a.W, a.H, a.W, a.C, 0, a.ld, 0, w_first.shape[0], out.ld,
                       n_valid=n_valid, act=act, out_dt=out.dt, has_bias=int(bias is not None), in_dt=a.dt), a, None,
             w_first, bias, out, stats)


def first_wgrad(a, g, dw_ptr, n_real, wstream=None, eng=None):
    """dW[n][c][tap] += sum_o g[o][n] * A[o][c*16 + tap]: lands in the reference (Cout, Cin, 4, 4) layout."""
    d = conv_desc(L.PG_CONV1X1, 1, 0, a.B, a.H, a.W, a.H, a.W, a.C, 0, a.ld, 0, g.C, g.C, out_dt=BF16, in_dt=BF16, ldw=1)
    (eng.wgrad_direct if eng is not None else run_wgrad)(d, a.b16, g, dw_ptr, a.C, n_real, a.C, wstream)



class LayerSpec:
    """One 4x4 convolution layer of either network."""

    def __init__(self, kind, stride, c1, c2, cout, bias, act, norm, dropout, wname, bname=None, norm_after_act=False,
                 bn=None):
        self.bn = bn          # norm_layer = nn.BatchNorm2d: state_dict prefix of the layer's BatchNorm2d module, else None
        self.kind, self.stride, self.c1, self.c2, self.cout = kind, stride, c1, c2, cout
        self.bias, self.act, self.norm, self.dropout = bias, act, norm, dropout
        self.wname, self.bname, self.norm_after_act = wname, bname, norm_after_act
        self.c1p, self.c2p, self.np = rup16(c1), (rup16(c2) if c2 else 0), rup16(cout)
        self.cin, self.cinp = c1 + c2, self.c1p + self.c2p


class PackedWeights:
    """bf16 operand copies of one layer's fp32 weight: forward form [Np][16][Cinp], dgrad form [Cinp][16][Np]."""

    def __init__(self, spec, device):
        self.spec = spec
        self.fwd_dt = Config.fwd_dt
        self.fwd = torch.empty((spec.np, 16, spec.cinp), device=device, dtype=TORCH_DT[self.fwd_dt])
        self.bwd = torch.empty((spec.cinp, 16, spec.np), device=device, dtype=torch.bfloat16)
        # 1-output-channel layers: bf16 copy of the master weight viewed as [Cin][16 taps] (tap-product data-gradient)
        self.taps_ok = spec.cout == 1 and spec.c1 == spec.c1p and spec.c2 == spec.c2p
        self.w16 = torch.empty((spec.cinp, 16), device=device, dtype=torch.bfloat16) if self.taps_ok else None

        # first layers (stride-2 conv reading the image): f16/bf16 copy of the master weight as [Np][Cin*16]
        self.wfirst = None

    def pack_w16(self, w):
        if self.wfirst is not None:
            K = self.spec.cin * 16
            L.call('pg_copy_f32_to_bf16_slice', w.data_ptr(), K, self.wfirst.data_ptr(), K, 0, K, self.spec.cout, self.fwd_dt,
                   _stream())
        if self.w16 is not None:
            n = self.spec.cinp
            L.call('pg_copy_f32_to_bf16_slice', w.data_ptr(), 16, self.w16.data_ptr(), 16, 0, 16, n, BF16, _stream())

    def jobs(self, w):
        """The pg_pack_weight calls of `pack` as (src, dst, N, Np, C1, C1p, C2, C2p, sn, sc, flip, dtype) tuples."""
        s, wp = self.spec, w.data_ptr()
        if s.kind == 'conv':
            return [(wp, self.fwd.data_ptr(), s.cout, s.np, s.c1, s.c1p, 0, 0, s.cin * 16, 16, 0, self.fwd_dt),
                    (wp, self.bwd.data_ptr(), s.cin, s.cinp, s.cout, s.np, 0, 0, 16, s.cin * 16,
                     1 if s.stride == 1 else 0, BF16)]
        out = [(wp, self.fwd.data_ptr(), s.cout, s.np, s.c1, s.c1p, s.c2, s.c2p, 16, s.cout * 16, 0, self.fwd_dt),
               (wp, self.bwd.data_ptr(), s.c1, s.c1p, s.cout, s.np, 0, 0, s.cout * 16, 16, 0, BF16)]
        if s.c2:
            out.append((wp + s.c1 * s.cout * 16 * 4, self.bwd.data_ptr() + s.c1p * 16 * s.np * 2, s.c2, s.c2p, s.cout,
                        s.np, 0, 0, s.cout * 16, 16, 0, BF16))
        return out

    def pack(self, w):
        s, st = self.spec, _stream()
        wp = w.data_ptr()
        if s.kind == 'conv':      # w: (cout, cin, 4, 4)
            L.call('pg_pack_weight', wp, self.fwd.data_ptr(), s.cout, s.np, s.c1, s.c1p, 0, 0, s.cin * 16, 16, 0, self.fwd_dt, st)
            # dgrad operand W'[ci][tap][co]; stride-1 layers run dgrad as a flipped stride-1 conv
            L.call('pg_pack_weight', wp, self.bwd.data_ptr(), s.cin, s.cinp, s.cout, s.np, 0, 0, 16, s.cin * 16,
                   1 if s.stride == 1 else 0, BF16, st)
        else:                     # convT, w: (cin_total, cout, 4, 4)
            L.call('pg_pack_weight', wp, self.fwd.data_ptr(), s.cout, s.np, s.c1, s.c1p, s.c2, s.c2p, 16, s.cout * 16, 0,
                   self.fwd_dt, st)
            # rows of the dgrad operand live in the padded-concat channel space
            L.call('pg_pack_weight', wp, self.bwd.data_ptr(), s.c1, s.c1p, s.cout, s.np, 0, 0, s.cout * 16, 16, 0, BF16, st)
            if s.c2:
                L.call('pg_pack_weight', wp + s.c1 * s.cout * 16 * 4, self.bwd.data_ptr() + s.c1p * 16 * s.np * 2, s.c2,
                       s.c2p, s.cout, s.np, 0, 0, s.cout * 16, 16, 0, BF16, st)


class NetEngine:
    """Common machinery: parameter lookup, weight packing with change detection."""

    def __init__(self, module, specs):
        self.module = module
        self.specs = specs
        self.packed = None
        self._stamp = None
        self._jobs = None
        self.seed = None   # device uint64 dropout counter
        self._tm_buf = self._tm_table = self._tm_jobs = None
        self._tm_off = 0

    def params(self):
        return dict(self.module.named_parameters())

    def device(self):
        return next(self.module.parameters()).device

    def mark_dirty(self):
        self._stamp = None

    def ensure_packed(self):
        ps = self.params()
        dev = self.device()
        if dev.type == 'cuda':
            # outside a Trainer step (module called on its own, one stream) the split-K slices are handed out from the
            # start again on every call; inside a step begin_step has done so and side streams are in play
            ensure_workspace(dev, rewind=not _KEEPING)
        if dev.type != 'cuda':
            raise RuntimeError('patchgan_b200: module parameters must live on a CUDA device (no CPU path)')
        stamp = tuple((ps[s.wname].data_ptr(), ps[s.wname]._version) for s in self.specs)
        if self.packed is None or self.packed[0].fwd.device != dev or self.packed[0].fwd_dt != Config.fwd_dt:
            self.packed = [PackedWeights(s, dev) for s in self.specs]
            self._jobs = None
            s0 = self.specs[0]
            if s0.kind == 'conv' and s0.stride == 2 and s0.c2 == 0:
                self.packed[0].wfirst = torch.zeros((s0.np, s0.cin * 16), device=dev, dtype=TORCH_DT[Config.fwd_dt])
            self.seed = torch.zeros(1, device=dev, dtype=torch.int64)
            self._stamp = None
        if stamp != self._stamp:
            ptrs = tuple(s_[0] for s_ in stamp)
            if self._jobs is None or self._jobs[0] != ptrs:
                self._jobs = (ptrs,) + self._build_jobs(ps, dev)
            _, table, njobs, ntiles = self._jobs
            L.call('pg_pack_weights_multi', table.data_ptr(), njobs, ntiles, _stream())
            self._stamp = stamp

    JOB_DT = np.dtype([('src', '<u8'), ('dst', '<u8'), ('sn', '<i8'), ('sc', '<i8'), ('N', '<i4'), ('Np', '<i4'),
                       ('C1', '<i4'), ('C1p', '<i4'), ('C2', '<i4'), ('C2p', '<i4'), ('flip', '<i4'), ('dt', '<i4'),
                       ('tile_begin', '<i4'), ('ctiles', '<i4')])

    def _build_jobs(self, ps, dev, layers=None):
        """Device table for pg_pack_weights_multi: every layer's forward + dgrad operand packs in one launch
        (layers = (first, last): only those layers)."""
        rows, tile = [], 0
        for pw in (self.packed if layers is None else self.packed[layers[0]:layers[1]]):
            w = ps[pw.spec.wname].detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError(f'{pw.spec.wname}: weights must be contiguous float32')
            for (src, dst, N, Np, C1, C1p, C2, C2p, sn, sc, flip, dt) in pw.jobs(w):
                ctiles = (C1p + C2p + 31) // 32
                rows.append((src, dst, sn, sc, N, Np, C1, C1p, C2, C2p, flip, dt, tile, ctiles))
                tile += ctiles * ((Np + 7) // 8)
            # flat operand copies (flip = 2): first-layer [N][Cin*16] and tap-product [Cin][16] keep the master layout
            for dst_t, dt in ((pw.wfirst, pw.fwd_dt), (pw.w16, BF16)):
                if dst_t is not None:
                    n = w.numel()
                    rows.append((w.data_ptr(), dst_t.data_ptr(), n, 0, 0, 0, 0, 0, 0, 0, 2, dt, tile, 1))
                    tile += (n + 4095) // 4096
        arr = np.array(rows, dtype=self.JOB_DT)
        table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        return table, len(rows), tile

    # ---- weight-gradients accumulated tap-major (TMA bulk reduce) and written to the reference layout in one launch
    GRAD_JOB_DT = np.dtype([('S', '<u8'), ('dst', '<u8'), ('ld_n', '<i8'), ('N', '<i4'), ('C', '<i4'), ('Ns', '<i4'),
                            ('Cs', '<i4'), ('tile_begin', '<i4'), ('ctiles', '<i4')])

    def begin_backward(self, group=False):
        """Start collecting weight-gradient jobs for this backward pass; zero the tap-major scratch.
        group=True: the weight-gradients are not launched one by one but collected and issued by flush_wgrads()."""
        self._tm_jobs = []
        self._tm_done = 0
        self._tm_off = 0
        self._wg_group = [] if (group and taps_enabled() and Config.group_wgrad) else None
        if taps_enabled():
            if getattr(self, '_tm_buf', None) is None or self._tm_buf.device != self.device():
                n = sum(p.numel() for p in self.module.parameters())
                self._tm_buf = torch.empty(n + 64 * len(self.specs) * 4 + 1024, device=self.device(), dtype=torch.float32)
                self._tm_table = None
            self._tm_buf.zero_()

    def wgrad(self, desc, a, g, dst_ptr, ld_n, n_real, c_real, wstream=None, g2=None, n_split=0):
        """Weight gradient of one (layer, source): dst[n*ld_n + c*16 + tap] (reference layout) receives it -- directly
        (atomics) or, by default, through the tap-major scratch + finalize_grads().
        g2 / n_split (grouped launches only): g is the virtual concat [g | g2] of a decoder layer's two sources."""
        jobs = getattr(self, '_tm_jobs', None)
        if not (taps_enabled() and jobs is not None and desc.mode == L.PG_CONV) or 'wgrad' in SKIP:
            return run_wgrad(desc, a, g, dst_ptr, ld_n, n_real, c_real, wstream)
        Ns, Cs = n_real, (c_real + 3) // 4 * 4
        need = 16 * Ns * Cs
        off = (self._tm_off + 63) // 64 * 64            # 256-byte aligned slices
        if off + need > self._tm_buf.numel():
            return run_wgrad(desc, a, g, dst_ptr, ld_n, n_real, c_real, wstream)
        self._tm_off = off + need
        sp = self._tm_buf.data_ptr() + off * 4
        jobs.append((sp, dst_ptr, ld_n, n_real, c_real, Ns, Cs))
        if self._wg_group is not None:
            self._wg_group.append((desc, a, g, g.ld, 1, sp, Ns, Ns, Cs, Cs, g2, n_split))    # (the Acts keep their storage alive)
            return
        wstream = pick_wstream(wstream)
        if wstream is not None:
            fork(wstream)
        with torch.cuda.stream(wstream if wstream is not None else torch.cuda.current_stream()):
            if L.PROFILER is not None:
                L.PROFILER.note(conv_flops(desc), desc_tag(desc))
            L.call('pg_conv_wgrad_tapmajor', ctypes.byref(desc), a.ptr, g.ptr, g.ld, sp, Ns, Cs, Config.impl, _stream())

    def wgrad_direct(self, desc, a, g, dw_ptr, ld_n, n_real, c_real, wstream=None):
        """A weight gradient accumulated straight into the reference layout (pointwise layers): grouped or launched now."""
        if getattr(self, '_wg_group', None) is not None and 'wgrad' not in SKIP:
            self._wg_group.append((desc, a, g, g.ld, 0, dw_ptr, ld_n, n_real, c_real, 0, None, 0))
            return
        run_wgrad(desc, a, g, dw_ptr, ld_n, n_real, c_real, wstream)

    def flush_wgrads(self, wstream=None):
        """Issue the collected weight-gradients as grouped launches (at most 24 jobs each) on `wstream` (forked here) or the
        current stream.  Their operands must have been produced on the current stream."""
        pend = getattr(self, '_wg_group', None)
        if not pend:
            return
        self._wg_group = []
        wstream = pick_wstream(wstream)
        if wstream is not None:
            fork(wstream)
        with torch.cuda.stream(wstream if wstream is not None else torch.cuda.current_stream()):
            for i in range(0, len(pend), 24):
                chunk = pend[i:i + 24]
                arr = (L.WgradJob * len(chunk))()
                for j, (desc, a, g, ldg, tm, dw, ld_n, n_real, c_real, Cs, g2, n_split) in enumerate(chunk):
                    ctypes.memmove(ctypes.byref(arr[j].desc), ctypes.byref(desc), ctypes.sizeof(L.ConvDesc))
                    arr[j].a, arr[j].g, arr[j].ldg, arr[j].tap_major, arr[j].dw = a.ptr, g.ptr, ldg, tm, dw
                    if g2 is not None:
                        arr[j].g2, arr[j].ldg2, arr[j].n_split = g2.ptr, g2.ld, n_split
                    arr[j].ld_n, arr[j].n_real, arr[j].c_real, arr[j].Cs = ld_n, n_real, c_real, Cs
                    if L.PROFILER is not None:
                        L.PROFILER.note(conv_flops(desc), 'group')
                L.call('pg_conv_wgrad_group', arr, len(chunk), _stream())

    def finalize_grads(self, partial=False):
        """Write the tap-major weight-gradients of this backward pass to their reference-layout destinations (one launch).
        Call on a stream that is ordered after the weight-gradient launches concerned.
        partial=True: only the jobs collected so far (and not yet written); the pass goes on collecting -- used to update
        the layers whose gradients are final while the backward of the remaining ones still runs."""
        jobs = getattr(self, '_tm_jobs', None)
        if jobs is None:
            return
        start = getattr(self, '_tm_done', 0)
        todo = jobs[start:]
        if partial:
            self._tm_done = len(jobs)
        else:
            self._tm_jobs = None
            self._tm_done = 0
        if not todo:
            return
        key = (start, tuple(todo))
        if self._tm_table is None:
            self._tm_table = {}
        ent = self._tm_table.get(start)
        if ent is None or ent[0] != key:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError('patchgan_b200: weight-gradient job table changed during CUDA-graph capture')
            rows, tile = [], 0
            for (sp, dst, ld_n, N, C, Ns, Cs) in todo:
                ctiles = (C + 31) // 32
                rows.append((sp, dst, ld_n, N, C, Ns, Cs, tile, ctiles))
                tile += N * ctiles
            arr = np.array(rows, dtype=self.GRAD_JOB_DT)
            ent = self._tm_table[start] = (key, torch.from_numpy(arr.view(np.uint8).copy()).to(self.device()), len(rows),
                                           tile)
        _, table, njobs, ntiles = ent
        L.call('pg_grad_finalize_multi', table.data_ptr(), njobs, ntiles, _stream())

    def repack(self):
        """Unconditional repack (used inside captured graphs right after the optimizer step)."""
        self._stamp = None
        self.ensure_packed()

    def layers_match_parameters(self):
        """True if parameter i of the module is the weight of layer i (no biases): layer ranges are parameter ranges."""
        return [n for n, _ in self.module.named_parameters()] == [s.wname for s in self.specs]

    def repack_layers(self, first, last, complete):
        """Repack the operand copies of layers [first, last) on the current stream (their optimizer update is ordered
        before on this stream).  complete=True on the call that finishes the set: the copies are then up to date."""
        ps = self.params()
        ptrs = tuple(ps[s.wname].data_ptr() for s in self.specs)
        cache = getattr(self, '_part_jobs', None)
        if cache is None or cache.get('ptrs') != ptrs:
            cache = self._part_jobs = {'ptrs': ptrs}
        ent = cache.get((first, last))
        if ent is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError('patchgan_b200: weight pack job table built during CUDA-graph capture')
            ent = cache[(first, last)] = self._build_jobs(ps, self.device(), (first, last))
        table, njobs, ntiles = ent
        if njobs:
            L.call('pg_pack_weights_multi', table.data_ptr(), njobs, ntiles, _stream())
        if complete:
            self._stamp = tuple((ps[s.wname].data_ptr(), ps[s.wname]._version) for s in self.specs)

    def bump_seed(self):
        L.call('pg_counter_add', self.seed.data_ptr(), 1, _stream())


# ------------------------------------------------------------------------------------------------
# shared per-layer helpers
# ------------------------------------------------------------------------------------------------

def pack_rows(x1, x2, dst, first_image):
    """NCHW float x1 [, x2] -> full NHWC rows [x1 | x2 | 0...] of `dst` (and its bf16 twin) starting at image
    `first_image`; dst.ld must be 16 or 32 (one launch, no pre-zeroing needed)."""
    B, C1, H, W = x1.shape
    C2 = x2.shape[1] if x2 is not None else 0
    off = first_image * H * W * dst.ld * 2
    L.call('pg_pack2_nchw_rows', x1.data_ptr(), C1, x2.data_ptr() if x2 is not None else None, C2, dst.ptr + off,
           dst.tw.ptr + off if dst.tw is not None else None, B, H, W, dst.ld, dst.dt, _stream())


def norm_fwd(x, sums, out, act, drop_p, seed, salt, b0=0):
    HW = x.H * x.W
    L.call('pg_norm_act_fwd', x.ptr, x.dt, sums.data_ptr() + b0 * x.C * 8 if sums is not None else None, out.ptr, out.dt,
           out.twptr,
           x.B, HW, x.C, x.ld, out.ld, act, drop_p, seed.data_ptr() if seed is not None else None, salt, _stream())


def instnorm_stats(x, sums=None, b0=0):
    """(sum, sum of squares) per (image, channel); `sums` (zeroed, for the whole batch) + first image index b0 when
    x is a batch slice."""
    if sums is None:
        sums = zeros((x.B, x.C, 2), x.t.device)
    L.call('pg_instnorm_stats', x.ptr, x.dt, x.B, x.H * x.W, x.C, x.ld, sums.data_ptr() + b0 * x.C * 8, _stream())
    return sums


def norm_bwd(x, sums, dy1, dy2, act, drop_p, seed, salt, xkind=0):
    """Backward through dropout/act/InstanceNorm: returns d(raw) as a new bf16 Act.
    xkind: 0 = x is the pre-norm conv output, X_IS_XHAT = x is the saved xhat, X_IS_OUTPUT = x is the block's output."""
    dev = x.t.device
    HW = x.H * x.W
    dx = new_act(x.B, x.H, x.W, x.C, dev)
    bsums = zeros((x.B, x.C, 2), dev)
    sp = seed.data_ptr() if seed is not None else None
    p2, l2 = (dy2.ptr, dy2.ld) if dy2 is not None else (None, 0)
    st = _stream()
    if taps_enabled():      # one call: small maps are a single launch
        L.call('pg_norm_act_bwd', x.ptr, x.dt | xkind, sums.data_ptr(), dy1.ptr, dy1.ld, p2, l2, bsums.data_ptr(), dx.ptr,
               dx.ld, x.B, HW, x.C, x.ld, act, drop_p, sp, salt, st)
        return dx
    L.call('pg_norm_act_bwd_reduce', x.ptr, x.dt | xkind, sums.data_ptr(), dy1.ptr, dy1.ld, p2, l2, bsums.data_ptr(), x.B,
           HW, x.C, x.ld, act, drop_p, sp, salt, st)
    L.call('pg_norm_act_bwd_apply', x.ptr, x.dt | xkind, sums.data_ptr(), dy1.ptr, dy1.ld, p2, l2, bsums.data_ptr(),
           dx.ptr, dx.ld, x.B, HW, x.C, x.ld, act, drop_p, sp, salt, st)
    return dx


BN_MOMENTUM = 0.1     # nn.BatchNorm2d default (unet.py:20,55 call norm_layer(output_filt) with defaults)


def batchnorm_fwd(raw, sums, bn, c_real, out, act, drop_p, seed, salt, training):
    """raw (fp32 conv output) + its per-(image, channel) sums -> out = dropout(act(BatchNorm2d(raw))).
    bn = (gamma, beta, running_mean, running_var, num_batches_tracked) tensors of the layer's BatchNorm2d module.
    `sums` is folded to batch statistics in place (kept for the backward)."""
    gamma, beta, rmean, rvar, nbt = bn
    st = _stream()
    L.call('pg_bn_fold_fwd', sums.data_ptr(), raw.B, raw.C, raw.H * raw.W, rmean.data_ptr(), rvar.data_ptr(), c_real,
           BN_MOMENTUM, 1 if training else 0, st)
    if training and nbt is not None:
        L.call('pg_counter_add', nbt.data_ptr(), 1, st)
    L.call('pg_norm_affine_act_fwd', raw.ptr, raw.dt, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), c_real, out.ptr,
           out.dt, out.twptr, raw.B, raw.H * raw.W, raw.C, raw.ld, out.ld, act, drop_p,
           seed.data_ptr() if seed is not None else None, salt, st)


def batchnorm_bwd(raw, sums, bn, c_real, dy1, dy2, act, drop_p, seed, salt, parts=None):
    """Backward of dropout / act / BatchNorm2d: returns d(raw) as a new bf16 Act and accumulates dgamma / dbeta.
    bn = (gamma, beta, dgamma, dbeta, training).  parts = [(first image, images)]: groups of images that were normalised
    together (separate discriminator calls batched into one pass); default: the whole batch."""
    gamma, beta, dgamma, dbeta, training = bn
    dev = raw.t.device
    HW = raw.H * raw.W
    dx = new_act(raw.B, raw.H, raw.W, raw.C, dev)
    bsums = zeros((raw.B, raw.C, 2), dev)
    sp = seed.data_ptr() if seed is not None else None
    p2, l2 = (dy2.ptr, dy2.ld) if dy2 is not None else (None, 0)
    st = _stream()
    L.call('pg_norm_affine_act_bwd_reduce', raw.ptr, raw.dt, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), c_real,
           dy1.ptr, dy1.ld, p2, l2, bsums.data_ptr(), raw.B, HW, raw.C, raw.ld, act, drop_p, sp, salt, st)
    for b0, nb in (parts or [(0, raw.B)]):
        L.call('pg_bn_fold_bwd', bsums.data_ptr() + b0 * raw.C * 8, nb, raw.C, dgamma.data_ptr() if dgamma is not None else None,
               dbeta.data_ptr() if dbeta is not None else None, c_real, 1 if training else 0, st)
    L.call('pg_norm_affine_act_bwd_apply', raw.ptr, raw.dt, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), c_real,
           dy1.ptr, dy1.ld, p2, l2, bsums.data_ptr(), dx.ptr, dx.ld, raw.B, HW, raw.C, raw.ld, act, drop_p, sp, salt, st)
    return dx


def act_bwd_out(y, dy, act):
    dx = new_act(y.B, y.H, y.W, y.C, y.t.device)
    L.call('pg_act_bwd_from_output', y.ptr, y.dt, y.ld, dy.ptr, dy.ld, dx.ptr, dx.ld, y.B * y.H * y.W, y.C, act,
           _stream())
    return dx


# ------------------------------------------------------------------------------------------------
# Generator
# ------------------------------------------------------------------------------------------------

class GeneratorEngine(NetEngine):
    def __init__(self, module):
        nf, inc, outc = module.nf, module.input_nc, module.output_nc
        filts = [nf, nf * 2, nf * 4, nf * 8, nf * 8, nf * 8, nf * 8]
        specs, prev = [], inc
        bn = bool(getattr(module, 'batchnorm', False))
        for i, f in enumerate(filts):
            specs.append(LayerSpec('conv', 2, prev, 0, f, False, module.activation, True, module.use_dropout,
                                   f'encoder.{i}.model.DownConv{i}.weight',
                                   bn=f'encoder.{i}.model.DownNorm{i}' if bn else None))
            prev = f
        enc_out = filts
        for i, f in enumerate(filts[:-1][::-1]):
            if i == 0:
                specs.append(LayerSpec('convT', 2, prev, 0, f, False, module.activation, False, False,
                                       f'decoder.{i}.model.UpConv{i}.weight'))
            else:
                specs.append(LayerSpec('convT', 2, prev, enc_out[6 - i], f, False, module.activation, True,
                                       module.use_dropout, f'decoder.{i}.model.UpConv{i}.weight',
                                       bn=f'decoder.{i}.model.UpNorm{i}' if bn else None))
            prev = f
        specs.append(LayerSpec('convT', 2, prev, enc_out[0], outc, False, module.final_act, False, False,
                               'decoder.6.model.UpConv6.weight'))
        super().__init__(module, specs)
        self.enc, self.dec = specs[:7], specs[7:]
        self.in_cp = rup16(inc)
        self.out_cp = rup16(outc)

    def bn_tensors(self, s):
        """(weight, bias, running_mean, running_var, num_batches_tracked) of layer s's BatchNorm2d module."""
        ps, bs = self.params(), dict(self.module.named_buffers())
        return (ps[s.bn + '.weight'], ps[s.bn + '.bias'], bs[s.bn + '.running_mean'], bs[s.bn + '.running_var'],
                bs.get(s.bn + '.num_batches_tracked'))

    def bn_names(self):
        """names of the BatchNorm2d affine parameters, in layer order (empty for InstanceNorm2d)"""
        return [s.bn + sfx for s in self.specs if s.bn is not None for sfx in ('.weight', '.bias')]

    def pack_input(self, x, twin=False):
        """NCHW float -> NHWC bf16 with channels zero-padded to 16."""
        B, C, H, W = x.shape
        if self.in_cp in (16, 32):
            a = new_act(B, H, W, self.in_cp, x.device, dt=Config.fwd_dt, twin=twin)
            pack_rows(x, None, a, 0)
            return a
        a = new_act(B, H, W, self.in_cp, x.device, dt=Config.fwd_dt, zero=True, twin=twin)
        L.call('pg_pack_nchw_f32_to_nhwc_bf16', x.data_ptr(), a.ptr, B, C, H, W, a.ld, 0, a.dt, _stream())
        if a.tw is not None:
            L.call('pg_pack_nchw_f32_to_nhwc_bf16', x.data_ptr(), a.tw.ptr, B, C, H, W, a.ld, 0, BF16, _stream())
        return a

    def forward(self, xin, training, save=True):
        """xin: Act (B,H,W,in_cp) bf16.  Returns (p: f32 Act (B,H,W,out_cp), ctx).
        ctx['enc'][i] = (input, pre-norm output or None, sums, output, dropout p, xhat or None);
        ctx['dec'][i] = (src1, src2, pre-norm output or None, sums or None, output, dropout p, xhat or None)."""
        self.ensure_packed()
        dev = xin.t.device
        B = xin.B
        ctx = {'enc': [], 'dec': [], 'training': training}
        h = xin
        enc_outs = []

        def norm_block(desc, src1, src2, w, s, Ho, Wo, salt):
            """conv (desc) -> InstanceNorm -> act -> dropout: one launch when the layer fits tensor memory, else three."""
            out = new_act(B, Ho, Wo, s.np, dev, dt=src1.dt, twin=save)
            dp = DROP_P if (training and s.dropout) else 0.0
            desc.ldo, desc.out_f32, desc.n_valid = out.ld, out.dt, s.cout
            fused = None
            if s.bn is None:
                fused = fused_conv_norm(desc, src1, src2, w, out, L.ACT[s.act], dp, self.seed, salt,
                                        save and (dp > 0 or s.act not in INVERTIBLE))
            if fused is not None:
                return None, fused[0], out, dp, fused[1]
            raw = new_act(B, Ho, Wo, s.np, dev, dt=F32)
            sums = zeros((B, s.np, 2), dev)
            desc.ldo, desc.out_f32, desc.n_valid = raw.ld, F32, s.np
            run_conv(desc, src1, src2, w, None, raw, sums)
            if s.bn is not None:
                # norm_layer = nn.BatchNorm2d: batch statistics (running ones in eval mode) + affine, three launches
                batchnorm_fwd(raw, sums, self.bn_tensors(s), s.cout, out, L.ACT[s.act], dp, self.seed, salt, training)
            else:
                norm_fwd(raw, sums, out, L.ACT[s.act], dp, self.seed, salt)
            return raw, sums, out, dp, None

        for i, s in enumerate(self.enc):
            Ho, Wo = (h.H, h.W) if h.im2col else (h.H // 2, h.W // 2)
            if Ho * Wo <= 1 and training:
                # aten::instance_norm raises for a single spatial element in training mode
                raise ValueError('Expected more than 1 spatial element when training (input too small for 7 '
                                 'stride-2 stages)')
            if h.im2col:
                w = self.packed[i].wfirst
                desc = conv_desc(L.PG_CONV1X1, 1, 0, B, h.H, h.W, h.H, h.W, h.C, 0, h.ld, 0, w.shape[0], s.np, in_dt=h.dt)
            else:
                w = self.packed[i].fwd
                desc = conv_desc(L.PG_CONV, 2, 1, B, h.H, h.W, Ho, Wo, h.C, 0, h.ld, 0, s.np, s.np, in_dt=h.dt)
            raw, sums, out, dp, xh = norm_block(desc, h, None, w, s, Ho, Wo, i)
            ctx['enc'].append((h, raw, sums, out, dp, xh) if save else None)
            enc_outs.append(out)
            h = out
        self.last_bottleneck = h          # encoder output (unet.py:119); UNet.forward(return_hidden=True) reads it
        for i, s in enumerate(self.dec):
            src1 = h
            src2 = enc_outs[6 - i] if i > 0 else None
            Ho, Wo = src1.H * 2, src1.W * 2
            pw = self.packed[7 + i]
            c2, ld2 = (src2.C, src2.ld) if src2 is not None else (0, 0)
            if s.norm:
                desc = conv_desc(L.PG_CONVT, 2, 1, B, src1.H, src1.W, Ho, Wo, src1.C, c2, src1.ld, ld2, s.np, s.np,
                                 in_dt=src1.dt)
                raw, sums, out, dp, xh = norm_block(desc, src1, src2, pw.fwd, s, Ho, Wo, 16 + i)
                ctx['dec'].append((src1, src2, raw, sums, out, dp, xh) if save else None)
            elif i < 6:
                out = new_act(B, Ho, Wo, s.np, dev, dt=src1.dt, twin=save)
                run_conv(conv_desc(L.PG_CONVT, 2, 1, B, src1.H, src1.W, Ho, Wo, src1.C, c2, src1.ld, ld2, s.np, out.ld,
                                   act=L.ACT[s.act], out_dt=out.dt, in_dt=src1.dt), src1, src2, pw.fwd, None, out)
                ctx['dec'].append((src1, src2, None, None, out, 0.0, None) if save else None)
            else:
                out = new_act(B, Ho, Wo, (s.cout + 3) // 4 * 4, dev, dt=F32)   # trimmed stride: real channels only
                fused = 0 if s.act == 'softmax' else L.ACT[s.act]
                if taps_enabled() and pw.taps_ok and s.act != 'softmax':
                    taps_forward(L.PG_CONVT, 2, 1, src1, src2, pw.fwd.data_ptr(), None, fused, out, 0)
                else:
                    run_conv(conv_desc(L.PG_CONVT, 2, 1, B, src1.H, src1.W, Ho, Wo, src1.C, c2, src1.ld, ld2, s.np,
                                       out.ld, n_valid=s.cout, act=fused, out_dt=F32, in_dt=src1.dt), src1, src2,
                             pw.fwd, None, out)
                if s.act == 'softmax':
                    L.call('pg_softmax_fwd', out.ptr, out.ptr, B * Ho * Wo, s.cout, out.ld, _stream())
                ctx['dec'].append((src1, src2, None, None, out, 0.0, None) if save else None)
            h = out
        return h, ctx

    def backward(self, ctx, d_raw, grads, need_dx=False, wstream=None, early=None, d_hidden=None, hooks=None):
        """d_raw: bf16 Act, gradient wrt the last ConvTranspose2d's output (pre final activation).
        d_hidden: optional bf16 Act, gradient wrt the encoder bottleneck returned by forward(return_hidden=True).
        grads: dict name -> zero-initialised float32 tensor in the reference layout (accumulated into).
        wstream: side stream for the weight-gradient launches (off the data-gradient critical path).
        early = (i, fn[, pull]): fn() is called once the data-gradient of encoder i has been issued.  At that point the
        weight-gradients of every layer but encoder 0 .. i-1 are on wstream and the operand copies of those layers have
        been read for the last time on the current stream.  pull=True also queues encoder i-1's weight-gradient before the
        grouped launch (its dY exists by then; its operand copies are still read by its own data-gradient afterwards) --
        measured slower on one GPU (the heavier group stretches the last one-launch kernel it runs beside), off by default.

        hooks: optional {('after_dec', j): fn, ('before_enc', i): fn}: fn() is called right after decoder j's / right before
        encoder i's data-gradient is issued (the data-parallel trainer places its all-reduces and SM reservations there).

        Every data-gradient convolution also runs the backward of the block that produced its input (activation, dropout,
        InstanceNorm; engine.dgrad_block_bwd), so the chain is one launch per layer."""
        hooks = hooks or {}
        dev = d_raw.t.device
        B = d_raw.B
        dskip = [None] * 7

        ps = self.params()

        def bn_of(s):
            if s.bn is None:
                return None
            gw, gb = grads.get(s.bn + '.weight'), grads.get(s.bn + '.bias')
            return (ps[s.bn + '.weight'], ps[s.bn + '.bias'], gw, gb, bool(ctx.get('training', True)))

        def block_of_dec(j):
            _, _, raw, sums, out, dp, xh = ctx['dec'][j]
            return Block(self.dec[j], raw, sums, out, dp, xh, 16 + j, bn_of(self.dec[j]))

        def block_of_enc(j):
            _, raw, sums, out, dp, xh = ctx['enc'][j]
            return Block(self.enc[j], raw, sums, out, dp, xh, j, bn_of(self.enc[j]))

        self.begin_backward(group=True)
        for i in range(6, -1, -1):
            s = self.dec[i]
            src1, src2 = ctx['dec'][i][0], ctx['dec'][i][1]
            pw = self.packed[7 + i]
            g = grads[s.wname]
            prod = block_of_dec(i - 1) if i >= 1 else block_of_enc(6)
            din = new_act(B, src1.H, src1.W, s.cinp, dev)
            if i == 6 and taps_enabled() and pw.taps_ok:
                # one output channel: tap products (gather dY once, then pointwise GEMMs)
                G6 = taps_gather(L.PG_CONVT, 2, 1, d_raw, 0, B, src1.H, src1.W)
                taps_wgrad(G6, src1.b16, g.data_ptr(), s.c1, wstream, self)
                if src2 is not None:
                    taps_wgrad(G6, src2.b16, g.data_ptr() + s.c1 * 16 * 4, s.c2, wstream, self)
                dd = conv_desc(L.PG_CONV1X1, 1, 0, B, G6.H, G6.W, G6.H, G6.W, 16, 0, 16, 0, s.cinp, din.ld)
                d_raw = dgrad_block_bwd(dd, G6, pw.w16, 16 * 2, din, s.c1p, prod, None, self.seed)
            else:
                # weight gradient: dW[ci][co][tap] = sum x[ci] * dY[co] -- PG_CONV geometry with A = dY, G = layer input
                merged = (self._wg_group is not None and src2 is not None and s.c1 == s.c1p and s.c2 == s.c2p and
                          s.c1p % 64 == 0 and s.c2p % 64 == 0)
                if merged:
                    # both sources of the concat in ONE job: the tap-shifted operand (dY, 16 taps) is read once, not twice
                    wd = conv_desc(L.PG_CONV, 2, 1, B, d_raw.H, d_raw.W, src1.H, src1.W, d_raw.C, 0, d_raw.ld, 0, s.cinp,
                                   s.cinp, out_dt=BF16, in_dt=BF16)
                    self.wgrad(wd, d_raw, src1.b16, g.data_ptr(), s.cout * 16, s.cin, s.cout, wstream, g2=src2.b16,
                               n_split=s.c1p)
                else:
                    wd = conv_desc(L.PG_CONV, 2, 1, B, d_raw.H, d_raw.W, src1.H, src1.W, d_raw.C, 0, d_raw.ld, 0, src1.C,
                                   src1.C, out_dt=BF16, in_dt=BF16)
                    self.wgrad(wd, d_raw, src1.b16, g.data_ptr(), s.cout * 16, s.c1, s.cout, wstream)
                if src2 is not None and not merged:
                    wd2 = conv_desc(L.PG_CONV, 2, 1, B, d_raw.H, d_raw.W, src2.H, src2.W, d_raw.C, 0, d_raw.ld, 0, src2.C,
                                    src2.C, out_dt=BF16, in_dt=BF16)
                    self.wgrad(wd2, d_raw, src2.b16, g.data_ptr() + s.c1 * s.cout * 16 * 4, s.cout * 16, s.c2, s.cout,
                               wstream)
                # data gradient: stride-2 conv of dY with W'[ci][tap][co]
                dd = conv_desc(L.PG_CONV, 2, 1, B, d_raw.H, d_raw.W, src1.H, src1.W, d_raw.C, 0, d_raw.ld, 0, s.cinp,
                               din.ld, c_valid=s.cout)
                d_raw = dgrad_block_bwd(dd, d_raw, pw.bwd, 16 * d_raw.C * 2, din, s.c1p, prod,
                                        d_hidden if i == 0 else None, self.seed)
            if i >= 1:
                dskip[6 - i] = din.slice(s.c1p, s.c2p)
            if ('after_dec', i) in hooks:
                hooks[('after_dec', i)]()
        # d_raw is now the gradient wrt encoder 6's convolution output
        dx = None
        pulled = -1                 # encoder layer whose weight-gradient was queued ahead of its turn (see `early`)

        def enc_wgrad(i, dy):
            s = self.enc[i]
            h, out = ctx['enc'][i][0], ctx['enc'][i][3]
            if h.im2col:
                first_wgrad(h, dy, grads[s.wname].data_ptr(), s.cout, wstream, self)
            else:
                wd = conv_desc(L.PG_CONV, 2, 1, B, h.H, h.W, out.H, out.W, h.C, 0, h.ld, 0, s.np, s.np, out_dt=BF16,
                               in_dt=BF16)
                self.wgrad(wd, h.b16, dy, grads[s.wname].data_ptr(), s.cin * 16, s.cout, s.cin, wstream)

        for i in range(6, -1, -1):
            s = self.enc[i]
            h, out = ctx['enc'][i][0], ctx['enc'][i][3]
            if i != pulled:
                enc_wgrad(i, d_raw)
            if ('before_enc', i) in hooks:
                hooks[('before_enc', i)]()
            if i > 0 or need_dx:
                Hi, Wi = (2 * h.H, 2 * h.W) if h.im2col else (h.H, h.W)
                din = new_act(B, Hi, Wi, s.cinp, dev)
                dd = conv_desc(L.PG_CONVT, 2, 1, B, out.H, out.W, Hi, Wi, d_raw.C, 0, d_raw.ld, 0, s.cinp, din.ld)
                if i > 0:
                    d_raw = dgrad_block_bwd(dd, d_raw, self.packed[i].bwd, 16 * d_raw.C * 2, din, s.cinp, block_of_enc(i - 1),
                                            dskip[i - 1], self.seed)
                else:
                    run_conv(dd, d_raw, None, self.packed[i].bwd, None, din)
                    dx = din
            if early is not None and i == early[0]:
                # d_raw is now dL/d(output of encoder i-1): that layer's weight-gradient joins this group, so that only the
                # layers below it are still open (self.early_late_layers of them)
                if len(early) > 2 and early[2] and i >= 2 and not ctx['enc'][i - 1][0].im2col:
                    enc_wgrad(i - 1, d_raw)
                    pulled = i - 1
                self.flush_wgrads(wstream)      # the weight-gradients of every layer still open: one grouped launch
                early[1]()
        self.flush_wgrads(wstream)              # the rest (everything, without `early`)
        if wstream is None:
            self.finalize_grads()       # (with a side stream the caller joins it first, then calls finalize_grads)
        return dx if need_dx else None


# ------------------------------------------------------------------------------------------------
# Discriminator
# ------------------------------------------------------------------------------------------------

class DiscriminatorEngine(NetEngine):
    def __init__(self, module):
        inc, ndf, nl, norm = module.input_nc, module.ndf, module.n_layers, module.norm
        bn = bool(norm) and bool(getattr(module, 'batchnorm', False))
        idx = 0
        specs = [LayerSpec('conv', 2, inc, 0, ndf, True, 'leakyrelu', False, False, 'model.0.weight', 'model.0.bias')]
        idx += 2
        mult = 1
        for n in range(1, nl):
            prev, mult = mult, min(2 ** n, 8)
            specs.append(LayerSpec('conv', 2, ndf * prev, 0, ndf * mult, False, 'tanh', norm, False,
                                   f'model.{idx}.weight', norm_after_act=True, bn=f'model.{idx + 2}' if bn else None))
            idx += 3 if norm else 2
        prev, mult = mult, min(2 ** nl, 8)
        specs.append(LayerSpec('conv', 1, ndf * prev, 0, ndf * mult, False, 'tanh', norm, False, f'model.{idx}.weight',
                               norm_after_act=True, bn=f'model.{idx + 2}' if bn else None))
        idx += 3 if norm else 2
        specs.append(LayerSpec('conv', 1, ndf * mult, 0, 1, True, 'sigmoid', False, False, f'model.{idx}.weight',
                               f'model.{idx}.bias'))
        super().__init__(module, specs)
        self.in_cp = rup16(inc)

    def bn_tensors(self, s):
        ps, bs = self.params(), dict(self.module.named_buffers())
        return (ps[s.bn + '.weight'], ps[s.bn + '.bias'], bs[s.bn + '.running_mean'], bs[s.bn + '.running_var'],
                bs.get(s.bn + '.num_batches_tracked'))

    def bn_names(self):
        return [s.bn + sfx for s in self.specs if s.bn is not None for sfx in ('.weight', '.bias')]

    def bn_repeat_update(self, ctx, b0, nb):
        """Running-statistics update of a REPEATED forward call over images b0 .. b0+nb-1 (trainer.py:98-99 runs D(fake) a
        second time: identical outputs, but BatchNorm2d in train mode updates its running buffers again).  The fold is
        idempotent on the already folded pairs."""
        if not self.module.training:
            return
        for li, s in enumerate(self.specs):
            if s.bn is None:
                continue
            h, t, sums, out = ctx[li]
            g, b, rm, rv, nbt = self.bn_tensors(s)
            st = _stream()
            L.call('pg_bn_fold_fwd', sums.data_ptr() + b0 * t.C * 8, nb, t.C, t.H * t.W, rm.data_ptr(), rv.data_ptr(), s.cout,
                   BN_MOMENTUM, 1, st)
            if nbt is not None:
                L.call('pg_counter_add', nbt.data_ptr(), 1, st)

    def new_input(self, B, H, W, device, twin=False, zero=True):
        return new_act(B, H, W, self.in_cp, device, dt=Config.fwd_dt, zero=zero, twin=twin)

    def forward(self, xin, save=True):
        """xin: Act (B,H,W,in_cp) bf16 -> (p: f32 Act (B,Ho,Wo,16) with the patch probabilities in channel 0, ctx)."""
        ctx = self.forward_begin(xin, save)
        self.forward_part(ctx, 0, xin.B)
        return ctx[-1][3], (ctx if save else [None] * len(ctx))

    def forward_begin(self, xin, save=True):
        """Allocate every layer's buffers for the whole batch of xin; no launches except the weight pack.
        Returns ctx = [(layer input, conv+act output, norm sums or None, layer output)] per layer."""
        self.ensure_packed()
        dev = xin.t.device
        B = xin.B
        ctx = []
        self._parts = []          # groups of images normalised together (BatchNorm2d): one per forward_part call
        h = xin
        last = len(self.specs) - 1
        for li, s in enumerate(self.specs):
            Ho = (h.H + 2 - 4) // s.stride + 1
            Wo = (h.W + 2 - 4) // s.stride + 1
            if h.im2col:
                Ho, Wo = h.H, h.W
            if Ho < 1 or Wo < 1:
                raise RuntimeError(f'Discriminator: input too small at layer {li} ({h.H}x{h.W})')
            if li == last:
                out = new_act(B, Ho, Wo, 4, dev, dt=F32)                         # one real channel, stride 4
                ctx.append((h, None, None, out))
            else:
                t = new_act(B, Ho, Wo, s.np, dev, dt=h.dt, twin=save)
                if s.norm:
                    sums = zeros((B, s.np, 2), dev)
                    out = new_act(B, Ho, Wo, s.np, dev, dt=h.dt, twin=save)
                    ctx.append((h, t, sums, out))
                else:
                    out = t
                    ctx.append((h, t, None, out))
            h = out
        return ctx

    def forward_part(self, ctx, b0, nb):
        """Run the layers for images b0 .. b0+nb-1 on the current stream (samples are independent: the only
        normalisation is per-sample InstanceNorm, disc.py:8)."""
        ps = self.params()
        last = len(self.specs) - 1
        self._parts.append((b0, nb))
        for li, s in enumerate(self.specs):
            h, t, sums, out = ctx[li]
            h, out = h.images(b0, nb), out.images(b0, nb)
            bias = ps[s.bname].detach() if s.bias else None
            if li == last and taps_enabled() and self.packed[li].taps_ok:
                taps_forward(L.PG_CONV, s.stride, 1, h, None, self.packed[li].fwd.data_ptr(), bias, L.ACT[s.act], out, 0)
            elif li == last:
                run_conv(conv_desc(L.PG_CONV, s.stride, 1, nb, h.H, h.W, out.H, out.W, h.C, 0, h.ld, 0, s.np, out.ld,
                                   n_valid=s.cout, act=L.ACT[s.act], out_dt=F32, has_bias=1, in_dt=h.dt), h, None,
                         self.packed[li].fwd, bias, out)
            else:
                t = t.images(b0, nb)
                if h.im2col:
                    first_conv(h, self.packed[li].wfirst, bias if s.bias else None, L.ACT[s.act], t, s.cout)
                else:
                    run_conv(conv_desc(L.PG_CONV, s.stride, 1, nb, h.H, h.W, t.H, t.W, h.C, 0, h.ld, 0, s.np, t.ld,
                                       n_valid=s.cout, act=L.ACT[s.act], out_dt=t.dt, has_bias=int(s.bias), in_dt=h.dt),
                             h, None, self.packed[li].fwd, bias, t)
                if s.norm and s.bn is not None:
                    # BatchNorm2d after the Tanh (disc.py:29-32): statistics over the images of THIS call
                    instnorm_stats(t, sums, b0)
                    gam, bet, rm, rv, nbt = self.bn_tensors(s)
                    st = _stream()
                    training = self.module.training
                    L.call('pg_bn_fold_fwd', sums.data_ptr() + b0 * t.C * 8, nb, t.C, t.H * t.W, rm.data_ptr(), rv.data_ptr(),
                           s.cout, BN_MOMENTUM, 1 if training else 0, st)
                    if training and nbt is not None:
                        L.call('pg_counter_add', nbt.data_ptr(), 1, st)
                    L.call('pg_norm_affine_act_fwd', t.ptr, t.dt, sums.data_ptr() + b0 * t.C * 8, gam.data_ptr(), bet.data_ptr(),
                           s.cout, out.ptr, out.dt, out.twptr, nb, t.H * t.W, t.C, t.ld, out.ld, 0, 0.0, None, 0, st)
                elif s.norm:
                    instnorm_stats(t, sums, b0)
                    norm_fwd(t, sums, out, 0, 0.0, None, 0, b0)

    def backward(self, ctx, d_raw, grads, need_dx, nb=None, wstream=None, dx_channels=None):
        """d_raw: bf16 Act (nb,Ho,Wo,16): gradient wrt the last conv's pre-sigmoid output.
        grads: dict of zero-initialised fp32 tensors to accumulate into, or None to skip weight gradients.
        nb: process only the first nb images of the saved batch.
        dx_channels = (first, count): the caller reads only these channels of the returned input gradient."""
        dev = d_raw.t.device
        B = d_raw.B if nb is None else nb
        din = None
        if grads is not None:
            self.begin_backward()
        for li in range(len(self.specs) - 1, -1, -1):
            s = self.specs[li]
            h, t, sums, out = ctx[li]
            h = h.first(B)
            pw = self.packed[li]
            last_taps = li == len(self.specs) - 1 and li > 0 and taps_enabled() and pw.taps_ok
            Gt = taps_gather(L.PG_CONV, s.stride, 1, d_raw, 0, B, h.H, h.W) if last_taps else None
            if grads is not None:
                if last_taps:
                    taps_wgrad(Gt, h.b16, grads[s.wname].data_ptr(), s.cin, wstream)
                elif h.im2col:
                    first_wgrad(h, d_raw, grads[s.wname].data_ptr(), s.cout, wstream)
                else:
                    wd = conv_desc(L.PG_CONV, s.stride, 1, B, h.H, h.W, d_raw.H, d_raw.W, h.C, 0, h.ld, 0, s.np, s.np,
                                   out_dt=BF16, in_dt=BF16)
                    self.wgrad(wd, h.b16, d_raw, grads[s.wname].data_ptr(), s.cin * 16, s.cout, s.cin, wstream)
                if s.bias:
                    with torch.cuda.stream(pick_wstream(wstream) if wstream is not None else torch.cuda.current_stream()):
                        L.call('pg_colsum', d_raw.ptr, B * d_raw.H * d_raw.W, d_raw.ld, s.cout,
                               grads[s.bname].data_ptr(), _stream())
            if li > 0 or need_dx:
                Hi, Wi = (2 * h.H, 2 * h.W) if h.im2col else (h.H, h.W)
                din = new_act(B, Hi, Wi, s.cinp, dev)
                nf, nv = 0, None
                if li == 0 and dx_channels is not None:
                    nf, nv = dx_channels[0], dx_channels[0] + dx_channels[1]
                # the activation backward of the previous layer (no norm in between) is fused into this data-gradient
                fuse_act = li > 0 and not self.specs[li - 1].norm and taps_enabled()
                yprev = ctx[li - 1][1].first(B) if fuse_act else None
                aprev = L.ACT[self.specs[li - 1].act] if fuse_act else 0
                if last_taps:
                    taps_dgrad(Gt, pw.w16, din, yprev, aprev)
                elif (li == 0 and dx_channels is not None and dx_channels[1] == 1 and s.stride == 2 and taps_enabled()
                        and s.cout == s.np):
                    # only the generated-mask channel of the input gradient is read: tap products of dY with the 16 x Cout
                    # slab W'[c = mask channel] of the data-gradient operand, scattered as a ConvTranspose2d.  The result
                    # gets a TRIMMED pixel stride (the channels up to the mask's, rounded to 4: 8 bytes per pixel instead
                    # of 32), which is 4x fewer sectors for the scatter here and for the reader (pg_gen_out_bwd)
                    din = new_act(B, Hi, Wi, (nv + 3) // 4 * 4, dev)
                    wslab = pw.bwd.data_ptr() + dx_channels[0] * 16 * s.np * 2
                    taps_forward(L.PG_CONVT, 2, 1, d_raw, None, wslab, None, 0, din, dx_channels[0])
                elif s.stride == 2:
                    dd = conv_desc(L.PG_CONVT, 2, 1, B, d_raw.H, d_raw.W, Hi, Wi, d_raw.C, 0, d_raw.ld, 0, s.cinp,
                                   din.ld, n_valid=nv, n_first=nf, c_valid=s.cout, act=aprev)
                    if fuse_act:
                        run_conv_dgrad_act(dd, d_raw, pw.bwd, din, yprev)
                    else:
                        run_conv(dd, d_raw, None, pw.bwd, None, din)
                else:
                    dd = conv_desc(L.PG_CONV, 1, 2, B, d_raw.H, d_raw.W, h.H, h.W, d_raw.C, 0, d_raw.ld, 0, s.cinp,
                                   din.ld, n_valid=nv, n_first=nf, c_valid=s.cout, act=aprev)
                    if fuse_act:
                        run_conv_dgrad_act(dd, d_raw, pw.bwd, din, yprev)
                    else:
                        run_conv(dd, d_raw, None, pw.bwd, None, din)
            if li > 0:
                ps = self.specs[li - 1]
                _, pt, psums, pout = ctx[li - 1]
                pt = pt.first(B)
                if ps.norm and ps.bn is not None:
                    pp = self.params()
                    gw = grads.get(ps.bn + '.weight') if grads is not None else None
                    gb = grads.get(ps.bn + '.bias') if grads is not None else None
                    parts = [(b0, nb_) for (b0, nb_) in getattr(self, '_parts', []) if b0 + nb_ <= B] or [(0, B)]
                    dt = batchnorm_bwd(pt, psums, (pp[ps.bn + '.weight'], pp[ps.bn + '.bias'], gw, gb, self.module.training),
                                       ps.cout, din, None, 0, 0.0, None, 0, parts)
                    d_raw = act_bwd_out(pt, dt, L.ACT[ps.act])
                elif ps.norm:
                    dt = norm_bwd(pt, psums, din, None, 0, 0.0, None, 0)
                    d_raw = act_bwd_out(pt, dt, L.ACT[ps.act])
                elif fuse_act:
                    d_raw = din                    # already multiplied by act'(previous output) in the epilogue
                else:
                    d_raw = act_bwd_out(pt, din, L.ACT[ps.act])
        if grads is not None and wstream is None:
            self.finalize_grads()
        return din if need_dx else None
