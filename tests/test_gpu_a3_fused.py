"""GPU parity of the one-launch conv + InstanceNorm (+ activation, dropout) kernels (pg_conv_norm_fwd,
pg_conv_dgrad_norm_bwd: accumulators resident in tensor memory across a grid barrier) against the numpy oracle's
conv -> instance_norm -> activation chain (unet.py:19-28, 53-66) and its autograd backward.

Operands are rounded to the 16-bit storage type before the oracle sees them, so what is left is fp32 accumulation order
and the 16-bit rounding of the result: 5e-3 norm-wise forward, 8e-3 backward (same bounds as the separate kernels in
tests/test_gpu_a_ops.py::test_instance_norm_act_forward_backward)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from patchgan_b200 import _lib as L
from patchgan_b200.engine import TORCH_DT, conv_desc, rup16
from tests.gpu_util import bf16_round, from_nhwc, pack_weight, relerr, stream, to_nhwc

pytestmark = pytest.mark.gpu


def rng(seed=0):
    return np.random.default_rng(seed)


def fused(kind, act, **kw):
    fn = L.FusedNorm()
    fn.kind, fn.act = kind, L.ACT[act]
    for k, v in kw.items():
        setattr(fn, k, v)
    return fn


def supported(d, fn, twin):
    return bool(L.lib().pg_conv_norm_supported(ctypes.byref(d), ctypes.byref(fn), int(twin)))


# (mode, B, C1, C2, Cout, H, W, act)   H, W = INPUT size of the convolution
FWD_CASES = [
    ('conv', 2, 32, 0, 64, 32, 32, 'leakyrelu'),      # 16x16 map: one image per tile
    ('conv', 3, 64, 0, 64, 8, 8, 'relu'),             # 4x4 map: eight images per tile, segmented statistics
    ('conv', 5, 64, 0, 32, 4, 4, 'tanh'),             # 2x2 bottleneck
    ('conv', 2, 16, 0, 32, 2, 4, 'leakyrelu'),        # rectangular 1x2 bottleneck
    ('conv', 2, 16, 0, 48, 24, 40, 'leakyrelu'),      # 12x20 map: clipped tiles, N = 48
    ('conv', 16, 16, 0, 32, 256, 256, 'leakyrelu'),   # enc0-sized: 2048 tiles, 14 TMEM slots per CTA
    ('convT', 2, 64, 64, 32, 8, 8, 'leakyrelu'),      # virtual concat, four parity classes
    ('convT', 2, 32, 32, 48, 2, 2, 'relu'),           # 2x2 lattice
    ('convT', 4, 32, 32, 32, 64, 64, 'tanh'),         # dec5-sized tiles (128x128 output)
    ('conv1x1', 2, 48, 0, 32, 64, 64, 'leakyrelu'),   # first layer as an im2col product, K = 48
]


@pytest.mark.parametrize('dt', [L.DT_F16, L.DT_BF16], ids=['f16', 'bf16'])
@pytest.mark.parametrize('case', FWD_CASES, ids=[str(c) for c in FWD_CASES])
def test_conv_instance_norm_act_forward(case, dt):
    mode, B, C1, C2, Co, H, W, act = case
    r = rng(11)
    x1 = bf16_round(r.standard_normal((B, C1, H, W)), dt)
    x2 = bf16_round(r.standard_normal((B, C2, H, W)), dt) if C2 else None
    Ci = C1 + C2
    Cop = rup16(Co)
    if mode == 'conv':
        w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Ci * 16), dt)
        raw = orc.conv2d_fwd(x1, w, None, 2)
        Ho, Wo = raw.shape[2], raw.shape[3]
        wd = pack_weight(w, Co, Cop, C1, C1, 0, 0, Ci * 16, 16, dt=dt)
        d = conv_desc(L.PG_CONV, 2, 1, B, H, W, Ho, Wo, C1, 0, C1, 0, Cop, Cop, n_valid=Co, out_dt=dt, in_dt=dt)
    elif mode == 'convT':
        w = bf16_round(r.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Ci * 4), dt)
        raw = orc.convT_fwd(np.concatenate([x1, x2], axis=1), w)
        Ho, Wo = 2 * H, 2 * W
        wd = pack_weight(w, Co, Cop, C1, C1, C2, C2, 16, Co * 16, dt=dt)
        d = conv_desc(L.PG_CONVT, 2, 1, B, H, W, Ho, Wo, C1, C2, C1, C2, Cop, Cop, n_valid=Co, out_dt=dt, in_dt=dt)
    else:
        w = bf16_round(r.standard_normal((Co, Ci)) / np.sqrt(Ci), dt)
        raw = np.einsum('bchw,nc->bnhw', x1, w).astype(np.float32)
        Ho, Wo = H, W
        wd = torch.zeros((Cop, Ci), dtype=TORCH_DT[dt], device='cuda')
        wd[:Co] = torch.from_numpy(w).cuda().to(TORCH_DT[dt])
        d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, W, H, W, C1, 0, C1, 0, Cop, Cop, n_valid=Co, out_dt=dt, in_dt=dt)
    xhat, _ = orc.instnorm_fwd(raw)
    ref = orc.act_fwd(act, xhat)
    x1d = to_nhwc(x1, dt=dt)
    x2d = to_nhwc(x2, dt=dt) if C2 else None
    out = torch.full((B, Ho, Wo, Cop), 7.0, device='cuda', dtype=TORCH_DT[dt])
    twin = torch.full((B, Ho, Wo, Cop), 7.0, device='cuda', dtype=torch.bfloat16)
    xh = torch.full((B, Ho, Wo, Cop), 7.0, device='cuda', dtype=TORCH_DT[dt])
    sums = torch.zeros((B, Cop, 2), device='cuda')
    sync = torch.zeros(4, device='cuda', dtype=torch.int32)
    kw = {}
    if dt == L.DT_BF16:      # with a scratch buffer the small-map cases take the split-K route, without it they split N
        ws = torch.empty(8 << 20, device='cuda', dtype=torch.uint8)
        kw = dict(ws=ws.data_ptr(), ws_bytes=ws.numel())
    fn = fused(L.FUSED_FWD, act, sums=sums.data_ptr(), sync=sync.data_ptr(), xhat=xh.data_ptr(), xhat_ld=Cop, **kw)
    assert supported(d, fn, True)
    L.call('pg_conv_norm_fwd', ctypes.byref(d), x1d.data_ptr(), x2d.data_ptr() if C2 else None, wd.data_ptr(),
           out.data_ptr(), twin.data_ptr(), ctypes.byref(fn), stream())
    torch.cuda.synchronize()
    assert relerr(from_nhwc(out, Co), ref) < 5e-3
    assert relerr(from_nhwc(xh, Co), xhat) < 5e-3
    # the bf16 twin is rounded from the same fp32 values (not from the 16-bit primary)
    assert relerr(from_nhwc(twin, Co), ref) < 5e-3
    if dt == L.DT_BF16:
        assert torch.equal(twin, out)
    # the sums the backward needs: (sum, sum of squares) of the convolution output
    got = sums.cpu().numpy()[:, :Co]
    assert np.allclose(got[..., 0], raw.sum(axis=(2, 3)), rtol=2e-3, atol=2e-2 * np.sqrt(Ho * Wo))
    assert np.allclose(got[..., 1], (raw.astype(np.float64) ** 2).sum(axis=(2, 3)), rtol=2e-3, atol=1e-3)
    if Cop > Co:
        assert float(out[..., Co:].float().abs().max()) == 0.0


def test_fused_forward_refuses_what_does_not_fit_tensor_memory():
    """cfg 5's first generator layer (4 x 512 x 512 x 64 outputs = 67 M accumulators) cannot stay resident: the query says
    so and the call fails loudly instead of computing something else."""
    d = conv_desc(L.PG_CONV, 2, 1, 4, 1024, 1024, 512, 512, 16, 0, 16, 0, 64, 64, out_dt=L.DT_F16, in_dt=L.DT_F16)
    fn = fused(L.FUSED_FWD, 'leakyrelu')
    assert not supported(d, fn, True)
    d1 = conv_desc(L.PG_CONV, 2, 1, 2, 2, 2, 1, 1, 16, 0, 16, 0, 64, 64, out_dt=L.DT_F16, in_dt=L.DT_F16)   # 1x1 map
    assert not supported(d1, fn, False)


def test_fused_forward_dropout_matches_the_separate_kernel_mask():
    """Dropout in the fused epilogue draws uniform(mix(seed, salt), pixel*C + channel) exactly like pg_norm_act_fwd, so
    either forward can be paired with either backward."""
    B, C, H, W, Co = 2, 32, 16, 16, 32
    r = rng(12)
    x = bf16_round(r.standard_normal((B, C, H, W)), L.DT_F16)
    w = bf16_round(r.standard_normal((Co, C, 4, 4)) / np.sqrt(C * 16), L.DT_F16)
    raw = orc.conv2d_fwd(x, w, None, 2)
    xhat, _ = orc.instnorm_fwd(raw)
    ref = orc.act_fwd('relu', xhat)
    Ho = H // 2
    d = conv_desc(L.PG_CONV, 2, 1, B, H, W, Ho, Ho, C, 0, C, 0, Co, Co, out_dt=L.DT_F16, in_dt=L.DT_F16)
    seed = torch.tensor([777], device='cuda', dtype=torch.int64)
    out = torch.empty((B, Ho, Ho, Co), device='cuda', dtype=torch.float16)
    sums = torch.zeros((B, Co, 2), device='cuda')
    sync = torch.zeros(4, device='cuda', dtype=torch.int32)
    fn = fused(L.FUSED_FWD, 'relu', sums=sums.data_ptr(), sync=sync.data_ptr(), drop_p=0.2, seed=seed.data_ptr(), salt=5)
    wd = pack_weight(w, Co, Co, C, C, 0, 0, C * 16, 16, dt=L.DT_F16)
    L.call('pg_conv_norm_fwd', ctypes.byref(d), to_nhwc(x, dt=L.DT_F16).data_ptr(), None, wd.data_ptr(), out.data_ptr(),
           None, ctypes.byref(fn), stream())
    # the same mask from the separate kernel: ones through act = none
    ones = torch.ones((B, Ho, Ho, Co), device='cuda', dtype=torch.float32)
    mask = torch.empty((B, Ho, Ho, Co), device='cuda', dtype=torch.float16)
    L.call('pg_norm_act_fwd', ones.data_ptr(), 1, None, mask.data_ptr(), 2, None, B, Ho * Ho, Co, Co, Co, 0, 0.2,
           seed.data_ptr(), 5, stream())
    torch.cuda.synchronize()
    m = from_nhwc(mask, Co)
    assert abs((m != 0).mean() - 0.8) < 0.03
    assert relerr(from_nhwc(out, Co), ref * m) < 5e-3


# (form, B, Cin, Cskip, Cout, H, W, act, dskip, via)  -- the LAYER whose data-gradient is taken: 'conv' = Conv2d(Cin -> Cout, s2)
# on an H x W input (data-gradient in PG_CONVT form), 'convT' = ConvTranspose2d(Cin + Cskip -> Cout) on an H x W input
# (data-gradient in PG_CONV form; only the first Cin channels of its input come from the normalised block).
# via: what the block's forward saved -- 'y' (output, invertible activation) or 'xhat'.
@pytest.fixture(params=[False, True], ids=['nsplit', 'ksplit'])
def scratch(request):
    return torch.empty(8 << 20, device='cuda', dtype=torch.uint8) if request.param else None


BWD_CASES = [
    ('conv', 2, 32, 0, 64, 32, 32, 'leakyrelu', True, 'y'),
    ('conv', 3, 64, 0, 64, 8, 8, 'relu', True, 'xhat'),
    ('conv', 5, 32, 0, 64, 4, 4, 'tanh', False, 'xhat'),
    ('conv', 2, 32, 0, 16, 2, 4, 'leakyrelu', True, 'y'),
    ('conv', 2, 48, 0, 32, 24, 40, 'leakyrelu', True, 'y'),
    ('convT', 2, 64, 64, 32, 8, 8, 'leakyrelu', False, 'y'),
    ('convT', 3, 32, 32, 16, 4, 4, 'relu', False, 'xhat'),
    ('convT', 2, 64, 0, 64, 2, 2, 'leakyrelu', False, 'y'),
    ('conv', 16, 32, 0, 64, 128, 128, 'leakyrelu', True, 'y'),     # enc0-sized gradient: 14 TMEM slots per CTA
]


@pytest.mark.parametrize('case', BWD_CASES, ids=[str(c) for c in BWD_CASES])
def test_conv_dgrad_instance_norm_backward(case, scratch):
    form, B, Ci, Cs, Co, H, W, act, with_skip, via = case
    fdt = L.DT_F16
    r = rng(13)
    raw = (r.standard_normal((B, Ci, H, W)) * 1.5 + 0.3).astype(np.float32)       # conv output of the normalised block
    xhat, rstd = orc.instnorm_fwd(raw)
    y = orc.act_fwd(act, xhat)
    # what the forward stored (16-bit); the oracle differentiates through the stored values
    y16, xh16 = bf16_round(y, fdt), bf16_round(xhat, fdt)
    Cit = Ci + Cs
    if form == 'conv':
        w = bf16_round(r.standard_normal((Co, Cit, 4, 4)) / np.sqrt(Co * 16))
        Ho, Wo = H // 2, W // 2
        dy = bf16_round(r.standard_normal((B, Co, Ho, Wo)))
        din_ref, _, _ = orc.conv2d_bwd(np.zeros((B, Cit, H, W), np.float32), w, dy, 2)
        wd = pack_weight(w, Cit, Cit, Co, rup16(Co), 0, 0, 16, Cit * 16)
        d = conv_desc(L.PG_CONVT, 2, 1, B, Ho, Wo, H, W, rup16(Co), 0, rup16(Co), 0, Cit, Cit, out_dt=L.DT_BF16)
    else:
        w = bf16_round(r.standard_normal((Cit, Co, 4, 4)) / np.sqrt(Co * 16))
        dy = bf16_round(r.standard_normal((B, Co, 2 * H, 2 * W)))
        din_ref, _ = orc.convT_bwd(np.zeros((B, Cit, H, W), np.float32), w, dy)
        wd = pack_weight(w, Cit, Cit, Co, rup16(Co), 0, 0, Co * 16, 16)
        d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * W, H, W, rup16(Co), 0, rup16(Co), 0, Cit, Cit, out_dt=L.DT_BF16)
    dskip = bf16_round(r.standard_normal((B, Ci, H, W))) if with_skip else None
    g = din_ref[:, :Ci] + (dskip if with_skip else 0)
    xh_used = xh16 if via == 'xhat' else np.where((y16 < 0) & (act == 'leakyrelu'), 5 * y16, y16)
    y_used = orc.act_fwd(act, xh_used)
    ref = orc.instnorm_bwd(orc.act_bwd(act, xh_used, y_used, g), xh_used, rstd)
    # forward sums (what pg_conv_norm_fwd leaves): sum and sum of squares of raw
    sums = torch.from_numpy(np.stack([raw.sum(axis=(2, 3)), (raw.astype(np.float64) ** 2).sum(axis=(2, 3))],
                                     axis=-1).astype(np.float32)).cuda()
    bsums = torch.zeros((B, Ci, 2), device='cuda')
    sync = torch.zeros(4, device='cuda', dtype=torch.int32)
    saved = to_nhwc(xh16 if via == 'xhat' else y16, dt=fdt)
    kw = dict(n_norm=Ci, sums=sums.data_ptr(), bsums=bsums.data_ptr(), sync=sync.data_ptr(), y_dtype=fdt)
    if via == 'xhat':
        kw.update(xhat=saved.data_ptr(), xhat_ld=Ci)
    else:
        kw.update(y=saved.data_ptr(), y_ld=Ci)
    dsk = to_nhwc(dskip) if with_skip else None
    if with_skip:
        kw.update(dskip=dsk.data_ptr(), dskip_ld=Ci)
    if scratch is not None:
        kw.update(ws=scratch.data_ptr(), ws_bytes=scratch.numel())
    fn = fused(L.FUSED_BWD, act, **kw)
    assert supported(d, fn, False)
    dx = torch.full((B, H, W, Cit), 7.0, device='cuda', dtype=torch.bfloat16)
    L.call('pg_conv_dgrad_norm_bwd', ctypes.byref(d), to_nhwc(dy).data_ptr(), wd.data_ptr(), dx.data_ptr(),
           ctypes.byref(fn), stream())
    torch.cuda.synchronize()
    got = from_nhwc(dx, Cit)
    assert relerr(got[:, :Ci], ref) < 8e-3
    if Cs:
        assert relerr(got[:, Ci:], din_ref[:, Ci:]) < 5e-3          # the skip half is the plain data-gradient


def test_separate_backward_kernels_accept_xhat_or_output_instead_of_raw():
    """pg_norm_act_bwd with the PG_X_IS_XHAT / PG_X_IS_OUTPUT flags (the tensors the fused forward leaves) gives what
    the raw-input form gives."""
    B, C, H, W = 2, 32, 16, 16
    r = rng(14)
    x = (r.standard_normal((B, C, H, W)) * 2 + 0.5).astype(np.float32)
    dy = bf16_round(r.standard_normal((B, C, H, W)))
    xhat, rstd = orc.instnorm_fwd(x)
    st = stream()
    xd = to_nhwc(x, C, f32=True)
    sums = torch.zeros((B, C, 2), device='cuda')
    L.call('pg_instnorm_stats', xd.data_ptr(), 1, B, H * W, C, C, sums.data_ptr(), st)
    d1 = to_nhwc(dy, C)
    for act, flag, saved in (('relu', 0x100, xhat), ('leakyrelu', 0x200, orc.act_fwd('leakyrelu', xhat)),
                             ('tanh', 0x100, xhat)):
        ref = orc.instnorm_bwd(orc.act_bwd(act, xhat, orc.act_fwd(act, xhat), dy), xhat, rstd)
        sd = to_nhwc(saved, C, f32=True)
        bs = torch.zeros((B, C, 2), device='cuda')
        dx = torch.empty((B, H, W, C), device='cuda', dtype=torch.bfloat16)
        L.call('pg_norm_act_bwd', sd.data_ptr(), 1 | flag, sums.data_ptr(), d1.data_ptr(), C, None, 0, bs.data_ptr(),
               dx.data_ptr(), C, B, H * W, C, C, L.ACT[act], 0.0, None, 0, st)
        torch.cuda.synchronize()
        assert relerr(from_nhwc(dx, C), ref) < 8e-3, act
