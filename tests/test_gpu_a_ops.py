"""GPU parity of the individual kernels against the numpy oracle, called through the C-ABI (ctypes).

Convolution inputs and weights are rounded to bf16 before the oracle sees them, so the only difference left is
fp32 accumulation order: tolerance 2e-3 norm-wise (bf16 output rounding is 2^-9 = 2e-3 per element, ~1e-3 rms).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc, rup16
from tests.gpu_util import act_of, bf16_round, from_nhwc, pack_weight, relerr, stream, to_nhwc

pytestmark = pytest.mark.gpu
IMPLS = [L.IMPL_SIMT, L.IMPL_TCGEN05]
IMPL_IDS = ['simt', 'tcgen05']
TOL = 3e-3


def rng(seed=0):
    return np.random.default_rng(seed)


def run_conv(desc, src1, src2, w, bias, out_dt, impl, Hout, Wout, N):
    from patchgan_b200.engine import TORCH_DT
    B = desc.B
    out_dt = L.DT_F32 if out_dt is True else (L.DT_BF16 if out_dt is False else out_dt)
    out = torch.full((B, Hout, Wout, desc.ldo), 7.0, device='cuda', dtype=TORCH_DT[out_dt])
    L.call('pg_conv_fwd', ctypes.byref(desc), src1.data_ptr(), src2.data_ptr() if src2 is not None else None,
           w.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(), None, impl, stream())
    torch.cuda.synchronize()
    return out


# (B, Cin, Cout, H, stride) -- covers BK=64/32/16, channel padding (3, 4, 1, 7), odd sizes 32->31->30, tiny maps
CONV_CASES = [
    (2, 3, 32, 64, 2), (2, 32, 64, 32, 2), (1, 64, 128, 16, 2), (2, 128, 256, 8, 2), (3, 256, 256, 4, 2),
    (2, 4, 64, 64, 2), (2, 256, 512, 32, 1), (2, 512, 1, 31, 1), (1, 16, 48, 20, 2), (2, 10, 16, 34, 1),
    (5, 64, 64, 2, 2),
]


DTS = [L.DT_BF16, L.DT_F16]
DT_IDS = ['bf16', 'f16']


@pytest.mark.parametrize('dt', DTS, ids=DT_IDS)
@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('case', CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_conv2d_forward(case, impl, dt):
    B, Ci, Co, H, s = case
    r = rng(1)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), dt)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Ci * 16), dt)
    b = r.standard_normal(Co).astype(np.float32)
    ref = orc.act_fwd('leakyrelu', orc.conv2d_fwd(x, w, b, s))
    Ho = ref.shape[2]
    Cip, Cop = rup16(Ci), rup16(Co)
    xd = to_nhwc(x, dt=dt)
    wd = pack_weight(w, Co, Cop, Ci, Cip, 0, 0, Ci * 16, 16, dt=dt)
    bd = torch.zeros(Cop, device='cuda')
    bd[:Co] = torch.from_numpy(b).cuda()
    d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Cip, 0, Cip, 0, Cop, Cop, n_valid=Co, act=L.ACT['leakyrelu'],
                  out_dt=L.DT_F32, has_bias=1, in_dt=dt)
    out = run_conv(d, xd, None, wd, bd, True, impl, Ho, Ho, Cop)
    assert relerr(from_nhwc(out, Co), ref) < TOL
    if Cop > Co:
        assert float(out[..., Co:].abs().max()) == 0.0


CONVT_CASES = [(2, 256, 0, 256, 2), (2, 256, 256, 256, 4), (1, 64, 64, 32, 16), (2, 32, 32, 1, 32), (2, 16, 16, 7, 16),
               (3, 128, 128, 64, 8), (2, 64, 0, 64, 5)]


@pytest.mark.parametrize('dt', DTS, ids=DT_IDS)
@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('case', CONVT_CASES, ids=[str(c) for c in CONVT_CASES])
def test_conv_transpose_forward_with_virtual_concat(case, impl, dt):
    B, C1, C2, Co, H = case
    r = rng(2)
    x1 = bf16_round(r.standard_normal((B, C1, H, H)), dt)
    x2 = bf16_round(r.standard_normal((B, C2, H, H)), dt) if C2 else None
    Ci = C1 + C2
    w = bf16_round(r.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Ci * 4), dt)
    xin = x1 if x2 is None else np.concatenate([x1, x2], axis=1)
    ref = orc.act_fwd('sigmoid', orc.convT_fwd(xin, w))
    Cop = rup16(Co)
    wd = pack_weight(w, Co, Cop, C1, rup16(C1), C2, rup16(C2) if C2 else 0, 16, Co * 16, dt=dt)
    x1d = to_nhwc(x1, dt=dt)
    x2d = to_nhwc(x2, dt=dt) if C2 else None
    d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, 2 * H, 2 * H, rup16(C1), rup16(C2) if C2 else 0, rup16(C1),
                  rup16(C2) if C2 else 0, Cop, Cop, n_valid=Co, act=L.ACT['sigmoid'], out_dt=dt, in_dt=dt)
    out = run_conv(d, x1d, x2d, wd, None, dt, impl, 2 * H, 2 * H, Cop)
    assert relerr(from_nhwc(out, Co), ref) < 5e-3     # 16-bit output
    if Cop > Co:
        assert float(out[..., Co:].float().abs().max()) == 0.0


DGRAD_CASES = [(2, 32, 64, 32, 2), (2, 256, 512, 32, 1), (2, 512, 1, 31, 1), (2, 4, 64, 64, 2), (1, 64, 64, 6, 1)]


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('case', DGRAD_CASES, ids=[str(c) for c in DGRAD_CASES])
def test_conv2d_data_gradient(case, impl):
    """dgrad(Conv2d s=2) as PG_CONVT, dgrad(Conv2d s=1) as flipped stride-1 pad-2 PG_CONV."""
    B, Ci, Co, H, s = case
    r = rng(3)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Co * 16))
    Ho = (H + 2 - 4) // s + 1
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
    ref, _, _ = orc.conv2d_bwd(x, w, dy, s)
    Cip, Cop = rup16(Ci), rup16(Co)
    wd = pack_weight(w, Ci, Cip, Co, Cop, 0, 0, 16, Ci * 16, flip=1 if s == 1 else 0)
    dyd = to_nhwc(dy)
    if s == 2:
        d = conv_desc(L.PG_CONVT, 2, 1, B, Ho, Ho, H, H, Cop, 0, Cop, 0, Cip, Cip, out_dt=L.DT_F32)
    else:
        d = conv_desc(L.PG_CONV, 1, 2, B, Ho, Ho, H, H, Cop, 0, Cop, 0, Cip, Cip, out_dt=L.DT_F32)
    out = run_conv(d, dyd, None, wd, None, True, impl, H, H, Cip)
    assert relerr(from_nhwc(out, Ci), ref) < TOL


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
def test_conv_transpose_data_gradient(impl):
    B, Ci, Co, H = 2, 96, 32, 8
    r = rng(4)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    w = bf16_round(r.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Co * 16))
    dy = bf16_round(r.standard_normal((B, Co, 2 * H, 2 * H)))
    ref, _ = orc.convT_bwd(x, w, dy)
    Cip, Cop = rup16(Ci), rup16(Co)
    wd = pack_weight(w, Ci, Cip, Co, Cop, 0, 0, Co * 16, 16)
    d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * H, H, H, Cop, 0, Cop, 0, Cip, Cip, out_dt=L.DT_F32)
    out = run_conv(d, to_nhwc(dy), None, wd, None, True, impl, H, H, Cip)
    assert relerr(from_nhwc(out, Ci), ref) < TOL


WGRAD_CASES = [(2, 3, 32, 64, 2), (2, 32, 64, 32, 2), (2, 256, 512, 16, 1), (2, 512, 1, 15, 1), (3, 64, 64, 4, 2),
               (2, 128, 256, 32, 2), (1, 64, 128, 31, 1), (2, 48, 96, 20, 2)]


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('dt', DTS, ids=DT_IDS)
@pytest.mark.parametrize('case', WGRAD_CASES, ids=[str(c) for c in WGRAD_CASES])
def test_conv2d_weight_gradient(case, dt, impl):
    B, Ci, Co, H, s = case
    r = rng(5)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), dt)
    w = r.standard_normal((Co, Ci, 4, 4)).astype(np.float32)
    Ho = (H + 2 - 4) // s + 1
    gdt = dt if impl == L.IMPL_TCGEN05 else L.DT_BF16        # tcgen05 kind::f16 needs one operand format
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)), gdt)
    _, ref, refb = orc.conv2d_bwd(x, w, dy, s, has_bias=True, need_dx=False)
    Cip, Cop = rup16(Ci), rup16(Co)
    dw = torch.zeros((Co, Ci, 4, 4), device='cuda')
    d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Cip, 0, Cip, 0, Cop, Cop, out_dt=gdt, in_dt=dt)
    xd, dyd = to_nhwc(x, dt=dt), to_nhwc(dy, dt=gdt)
    L.call('pg_conv_wgrad', ctypes.byref(d), xd.data_ptr(), dyd.data_ptr(), Cop, dw.data_ptr(), Ci * 16, Co, Ci,
           impl, stream())
    torch.cuda.synchronize()
    assert relerr(dw.cpu().numpy(), ref) < 1e-4
    if gdt == L.DT_BF16:
        db = torch.zeros(Co, device='cuda')
        L.call('pg_colsum', dyd.data_ptr(), B * Ho * Ho, Cop, Co, db.data_ptr(), stream())
        torch.cuda.synchronize()
        assert relerr(db.cpu().numpy(), refb) < 1e-4


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
def test_conv_transpose_weight_gradient_two_sources(impl):
    B, C1, C2, Co, H = 2, 32, 48, 16, 8
    r = rng(6)
    xdt = L.DT_BF16 if impl == L.IMPL_TCGEN05 else L.DT_F16
    x1 = bf16_round(r.standard_normal((B, C1, H, H)), xdt)
    x2 = bf16_round(r.standard_normal((B, C2, H, H)), xdt)
    w = r.standard_normal((C1 + C2, Co, 4, 4)).astype(np.float32)
    dy = bf16_round(r.standard_normal((B, Co, 2 * H, 2 * H)))
    _, ref = orc.convT_bwd(np.concatenate([x1, x2], axis=1), w, dy, need_dx=False)
    dw = torch.zeros((C1 + C2, Co, 4, 4), device='cuda')
    dyd = to_nhwc(dy)
    for (xs, C, off) in ((x1, C1, 0), (x2, C2, C1)):
        xd = to_nhwc(xs, dt=xdt)
        d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * H, H, H, rup16(Co), 0, rup16(Co), 0, rup16(C), rup16(C),
                      out_dt=xdt, in_dt=L.DT_BF16)
        L.call('pg_conv_wgrad', ctypes.byref(d), dyd.data_ptr(), xd.data_ptr(), rup16(C),
               dw.data_ptr() + off * Co * 16 * 4, Co * 16, C, Co, impl, stream())
    torch.cuda.synchronize()
    assert relerr(dw.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize('act', ['leakyrelu', 'relu', 'tanh'])
@pytest.mark.parametrize('shape', [(2, 32, 16, 16), (3, 64, 2, 2), (1, 16, 31, 31), (2, 24, 8, 8)])
def test_instance_norm_act_forward_backward(shape, act):
    B, C, H, W = shape
    r = rng(7)
    x = (r.standard_normal(shape) * 2 + 0.5).astype(np.float32)
    dy1 = bf16_round(r.standard_normal(shape))
    dy2 = bf16_round(r.standard_normal(shape))
    xhat, rstd = orc.instnorm_fwd(x)
    y = orc.act_fwd(act, xhat)
    dx_ref = orc.instnorm_bwd(orc.act_bwd(act, xhat, y, dy1 + dy2), xhat, rstd)
    Cp = (C + 7) // 8 * 8
    xd = to_nhwc(x, Cp, f32=True)
    sums = torch.zeros((B, Cp, 2), device='cuda')
    st = stream()
    L.call('pg_instnorm_stats', xd.data_ptr(), 1, B, H * W, Cp, Cp, sums.data_ptr(), st)
    yd = torch.empty((B, H, W, Cp), device='cuda', dtype=torch.bfloat16)
    y2 = torch.empty_like(yd)
    L.call('pg_norm_act_fwd', xd.data_ptr(), 1, sums.data_ptr(), yd.data_ptr(), 0, y2.data_ptr(), B, H * W, Cp, Cp, Cp,
           L.ACT[act], 0.0, None, 0, st)
    d1, d2 = to_nhwc(dy1, Cp), to_nhwc(dy2, Cp)
    bs = torch.zeros((B, Cp, 2), device='cuda')
    dx = torch.empty((B, H, W, Cp), device='cuda', dtype=torch.bfloat16)
    L.call('pg_norm_act_bwd_reduce', xd.data_ptr(), 1, sums.data_ptr(), d1.data_ptr(), Cp, d2.data_ptr(), Cp,
           bs.data_ptr(), B, H * W, Cp, Cp, L.ACT[act], 0.0, None, 0, st)
    L.call('pg_norm_act_bwd_apply', xd.data_ptr(), 1, sums.data_ptr(), d1.data_ptr(), Cp, d2.data_ptr(), Cp,
           bs.data_ptr(), dx.data_ptr(), Cp, B, H * W, Cp, Cp, L.ACT[act], 0.0, None, 0, st)
    # the one-call entry point (single launch for maps of up to 1024 pixels, else the two passes above)
    bs2 = torch.zeros((B, Cp, 2), device='cuda')
    dx2 = torch.full((B, H, W, Cp), 7.0, device='cuda', dtype=torch.bfloat16)
    L.call('pg_norm_act_bwd', xd.data_ptr(), 1, sums.data_ptr(), d1.data_ptr(), Cp, d2.data_ptr(), Cp, bs2.data_ptr(),
           dx2.data_ptr(), Cp, B, H * W, Cp, Cp, L.ACT[act], 0.0, None, 0, st)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(yd, C), y) < 5e-3
    assert torch.equal(yd, y2)                      # bf16 twin
    assert relerr(from_nhwc(dx, C), dx_ref) < 8e-3
    assert relerr(from_nhwc(dx2, C), dx_ref) < 8e-3


def test_dropout_statistics_and_backward_mask_reuse():
    """RNG streams cannot match torch's Philox: check rate 0.2, scale 1/0.8, and that backward regenerates the
    same mask from (seed, salt, index)."""
    B, C, H, W = 2, 32, 64, 64
    x = torch.ones((B, H, W, C), device='cuda', dtype=torch.float32)
    seed = torch.tensor([12345], device='cuda', dtype=torch.int64)
    y = torch.empty((B, H, W, C), device='cuda', dtype=torch.bfloat16)
    st = stream()
    L.call('pg_norm_act_fwd', x.data_ptr(), 1, None, y.data_ptr(), 0, None, B, H * W, C, C, C, 0, 0.2, seed.data_ptr(), 3,
           st)
    dy = torch.ones((B, H, W, C), device='cuda', dtype=torch.bfloat16)
    dx = torch.empty_like(dy)
    L.call('pg_norm_act_bwd_apply', x.data_ptr(), 1, None, dy.data_ptr(), C, None, 0, None, dx.data_ptr(), C, B, H * W, C,
           C, 0, 0.2, seed.data_ptr(), 3, st)
    y2 = torch.empty_like(y)
    L.call('pg_counter_add', seed.data_ptr(), 1, st)
    L.call('pg_norm_act_fwd', x.data_ptr(), 1, None, y2.data_ptr(), 0, None, B, H * W, C, C, C, 0, 0.2, seed.data_ptr(), 3,
           st)
    torch.cuda.synchronize()
    yf = y.float()
    keep = (yf != 0).float().mean().item()
    assert abs(keep - 0.8) < 0.01
    assert torch.allclose(yf[yf != 0], torch.tensor(1.25, device='cuda'))
    assert torch.equal(dx, y)                 # same mask, same scale in backward
    assert not torch.equal(y2, y)             # a new seed draws a new mask


@pytest.mark.parametrize('loss_type', ['tversky', 'weighted_bce', 'MAE'])
@pytest.mark.parametrize('C', [1, 3])
def test_segmentation_losses_and_gradients(loss_type, C):
    B, H, W = 3, 32, 32
    r = rng(8)
    p = (0.02 + 0.96 * r.random((B, C, H, W))).astype(np.float32)
    t = (r.random((B, C, H, W)) > 0.6).astype(np.float32)
    tr = orc.Trainer(None, None) if False else None
    o = orc.Trainer.__new__(orc.Trainer)
    o.loss_type, o.seg_alpha, o.tversky_beta, o.tversky_gamma = loss_type, 200, 0.75, 0.75
    ref_val, ref_grad = o.seg_loss(t, p)
    pd = to_nhwc(p, 16, f32=True)
    td = torch.from_numpy(t).cuda()
    st = stream()
    lt = L.LOSS[loss_type]
    part = torch.zeros((B, 8), device='cuda')
    coef = torch.zeros((B, 4), device='cuda')
    losses = torch.zeros(8, device='cuda')
    chsum = torch.zeros((B, C), device='cuda')
    L.call('pg_target_chsum', td.data_ptr(), chsum.data_ptr(), B, C, H * W, st)
    L.call('pg_seg_loss_partials', pd.data_ptr(), 16, td.data_ptr(), chsum.data_ptr(), part.data_ptr(), B, C, H * W, lt,
           st)
    L.call('pg_seg_loss_finalize', part.data_ptr(), coef.data_ptr(), losses.data_ptr(), 0, B, C, H * W, lt, 0.75, 0.75,
           200.0, st)
    dx = torch.empty((B, H, W, 16), device='cuda', dtype=torch.bfloat16)
    L.call('pg_gen_out_bwd', pd.data_ptr(), 16, td.data_ptr(), chsum.data_ptr(), coef.data_ptr(), None, 0, 0,
           dx.data_ptr(), 16, B, C, H * W, lt, 0, 0.75, st)
    torch.cuda.synchronize()
    assert abs(losses[0].item() - ref_val) <= 1e-4 * abs(ref_val)
    assert relerr(from_nhwc(dx, C), ref_grad) < 5e-3


def test_bce_against_constant_labels():
    r = rng(9)
    n = 2 * 30 * 30
    p = (0.01 + 0.98 * r.random(n)).astype(np.float32)
    p[:4] = [0.0, 1.0, 1e-30, 1 - 1e-7]       # exercise the -100 log clamp
    pd = torch.zeros((n, 16), device='cuda')
    pd[:, 0] = torch.from_numpy(p).cuda()
    for label in (0.0, 1.0):
        t = np.full(n, label, dtype=np.float32)
        ref = orc.bce_fwd(p, t)
        ref_g = orc.bce_bwd(p, t, gout=0.5) * (p.astype(np.float64) * (1 - p.astype(np.float64)))
        losses = torch.zeros(8, device='cuda')
        dz = torch.full((n, 16), 9.0, device='cuda', dtype=torch.bfloat16)
        L.call('pg_bce_const', pd.data_ptr(), 16, label, 0.5, losses.data_ptr(), 2, dz.data_ptr(), 16, n, stream())
        torch.cuda.synchronize()
        assert abs(losses[2].item() - ref) <= 1e-4 * abs(ref)
        assert relerr(dz[:, 0].float().cpu().numpy(), ref_g) < 5e-3
        assert float(dz[:, 1:].float().abs().max()) == 0.0


def test_fused_adam_matches_oracle():
    r = rng(10)
    n = 10007
    p0 = r.standard_normal(n).astype(np.float32)
    params = {'w': p0.copy()}
    opt = orc.Adam(params, 1e-3)
    pd = torch.zeros(n + 1, device='cuda')[:n]
    pd.copy_(torch.from_numpy(p0))
    pbuf = torch.from_numpy(p0).cuda()
    m = torch.zeros(n, device='cuda')
    v = torch.zeros(n, device='cuda')
    hyper = torch.tensor([1e-3, 0, 0, 0], device='cuda')
    step = torch.zeros(1, device='cuda', dtype=torch.int32)
    for it in range(3):
        g = (r.standard_normal(n) * (10.0 ** r.integers(-6, 2))).astype(np.float32)
        opt.step({'w': g})
        gd = torch.from_numpy(g).cuda()
        L.call('pg_adam_step', pbuf.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(),
               step.data_ptr(), 0.9, 0.999, 1e-8, 1.0, stream())
    torch.cuda.synchronize()
    assert int(step.item()) == 3
    np.testing.assert_allclose(pbuf.cpu().numpy(), params['w'], rtol=0, atol=2e-6)


def test_layout_roundtrip_and_softmax():
    r = rng(11)
    x = r.standard_normal((2, 5, 9, 7)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    buf = torch.zeros((2, 9, 7, 16), device='cuda', dtype=torch.bfloat16)
    L.call('pg_pack_nchw_f32_to_nhwc_bf16', xd.data_ptr(), buf.data_ptr(), 2, 5, 9, 7, 16, 3, L.DT_BF16, stream())
    back = torch.empty((2, 5, 9, 7), device='cuda')
    L.call('pg_unpack_nhwc_to_nchw_f32', buf.data_ptr(), 0, back.data_ptr(), 2, 5, 9, 7, 16, 3, stream())
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), bf16_round(x))          # bit-exact byte movement
    assert float(buf[..., :3].float().abs().max()) == 0.0 and float(buf[..., 8:].float().abs().max()) == 0.0
    xs = to_nhwc(x, 16, f32=True)
    L.call('pg_softmax_fwd', xs.data_ptr(), xs.data_ptr(), 2 * 9 * 7, 5, 16, stream())
    torch.cuda.synchronize()
    assert relerr(from_nhwc(xs, 5), orc.act_fwd('softmax', x)) < 1e-5


def test_inference_tiling_bit_exact():
    r = rng(12)
    for (Hh, Ww) in ((300, 300), (260, 260)):
        img = r.random((3, Hh, Ww), dtype=np.float32)
        size, overlap = 128, 0.9
        eff = int(overlap * size)
        ncy, ncx = int(np.ceil(Hh / eff)), int(np.ceil(Ww / eff))
        ref = orc.n_crop(img, size, overlap)
        imd = torch.from_numpy(img).cuda()
        crops = torch.empty((ncx * ncy, 3, size, size), device='cuda')
        L.call('pg_ncrop', imd.data_ptr(), crops.data_ptr(), 3, Hh, Ww, size, eff, ncy, ncx, stream())
        torch.cuda.synchronize()
        assert np.array_equal(crops.cpu().numpy(), ref)
        masks = r.random((ncx * ncy, 4, size, size), dtype=np.float32)
        md = torch.from_numpy(masks).cuda()
        mo = torch.empty((4, Hh, Ww), device='cuda')
        am = torch.empty((Hh, Ww), device='cuda', dtype=torch.int32)
        L.call('pg_build_mask', md.data_ptr(), mo.data_ptr(), am.data_ptr(), 4, Hh, Ww, size, eff, ncy, ncx, 0.0, stream())
        torch.cuda.synchronize()
        assert np.array_equal(am.cpu().numpy(), orc.build_mask(masks, size, (Hh, Ww), 0.0, overlap))
        m1 = md[:, :1].contiguous()
        L.call('pg_build_mask', m1.data_ptr(), mo.data_ptr(), None, 1, Hh, Ww, size, eff, ncy, ncx, 0.5, stream())
        torch.cuda.synchronize()
        assert np.array_equal(mo[0].cpu().numpy(), orc.build_mask(masks[:, :1], size, (Hh, Ww), 0.5, overlap)
                              .astype(np.float32))


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('case', [(2, 32, 64, 32, 2), (2, 256, 512, 16, 1), (3, 64, 64, 4, 2), (2, 128, 256, 32, 2),
                                  (1, 64, 128, 31, 1), (2, 48, 96, 20, 2), (2, 16, 7, 12, 2),
                                  (8, 128, 256, 32, 2), (3, 256, 512, 24, 1), (4, 128, 384, 20, 1)], ids=str)
def test_conv2d_weight_gradient_tap_major_and_finalize(case, impl):
    """pg_conv_wgrad_tapmajor (TMA bulk-reduce epilogue on the tcgen05 path) + pg_grad_finalize_multi == reference layout."""
    from patchgan_b200.engine import NetEngine
    B, Ci, Co, H, s = case
    r = rng(7)
    x = bf16_round(r.standard_normal((B, Ci, H, H)))
    w = r.standard_normal((Co, Ci, 4, 4)).astype(np.float32)
    Ho = (H + 2 - 4) // s + 1
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
    _, ref, _ = orc.conv2d_bwd(x, w, dy, s, has_bias=True, need_dx=False)
    Cip, Cop = rup16(Ci), rup16(Co)
    Cs = (Ci + 3) // 4 * 4
    S = torch.zeros((16, Co, Cs), device='cuda')
    d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Cip, 0, Cip, 0, Cop, Cop, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    xd, dyd = to_nhwc(x), to_nhwc(dy)
    for _ in range(2):       # two accumulating calls: the scratch must hold the sum
        L.call('pg_conv_wgrad_tapmajor', ctypes.byref(d), xd.data_ptr(), dyd.data_ptr(), Cop, S.data_ptr(), Co, Cs, impl, stream())
    dw = torch.full((Co, Ci, 4, 4), 7.0, device='cuda')
    job = np.array([(S.data_ptr(), dw.data_ptr(), Ci * 16, Co, Ci, Co, Cs, 0, (Ci + 31) // 32)], dtype=NetEngine.GRAD_JOB_DT)
    table = torch.from_numpy(job.view(np.uint8).copy()).cuda()
    L.call('pg_grad_finalize_multi', table.data_ptr(), 1, Co * ((Ci + 31) // 32), stream())
    torch.cuda.synchronize()
    assert relerr(dw.cpu().numpy(), 2 * ref) < 1e-4


@pytest.mark.parametrize('impl', IMPLS, ids=IMPL_IDS)
@pytest.mark.parametrize('act', ['tanh', 'leakyrelu'])
@pytest.mark.parametrize('case', [(2, 32, 64, 32, 2), (2, 256, 512, 32, 1), (3, 64, 128, 16, 2)], ids=str)
def test_data_gradient_with_fused_activation_backward(case, act, impl):
    """pg_conv_dgrad_act: dx = dgrad(dy) * act'(y) with y the saved activation output in front of the convolution."""
    B, Ci, Co, H, s = case
    r = rng(8)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Co * 16))
    Ho = (H + 2 - 4) // s + 1
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
    dxc, _, _ = orc.conv2d_bwd(x, w, dy, s)
    y = bf16_round(orc.act_fwd(act, r.standard_normal((B, Ci, H, H)).astype(np.float32)), L.DT_F16)
    ref = dxc * orc.act_bwd_from_output(act, y) if hasattr(orc, 'act_bwd_from_output') else None
    if ref is None:
        g = {'tanh': 1 - y * y, 'leakyrelu': np.where(y > 0, 1.0, 0.2)}[act]
        ref = dxc * g
    wd = pack_weight(w, Ci, Ci, Co, Co, 0, 0, 16, Ci * 16, flip=1 if s == 1 else 0)
    if s == 2:
        d = conv_desc(L.PG_CONVT, 2, 1, B, Ho, Ho, H, H, Co, 0, Co, 0, Ci, Ci, out_dt=L.DT_BF16, act=L.ACT[act])
    else:
        d = conv_desc(L.PG_CONV, 1, 2, B, Ho, Ho, H, H, Co, 0, Co, 0, Ci, Ci, out_dt=L.DT_BF16, act=L.ACT[act])
    out = torch.full((B, H, H, Ci), 7.0, device='cuda', dtype=torch.bfloat16)
    yd = to_nhwc(y, dt=L.DT_F16)
    L.call('pg_conv_dgrad_act', ctypes.byref(d), to_nhwc(dy).data_ptr(), wd.data_ptr(), out.data_ptr(), yd.data_ptr(), Ci, L.DT_F16,
           impl, stream())
    torch.cuda.synchronize()
    assert relerr(from_nhwc(out, Ci), ref) < 6e-3       # bf16 output (+ one extra bf16 rounding on the unfused path)
