"""GPU parity of the CTA-pair (tcgen05.mma.cta_group::2) variant of the persistent convolution kernel against the numpy oracle.

The variant is chosen for wide layers with an even number (>= 148) of 128-pixel tiles: the discriminator's layers at the
benchmark batch sizes.  Shapes here are the smallest that qualify (the oracle finishes them in seconds); every test asserts
that the pair kernel actually ran (pg_pair_launch_count).  Inputs and weights are rounded to the 16-bit operand type before
the oracle sees them, so the tolerance covers fp32 accumulation order and the 16-bit output rounding only."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc, rup16
from tests.gpu_util import bf16_round, from_nhwc, pack_weight, relerr, stream, to_nhwc
from tests.test_gpu_a_ops import run_conv

pytestmark = pytest.mark.gpu
TC = L.IMPL_TCGEN05


def pair_count():
    return L.lib().pg_pair_launch_count()


@pytest.fixture(autouse=True)
def every_eligible_shape():
    """The default takes the pair kernel only where it measured faster; these tests take it wherever it is legal."""
    L.lib().pg_set_pair_mode(2)
    yield
    L.lib().pg_set_pair_mode(-1)


# (B, Cin, Cout, H, stride): tile width 128 and 256, 64-/128-/256-channel inputs, the stride-1 31x31 layer (masked tile rows)
@pytest.mark.parametrize('dt', [L.DT_BF16, L.DT_F16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('case', [(5, 64, 128, 128, 2), (5, 64, 256, 128, 2), (10, 128, 512, 32, 1), (20, 128, 128, 64, 2),
                                  (6, 32, 384, 128, 2)], ids=str)
def test_conv2d_forward_pair(case, dt):
    B, Ci, Co, H, s = case
    r = np.random.default_rng(11)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), dt)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Ci * 16), dt)
    b = r.standard_normal(Co).astype(np.float32)
    ref = orc.act_fwd('leakyrelu', orc.conv2d_fwd(x, w, b, s))
    Ho = ref.shape[2]
    xd = to_nhwc(x, dt=dt)
    wd = pack_weight(w, Co, Co, Ci, Ci, 0, 0, Ci * 16, 16, dt=dt)
    bd = torch.from_numpy(b).cuda()
    d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, n_valid=Co, act=L.ACT['leakyrelu'],
                  out_dt=dt, has_bias=1, in_dt=dt)
    n0 = pair_count()
    out = run_conv(d, xd, None, wd, bd, dt, TC, Ho, Ho, Co)
    assert pair_count() == n0 + 1, 'the CTA-pair kernel was not chosen for this shape'
    assert relerr(from_nhwc(out, Co), ref) < 5e-3


# (B, C1, C2, Cout, H): ConvTranspose2d = 4 parity classes, virtual concat of two sources
@pytest.mark.parametrize('case', [(6, 128, 0, 128, 32), (6, 64, 64, 128, 32), (3, 64, 0, 256, 64)], ids=str)
def test_conv_transpose_forward_pair(case):
    B, C1, C2, Co, H = case
    dt = L.DT_F16
    r = np.random.default_rng(12)
    x1 = bf16_round(r.standard_normal((B, C1, H, H)), dt)
    x2 = bf16_round(r.standard_normal((B, C2, H, H)), dt) if C2 else None
    Ci = C1 + C2
    w = bf16_round(r.standard_normal((Ci, Co, 4, 4)) / np.sqrt(Ci * 4), dt)
    xin = x1 if x2 is None else np.concatenate([x1, x2], axis=1)
    ref = orc.act_fwd('tanh', orc.convT_fwd(xin, w))
    wd = pack_weight(w, Co, Co, C1, C1, C2, C2, 16, Co * 16, dt=dt)
    d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, 2 * H, 2 * H, C1, C2, C1, C2, Co, Co, n_valid=Co, act=L.ACT['tanh'],
                  out_dt=dt, in_dt=dt)
    n0 = pair_count()
    out = run_conv(d, to_nhwc(x1, dt=dt), to_nhwc(x2, dt=dt) if C2 else None, wd, None, dt, TC, 2 * H, 2 * H, Co)
    assert pair_count() == n0 + 1, 'the CTA-pair kernel was not chosen for this shape'
    assert relerr(from_nhwc(out, Co), ref) < 5e-3


@pytest.mark.parametrize('case', [(5, 128, 128, 128, 2), (20, 128, 256, 32, 1)], ids=str)
def test_data_gradient_with_activation_backward_pair(case):
    """dX = dgrad(dY) * act'(saved output), bf16 operands, 16-bit output + bf16 twin: the discriminator's backward chain."""
    B, Ci, Co, H, s = case
    r = np.random.default_rng(13)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Co * 16))
    Ho = (H + 2 - 4) // s + 1
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
    y_prev = bf16_round(r.standard_normal((B, Ci, H, H)), L.DT_F16)           # saved OUTPUT of the previous LeakyReLU
    ref, _, _ = orc.conv2d_bwd(x, w, dy, s)
    ref = ref * np.where(y_prev > 0, 1.0, 0.2).astype(np.float32)
    wd = pack_weight(w, Ci, Ci, Co, Co, 0, 0, 16, Ci * 16, flip=1 if s == 1 else 0)
    if s == 2:
        d = conv_desc(L.PG_CONVT, 2, 1, B, Ho, Ho, H, H, Co, 0, Co, 0, Ci, Ci, out_dt=L.DT_BF16, act=L.ACT['leakyrelu'])
    else:
        d = conv_desc(L.PG_CONV, 1, 2, B, Ho, Ho, H, H, Co, 0, Co, 0, Ci, Ci, out_dt=L.DT_BF16, act=L.ACT['leakyrelu'])
    yd = to_nhwc(y_prev, dt=L.DT_F16)
    out = torch.full((B, H, H, Ci), 7.0, device='cuda', dtype=torch.bfloat16)
    n0 = pair_count()
    L.call('pg_conv_dgrad_act', ctypes.byref(d), to_nhwc(dy).data_ptr(), wd.data_ptr(), out.data_ptr(), yd.data_ptr(), Ci,
           L.DT_F16, TC, stream())
    torch.cuda.synchronize()
    assert pair_count() == n0 + 1, 'the CTA-pair kernel was not chosen for this shape'
    assert relerr(from_nhwc(out, Ci), ref) < 6e-3


@pytest.mark.parametrize('tap_major', [0, 1], ids=['direct', 'tapmajor'])
@pytest.mark.parametrize('case', [(2, 256, 512, 16, 1), (4, 128, 256, 32, 2), (3, 256, 256, 25, 1)], ids=str)
def test_weight_gradient_pair(case, tap_major):
    """wgrad_tc_pair_kernel: two n-tiles per cluster, each CTA loads half of every activation tile (N % 256 == 0, C >= 128)."""
    B, Ci, Co, H, s = case
    r = np.random.default_rng(14)
    x = bf16_round(r.standard_normal((B, Ci, H, H)))
    Ho = (H + 2 - 4) // s + 1
    dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
    _, ref, _ = orc.conv2d_bwd(x, np.zeros((Co, Ci, 4, 4), np.float32), dy, s, has_bias=True, need_dx=False)
    d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    xd, dyd = to_nhwc(x), to_nhwc(dy)
    n0 = pair_count()
    if tap_major:
        S = torch.zeros((16, Co, Ci), device='cuda')
        L.call('pg_conv_wgrad_tapmajor', ctypes.byref(d), xd.data_ptr(), dyd.data_ptr(), Co, S.data_ptr(), Co, Ci, TC, stream())
        torch.cuda.synchronize()
        got = S.cpu().numpy().reshape(4, 4, Co, Ci).transpose(2, 3, 0, 1)
    else:
        dw = torch.zeros((Co, Ci, 4, 4), device='cuda')
        L.call('pg_conv_wgrad', ctypes.byref(d), xd.data_ptr(), dyd.data_ptr(), Co, dw.data_ptr(), Ci * 16, Co, Ci, TC, stream())
        torch.cuda.synchronize()
        got = dw.cpu().numpy()
    assert pair_count() == n0 + 1, 'the CTA-pair weight-gradient kernel was not chosen for this shape'
    assert relerr(got, ref) < 1e-4
