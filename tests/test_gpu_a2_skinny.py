"""GPU parity of the streaming ("skinny") kernels for layers with one real channel on one side -- the generator's
output ConvTranspose2d (unet.py:106-107), the discriminator's last Conv2d (disc.py:45), the mask-channel data-gradient
of the discriminator's first layer (trainer.py:84-89) -- against the numpy oracle, through the C-ABI with
impl = PG_IMPL_SKINNY (PG_IMPL_AUTO picks the same kernels; the tcgen05 / SIMT paths are covered in test_gpu_a_ops)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from patchgan_b200 import _lib as L
from patchgan_b200.engine import TORCH_DT, conv_desc, rup16
from tests.gpu_util import bf16_round, from_nhwc, pack_weight, relerr, stream, to_nhwc

pytestmark = pytest.mark.gpu
SK = L.IMPL_SKINNY


def rng(seed=0):
    return np.random.default_rng(seed)


def run_conv(desc, src1, src2, w, bias, out_dt, Hout, Wout, impl=SK):
    out = torch.full((desc.B, Hout, Wout, desc.ldo), 7.0, device='cuda', dtype=TORCH_DT[out_dt])
    L.call('pg_conv_fwd', ctypes.byref(desc), src1.data_ptr(), src2.data_ptr() if src2 is not None else None,
           w.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(), None, impl, stream())
    torch.cuda.synchronize()
    assert L.lib().pg_last_conv_impl() == SK
    return out


@pytest.mark.parametrize('dt', [L.DT_BF16, L.DT_F16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('case', [(2, 512, 31, 4), (3, 64, 6, 16), (1, 256, 17, 4)], ids=str)
def test_last_conv_one_output_channel(case, dt):
    """Conv2d(C -> 1, k4, s1, p1) + bias + sigmoid, trimmed (ldo 4) and full (ldo 16) output rows."""
    B, Ci, H, ldo = case
    r = rng(11)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), dt)
    w = bf16_round(r.standard_normal((1, Ci, 4, 4)) / np.sqrt(Ci * 16), dt)
    b = r.standard_normal(1).astype(np.float32)
    ref = orc.act_fwd('sigmoid', orc.conv2d_fwd(x, w, b, 1))
    Ho = ref.shape[2]
    wd = pack_weight(w, 1, 16, Ci, Ci, 0, 0, Ci * 16, 16, dt=dt)
    bd = torch.from_numpy(b).cuda()
    d = conv_desc(L.PG_CONV, 1, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, 16, ldo, n_valid=1, act=L.ACT['sigmoid'],
                  out_dt=L.DT_F32, has_bias=1, in_dt=dt)
    out = run_conv(d, to_nhwc(x, dt=dt), None, wd, bd, L.DT_F32, Ho, Ho)
    assert relerr(from_nhwc(out, 1), ref) < 1e-4


@pytest.mark.parametrize('dt', [L.DT_BF16, L.DT_F16], ids=['bf16', 'f16'])
@pytest.mark.parametrize('case', [(2, 32, 32, 32), (1, 64, 0, 6), (3, 16, 48, 10)], ids=str)
def test_output_conv_transpose_one_channel_virtual_concat(case, dt):
    B, C1, C2, H = case
    r = rng(12)
    x1 = bf16_round(r.standard_normal((B, C1, H, H)), dt)
    x2 = bf16_round(r.standard_normal((B, C2, H, H)), dt) if C2 else None
    Ci = C1 + C2
    w = bf16_round(r.standard_normal((Ci, 1, 4, 4)) / np.sqrt(Ci * 4), dt)
    xin = x1 if x2 is None else np.concatenate([x1, x2], axis=1)
    ref = orc.act_fwd('sigmoid', orc.convT_fwd(xin, w))
    wd = pack_weight(w, 1, 16, C1, C1, C2, C2, 16, 16, dt=dt)
    d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, 2 * H, 2 * H, C1, C2, C1, C2, 16, 4, n_valid=1, act=L.ACT['sigmoid'],
                  out_dt=L.DT_F32, in_dt=dt)
    out = run_conv(d, to_nhwc(x1, dt=dt), to_nhwc(x2, dt=dt) if C2 else None, wd, None, L.DT_F32, 2 * H, 2 * H)
    assert relerr(from_nhwc(out, 1), ref) < 1e-4


def test_first_conv_data_gradient_of_one_input_channel():
    """dgrad(Conv2d(4 -> 64, s2)) restricted to input channel 3 (the generated mask): PG_CONVT with n_first = 3."""
    B, Ci, Co, H = 2, 4, 64, 32
    r = rng(13)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    w = bf16_round(r.standard_normal((Co, Ci, 4, 4)) / np.sqrt(Co * 16))
    dy = bf16_round(r.standard_normal((B, Co, H // 2, H // 2)))
    ref, _, _ = orc.conv2d_bwd(x, w, dy, 2)
    wd = pack_weight(w, Ci, 16, Co, Co, 0, 0, 16, Ci * 16)
    d = conv_desc(L.PG_CONVT, 2, 1, B, H // 2, H // 2, H, H, Co, 0, Co, 0, 16, 16, n_valid=4, out_dt=L.DT_BF16, n_first=3)
    out = run_conv(d, to_nhwc(dy), None, wd, None, L.DT_BF16, H, H)
    got = out[..., 3].float().cpu().numpy()
    assert relerr(got, ref[:, 3]) < 5e-3      # bf16 output
    assert float((out[..., 0].float() - 7.0).abs().max()) == 0.0     # channels below n_first are left alone


@pytest.mark.parametrize('case', [(2, 512, 31, 1), (2, 64, 12, 2), (1, 96, 9, 2)], ids=str)
def test_data_gradient_from_one_channel(case):
    """dgrad of a 1-output-channel layer: Conv2d s1 (flipped stride-1 pad-2 PG_CONV) and ConvTranspose2d (stride-2 PG_CONV)."""
    B, Ci, H, s = case
    r = rng(14)
    x = r.standard_normal((B, Ci, H, H)).astype(np.float32)
    if s == 1:
        w = bf16_round(r.standard_normal((1, Ci, 4, 4)) / 4)
        dy = bf16_round(r.standard_normal((B, 1, H - 1, H - 1)))
        ref, _, _ = orc.conv2d_bwd(x, w, dy, 1)
        wd = pack_weight(w, Ci, Ci, 1, 16, 0, 0, 16, Ci * 16, flip=1)
        d = conv_desc(L.PG_CONV, 1, 2, B, H - 1, H - 1, H, H, 16, 0, 16, 0, Ci, Ci, out_dt=L.DT_BF16, c_valid=1)
    else:
        w = bf16_round(r.standard_normal((Ci, 1, 4, 4)) / 4)
        dy = bf16_round(r.standard_normal((B, 1, 2 * H, 2 * H)))
        ref, _ = orc.convT_bwd(x, w, dy)
        wd = pack_weight(w, Ci, Ci, 1, 16, 0, 0, 16, 16)
        d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * H, H, H, 16, 0, 16, 0, Ci, Ci, out_dt=L.DT_BF16, c_valid=1)
    out = run_conv(d, to_nhwc(dy), None, wd, None, L.DT_BF16, H, H)
    assert relerr(from_nhwc(out, Ci), ref) < 5e-3


@pytest.mark.parametrize('case', [(2, 512, 15), (3, 64, 7), (1, 40, 30)], ids=str)
def test_weight_gradient_one_output_channel(case):
    B, Ci, H = case
    r = rng(15)
    x = bf16_round(r.standard_normal((B, Ci, H, H)))
    w = r.standard_normal((1, Ci, 4, 4)).astype(np.float32)
    dy = bf16_round(r.standard_normal((B, 1, H - 1, H - 1)))
    _, ref, _ = orc.conv2d_bwd(x, w, dy, 1, has_bias=True, need_dx=False)
    Cip = rup16(Ci)
    dw = torch.zeros((1, Ci, 4, 4), device='cuda')
    d = conv_desc(L.PG_CONV, 1, 1, B, H, H, H - 1, H - 1, Cip, 0, Cip, 0, 16, 16, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    L.call('pg_conv_wgrad', ctypes.byref(d), to_nhwc(x).data_ptr(), to_nhwc(dy).data_ptr(), 16, dw.data_ptr(), Ci * 16, 1, Ci,
           SK, stream())
    torch.cuda.synchronize()
    assert L.lib().pg_last_conv_impl() == SK
    assert relerr(dw.cpu().numpy(), ref) < 1e-4


def test_weight_gradient_conv_transpose_one_output_channel_two_sources():
    B, C1, C2, H = 2, 32, 48, 8
    r = rng(16)
    x1 = bf16_round(r.standard_normal((B, C1, H, H)))
    x2 = bf16_round(r.standard_normal((B, C2, H, H)))
    w = r.standard_normal((C1 + C2, 1, 4, 4)).astype(np.float32)
    dy = bf16_round(r.standard_normal((B, 1, 2 * H, 2 * H)))
    _, ref = orc.convT_bwd(np.concatenate([x1, x2], axis=1), w, dy, need_dx=False)
    dw = torch.zeros((C1 + C2, 1, 4, 4), device='cuda')
    dyd = to_nhwc(dy)
    for (xs, C, off) in ((x1, C1, 0), (x2, C2, C1)):
        xd = to_nhwc(xs)
        d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * H, H, H, 16, 0, 16, 0, rup16(C), rup16(C), out_dt=L.DT_BF16,
                      in_dt=L.DT_BF16)
        L.call('pg_conv_wgrad', ctypes.byref(d), dyd.data_ptr(), xd.data_ptr(), rup16(C), dw.data_ptr() + off * 16 * 4, 16, C, 1,
               SK, stream())
        assert L.lib().pg_last_conv_impl() == SK
    torch.cuda.synchronize()
    assert relerr(dw.cpu().numpy(), ref) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# the same layers as pointwise tap products on the tensor cores (PG_CONV1X1 + pg_taps_scatter / pg_taps_gather):
# this is what the engine runs
# ---------------------------------------------------------------------------------------------------------------
from patchgan_b200 import engine as E
from tests.gpu_util import act_of


@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TCGEN05], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', [(2, 64, 24, 48), (1, 16, 9, 16), (3, 512, 31, 16), (2, 32, 130, 16)], ids=str)
def test_pointwise_conv(case, impl):
    """PG_CONV1X1: out[pix][n] = sum_c in[pix][c] W[n][c]."""
    B, C, H, N = case
    r = rng(21)
    x = bf16_round(r.standard_normal((B, C, H, H)), L.DT_F16)
    w = bf16_round(r.standard_normal((N, C)) / np.sqrt(C), L.DT_F16)
    ref = np.einsum('bchw,nc->bnhw', x.astype(np.float64), w.astype(np.float64))
    wd = torch.from_numpy(w).cuda().half().contiguous()
    d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, C, 0, C, 0, N, N, out_dt=L.DT_F32, in_dt=L.DT_F16)
    out = torch.full((B, H, H, N), 7.0, device='cuda')
    L.call('pg_conv_fwd', ctypes.byref(d), to_nhwc(x, dt=L.DT_F16).data_ptr(), None, wd.data_ptr(), None, out.data_ptr(), None,
           impl, stream())
    torch.cuda.synchronize()
    assert relerr(from_nhwc(out, N), ref) < 1e-4


@pytest.mark.parametrize('impl', [L.IMPL_SIMT, L.IMPL_TCGEN05], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', [(2, 64, 24), (3, 512, 31), (1, 32, 70)], ids=str)
def test_pointwise_weight_gradient_tap_layout(case, impl):
    """PG_CONV1X1 wgrad with ld_n = 1, ldw = 16: dw[c*16 + tap] += sum_q x[q][c] G[q][tap]."""
    B, C, H = case
    r = rng(22)
    x = bf16_round(r.standard_normal((B, C, H, H)))
    g = bf16_round(r.standard_normal((B, 16, H, H)))
    ref = np.einsum('bchw,bthw->ct', x.astype(np.float64), g.astype(np.float64))
    dw = torch.zeros((C, 16), device='cuda')
    d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, C, 0, C, 0, 16, 16, out_dt=L.DT_BF16, in_dt=L.DT_BF16, ldw=16)
    L.call('pg_conv_wgrad', ctypes.byref(d), to_nhwc(x).data_ptr(), to_nhwc(g).data_ptr(), 16, dw.data_ptr(), 1, 16, C, impl,
           stream())
    torch.cuda.synchronize()
    assert relerr(dw.cpu().numpy(), ref) < 1e-4


def test_tap_products_discriminator_last_layer():
    """Conv2d(C -> 1, s1) forward / dgrad / wgrad through the engine's tap-product helpers vs the oracle."""
    B, Ci, H = 2, 512, 31
    r = rng(23)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), L.DT_F16)
    w = r.standard_normal((1, Ci, 4, 4)).astype(np.float32) / np.sqrt(Ci * 16)
    b = r.standard_normal(1).astype(np.float32)
    wq = bf16_round(w, L.DT_F16)
    ref = orc.act_fwd('sigmoid', orc.conv2d_fwd(x, wq, b, 1))
    Ho = H - 1
    xa = act_of(to_nhwc(x, dt=L.DT_F16))
    wd = pack_weight(w, 1, 16, Ci, Ci, 0, 0, Ci * 16, 16, dt=L.DT_F16)
    out = act_of(torch.full((B, Ho, Ho, 4), 7.0, device='cuda'))
    E.taps_forward(L.PG_CONV, 1, 1, xa, None, wd.data_ptr(), torch.from_numpy(b).cuda(), L.ACT['sigmoid'], out, 0)
    torch.cuda.synchronize()
    assert relerr(out.t[..., 0].cpu().numpy(), ref[:, 0]) < 1e-4
    # backward
    dy = bf16_round(r.standard_normal((B, 1, Ho, Ho)))
    xb = bf16_round(x)
    wb = bf16_round(w)
    rdx, rdw, _ = orc.conv2d_bwd(xb, wb, dy, 1)
    dya = act_of(to_nhwc(dy))
    G = E.taps_gather(L.PG_CONV, 1, 1, dya, 0, B, H, H)
    w16 = torch.from_numpy(wb.reshape(Ci, 16)).cuda().bfloat16().contiguous()
    din = act_of(torch.empty((B, H, H, Ci), device='cuda', dtype=torch.bfloat16))
    E.taps_dgrad(G, w16, din)
    dw = torch.zeros((1, Ci, 4, 4), device='cuda')
    E.taps_wgrad(G, act_of(to_nhwc(xb)), dw.data_ptr(), Ci)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(din.t, Ci), rdx) < 5e-3
    assert relerr(dw.cpu().numpy(), rdw) < 1e-4


def test_tap_products_generator_output_layer():
    """ConvTranspose2d(C1 + C2 -> 1) forward / dgrad / wgrad through the tap-product helpers vs the oracle."""
    B, C1, C2, H = 2, 32, 32, 12
    Ci = C1 + C2
    r = rng(24)
    x = bf16_round(r.standard_normal((B, Ci, H, H)), L.DT_F16)
    w = r.standard_normal((Ci, 1, 4, 4)).astype(np.float32) / np.sqrt(Ci * 4)
    ref = orc.act_fwd('sigmoid', orc.convT_fwd(x, bf16_round(w, L.DT_F16)))
    x1 = act_of(to_nhwc(x[:, :C1], dt=L.DT_F16))
    x2 = act_of(to_nhwc(x[:, C1:], dt=L.DT_F16))
    wd = pack_weight(w, 1, 16, C1, C1, C2, C2, 16, 16, dt=L.DT_F16)
    out = act_of(torch.full((B, 2 * H, 2 * H, 4), 7.0, device='cuda'))
    E.taps_forward(L.PG_CONVT, 2, 1, x1, x2, wd.data_ptr(), None, L.ACT['sigmoid'], out, 0)
    torch.cuda.synchronize()
    assert relerr(out.t[..., 0].cpu().numpy(), ref[:, 0]) < 1e-4
    dy = bf16_round(r.standard_normal((B, 1, 2 * H, 2 * H)))
    xb, wb = bf16_round(x), bf16_round(w)
    rdx, rdw = orc.convT_bwd(xb, wb, dy)
    G = E.taps_gather(L.PG_CONVT, 2, 1, act_of(to_nhwc(dy)), 0, B, H, H)
    w16 = torch.from_numpy(wb.reshape(Ci, 16)).cuda().bfloat16().contiguous()
    din = act_of(torch.empty((B, H, H, Ci), device='cuda', dtype=torch.bfloat16))
    E.taps_dgrad(G, w16, din)
    dw = torch.zeros((Ci, 1, 4, 4), device='cuda')
    E.taps_wgrad(G, act_of(to_nhwc(xb[:, :C1])), dw.data_ptr(), C1)
    E.taps_wgrad(G, act_of(to_nhwc(xb[:, C1:])), dw.data_ptr() + C1 * 16 * 4, C2)
    torch.cuda.synchronize()
    assert relerr(from_nhwc(din.t, Ci), rdx) < 5e-3
    assert relerr(dw.cpu().numpy(), rdw) < 1e-4
