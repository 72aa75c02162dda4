"""pg_conv_wgrad_group: several weight gradients in one launch give what the one-launch-each calls give (same kernel body,
same tensor maps; only the pixel-tile split counts differ) and match the numpy oracle."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import patchgan_oracle as orc
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc, rup16
from tests.gpu_util import bf16_round, relerr, stream, to_nhwc

pytestmark = pytest.mark.gpu


def test_grouped_weight_gradients_match_oracle_and_single_launches():
    r = np.random.default_rng(31)
    # (B, Cin, Cout, H, stride): conv layers of different sizes, incl. one that is split over many CTAs
    cases = [(2, 32, 64, 32, 2), (2, 64, 64, 8, 2), (1, 128, 256, 16, 2), (2, 256, 128, 4, 2), (4, 16, 32, 64, 2), (2, 64, 128, 33, 1)]
    jobs = (L.WgradJob * len(cases))()
    keep, refs, outs = [], [], []
    for j, (B, Ci, Co, H, s) in enumerate(cases):
        x = bf16_round(r.standard_normal((B, Ci, H, H)))
        Ho = (H + 2 - 4) // s + 1
        dy = bf16_round(r.standard_normal((B, Co, Ho, Ho)))
        _, dw_ref, _ = orc.conv2d_bwd(x, np.zeros((Co, Ci, 4, 4), np.float32), dy, s, need_dx=False)
        refs.append(dw_ref)
        xd, dyd = to_nhwc(x), to_nhwc(dy)
        Cip, Cop = rup16(Ci), rup16(Co)
        d = conv_desc(L.PG_CONV, s, 1, B, H, H, Ho, Ho, Cip, 0, Cip, 0, Cop, Cop, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
        tap_major = j % 2            # alternate the two epilogues
        if tap_major:
            dw = torch.zeros((16, Co, Ci), device='cuda')
            ld_n, n_real, c_real, Cs = Co, Co, Ci, Ci
        else:
            dw = torch.zeros((Co, Ci, 4, 4), device='cuda')
            ld_n, n_real, c_real, Cs = Ci * 16, Co, Ci, 0
        ctypes.memmove(ctypes.byref(jobs[j].desc), ctypes.byref(d), ctypes.sizeof(L.ConvDesc))
        jobs[j].a, jobs[j].g, jobs[j].ldg, jobs[j].tap_major, jobs[j].dw = xd.data_ptr(), dyd.data_ptr(), Cop, tap_major, dw.data_ptr()
        jobs[j].ld_n, jobs[j].n_real, jobs[j].c_real, jobs[j].Cs = ld_n, n_real, c_real, Cs
        keep += [xd, dyd]
        outs.append((dw, tap_major))
    L.call('pg_conv_wgrad_group', jobs, len(cases), stream())
    torch.cuda.synchronize()
    # one more launch: a job whose gradient-side operand is the virtual concat of two tensors (64 | 64 channels)
    B, Ci, Co, H = 2, 32, 128, 16
    x = bf16_round(r.standard_normal((B, Ci, H, H)))
    dy = bf16_round(r.standard_normal((B, Co, H // 2, H // 2)))
    _, dw_ref, _ = orc.conv2d_bwd(x, np.zeros((Co, Ci, 4, 4), np.float32), dy, 2, need_dx=False)
    xd, d1, d2 = to_nhwc(x), to_nhwc(dy[:, :64]), to_nhwc(dy[:, 64:])
    cat = (L.WgradJob * 1)()
    d = conv_desc(L.PG_CONV, 2, 1, B, H, H, H // 2, H // 2, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    ctypes.memmove(ctypes.byref(cat[0].desc), ctypes.byref(d), ctypes.sizeof(L.ConvDesc))
    S = torch.zeros((16, Co, Ci), device='cuda')
    cat[0].a, cat[0].g, cat[0].ldg, cat[0].tap_major, cat[0].dw = xd.data_ptr(), d1.data_ptr(), 64, 1, S.data_ptr()
    cat[0].ld_n, cat[0].n_real, cat[0].c_real, cat[0].Cs = Co, Co, Ci, Ci
    cat[0].g2, cat[0].ldg2, cat[0].n_split = d2.data_ptr(), 64, 64
    L.call('pg_conv_wgrad_group', cat, 1, stream())
    torch.cuda.synchronize()
    assert relerr(S.cpu().numpy().transpose(1, 2, 0).reshape(dw_ref.shape), dw_ref) < 1e-4
    for (dw, tm), ref in zip(outs, refs):
        got = dw.cpu().numpy()
        if tm:
            got = got.transpose(1, 2, 0).reshape(ref.shape)        # S[tap][n][c] -> (n, c, kh, kw)
        assert relerr(got, ref) < 1e-4


@pytest.mark.parametrize('C', [512, 64, 34])
@pytest.mark.parametrize('act', ['tanh', 'leakyrelu', None])
def test_tap_product_data_gradient_with_fused_activation_backward(act, C):
    """pg_taps_dgrad_act: dx[q][c] = (sum_tap G[q][tap] W[c][tap]) * act'(y[q][c]) against numpy."""
    r = np.random.default_rng(32)
    nq = 2 * 31 * 31 + 5
    G = bf16_round(r.standard_normal((nq, 16)))
    W = bf16_round(r.standard_normal((C, 16)) / 4)
    yv = bf16_round(np.tanh(r.standard_normal((nq, C))), L.DT_F16)
    ref = G @ W.T
    if act == 'tanh':
        ref = ref * (1 - yv * yv)
    elif act == 'leakyrelu':
        ref = ref * np.where(yv > 0, 1.0, 0.2)
    Gd = torch.from_numpy(G).cuda().to(torch.bfloat16)
    Wd = torch.from_numpy(W).cuda().to(torch.bfloat16)
    yd = torch.from_numpy(yv).cuda().to(torch.float16)
    dx = torch.full((nq, C), 7.0, device='cuda', dtype=torch.bfloat16)
    L.call('pg_taps_dgrad_act', Gd.data_ptr(), Wd.data_ptr(), dx.data_ptr(), C, yd.data_ptr() if act else None, C, L.DT_F16,
           L.ACT[act], nq, C, stream())
    torch.cuda.synchronize()
    assert relerr(dx.float().cpu().numpy(), ref) < 5e-3
