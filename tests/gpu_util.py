"""Test plumbing: numpy NCHW <-> device NHWC bf16, raw C-ABI calls."""
import ctypes

import numpy as np
import torch

from patchgan_b200 import _lib as L
from patchgan_b200.engine import Act, conv_desc, rup16, TORCH_DT


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def bf16_round(a, dt=L.DT_BF16):
    """round to the 16-bit type `dt` and back to float32"""
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(TORCH_DT[dt]).float().numpy()


def to_nhwc(x, cp=None, f32=False, dt=L.DT_BF16):
    """numpy NCHW -> device NHWC (16-bit `dt` or f32) with channels zero-padded to cp."""
    B, C, H, W = x.shape
    cp = cp or rup16(C)
    t = torch.zeros((B, H, W, cp), dtype=torch.float32)
    t[..., :C] = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
    t = t.cuda()
    return t if f32 else t.to(TORCH_DT[dt])


def from_nhwc(t, C):
    return np.ascontiguousarray(t[..., :C].float().cpu().numpy().transpose(0, 3, 1, 2))


def act_of(t, C=None):
    B, H, W, Cp = t.shape
    dt = {torch.bfloat16: L.DT_BF16, torch.float32: L.DT_F32, torch.float16: L.DT_F16}[t.dtype]
    return Act(t, B, H, W, C or Cp, ld=Cp, dt=dt)


def pack_weight(w, N, Np, C1, C1p, C2, C2p, sn, sc, flip=0, dt=L.DT_BF16):
    """w: numpy float32 (any shape, reference layout) -> device 16-bit [Np,16,C1p+C2p]"""
    wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).cuda()
    out = torch.empty((Np, 16, C1p + C2p), dtype=TORCH_DT[dt], device='cuda')
    L.call('pg_pack_weight', wd.data_ptr(), out.data_ptr(), N, Np, C1, C1p, C2, C2p, sn, sc, flip, dt, stream())
    return out


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
