"""Test plumbing: numpy NCHW <-> device NHWC bf16, raw C-ABI calls."""
import ctypes

import numpy as np
import torch

from patchgan_b200 import _lib as L
from patchgan_b200.engine import Act, conv_desc, rup16


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def to_nhwc(x, cp=None, f32=False):
    """numpy NCHW -> device NHWC (bf16 or f32) with channels zero-padded to cp."""
    B, C, H, W = x.shape
    cp = cp or rup16(C)
    t = torch.zeros((B, H, W, cp), dtype=torch.float32)
    t[..., :C] = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
    t = t.cuda()
    return t if f32 else t.to(torch.bfloat16)


def from_nhwc(t, C):
    return np.ascontiguousarray(t[..., :C].float().cpu().numpy().transpose(0, 3, 1, 2))


def act_of(t, C=None):
    B, H, W, Cp = t.shape
    return Act(t, B, H, W, C or Cp, ld=Cp, f32=(t.dtype == torch.float32))


def pack_weight(w, N, Np, C1, C1p, C2, C2p, sn, sc, flip=0):
    """w: numpy float32 (any shape, reference layout) -> device bf16 [Np,16,C1p+C2p]"""
    wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).cuda()
    out = torch.empty((Np, 16, C1p + C2p), dtype=torch.bfloat16, device='cuda')
    L.call('pg_pack_weight', wd.data_ptr(), out.data_ptr(), N, Np, C1, C1p, C2, C2p, sn, sc, flip, stream())
    return out


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
