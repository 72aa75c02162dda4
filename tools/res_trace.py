"""Per-CTA phase timing of conv_res_kernel (pg_conv_norm_fwd / pg_conv_dgrad_norm_bwd) for the cfg 3 generator layers:
   python tools/res_trace.py
For every shape: CUDA-event time of a warm launch, then one traced launch (pg_debug_set_trace): %globaltimer stamps per CTA
  0 start, 1 setup done, 2 first tile accumulated, 3 last tile accumulated, 4 phase 1 done, 5 grid barrier passed,
  8 phase 2 done, 6 exit; split-K launches also 10 partial sums + statistics done, 11 second grid barrier passed.  PG_TC_DEBUG=1 prints the plan (tile width, K-split) of every launch."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc

B = int(os.environ.get('RES_B', '16'))
# (kind, mode, H (conv input / lattice), C1, C2, N, n_norm)
FWD = [('conv1x1', 128, 48, 0, 32), ('conv', 128, 32, 0, 64), ('conv', 64, 64, 0, 128), ('conv', 32, 128, 0, 256),
       ('conv', 16, 256, 0, 256), ('conv', 8, 256, 0, 256), ('conv', 4, 256, 0, 256),
       ('convT', 4, 256, 256, 256), ('convT', 8, 256, 256, 256), ('convT', 16, 256, 256, 128), ('convT', 32, 128, 128, 64),
       ('convT', 64, 64, 64, 32)]
# backward: dgrad geometry.  ('conv1x1', H, K=16, N, n_norm) = dec6 taps dgrad; ('conv', Hout lattice...) = dgrad of convT;
# ('convT', ...) = dgrad of conv
BWD = [('conv1x1', 128, 16, 32, 32), ('conv', 64, 32, 128, 64), ('conv', 32, 64, 256, 128), ('conv', 16, 128, 512, 256),
       ('conv', 8, 256, 512, 256), ('conv', 2, 256, 256, 256),
       ('convT', 2, 256, 256, 256), ('convT', 4, 256, 256, 256), ('convT', 8, 256, 256, 256), ('convT', 16, 256, 128, 128),
       ('convT', 32, 128, 64, 64), ('convT', 64, 64, 32, 32)]


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def report(tag, launch, flops):
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    trace = torch.zeros(1 << 16, device='cuda', dtype=torch.int64)
    L.lib().pg_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    launch(); torch.cuda.synchronize()
    L.lib().pg_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(-1, 16)
    t = t[t[:, 0] != 0]
    t0 = t[:, 0].min()
    r = lambda a: f'{np.median(a) / 1e3:6.2f}/{np.max(a) / 1e3:6.2f}'
    span = (t[:, 6].max() - t0) / 1e3
    print(f'{tag:44s} ctas {len(t):3d} tiles/cta {int(t[:, 9].max()):2d} event {us:6.1f}us span {span:6.1f}us {flops / span / 1e6:7.1f} TF/s | '
          f'start {r(t[:, 0] - t0)} setup {r(t[:, 1] - t[:, 0])} first-tile {r(t[:, 2] - t[:, 1])} mainloop {r(t[:, 3] - t[:, 2])} '
          f'phase1-tail {r(t[:, 4] - t[:, 3])} barrier {r(t[:, 5] - t[:, 4])} phase2 {r(t[:, 8] - t[:, 5])} exit {r(t[:, 6] - t[:, 8])}'
          + (f' | split-K: sum+stats {r(t[:, 10] - t[:, 5])} barrier2 {r(t[:, 11] - t[:, 10])} finish {r(t[:, 8] - t[:, 11])}'
             if t[:, 10].max() > 0 else ''), flush=True)


def run_fwd(mode, H, C1, C2, N):
    dt = L.DT_F16
    if mode == 'conv1x1':
        d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, C1, 0, C1, 0, N, N, out_dt=dt, in_dt=dt)
        Ho, flops, wk = H, 2.0 * B * H * H * C1 * N, C1
    elif mode == 'conv':
        Ho = H // 2
        d = conv_desc(L.PG_CONV, 2, 1, B, H, H, Ho, Ho, C1, 0, C1, 0, N, N, out_dt=dt, in_dt=dt)
        flops, wk = 2.0 * B * Ho * Ho * C1 * N * 16, 16 * C1
    else:
        Ho = 2 * H
        d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, Ho, Ho, C1, C2, C1, C2, N, N, out_dt=dt, in_dt=dt)
        flops, wk = 2.0 * B * H * H * (C1 + C2) * N * 16, 16 * (C1 + C2)
    x1 = torch.randn((B, H, H, C1), device='cuda', dtype=torch.float16)
    x2 = torch.randn((B, H, H, C2), device='cuda', dtype=torch.float16) if C2 else None
    w = torch.randn((N, wk), device='cuda', dtype=torch.float16) * 0.05
    out = torch.empty((B, Ho, Ho, N), device='cuda', dtype=torch.float16)
    twin = torch.empty((B, Ho, Ho, N), device='cuda', dtype=torch.bfloat16)
    sums = torch.zeros((B, N, 2), device='cuda')
    sync = torch.zeros(16, device='cuda', dtype=torch.int32)
    fn = L.FusedNorm()
    fn.kind, fn.act, fn.sums, fn.sync = L.FUSED_FWD, L.ACT['leakyrelu'], sums.data_ptr(), sync.data_ptr()
    fn.ws, fn.ws_bytes = WS.data_ptr(), WS.numel()
    if not L.lib().pg_conv_norm_supported(ctypes.byref(d), ctypes.byref(fn), 1):
        print('fwd', mode, H, 'unsupported')
        return

    def launch():
        sums.zero_(); sync.zero_()
        L.call('pg_conv_norm_fwd', ctypes.byref(d), x1.data_ptr(), x2.data_ptr() if C2 else None, w.data_ptr(), out.data_ptr(),
               twin.data_ptr(), ctypes.byref(fn), stream())
    report(f'fwd {mode} {H}x{H} C{C1}+{C2} N{N}', launch, flops)


def run_bwd(mode, H, C, N, n_norm):
    if mode == 'conv1x1':
        d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, C, 0, C, 0, N, N)
        Ho, flops, wk = H, 2.0 * B * H * H * C * N, C
    elif mode == 'conv':          # dgrad of a ConvTranspose2d: stride-2 conv over dY (2H x 2H) -> H x H
        d = conv_desc(L.PG_CONV, 2, 1, B, 2 * H, 2 * H, H, H, C, 0, C, 0, N, N)
        Ho, flops, wk = H, 2.0 * B * H * H * C * N * 16, 16 * C
    else:                         # dgrad of a Conv2d s2: convT-form over dY (H x H) -> 2H x 2H
        d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, 2 * H, 2 * H, C, 0, C, 0, N, N)
        Ho, flops, wk = 2 * H, 2.0 * B * H * H * C * N * 16, 16 * C
    Hi = 2 * H if mode == 'conv' else H
    dy = torch.randn((B, Hi, Hi, C), device='cuda', dtype=torch.bfloat16)
    w = torch.randn((N, wk), device='cuda', dtype=torch.bfloat16) * 0.05
    dx = torch.empty((B, Ho, Ho, N), device='cuda', dtype=torch.bfloat16)
    y = torch.randn((B, Ho, Ho, n_norm), device='cuda', dtype=torch.float16)
    dskip = torch.randn((B, Ho, Ho, n_norm), device='cuda', dtype=torch.bfloat16)
    sums = torch.rand((B, n_norm, 2), device='cuda') + 1.0
    sums[..., 1] += Ho * Ho
    ws = torch.zeros(B * n_norm * 2 + 16, device='cuda')
    fn = L.FusedNorm()
    fn.kind, fn.act, fn.n_norm, fn.sums = L.FUSED_BWD, L.ACT['leakyrelu'], n_norm, sums.data_ptr()
    fn.bsums, fn.sync = ws.data_ptr(), ws.data_ptr() + B * n_norm * 2 * 4
    fn.y, fn.y_ld, fn.y_dtype = y.data_ptr(), n_norm, L.DT_F16
    fn.ws, fn.ws_bytes = WS.data_ptr(), WS.numel()
    if n_norm == N:
        fn.dskip, fn.dskip_ld = dskip.data_ptr(), n_norm
    if not L.lib().pg_conv_norm_supported(ctypes.byref(d), ctypes.byref(fn), 0):
        print('bwd', mode, H, 'unsupported')
        return

    def launch():
        ws.zero_()
        L.call('pg_conv_dgrad_norm_bwd', ctypes.byref(d), dy.data_ptr(), w.data_ptr(), dx.data_ptr(), ctypes.byref(fn), stream())
    report(f'bwd {mode} {H}x{H} C{C} N{N} norm{n_norm}', launch, flops)


if __name__ == '__main__':
    WS = torch.empty(8 << 20, device='cuda', dtype=torch.uint8)
    for s in FWD:
        run_fwd(*s)
    for s in BWD:
        run_bwd(*s)
