"""Per-CTA phase timing of conv_tc_kernel (pg_debug_set_trace) for a list of layer shapes.
   python tools/conv_trace.py            # cfg3 layer shapes"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc, ensure_workspace

SHAPES = [  # mode stride pad B H Ci Co outdt
    ('conv', 2, 1, 16, 256, 16, 32), ('conv', 2, 1, 16, 128, 32, 64), ('conv', 2, 1, 16, 64, 64, 128),
    ('conv', 2, 1, 16, 32, 128, 256), ('conv', 2, 1, 16, 16, 256, 256), ('conv', 2, 1, 16, 4, 256, 256),
    ('convT', 2, 1, 16, 16, 512, 128), ('convT', 2, 1, 16, 64, 128, 32), ('convT', 2, 1, 16, 128, 64, 16),
    ('conv', 2, 1, 32, 256, 16, 64), ('conv', 2, 1, 32, 128, 64, 128), ('conv', 2, 1, 32, 64, 128, 256),
    ('conv', 1, 1, 32, 32, 256, 512), ('conv', 1, 1, 32, 31, 512, 16),
]

OUT16 = os.environ.get('PROBE_OUT', 'f32') == 'f16'     # 16-bit output + bf16 twin like the discriminator layers


def run(mode, stride, pad, B, H, Ci, Co):
    dev = 'cuda'
    if mode == 'conv1x1':
        Ho = H
        d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16, in_dt=L.DT_F16)
        flops = 2.0 * B * H * H * Ci * Co
    elif mode == 'conv':
        Ho = (H + 2 * pad - 4) // stride + 1
        d = conv_desc(L.PG_CONV, stride, pad, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16 if OUT16 else L.DT_F32, in_dt=L.DT_F16, act=2 if OUT16 else 0)
        flops = 2.0 * B * Ho * Ho * Ci * Co * 16
    else:
        Ho = 2 * H
        d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16 if OUT16 else L.DT_F32, in_dt=L.DT_F16, act=2 if OUT16 else 0)
        flops = 2.0 * B * H * H * Ci * Co * 16
    x = torch.randn((B, H, H, Ci), device=dev, dtype=torch.float16)
    w = torch.randn((Co, 16, Ci), device=dev, dtype=torch.float16)
    out = torch.empty((B, Ho, Ho, Co), device=dev, dtype=torch.float16 if (mode == 'conv1x1' or OUT16) else torch.float32)
    twin = torch.empty((B, Ho, Ho, Co), device=dev, dtype=torch.bfloat16) if OUT16 else None
    trace = torch.zeros(1 << 20, device=dev, dtype=torch.int64)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        L.call('pg_conv_fwd', ctypes.byref(d), x.data_ptr(), None, w.data_ptr(), None, out.data_ptr(), twin.data_ptr() if OUT16 else None, L.IMPL_TCGEN05, st)
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    us_plain = e0.elapsed_time(e1) * 1e3
    L.lib().pg_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    call(); torch.cuda.synchronize()
    L.lib().pg_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(-1, 16)
    t = t[t[:, 0] != 0]
    n = len(t)
    t0 = t[:, 0].min()
    if n and t[0, 15] == 1:            # persistent kernel: per-CTA totals (cycles -> us at 1.9 GHz)
        cyc = lambda a: f'{np.median(a)/1.9e3:6.2f}/{np.max(a)/1.9e3:6.2f}'
        r = lambda a: f'{np.median(a)/1e3:6.2f}/{np.max(a)/1e3:6.2f}'
        span = (t[:, 6].max() - t0) / 1e3
        tiles = t[:, 13]
        print(f'{mode} s{stride} B{B} {H}x{H} C{Ci}->N{Co} [persistent]: ctas {n} event {us_plain:7.1f}us span {span:7.1f}us '
              f'{flops/span/1e6:7.1f} TF/s | tiles/cta {np.median(tiles):.0f} start(med/max) {r(t[:,0]-t0)} setup {r(t[:,1]-t[:,0])} '
              f'first-acc {r(t[:,2]-t[:,1])} life {r(t[:,6]-t[:,0])} | us per CTA: producer-wait-ring {cyc(t[:,8])} issuer-wait-operands {cyc(t[:,9])} '
              f'issuer-wait-drain {cyc(t[:,10])} epi-wait-acc {cyc(t[:,11])} epi-busy {cyc(t[:,12])} | epi-busy per tile '
              f'{np.median(t[:,12]/np.maximum(tiles,1))/1.9e3:5.2f} us')
        return
    span = (t[:, 6].max() - t0) / 1e3
    r = lambda a: f'{np.median(a)/1e3:6.2f}/{np.max(a)/1e3:6.2f}'
    sm = t[:, 7]
    print(f'{mode} s{stride} B{B} {H}x{H} C{Ci}->N{Co}: ctas {n} sms {len(set(sm.tolist()))} event {us_plain:7.1f}us span {span:7.1f}us {flops/span/1e6:7.1f} TF/s | '
          f'start(med/max) {r(t[:,0]-t0)} setup {r(t[:,1]-t[:,0])} first-full {r(t[:,2]-t[:,1])} mainloop {r(t[:,3]-t[:,2])} '
          f'mma-waited {np.median(t[:,8])/1.9e3:6.2f} prod-waited {np.median(t[:,9])/1.9e3:6.2f} prod-done {r(t[:,10]-t[:,1])} push {r(t[:,11]-t[:,4])} sync {r(t[:,12]-t[:,11])} reduce {r(t[:,13]-t[:,12])} acc-wait {r(t[:,4]-t[:,3])} epi {r(t[:,5]-t[:,4])} exit {r(t[:,6]-t[:,5])} cta-life {r(t[:,6]-t[:,0])}')

if __name__ == '__main__':
    ensure_workspace(torch.device('cuda', 0))
    pass
    if len(sys.argv) >= 2 and sys.argv[1] == 'q':      # persistent-kernel shapes of cfg 3
        for s in [('conv1x1', 1, 0, 32, 128, 64, 64), ('conv1x1', 1, 0, 16, 128, 48, 32), ('conv1x1', 1, 0, 16, 128, 64, 16),
                  ('conv', 2, 1, 16, 128, 32, 64), ('conv', 2, 1, 16, 64, 64, 128), ('conv', 2, 1, 32, 128, 64, 128),
                  ('convT', 2, 1, 16, 32, 256, 64), ('convT', 2, 1, 16, 64, 128, 32), ('convT', 2, 1, 32, 64, 128, 64),
                  ('convT', 2, 1, 32, 32, 256, 128), ('conv1x1', 1, 0, 32, 31, 16, 512)]:
            run(*s)
        sys.exit(0)
    if len(sys.argv) >= 2 and sys.argv[1] == 'd':      # the discriminator's wide layers (CTA-pair candidates)
        for s in [('conv', 2, 1, 32, 128, 64, 128), ('conv', 2, 1, 32, 64, 128, 256), ('conv', 1, 1, 32, 32, 256, 512),
                  ('convT', 2, 1, 32, 32, 256, 128), ('convT', 2, 1, 32, 64, 128, 64), ('conv', 1, 2, 32, 31, 512, 256)]:
            run(*s)
        sys.exit(0)
    sel = SHAPES if len(sys.argv) < 2 else ([('conv1x1', 1, 0, 32, 128, 64, 64), ('conv1x1', 1, 0, 16, 128, 48, 32), ('conv1x1', 1, 0, 16, 128, 64, 16), ('convT', 2, 1, 32, 64, 128, 64)] if sys.argv[1] == 'p' else [s for s in SHAPES if s[3] == 32])
    for s in sel:
        run(*s)
