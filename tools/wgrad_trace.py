"""Per-CTA phase timing of wgrad_tc_kernel (pg_debug_set_trace) for the cfg3 layer shapes."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc

SHAPES = [  # stride B H Ci N    (PG_CONV geometry: A = [B,H,H,Ci] input, G = [B,Ho,Ho,N] output gradient)
    (2, 16, 256, 16, 32), (2, 16, 128, 32, 64), (2, 16, 64, 64, 128), (2, 16, 32, 128, 256), (2, 16, 16, 256, 256),
    (2, 16, 4, 256, 256), (2, 32, 256, 16, 64), (2, 32, 128, 64, 128), (2, 32, 64, 128, 256), (1, 32, 32, 256, 512),
    (1, 32, 31, 512, 16),
]

def run(stride, B, H, Ci, N):
    dev = 'cuda'
    Ho = (H + 2 - 4) // stride + 1
    d = conv_desc(L.PG_CONV, stride, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, N, N, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    flops = 2.0 * B * Ho * Ho * Ci * N * 16
    a = torch.randn((B, H, H, Ci), device=dev, dtype=torch.bfloat16)
    g = torch.randn((B, Ho, Ho, N), device=dev, dtype=torch.bfloat16)
    dw = torch.zeros((N, Ci, 16), device=dev, dtype=torch.float32)
    trace = torch.zeros(1 << 20, device=dev, dtype=torch.int64)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        L.call('pg_conv_wgrad', ctypes.byref(d), a.data_ptr(), g.data_ptr(), N, dw.data_ptr(), Ci * 16, N, Ci, L.IMPL_TCGEN05, st)
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
    t = t[t[:, 6] != 0]
    n = len(t)
    t0 = t[:, 0].min()
    span = (t[:, 6].max() - t0) / 1e3
    r = lambda x: f'{np.median(x)/1e3:6.2f}/{np.max(x)/1e3:6.2f}'
    w = t[t[:, 3] != 0]
    print(f'wgrad s{stride} B{B} {H}x{H} C{Ci} N{N}: ctas {n} sms {len(set(t[:,7].tolist()))} event {us_plain:7.1f}us span {span:7.1f}us {flops/span/1e6:7.1f} TF/s | '
          f'start {r(t[:,0]-t0)} setup {r(t[:,1]-t[:,0])} first-full {r(w[:,2]-w[:,1])} mainloop {r(w[:,3]-w[:,2])} '
          f'acc-wait {r(w[:,4]-w[:,3])} epi {r(w[:,5]-w[:,4])} life {r(t[:,6]-t[:,0])}')

if __name__ == '__main__':
    for s in SHAPES:
        run(*s)
