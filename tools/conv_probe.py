"""Time one convolution shape through the C-ABI (CUDA events, L2-warm), e.g.
   python tools/conv_probe.py conv 2 16 4 256 256     # mode stride B H Cin Cout"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc

ACT = int(os.environ.get('PROBE_ACT', '0'))
OUT16 = os.environ.get('PROBE_OUT', 'f32') == 'f16'  # 16-bit output + bf16 twin like the discriminator layers      # PgAct code of the fused activation (pointwise mode)


def run(mode, stride, B, H, Ci, Co, reps=50):
    dev = 'cuda'
    if mode == 'conv1x1':
        Ho = H
        d = conv_desc(L.PG_CONV1X1, 1, 0, B, H, H, H, H, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16, in_dt=L.DT_F16, act=ACT)
        flops = 2.0 * B * H * H * Ci * Co
    elif mode == 'conv':
        Ho = (H + 2 - 4) // stride + 1
        d = conv_desc(L.PG_CONV, stride, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16 if OUT16 else L.DT_F32, in_dt=L.DT_F16, act=ACT)
        flops = 2.0 * B * Ho * Ho * Ci * Co * 16
    else:
        Ho = 2 * H
        d = conv_desc(L.PG_CONVT, 2, 1, B, H, H, Ho, Ho, Ci, 0, Ci, 0, Co, Co, out_dt=L.DT_F16 if OUT16 else L.DT_F32, in_dt=L.DT_F16, act=ACT)
        flops = 2.0 * B * H * H * Ci * Co * 16
    x = torch.randn((B, H, H, Ci), device=dev, dtype=torch.float16)
    w = torch.randn((Co, 16, Ci), device=dev, dtype=torch.float16)
    out = torch.empty((B, Ho, Ho, Co), device=dev, dtype=torch.float16 if (mode == 'conv1x1' or OUT16) else torch.float32)
    twin = torch.empty((B, Ho, Ho, Co), device=dev, dtype=torch.bfloat16) if OUT16 else None
    def call():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.call('pg_conv_fwd', ctypes.byref(d), x.data_ptr(), None, w.data_ptr(), None, out.data_ptr(), twin.data_ptr() if twin is not None else None, L.IMPL_TCGEN05, st)
    for _ in range(5): call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): call()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f'{mode} s{stride} B{B} {H}x{H} C{Ci}->N{Co}: {us:8.2f} us  {flops / us / 1e6:8.1f} TFLOP/s  (PG_TC_STAGES={os.environ.get("PG_TC_STAGES","-")} OCC={os.environ.get("PG_TC_OCC","-")})')

if __name__ == '__main__':
    a = sys.argv[1:]
    run(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]))
