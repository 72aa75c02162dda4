"""Time the streaming (1-real-channel) kernels at the cfg3 shapes through the C-ABI (CUDA events, graph of 20 reps)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from patchgan_b200 import _lib as L
from patchgan_b200.engine import conv_desc

dev = 'cuda'
SK = L.IMPL_SKINNY
IMPL = int(os.environ.get('PROBE_IMPL', SK))


def timeit(name, fn, bytes_moved, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f'{name:44s} {us:8.1f} us   {bytes_moved / us / 1e3:8.1f} GB/s (algorithmic bytes {bytes_moved/1e6:.1f} MB)', flush=True)


def st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def fwd(name, d, src1, src2, w, bias, out, bytes_moved):
    def fn():
        L.call('pg_conv_fwd', ctypes.byref(d), src1.data_ptr(), src2.data_ptr() if src2 is not None else None, w.data_ptr(),
               bias.data_ptr() if bias is not None else None, out.data_ptr(), None, IMPL, st())
    timeit(name, fn, bytes_moved)


def main():
    h = torch.float16
    # generator output layer: convT (32+32 -> 1), 128 -> 256, B16
    x1 = torch.randn((16, 128, 128, 32), device=dev, dtype=h); x2 = torch.randn_like(x1)
    w = torch.randn((16, 16, 64), device=dev, dtype=h)
    out = torch.empty((16, 256, 256, 4), device=dev, dtype=torch.float32)
    d = conv_desc(L.PG_CONVT, 2, 1, 16, 128, 128, 256, 256, 32, 32, 32, 32, 16, 4, n_valid=1, act=4, out_dt=L.DT_F32, in_dt=L.DT_F16)
    fwd('gen out convT fwd (64->1, 128->256, B16)', d, x1, x2, w, None, out, x1.numel() * 4 + 16 * 256 * 256 * 4)
    # disc last layer: conv s1 512 -> 1, 31 -> 30, B32
    x = torch.randn((32, 31, 31, 512), device=dev, dtype=h)
    w = torch.randn((16, 16, 512), device=dev, dtype=h)
    bias = torch.zeros(16, device=dev)
    out = torch.empty((32, 30, 30, 4), device=dev, dtype=torch.float32)
    d = conv_desc(L.PG_CONV, 1, 1, 32, 31, 31, 30, 30, 512, 0, 512, 0, 16, 4, n_valid=1, act=4, out_dt=L.DT_F32, has_bias=1, in_dt=L.DT_F16)
    fwd('disc last conv fwd (512->1, 31->30, B32)', d, x, None, w, bias, out, x.numel() * 2)
    # dgrad of disc first layer, mask channel only: convT-form 64 -> ch 3, 128 -> 256, B16
    g = torch.randn((16, 128, 128, 64), device=dev, dtype=torch.bfloat16)
    w = torch.randn((16, 16, 64), device=dev, dtype=torch.bfloat16)
    out = torch.empty((16, 256, 256, 16), device=dev, dtype=torch.bfloat16)
    d = conv_desc(L.PG_CONVT, 2, 1, 16, 128, 128, 256, 256, 64, 0, 64, 0, 16, 16, n_valid=4, out_dt=L.DT_BF16, n_first=3)
    fwd('disc first dgrad ch3 (64->1, 128->256, B16)', d, g, None, w, None, out, g.numel() * 2 + 16 * 256 * 256 * 2)
    # dgrad of disc last layer: 1 -> 512, 30 -> 31, B32
    dy = torch.randn((32, 30, 30, 16), device=dev, dtype=torch.bfloat16)
    w = torch.randn((512, 16, 16), device=dev, dtype=torch.bfloat16)
    out = torch.empty((32, 31, 31, 512), device=dev, dtype=torch.bfloat16)
    d = conv_desc(L.PG_CONV, 1, 2, 32, 30, 30, 31, 31, 16, 0, 16, 0, 512, 512, out_dt=L.DT_BF16, c_valid=1)
    fwd('disc last dgrad (1->512, 30->31, B32)', d, dy, None, w, None, out, out.numel() * 2)
    # dgrad of generator output layer: conv s2 form 1 -> 64, 256 -> 128, B16
    dy = torch.randn((16, 256, 256, 16), device=dev, dtype=torch.bfloat16)
    w = torch.randn((64, 16, 16), device=dev, dtype=torch.bfloat16)
    out = torch.empty((16, 128, 128, 64), device=dev, dtype=torch.bfloat16)
    d = conv_desc(L.PG_CONV, 2, 1, 16, 256, 256, 128, 128, 16, 0, 16, 0, 64, 64, out_dt=L.DT_BF16, c_valid=1)
    fwd('gen out dgrad (1->64, 256->128, B16)', d, dy, None, w, None, out, out.numel() * 2 + dy.numel() * 2)
    # wgrads
    a = torch.randn((32, 31, 31, 512), device=dev, dtype=torch.bfloat16)
    g = torch.randn((32, 30, 30, 16), device=dev, dtype=torch.bfloat16)
    dw = torch.zeros((1, 512, 16), device=dev)
    d = conv_desc(L.PG_CONV, 1, 1, 32, 31, 31, 30, 30, 512, 0, 512, 0, 16, 16, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    def f1():
        L.call('pg_conv_wgrad', ctypes.byref(d), a.data_ptr(), g.data_ptr(), 16, dw.data_ptr(), 512 * 16, 1, 512, IMPL, st())
    timeit('disc last wgrad (512x1, B32)', f1, a.numel() * 2)
    a2 = torch.randn((16, 256, 256, 16), device=dev, dtype=torch.bfloat16)
    g2 = torch.randn((16, 128, 128, 32), device=dev, dtype=torch.bfloat16)
    dw2 = torch.zeros((32, 1, 16), device=dev)
    d2 = conv_desc(L.PG_CONV, 2, 1, 16, 256, 256, 128, 128, 16, 0, 16, 0, 32, 32, out_dt=L.DT_BF16, in_dt=L.DT_BF16)
    def f2():
        L.call('pg_conv_wgrad', ctypes.byref(d2), a2.data_ptr(), g2.data_ptr(), 32, dw2.data_ptr(), 16, 32, 1, IMPL, st())
    timeit('gen out wgrad (32x1 per source, B16)', f2, g2.numel() * 2 + a2.numel() * 2)


if __name__ == '__main__':
    main()
