#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a5_pair.py tests/test_gpu_a_ops.py -k "pair or conv" > gpurun_out/pair_tests.log 2>&1; echo "pair tests rc=$?"; tail -n 3 gpurun_out/pair_tests.log
for pr in 1 0; do
  echo "== PG_TC_PAIR=$pr"
  PROBE_OUT=f16 PG_TC_PAIR=$pr timeout 300 python tools/conv_trace.py d 2>&1 | tail -n 8 | cut -c1-420
done
timeout 300 python - <<'PY'
import ctypes, torch
from patchgan_b200 import _lib as L
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
B, H = 16, 256
x = torch.rand((B, 3, H, H), device='cuda'); y = torch.rand((B, 1, H, H), device='cuda')
rows = B * (H // 2) ** 2
a = torch.empty((rows, 48), device='cuda', dtype=torch.float16); a2 = torch.empty_like(a, dtype=torch.bfloat16)
us = t(lambda: L.call('pg_im2col_s2', x.data_ptr(), 3 * H * H, H * H, H, 1, 3, B, H, H, a.data_ptr(), a2.data_ptr(), 48, 0, L.DT_F16, st))
print(f'im2col_s2 G input: {us:.1f} us  {(x.numel() * 4 + 2 * a.numel() * 2) / us / 1e3:.0f} GB/s')
d = torch.empty((2 * rows, 64), device='cuda', dtype=torch.float16); d2 = torch.empty_like(d, dtype=torch.bfloat16)
us = t(lambda: L.call('pg_im2col_s2_pair', x.data_ptr(), 3, y.data_ptr(), 1, B, H, H, d.data_ptr(), d2.data_ptr(), 64, L.DT_F16, st))
print(f'im2col_s2_pair D input: {us:.1f} us  {((x.numel() + y.numel()) * 4 + 2 * d.numel() * 2 * 7 / 8) / us / 1e3:.0f} GB/s')
g = torch.rand((B, H, H, 4), device='cuda')
us = t(lambda: L.call('pg_im2col_s2', g.data_ptr(), H * H * 4, 1, H * 4, 4, 1, B, H, H, d.data_ptr(), d2.data_ptr(), 64, 3, L.DT_F16, st))
print(f'im2col_s2 mask column: {us:.1f} us')
PY
