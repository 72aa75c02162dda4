#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a5_pair.py tests/test_gpu_a_ops.py -k "pair or weight_gradient" > gpurun_out/wgpair_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/wgpair_tests.log
for pr in 1 0; do
  echo "== PG_WG_PAIR=$pr"
  PG_WG_PAIR=$pr timeout 300 python tools/wgrad_trace.py 2>&1 | tail -n 8 | cut -c1-300
done
for pr in 1 0; do PG_WG_PAIR=$pr timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>/dev/null | cut -c1-110; done
