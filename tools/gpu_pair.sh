#!/bin/bash
# CTA-pair (cta_group::2) kernel: parity tests, then the per-CTA role trace of the discriminator's wide layers with the
# pair kernel on every eligible shape (PG_TC_PAIR=2) and off (0).   gpurun -- 'bash tools/gpu_pair.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a5_pair.py > gpurun_out/pair_tests.log 2>&1; echo "pair tests rc=$?"; tail -n 3 gpurun_out/pair_tests.log
for pr in 2 0; do
  echo "== PG_TC_PAIR=$pr"
  PROBE_OUT=f16 PG_TC_PAIR=$pr timeout 300 python tools/conv_trace.py d 2>&1 | tail -n 8 | cut -c1-420
done
