#!/bin/bash
# conv+norm kernel: fused-op tests, then the per-CTA phase trace of the cfg3 generator layers for the planner's own choice
# and for forced K-splits (PG_RES_S)
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a3_fused.py > gpurun_out/res_tests.log 2>&1; echo "fused tests rc=$?"; tail -n 3 gpurun_out/res_tests.log
for S in 0 4 8 16 32; do
  echo "== PG_RES_S=$S"
  PG_RES_S=$S timeout 300 python tools/res_trace.py 2>&1 | grep -E "^(fwd|bwd)" | cut -c1-60,75-110,118-520 > gpurun_out/res_trace_S$S.txt
  cat gpurun_out/res_trace_S$S.txt | cut -c1-330
done
