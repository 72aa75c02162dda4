#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a4_wgrad_group.py tests/test_gpu_b_models.py -k "tap_product or discriminator" > gpurun_out/r2l_t.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2l_t.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>/dev/null | cut -c1-110
for S in 2 4; do echo "== PG_RES_S=$S"; PG_RES_S=$S timeout 200 python tools/res_trace.py 2>&1 | grep -v "^conv_" | cut -c1-175; done
echo "== PG_RES_BN=64"; PG_RES_BN=64 timeout 200 python tools/res_trace.py 2>&1 | grep -v "^conv_" | cut -c1-175
