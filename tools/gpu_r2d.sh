#!/bin/bash
# round 2, call D: fused kernels v3 (sleeping waits, direct stores, split-K mode): tests, trace, bench
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a3_fused.py > gpurun_out/r2d_fused.log 2>&1; echo "fused ops rc=$?"
PG_TC_DEBUG=1 timeout 300 python tools/res_trace.py > gpurun_out/r2d_trace.log 2>&1; echo "trace rc=$?"
$T tests/test_gpu_b_models.py tests/test_gpu_c_step.py tests/test_gpu_c2_benchshapes.py tests/test_gpu_d_api.py > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail gpurun_out/r2d_detail.json > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -n 4 gpurun_out/r2d_fused.log gpurun_out/r2d_tests.log
grep -v "^conv_tc\|^wgrad" gpurun_out/r2d_trace.log | cut -c1-400
cut -c1-200 gpurun_out/r2d_bench.json
