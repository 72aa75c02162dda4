#!/bin/bash
# round 2, call C: per-CTA phase trace of the fused kernels
mkdir -p gpurun_out
timeout 300 python tools/res_trace.py > gpurun_out/r2c_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_gpu_c_step.py tests/test_gpu_d_api.py tests/test_gpu_b_models.py > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
cat gpurun_out/r2c_trace.log
tail -n 5 gpurun_out/r2c_tests.log
