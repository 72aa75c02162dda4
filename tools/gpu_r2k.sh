#!/bin/bash
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a4_wgrad_group.py tests/test_gpu_b_models.py tests/test_gpu_c_step.py tests/test_gpu_c2_benchshapes.py > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/r2k_tests.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>/dev/null | cut -c1-120; done
timeout 300 python tools/timeline.py > gpurun_out/r2k_timeline.txt 2>&1; sed -n '1,2p;52,80p' gpurun_out/r2k_timeline.txt | cut -c1-120
