#!/bin/bash
# round 2, call H: grouped weight-gradient launches
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a4_wgrad_group.py tests/test_gpu_a_ops.py -k "wgrad or weight" > gpurun_out/r2h_ops.log 2>&1; echo "ops rc=$?"
$T tests/test_gpu_b_models.py tests/test_gpu_c_step.py tests/test_gpu_c2_benchshapes.py > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_$tag.json 2> gpurun_out/r2h_$tag.err; echo "$tag rc=$? $(cut -c1-110 gpurun_out/r2h_$tag.json)"; }
run group X=1
run nogroup PATCHGAN_B200_GROUP_WGRAD=0
run group_early PATCHGAN_B200_DREAL_LATE=0
timeout 300 python tools/timeline.py > gpurun_out/r2h_timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 4 gpurun_out/r2h_ops.log gpurun_out/r2h_tests.log
tail -n 75 gpurun_out/r2h_timeline.txt | cut -c1-130
