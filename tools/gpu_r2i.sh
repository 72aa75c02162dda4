#!/bin/bash
# round 2, call I: deeper wgrad rings, CUDA-core tap-product dgrad
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a4_wgrad_group.py tests/test_gpu_a_ops.py tests/test_gpu_a2_skinny.py > gpurun_out/r2i_ops.log 2>&1; echo "ops rc=$?"
$T tests/test_gpu_b_models.py tests/test_gpu_c_step.py tests/test_gpu_c2_benchshapes.py > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; echo "$tag rc=$? $(cut -c1-110 gpurun_out/r2i_$tag.json)"; }
run base X=1
run g2 PG_WG_GSTAGES=2
run g3 PG_WG_GSTAGES=3
timeout 300 python tools/timeline.py > gpurun_out/r2i_timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 4 gpurun_out/r2i_ops.log gpurun_out/r2i_tests.log
sed -n '1,2p;30,80p' gpurun_out/r2i_timeline.txt | cut -c1-130
