#!/bin/bash
# round 2, call A: the fused conv + InstanceNorm kernels (op tests first, then the networks, then the step)
mkdir -p gpurun_out
export PG_TC_DEBUG=1
T="timeout 300 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a3_fused.py -k "forward and f16" > gpurun_out/r2a_fwd16.log 2>&1; echo "fwd16 rc=$?"
$T tests/test_gpu_a3_fused.py -k "forward and bf16" > gpurun_out/r2a_fwdbf.log 2>&1; echo "fwdbf rc=$?"
$T tests/test_gpu_a3_fused.py -k "backward" > gpurun_out/r2a_bwd.log 2>&1; echo "bwd rc=$?"
$T tests/test_gpu_a3_fused.py -k "not forward and not backward or refuses or dropout" > gpurun_out/r2a_misc.log 2>&1; echo "misc rc=$?"
unset PG_TC_DEBUG
$T tests/test_gpu_b_models.py > gpurun_out/r2a_models.log 2>&1; echo "models rc=$?"
$T tests/test_gpu_c_step.py > gpurun_out/r2a_step.log 2>&1; echo "step rc=$?"
PATCHGAN_B200_FUSED_BWD=0 $T tests/test_gpu_c_step.py -k "oracle_and_reference or rectangular" > gpurun_out/r2a_step_nobwd.log 2>&1; echo "step(fused fwd only) rc=$?"
timeout 600 python -m pytest -q -p no:cacheprovider tests -m gpu > gpurun_out/r2a_all.log 2>&1; echo "all rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
PATCHGAN_B200_FUSED_FWD=0 PATCHGAN_B200_FUSED_BWD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_legacy.json 2> gpurun_out/r2a_bench_legacy.err; echo "bench legacy rc=$?"
tail -3 gpurun_out/r2a_*.log
cat gpurun_out/r2a_bench.json | cut -c1-600
