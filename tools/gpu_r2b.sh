#!/bin/bash
# round 2, call B: 16-warp epilogue of the fused kernels; benchmark-shape step tests; per-launch timings fused vs separate
mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -x -p no:cacheprovider"
$T tests/test_gpu_a3_fused.py > gpurun_out/r2b_fused.log 2>&1; echo "fused ops rc=$?"
$T tests/test_gpu_b_models.py > gpurun_out/r2b_models.log 2>&1; echo "models rc=$?"
$T tests/test_gpu_c_step.py > gpurun_out/r2b_step.log 2>&1; echo "step rc=$?"
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_gpu_d_api.py > gpurun_out/r2b_api.log 2>&1; echo "api rc=$?"
timeout 900 python -m pytest -q -p no:cacheprovider tests/test_gpu_c2_benchshapes.py -s > gpurun_out/r2b_shapes.log 2>&1; echo "benchshapes rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail gpurun_out/r2b_detail.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
PATCHGAN_B200_FUSED_FWD=0 PATCHGAN_B200_FUSED_BWD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail gpurun_out/r2b_detail_legacy.json > gpurun_out/r2b_bench_legacy.json 2> gpurun_out/r2b_bench_legacy.err; echo "bench legacy rc=$?"
PATCHGAN_B200_FUSED_BWD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_fwdonly.json 2> gpurun_out/r2b_bench_fwdonly.err; echo "bench fwd-only rc=$?"
for f in gpurun_out/r2b_*.log; do echo "== $f"; tail -n 3 $f; done
cut -c1-200 gpurun_out/r2b_bench.json gpurun_out/r2b_bench_legacy.json gpurun_out/r2b_bench_fwdonly.json
