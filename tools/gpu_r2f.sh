#!/bin/bash
# round 2, call F: schedule experiments (weight-gradient streams, late D(real) forward)
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err; echo "$tag rc=$? $(cut -c1-110 gpurun_out/r2f_$tag.json)"; }
run base X=1
run ws2 PATCHGAN_B200_WSTREAMS=2
run ws3 PATCHGAN_B200_WSTREAMS=3
run late PATCHGAN_B200_DREAL_LATE=1
run late_ws3 PATCHGAN_B200_DREAL_LATE=1 PATCHGAN_B200_WSTREAMS=3
run legacy PATCHGAN_B200_FUSED_FWD=0 PATCHGAN_B200_FUSED_BWD=0
run fwdonly PATCHGAN_B200_FUSED_BWD=0
run fwdonly_late PATCHGAN_B200_FUSED_BWD=0 PATCHGAN_B200_DREAL_LATE=1
PATCHGAN_B200_WSTREAMS=3 PATCHGAN_B200_DREAL_LATE=1 timeout 600 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_c_step.py > gpurun_out/r2f_step.log 2>&1; echo "step tests (ws3, late) rc=$?"
tail -n 3 gpurun_out/r2f_step.log
