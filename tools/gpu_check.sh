#!/bin/bash
# Full GPU suite, then bench lines (cfg3 twice) and the in-graph timeline.  gpurun -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/${TAG}_tests.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>gpurun_out/${TAG}_bench$i.err | tee gpurun_out/${TAG}_bench$i.json | cut -c1-120; done
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1; head -n 2 gpurun_out/${TAG}_timeline.txt
