#!/bin/bash
TAG=${1:-q}
O=gpurun_out
python -m pytest tests/test_gpu_a2_skinny.py tests/test_gpu_c_step.py -x -q 2>&1 | tail -5 | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
