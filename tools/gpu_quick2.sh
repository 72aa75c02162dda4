#!/bin/bash
TAG=${1:-q}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_b_models.py tests/test_gpu_c_step.py -x -q 2>&1 | tail -5 | cut -c1-300
PG_TC_DEBUG=1 timeout 120 python tools/conv_trace.py 2>&1 | grep -v "^conv_tc: grid" | cut -c1-250 | head -8
PG_TC_DEBUG=1 timeout 120 python tools/conv_trace.py 2>&1 | grep "splits [2-8]" | sort -u | head
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
