#!/bin/bash
# quick GPU check: tests + bench (streams on / off)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_$TAG.log | cut -c1-300
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
PATCHGAN_B200_STREAMS=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_nostreams.json 2> $O/bench_${TAG}_nostreams.err; echo "bench rc=$?"
cut -c1-400 $O/bench_${TAG}_nostreams.json
