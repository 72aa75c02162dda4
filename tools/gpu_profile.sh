#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list of one eager step, ncu --set full of the dominant conv kernel.
# usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_$TAG.log
tail -3 $O/pytest_$TAG.log
python bench.py --steps 20 --warmup 5 --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cat $O/bench_$TAG.json
# launch list: one eager step (graphs off so that every kernel is its own launch)
export PATCHGAN_B200_GRAPH=0
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_$TAG.log 2>&1
echo "ncu launches rc=$?"
# full capture of the dominant kernel (d3 forward shape: conv s1 B32 32x32 C256->N512), 3 launches
python tools/conv_probe.py conv 1 32 32 256 512 > $O/probe_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 3 -o $O/prof_$TAG -f \
    python tools/conv_probe.py conv 1 32 32 256 512 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
cat $O/probe_$TAG.log
