#!/bin/bash
# One gpurun call: GPU tests, bench lines (ours + reference arm), ncu launch list of one eager step (duration + DRAM bytes
# of every launch), ncu --set full of the dominant tensor-core kernel (3 launches: the largest discriminator layer).
#   usage (here): gpurun --timeout 1800 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_$TAG.log
python bench.py --steps 30 --warmup 5 --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cat $O/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-300 $O/bench_ref_$TAG.json
# launch list: 4 eager steps (graphs / side streams off so that every kernel is its own serial launch); the last step is kept
python tools/step_once.py 4 > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file $O/launches_$TAG.csv python tools/step_once.py 4 > $O/ncu_$TAG.log 2>&1
echo "ncu launches rc=$?"; cat $O/plain_$TAG.log
# full capture of the dominant kernel at its largest shape (discriminator conv 256 -> 512, stride 1, B32)
python tools/conv_probe.py conv 1 32 32 256 512 > $O/probe_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 3 -o $O/prof_$TAG -f \
    python tools/conv_probe.py conv 1 32 32 256 512 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; cat $O/probe_$TAG.log
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $O/prof_$TAG.ncu-rep --page details --csv > $O/prof_${TAG}_details.csv 2>/dev/null
ls -la $O | tail -5
