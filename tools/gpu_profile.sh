#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, bench line (ours + CPU arm), ncu launch list of an eager
# step, ncu --set full of the step's main kernels.  tools/summarize_ncu.py turns the CSVs into profiles/r02_*.json.
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 5 --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 $O/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-300 $O/bench_ref_$TAG.json
timeout 300 python tools/timeline.py > $O/timeline_$TAG.txt 2>&1; echo "timeline rc=$?"
timeout 300 python tools/step_once.py 4 > $O/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
    --log-file $O/launches_$TAG.csv python tools/step_once.py 4 > $O/ncu_$TAG.log 2>&1
echo "ncu launches rc=$?"; cat $O/plain_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'conv_res_kernel|wgrad_group_kernel|wgrad_tc_kernel|conv_tc_pers_kernel|conv_tc_kernel|adam_kernel|seg_loss_partials|taps_dgrad_act|prep_batch|gen_out_bwd|im2col_s2' \
    -c 110 -o $O/full_$TAG -f python tools/step_once.py 1 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $O/full_$TAG.ncu-rep --page raw --csv > $O/full_${TAG}_raw.csv 2>/dev/null
rm -f $O/full_$TAG.ncu-rep
ls -la $O | grep $TAG
