O=gpurun_out
TAG=${1:-f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | cut -c1-200
python bench.py --steps 30 --warmup 5 --detail $O/detail_$TAG.json > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cut -c1-200 $O/bench_$TAG.json
python bench.py --config cfg2 --steps 20 --warmup 5 > $O/bench_cfg2_$TAG.json 2> $O/bench_cfg2_$TAG.err; echo "cfg2 rc=$?"; cat $O/bench_cfg2_$TAG.json | cut -c1-900; tail -3 $O/bench_cfg2_$TAG.err
python tools/conv_probe.py conv 1 32 32 256 512 > $O/probe_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 3 -o $O/prof_$TAG -f \
    python tools/conv_probe.py conv 1 32 32 256 512 > $O/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; cat $O/probe_$TAG.log
ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_${TAG}_raw.csv 2>/dev/null
rm -f $O/prof_$TAG.ncu-rep
