O=gpurun_out
for c in cfg4 cfg5; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
  cut -c1-330 $O/bench_$c.json; tail -3 $O/bench_$c.err | cut -c1-300
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
