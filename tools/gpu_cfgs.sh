#!/bin/bash
# Single-GPU bench lines of the other configurations (cfg4, cfg5 training; cfg2 inference).
mkdir -p gpurun_out
for c in cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc=$? $(cut -c1-130 gpurun_out/bench_$c.json)"
done
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 rc=$? $(cut -c1-130 gpurun_out/bench_cfg2.json)"
