#!/bin/bash
mkdir -p gpurun_out
for c in cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --detail gpurun_out/r2m_detail_$c.json > gpurun_out/r2m_$c.json 2> gpurun_out/r2m_$c.err; echo "$c rc=$? $(cut -c1-150 gpurun_out/r2m_$c.json)"; tail -n 3 gpurun_out/r2m_$c.err
done
timeout 300 python bench.py --config cfg2 --steps 20 --warmup 5 > gpurun_out/r2m_cfg2.json 2> gpurun_out/r2m_cfg2.err; echo "cfg2 rc=$? $(cut -c1-150 gpurun_out/r2m_cfg2.json)"
