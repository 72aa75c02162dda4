#!/bin/bash
mkdir -p gpurun_out
for pr in 1 0; do
  echo "== PG_TC_PAIR=$pr"
  PROBE_OUT=f16 PG_TC_PAIR=$pr timeout 300 python tools/conv_trace.py d 2>&1 | tail -n 8
done
