#!/bin/bash
# CTA-pair kernel: parity tests, then A/B of the discriminator layer shapes with PG_TC_PAIR=1/0
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_a5_pair.py > gpurun_out/pair_tests.log 2>&1; echo "pair tests rc=$?"; tail -n 5 gpurun_out/pair_tests.log
for pr in 1 0; do
  echo "== PG_TC_PAIR=$pr"
  for sh in "conv 2 32 128 64 128" "conv 2 32 64 128 256" "conv 1 32 32 256 512" "convT 2 32 32 256 128" "convT 2 32 64 128 64" "convT 2 16 32 256 128"; do
    PROBE_OUT=f16 PROBE_ACT=2 PG_TC_PAIR=$pr timeout 120 python tools/conv_probe.py $sh 2>&1 | tail -n 1
  done
done
