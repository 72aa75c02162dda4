#!/bin/bash
# round 2, call G (2 GPUs): data-parallel parity test (raw NCCL in one graph / process group split graphs) + 2-GPU bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 700 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_e_dp.py > gpurun_out/r2g_dp.log 2>&1; echo "dp tests rc=$?"
tail -n 25 gpurun_out/r2g_dp.log | cut -c1-300
for raw in 1 0; do
  PATCHGAN_B200_RAW_NCCL=$raw timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$raw bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2g_bench2_raw$raw.json 2> gpurun_out/r2g_bench2_raw$raw.err; echo "bench2 raw=$raw rc=$? $(cut -c1-120 gpurun_out/r2g_bench2_raw$raw.json)"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench1.json 2> gpurun_out/r2g_bench1.err; echo "bench1 rc=$? $(cut -c1-120 gpurun_out/r2g_bench1.json)"
