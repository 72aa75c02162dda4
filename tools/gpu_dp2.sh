#!/bin/bash
# Data-parallel check on N GPUs: the 2-rank parity tests, rank 0's in-graph timeline, a bench line.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_e_dp.py > gpurun_out/dp_tests.log 2>&1; echo "dp tests rc=$?"; tail -n 3 gpurun_out/dp_tests.log
bash tools/gpu_dp_timeline.sh $N
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>gpurun_out/dp_bench.err | tee gpurun_out/dp_bench_${N}gpu.json | cut -c1-140
