#!/bin/bash
# In-graph timeline of a data-parallel step on N GPUs (rank 0's view).  gpurun --gpus N -- 'bash tools/gpu_dp_timeline.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tools/timeline.py > gpurun_out/dp_timeline_${N}gpu.txt 2> gpurun_out/dp_timeline_${N}gpu.err
echo "timeline rc=$?"; head -n 1 gpurun_out/dp_timeline_${N}gpu.txt; grep -E "Error|Traceback" gpurun_out/dp_timeline_${N}gpu.err | head -n 3
