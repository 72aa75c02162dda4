#!/bin/bash
# DP A/B on N GPUs: SM-reservation toggle on/off, NCCL CTA budgets
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_e_dp.py > gpurun_out/dp_tests.log 2>&1; echo "dp tests rc=$?"; tail -n 3 gpurun_out/dp_tests.log
fi
run() {
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-infer 2>gpurun_out/dp3.err | cut -c1-125
  echo "   ^ $1 rc=${PIPESTATUS[0]}"; grep -E "Error|Traceback|timeout" gpurun_out/dp3.err | head -n 2
}
run "PATCHGAN_B200_DP_SMTOGGLE=1"
run "PATCHGAN_B200_DP_SMTOGGLE=0"
run "PATCHGAN_B200_DP_SMTOGGLE=1 PATCHGAN_B200_NCCL_SMS=32"
run "PATCHGAN_B200_DP_SMTOGGLE=1 PATCHGAN_B200_NCCL_SMS=24"
