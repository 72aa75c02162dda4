#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() {  # tag envs args
  local tag=$1 envs=$2; shift 2
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-infer "$@" \
      > gpurun_out/scale_${tag}_${N}gpu.json 2> gpurun_out/scale_${tag}_${N}gpu.err
  echo "$tag N=$N rc=$? $(cut -c1-130 gpurun_out/scale_${tag}_${N}gpu.json)"
  grep -E "Error|Traceback|timeout" gpurun_out/scale_${tag}_${N}gpu.err | head -n 2
}
run cfg3 "X=1" --config cfg3
if [ "$N" = "8" ]; then run cfg3_sms32 "PATCHGAN_B200_NCCL_SMS=32" --config cfg3; fi
run cfg4 "X=1" --config cfg4
run cfg5 "X=1" --config cfg5
