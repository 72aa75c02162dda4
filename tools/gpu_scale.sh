#!/bin/bash
# Data-parallel bench lines for every config on N GPUs of one box:  gpurun --gpus N -- 'bash tools/gpu_scale.sh N [cfgs...]'
# Writes gpurun_out/scale_<cfg>_<N>gpu.json (one JSON line each); cfg3 is also run through the process-group
# (non-graph-captured) all-reduce path as scale_cfg3_pg_<N>gpu.json for comparison.
N=${1:-2}; shift
CFGS=${@:-cfg3 cfg4 cfg5}
mkdir -p gpurun_out
run() {  # tag, extra env, bench args
  local tag=$1 envs=$2; shift 2
  env $envs timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-infer "$@" \
      > gpurun_out/scale_${tag}_${N}gpu.json 2> gpurun_out/scale_${tag}_${N}gpu.err
  echo "$tag N=$N rc=$? $(cut -c1-130 gpurun_out/scale_${tag}_${N}gpu.json)"
  grep -E "Error|error|Traceback|timeout" gpurun_out/scale_${tag}_${N}gpu.err | head -n 3
}
for c in $CFGS; do
  run $c "X=1" --config $c
done
case " $CFGS " in *" cfg3 "*) run cfg3_pg "PATCHGAN_B200_RAW_NCCL=0" --config cfg3;; esac
