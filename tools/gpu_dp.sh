O=gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_$name.json 2> $O/bench_2gpu_$name.err
  echo "$name rc=$? $(cut -c1-200 $O/bench_2gpu_$name.json)"
  grep -E "Error|error|rror:" $O/bench_2gpu_$name.err | head -3 | cut -c1-300
}
run default X=1
