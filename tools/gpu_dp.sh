O=gpurun_out
N=${1:-2}
timeout -k 5 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
echo "N=$N rc=$? $(cut -c1-260 $O/bench_${N}gpu.json)"
grep -E "Error|rror:" $O/bench_${N}gpu.err | head -3 | cut -c1-300
timeout -k 5 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
