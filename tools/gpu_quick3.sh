O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_a_ops.py tests/test_gpu_a2_skinny.py -x -q 2>&1 | tail -4 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_b_models.py tests/test_gpu_c_step.py -x -q 2>&1 | tail -4 | cut -c1-300
PG_TC_DEBUG=1 timeout 120 python tools/skinny_probe.py 2>&1 | grep -c persistent
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_pers.json > $O/bench_pers.json 2> $O/bench_pers.err; echo "bench rc=$?"
cut -c1-200 $O/bench_pers.json; tail -3 $O/bench_pers.err | cut -c1-300
PG_TC_PERSIST=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-130
