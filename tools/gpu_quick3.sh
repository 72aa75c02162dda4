O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_a_ops.py -x -q -k "tap_major" 2>&1 | tail -4 | cut -c1-300
PG_TC_DEBUG=1 timeout 100 python tools/wgrad_trace.py 2>&1 | grep "NT 2" | sort -u | head -5
timeout 300 python -m pytest tests/test_gpu_c_step.py -x -q 2>&1 | tail -3 | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_nt2.json > $O/bench_nt2.json 2> $O/bench_nt2.err; echo "bench rc=$?"
cut -c1-140 $O/bench_nt2.json; tail -3 $O/bench_nt2.err | cut -c1-300
PG_WG_NT2=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
