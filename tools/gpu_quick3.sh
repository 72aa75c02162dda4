O=gpurun_out
for m in 4 2 1; do
PG_TC_PERSIST_MIN=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --detail $O/detail_pers$m.json 2>/dev/null | cut -c1-130
done
