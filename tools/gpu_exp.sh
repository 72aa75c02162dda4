#!/bin/bash
# GPU tests once, then one short bench per environment setting given as arguments, e.g.
#   gpurun --timeout 600 -- 'timeout 560 bash tools/gpu_exp.sh TAG "" "PG_TC_DEFER=0" "PG_TC_NACC=1"'
TAG=$1; shift
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_$TAG.log | cut -c1-300
i=0
for envs in "$@"; do
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$i.json 2> $O/bench_${TAG}_$i.err; rc=$?
  echo "[$envs] rc=$rc $(python - <<PY
import json
try:
    d=json.load(open('$O/bench_${TAG}_$i.json'))
    print('img/s', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sync', d['e2e'].get('sync_batch_value'))
except Exception as ex:
    print('no json', ex)
PY
)"
  [ $rc -ne 0 ] && tail -5 $O/bench_${TAG}_$i.err
  i=$((i+1))
done
