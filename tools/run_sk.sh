python tools/skinny_probe.py > gpurun_out/skprobe.log 2>&1 && cat gpurun_out/skprobe.log &&
ncu --set full --clock-control none --import-source on -k regex:"fewout|fewin|wgrad1" -c 7 -o gpurun_out/prof_sk -f python tools/skinny_probe.py > gpurun_out/ncu_sk.log 2>&1
echo "ncu rc=$?"
PROBE_IMPL=2 python tools/skinny_probe.py 2>&1 | sed 's/^/TC: /'
