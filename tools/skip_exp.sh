for sk in wgrad dstep gbwd "wgrad,dstep" "wgrad,dstep,gbwd,adam"; do
  PATCHGAN_B200_SKIP=$sk python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /tmp/o.json 2> /tmp/e.txt
  echo "skip=$sk rc=$? $(cut -c1-140 /tmp/o.json)"; tail -2 /tmp/e.txt | cut -c1-300
done
