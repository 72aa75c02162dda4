#!/bin/bash
# ncu --set full with source counters of ONE launch of the conv kernel for a probe shape; CSV pages come back in gpurun_out/
#   bash tools/gpu_ncu_src.sh TAG conv1x1 1 32 128 64 64
TAG=$1; shift
O=gpurun_out
mkdir -p $O
python tools/conv_probe.py "$@" > $O/probe_$TAG.log 2>&1 || { cat $O/probe_$TAG.log; exit 1; }
cat $O/probe_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o $O/src_$TAG -f \
    python tools/conv_probe.py "$@" > $O/ncu_src_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i $O/src_$TAG.ncu-rep --page source --csv > $O/src_${TAG}_source.csv 2>/dev/null
ncu -i $O/src_$TAG.ncu-rep --page details --csv > $O/src_${TAG}_details.csv 2>/dev/null
rm -f $O/src_$TAG.ncu-rep
ls -la $O | grep src_$TAG
