#!/bin/bash
# ncu --set full of the launches of one (regex) kernel family inside eager training steps; details CSV comes back.
#   bash tools/gpu_ncu_kernel.sh TAG 'pack_weight_multi|grad_finalize_multi' [skip] [count]
TAG=$1; RE=$2; SKIP=${3:-4}; CNT=${4:-4}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $O/k_$TAG -f \
    python tools/step_once.py 4 > $O/ncu_k_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i $O/k_$TAG.ncu-rep --page details --csv > $O/k_${TAG}_details.csv 2>/dev/null
ncu -i $O/k_$TAG.ncu-rep --page source --csv > $O/k_${TAG}_source.csv 2>/dev/null
rm -f $O/k_$TAG.ncu-rep
ls -la $O | grep k_$TAG
