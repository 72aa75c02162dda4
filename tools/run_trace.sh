PG_TC_DEBUG=1 python tools/conv_trace.py ${1:-p} > gpurun_out/trace_${1:-p}.log 2>&1
grep "^conv_tc" gpurun_out/trace_${1:-p}.log | sort -u
grep -v "^conv_tc:" gpurun_out/trace_${1:-p}.log | cut -c 1-700
