PG_TC_DEBUG=1 python tools/conv_trace.py > gpurun_out/trace_split.log 2>&1
grep "splits [2-8]" gpurun_out/trace_split.log | sort -u
grep -v "^conv_tc:" gpurun_out/trace_split.log | sed -n 4,7p | cut -c 1-130,200-560
