PG_TC_DEBUG=1 python tools/conv_trace.py p > gpurun_out/trace_p.log 2>&1
grep "^conv_tc" gpurun_out/trace_p.log | sort -u
grep -v "^conv_tc:" gpurun_out/trace_p.log | cut -c 1-330
