python tools/conv_trace.py 2>&1 | grep -v "^conv_tc:"
