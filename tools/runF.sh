mkdir -p gpurun_out
XT_TRACE=1 timeout 100 python tools/trace_c5.py 32768 > gpurun_out/F_trace5.log 2>&1
tail -26 gpurun_out/F_trace5.log | cut -c1-200
