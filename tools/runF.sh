mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_symeig.py -x -q --timeout 120 > gpurun_out/F_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/F_pytest.log
XT_TRACE=1 timeout 200 python tools/trace_c5.py 65536 > gpurun_out/F_trace5.log 2>&1
tail -3 gpurun_out/F_pytest.log; grep -E "ms \{" gpurun_out/F_trace5.log; tail -20 gpurun_out/F_trace5.log | cut -c1-60 | sed -n 6,11p
