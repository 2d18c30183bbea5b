mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_symeig.py -x -q --timeout 120 > gpurun_out/F_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/F_pytest.log
XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/F_trace.log 2>&1
tail -6 gpurun_out/F_pytest.log; grep "device span" gpurun_out/F_trace.log | tail -2
