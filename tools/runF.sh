mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_symeig.py -x -q > gpurun_out/F_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/F_pytest.log
XT_TRACE=1 timeout 120 python tools/trace_c2.py > gpurun_out/F_trace.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/F_bench.log 2>&1
tail -4 gpurun_out/F_pytest.log; grep barrier gpurun_out/F_trace.log | tail -4; tail -19 gpurun_out/F_trace.log | cut -c1-20,100-; tail -1 gpurun_out/F_bench.log | cut -c1-400
