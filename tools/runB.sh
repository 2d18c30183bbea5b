mkdir -p gpurun_out
timeout 120 python tools/check_eigh.py > gpurun_out/B_check.log 2>&1
timeout 600 python -m pytest tests/test_gpu_symeig.py -x -q > gpurun_out/B_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/B_pytest.log
timeout 120 python tests/gpu_eigh_phases.py > gpurun_out/B_eigh_phases.log 2>&1
XT_TRACE=1 timeout 120 python tools/trace_c2.py > gpurun_out/B_trace.log 2>&1
grep -v "e-1[3-6]  resid .e-1[4-6]" gpurun_out/B_check.log | head; tail -5 gpurun_out/B_pytest.log; grep -v phases gpurun_out/B_eigh_phases.log; tail -19 gpurun_out/B_trace.log
