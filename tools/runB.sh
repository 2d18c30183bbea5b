mkdir -p gpurun_out
timeout 100 python tools/check_eigh.py > gpurun_out/B_check.log 2>&1
timeout 300 python -m pytest tests/test_gpu_symeig.py -x -q --timeout 120 > gpurun_out/B_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/B_pytest.log
timeout 100 python tests/gpu_eigh_phases.py > gpurun_out/B_eigh_phases.log 2>&1
XT_LAG1_M=0 XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/B_trace_lag2.log 2>&1
XT_LAG1_M=96 XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/B_trace_lag1.log 2>&1
XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/B_trace_auto.log 2>&1
awk '{ if ($7+0 > 1e-10 || $9+0 > 1e-10) print }' gpurun_out/B_check.log | head -5; tail -3 gpurun_out/B_pytest.log; grep -v phases gpurun_out/B_eigh_phases.log
for f in lag2 lag1 auto; do echo "== $f"; grep -E "device span|niter" gpurun_out/B_trace_$f.log | tail -2; tail -19 gpurun_out/B_trace_$f.log | cut -c1-95 | head -14; done
