mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/P_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/P_pytest.log; tail -3 gpurun_out/P_pytest.log
timeout 500 python bench.py --steps 30 --warmup 3 > gpurun_out/P_bench.log 2>&1; tail -1 gpurun_out/P_bench.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/P_bench_ref.log 2>&1; tail -1 gpurun_out/P_bench_ref.log | cut -c1-200
XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/P_trace.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/P_launches.csv python tools/trace_c2.py > gpurun_out/P_ncu1.log 2>&1
