mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/P_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/P_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/P_bench.log 2>&1
XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/P_trace.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/P_launches.csv python tools/trace_c2.py > gpurun_out/P_ncu1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:expand_fused -s 9 -c 1 -o gpurun_out/P_expand10 -f python tools/trace_c2.py > gpurun_out/P_ncu2.log 2>&1
timeout 200 python tests/gpu_eigh_phases.py > gpurun_out/P_eigh_phases.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/P_bench_ref.log 2>&1
tail -3 gpurun_out/P_pytest.log; tail -1 gpurun_out/P_bench.log | cut -c1-300; tail -1 gpurun_out/P_bench_ref.log | cut -c1-400
