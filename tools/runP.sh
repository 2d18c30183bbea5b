mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/P_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/P_pytest.log; tail -3 gpurun_out/P_pytest.log
timeout 500 python bench.py --steps 30 --warmup 3 > gpurun_out/P_bench.log 2>&1; tail -1 gpurun_out/P_bench.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/P_bench_ref.log 2>&1; tail -1 gpurun_out/P_bench_ref.log | cut -c1-300
XT_TRACE=1 timeout 100 python tools/trace_c2.py > gpurun_out/P_trace.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/P_launches.csv python tools/trace_c2.py > gpurun_out/P_ncu1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mv_tma_kernel -s 20 -c 1 -o gpurun_out/P_mv -f python tools/trace_c2.py > gpurun_out/P_ncu2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:expand_fused -s 22 -c 1 -o gpurun_out/P_expand -f python tools/trace_c2.py > gpurun_out/P_ncu3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:rr_kernel -s 23 -c 1 -o gpurun_out/P_rr -f python tools/trace_c2.py > gpurun_out/P_ncu4.log 2>&1
timeout 200 python tests/gpu_eigh_phases.py > gpurun_out/P_eigh_phases.log 2>&1
ls -la gpurun_out/*.ncu-rep
