mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_rootfinder.py tests/test_gpu_baseline_sizes.py -m gpu -x -q --timeout 300 > gpurun_out/P_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/P_pytest.log; tail -5 gpurun_out/P_pytest.log
XT_NO_SOLVE_SLICES=1 timeout 300 python tests/gpu_bench_solve_pdl.py 2>&1 | sed 's/^pdl  /1slice/' > gpurun_out/P_one.log
timeout 300 python tests/gpu_bench_solve_pdl.py 2>&1 | sed 's/^pdl  /sliced/' > gpurun_out/P_sliced.log
paste -d'\n' gpurun_out/P_one.log gpurun_out/P_sliced.log
