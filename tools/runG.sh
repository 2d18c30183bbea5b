mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_rootfinder.py -x -q --timeout 100 > gpurun_out/G_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/G_pytest.log
timeout 150 python tests/gpu_bench_c4.py > gpurun_out/G_c4.log 2>&1
timeout 200 python -m pytest tests/test_gpu_solve.py -x -q --timeout 100 > gpurun_out/G_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/G_pytest2.log
tail -25 gpurun_out/G_pytest.log; tail -5 gpurun_out/G_c4.log; tail -3 gpurun_out/G_pytest2.log
