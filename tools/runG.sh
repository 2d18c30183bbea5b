mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solve.py tests/test_gpu_rootfinder.py -x -q --timeout 100 > gpurun_out/G_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/G_pytest.log
tail -25 gpurun_out/G_pytest.log
