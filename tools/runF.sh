mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_matvec.py -x -q --timeout 120 > gpurun_out/F_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/F_pytest.log
timeout 100 python tests/gpu_matvec_k16.py > gpurun_out/F_k16.log 2>&1
tail -12 gpurun_out/F_pytest.log; cat gpurun_out/F_k16.log
