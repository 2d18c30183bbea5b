mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_symeig.py -x -q --timeout 120 > gpurun_out/F_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/F_pytest.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/F_bench.log 2>&1
timeout 100 python tests/gpu_host_overhead.py > gpurun_out/F_host.log 2>&1
tail -2 gpurun_out/F_pytest.log; tail -1 gpurun_out/F_bench.log | cut -c60-170; head -40 gpurun_out/F_host.log
