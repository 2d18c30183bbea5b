mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dist_c5.py 32768 16 > gpurun_out/H_c5_2gpu.log 2>&1
tail -4 gpurun_out/H_c5_2gpu.log
timeout 100 python tests/gpu_dist_c5.py 32768 16 > gpurun_out/H_c5_1gpu.log 2>&1
tail -3 gpurun_out/H_c5_1gpu.log
