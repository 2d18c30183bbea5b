mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/H_bench2.log 2>&1
tail -1 gpurun_out/H_bench2.log | cut -c1-420
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/H_bench2_ref.log 2>&1
tail -1 gpurun_out/H_bench2_ref.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dist_c5.py 32768 16 > gpurun_out/H_c5_2gpu.log 2>&1
tail -3 gpurun_out/H_c5_2gpu.log
