mkdir -p gpurun_out
W=${1:-1}
if [ "$W" = "1" ]; then
  timeout 280 python tests/gpu_dist_c5.py 65536 16 > gpurun_out/C5_w1.log 2>&1
else
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29521 tests/gpu_dist_c5.py 65536 16 > gpurun_out/C5_w$W.log 2>&1
fi
tail -3 gpurun_out/C5_w$W.log
