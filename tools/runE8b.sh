mkdir -p gpurun_out
W=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531"
timeout 150 $TR tests/gpu_dist_c5.py 65536 16 sharded 32 6 > gpurun_out/E2_w${W}_sharded32.log 2>&1; grep "^C5\|rror" gpurun_out/E2_w${W}_sharded32.log | cut -c1-450
XT_SHARDED_LAG=1 timeout 150 $TR tests/gpu_dist_c5.py 65536 16 sharded 32 6 > gpurun_out/E2_w${W}_lag1.log 2>&1; grep "^C5\|rror" gpurun_out/E2_w${W}_lag1.log | cut -c1-450
