mkdir -p gpurun_out
W=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531"
timeout 200 $TR tests/gpu_dist_c5.py 65536 16 sharded 32 6 > gpurun_out/E_w${W}_sharded32.log 2>&1; grep "^C5\|rror" gpurun_out/E_w${W}_sharded32.log | cut -c1-450
timeout 300 $TR bench.py --gpus $W --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/E_w${W}_bench.log 2>&1; tail -1 gpurun_out/E_w${W}_bench.log | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/E_w${W}_bench.log').read().strip().splitlines()[-1])
    m=d.get('multi_gpu') or {}
    print({k:(v.get('ms_per_solve'), v.get('iters'), v.get('ms_per_application'), v.get('frac_of_aggregate_hbm_peak'), v.get('error')) for k,v in m.items()})
    print('e2e', d['e2e'], d['config'].get('host_cpus_bound_per_rank'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/E_w${W}_bench.log').read()[-2000:])
PY
