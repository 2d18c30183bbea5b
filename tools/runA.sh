set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/A_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/A_pytest.log
python tests/gpu_eigh_phases.py > gpurun_out/A_eigh_phases.log 2>&1
XT_TRACE=1 python - > gpurun_out/A_trace.log 2>&1 <<'PY'
import torch, oracle, xitorch_b200 as xt
n, neig = 16384, 8
A = oracle.make_herm(n, neig, torch.float32, seed=7).cuda()
op = xt.LinearOperator.m(A, is_hermitian=True)
for i in range(3):
    info = {}
    xt.linalg.symeig(op, neig=neig, mode="lowest", method="davidson", min_eps=1e-4, info=info)
    torch.cuda.synchronize()
    print(info)
PY
python bench.py --steps 20 --warmup 3 > gpurun_out/A_bench.log 2>&1
tail -3 gpurun_out/A_pytest.log; cat gpurun_out/A_eigh_phases.log; tail -40 gpurun_out/A_trace.log; tail -2 gpurun_out/A_bench.log
