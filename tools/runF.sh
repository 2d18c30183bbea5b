mkdir -p gpurun_out
XT_TRACE=1 timeout 200 python tools/trace_c5.py 65536 > gpurun_out/F_trace5.log 2>&1
tail -24 gpurun_out/F_trace5.log | cut -c1-210
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from xitorch_b200 import _dense
n = 65536
A = torch.empty(n, n, device='cuda').normal_()
X = torch.randn(n, 16, device='cuda')
for impl in (0, 3, 4):
    y = _dense.block_matvec(A, X, impl=impl); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): _dense.block_matvec(A, X, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("N=65536 k=16 impl=%d: %.2f ms  %.0f GB/s" % (impl, ms, 4.0 * n * n / ms / 1e6))
PY
