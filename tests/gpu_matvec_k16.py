"""timing of the k = 16 block matvec layouts at N = 16384 (not a pytest file)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense
n = 16384
g = torch.Generator().manual_seed(0)
A = torch.randn(n, n, generator=g).cuda()
for k in (16, 12, 9):
    X = torch.randn(n, k, generator=g).cuda()
    ref = None
    for impl, name in ((6, "tensor-core 3xTF32"), (3, "SIMT row-slice"), (4, "SIMT column-slice")):
        y = _dense.block_matvec(A, X, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            _dense.block_matvec(A, X, impl=impl + (256 if i & 1 else 0))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if ref is None:
            ref = A.double() @ X.double(); scale = (A.double().abs() @ X.double().abs()).max().item()
        print("k=%2d %-22s %.1f us  %.0f GB/s   max err / (|A||x|) = %.2e" % (k, name, ms * 1e3, 4.0 * n * n / ms / 1e6,
              (y.double() - ref).abs().max().item() / scale))
