"""L2 carry-over sweep (not a pytest file): alternate forward / reversed passes over the same A, keep the last
XT_MV_L2_KEEP_MB megabytes of each pass in L2 (evict-last) and time the pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense

n, k = 16384, 8
A = torch.randn(n, n, device="cuda")
X = torch.randn(n, k, device="cuda")
ref = A[-200:].double() @ X.double()
for alternate in (0, 1):
    for keep in (0, 16, 32, 48, 64, 80, 96, 112, 128):
        os.environ["XT_MV_L2_KEEP_MB"] = str(keep)
        for i in range(4):
            y = _dense.block_matvec(A, X, impl=3 + (256 if (alternate and i & 1) else 0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            y = _dense.block_matvec(A, X, impl=3 + (256 if (alternate and i & 1) else 0))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        err = ((y[-200:].double() - ref).abs().max() / ref.abs().max()).item()
        print("alternate=%d keep=%3d MB: %.1f us  %.0f GB/s (algorithmic)  relerr %.1e"
              % (alternate, keep, ms * 1e3, 4 * n * n / ms / 1e6, err), flush=True)
