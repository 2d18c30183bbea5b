import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense
nr = nc = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A = torch.randn(nr, nc, device="cuda"); X = torch.randn(nc, 16, device="cuda")
for i in range(2): _dense.block_matvec(A, X, impl=7)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): _dense.block_matvec(A, X, impl=7)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("dbg=%s %d x %d: %.3f ms %.0f GB/s" % (os.environ.get("XT_TC5_DBG", "0"), nr, nc, ms, 4.0 * nr * nc / ms / 1e6), flush=True)
