"""tcgen05 block matvec (impl = 7, fp32 k = 16): accuracy against fp64 and timing against the SIMT layouts.
    python tests/gpu_tc5.py [quick]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense

torch.manual_seed(0)
dev = "cuda"

def check(nr, nc, impl=7, E=False):
    A = torch.randn(nr, nc, device=dev)
    X = torch.randn(nc, 16, device=dev)
    y = _dense.block_matvec(A, X, impl=impl)
    torch.cuda.synchronize()
    ref = A.double() @ X.double()
    bound = (A.double().abs() @ X.double().abs())
    err = ((y.double() - ref).abs() / bound).max().item()
    print("  %6d x %6d impl=%d: max err / (|A||x|) = %.2e  %s" % (nr, nc, impl, err, "OK" if err <= 2e-6 else "FAIL"), flush=True)
    return err

print("accuracy:")
for nr, nc in [(128, 64), (128, 256), (256, 1024), (1000, 1000), (4096, 4096), (8192, 16384), (148 * 128 + 5, 2048 + 40)]:
    check(nr, nc)
    check(nr, nc, impl=3)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(0)
print("timing (A larger than L2):")
for nr, nc in [(16384, 16384), (8192, 65536), (65536, 65536)]:
    A = torch.randn(nr, nc, device=dev)
    X = torch.randn(nc, 16, device=dev)
    for impl, name in ((7, "tcgen05 3xTF32"), (3, "SIMT row-slice"), (0, "auto")):
        for i in range(3):
            _dense.block_matvec(A, X, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for i in range(reps):
            _dense.block_matvec(A, X, impl=impl + (256 if i & 1 else 0))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("  %6d x %6d %-16s %.3f ms  %.0f GB/s" % (nr, nc, name, ms, 4.0 * nr * nc / (ms * 1e-3) / 1e9), flush=True)
    del A, X
