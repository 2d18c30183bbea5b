"""first-light script for the GPU box: matvec timing + quick solver sanity (not a pytest file)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense
import xitorch_b200 as xt

dev = "cuda"
print(torch.cuda.get_device_name(0))
for dtype, n, k in [(torch.float32, 16384, 8), (torch.float32, 16384, 1), (torch.float32, 16384, 16),
                    (torch.bfloat16, 16384, 1), (torch.float64, 8192, 8), (torch.float32, 4096, 1)]:
    A = torch.randn(n, n, device=dev).to(dtype)
    vdt = torch.float64 if dtype == torch.float64 else torch.float32
    X = torch.randn(n, k, device=dev, dtype=vdt)
    for impl in (1, 2):
        for _ in range(3):
            y = _dense.block_matvec(A, X, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            y = _dense.block_matvec(A, X, impl=impl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gb = A.numel() * A.element_size() / 1e9
        ref = (A.double()[:64] @ X.double())
        err = (y[:64].double() - ref).abs().max().item() / ref.abs().max().item()
        print("matvec %-8s n=%5d k=%2d impl=%d: %.3f ms  %.0f GB/s  relerr %.2e" % (str(dtype)[6:], n, k, impl, ms, gb / ms * 1e3, err), flush=True)
