"""first-light script for the GPU box: matvec timing + quick solver sanity (not a pytest file)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense
import xitorch_b200 as xt

dev = "cuda"
print(torch.cuda.get_device_name(0))
for dtype, n, k in [(torch.float32, 16384, 8), (torch.float32, 16384, 1), (torch.float32, 16384, 2), (torch.float32, 16384, 4),
                    (torch.float32, 16384, 16), (torch.bfloat16, 16384, 1), (torch.bfloat16, 16384, 8),
                    (torch.float64, 8192, 8), (torch.float64, 8192, 1), (torch.float32, 4096, 1)]:
    A = torch.randn(n, n, device=dev).to(dtype)
    vdt = torch.float64 if dtype == torch.float64 else torch.float32
    X = torch.randn(n, k, device=dev, dtype=vdt)
    for impl in ((3, 4) if (dtype == torch.float32 and k >= 8) else (1,)):
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
        ref = (A.double()[-300:] @ X.double())
        err = (y[-300:].double() - ref).abs().max().item() / ref.abs().max().item()
        print("matvec %-8s n=%5d k=%2d impl=%d: %.3f ms  %.0f GB/s  relerr %.2e" % (str(dtype)[6:], n, k, impl, ms, gb / ms * 1e3, err), flush=True)

import oracle, time, warnings
# C2 first light
n, neig = 16384, 8
A = oracle.make_herm(n, neig, torch.float32).to(dev)
op = xt.LinearOperator.m(A, is_hermitian=True)
for method in ("davidson", "lanczos"):
    for rep in range(2):
        info = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ev, vec = xt.linalg.symeig(op, neig=neig, method=method, min_eps=1e-4, info=info)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(method, "N=16384 neig=8 fp32:", info, "%.2f ms -> %.0f it/s" % ((t1 - t0) * 1e3, info["niter"] / (t1 - t0)), ev.cpu().numpy(), flush=True)
# C1
A1 = oracle.make_spd_c1(256).to(dev)
B1 = torch.randn(256, 3, dtype=torch.float64, device=dev)
for method in ("cg", "bicgstab", "gmres"):
    info = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = xt.linalg.solve(xt.LinearOperator.m(A1, True), B1, method=method, posdef=True, info=info)
    print(method, "C1:", info, "resid", (A1 @ x - B1).norm().item() / B1.norm().item(), flush=True)
