"""timing of the linear solvers on one GPU (not a pytest file): C1 and large dense SPD systems."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
import xitorch_b200 as xt
from xitorch_b200 import _lib
dev = "cuda"
def run(tag, A, B, method, herm, **opts):
    op = xt.LinearOperator.m(A, is_hermitian=herm)
    for rep in range(3):
        info = {}
        _lib.profile_reset(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x = xt.linalg.solve(op, B, method=method, info=info, **opts)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        mv_ms, n_mv, n_l = _lib.profile_read(); _lib.profile_reset(False)
    res = ((A.double() @ x.double() - B.double()).norm() / B.double().norm()).item()
    per = A.numel() * A.element_size()
    print("%-34s %-8s niter=%3d conv=%s  %.2f ms  %.3f ms/iter  matvec avg %.3f ms (%.0f GB/s, %d launches, %d kernels)  relres %.1e"
          % (tag, method, info["niter"], info["converged"], (t1 - t0) * 1e3, (t1 - t0) * 1e3 / max(info["niter"], 1),
             mv_ms / max(n_mv, 1), per / (mv_ms / max(n_mv, 1) * 1e-3) / 1e9 if n_mv else 0, n_mv, n_l, res), flush=True)
A1 = oracle.make_spd_c1(256).to(dev); torch.manual_seed(123); B1 = (A1 @ torch.randn(256, 3, dtype=torch.float64, device=dev))
for m in ("cg", "bicgstab", "gmres"):
    run("C1 n=256 fp64 ncols=3", A1, B1, m, True, posdef=True)
A = oracle.make_herm(16384, 8, torch.float32).to(dev)
g = torch.Generator(device=dev); g.manual_seed(5)
for nc in (1, 8):
    B = torch.randn(16384, nc, device=dev, generator=g)
    for m in ("cg", "bicgstab", "gmres"):
        run("n=16384 fp32 ncols=%d (cond~30)" % nc, A, B, m, True, posdef=True, rtol=1e-6)
