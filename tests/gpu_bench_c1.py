"""BASELINE configs[0] (cg on a 256 x 256 SPD fp64 operator): time per iteration of the on-chip cluster kernel against the
general two-launch path.   python tests/gpu_bench_c1.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle, xitorch_b200 as xt
n = 256
A = oracle.make_spd_c1(n)
torch.manual_seed(123)
B = A @ torch.randn(n, 3, dtype=torch.float64)
op = xt.LinearOperator.m(A.cuda(), is_hermitian=True)
Bd = B.cuda()
for label, env in (("on-chip cluster kernel", None), ("general path (matvec + step kernels)", "1")):
    if env: os.environ["XT_NO_SMALL_CG"] = env
    info = {}
    for _ in range(5):
        xt.linalg.solve(op, Bd, method="cg", posdef=True, rtol=1e-10, atol=1e-12, info=info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        xt.linalg.solve(op, Bd, method="cg", posdef=True, rtol=1e-10, atol=1e-12, info=info)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("C1 cg n=256 fp64 ncols=3 %-40s %d iterations  %.3f ms per solve  %.2f us per iteration" %
          (label, info["niter"], ms, ms * 1e3 / info["niter"]), flush=True)
