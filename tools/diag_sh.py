import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle, xitorch_b200 as xt
from xitorch_b200 import dist as xd
for (n, neig, dt, eps, mb, keep) in [(2048, 4, torch.float64, 1e-8, None, None), (2048, 4, torch.float64, 1e-6, None, None),
                                     (2048, 4, torch.float64, 1e-8, 128, None), (2048, 4, torch.float64, 1e-8, 64, 32),
                                     (2048, 4, torch.float32, 1e-4, None, None), (2048, 8, torch.float64, 1e-8, None, None)]:
    A = oracle.make_herm(n, neig, dt, seed=11).cuda()
    info = {}
    ev, vec = xd.symeig_row_partitioned(A, n, neig, min_eps=eps, info=info, engine="sharded", max_basis=mb, restart_keep=keep, max_niter=300)
    ref = torch.linalg.eigvalsh(A.double())[:neig]
    info1 = {}
    ev1, _ = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="lanczos", min_eps=eps, info=info1, max_basis=mb)
    print(n, neig, dt, eps, mb, keep, info, "relerr %.2e" % ((ev.double() - ref).abs() / ref.abs()).max().item(), "| dense engine:", info1, flush=True)
