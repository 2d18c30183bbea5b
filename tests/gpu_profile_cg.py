"""a few cg / bicgstab iterations on one GPU for a kernel launch list under ncu (not a pytest file).
N, NCOLS, METHOD, NITER from the environment."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
import xitorch_b200 as xt
n, nc = int(os.environ.get("N", "1024")), int(os.environ.get("NCOLS", "8"))
method, niter = os.environ.get("METHOD", "cg"), int(os.environ.get("NITER", "12"))
A = oracle.make_herm(n, 8, torch.float32).cuda()
A = A + (abs(torch.linalg.eigvalsh(A.double().cpu())[0].item()) + 1.0) * torch.eye(n, device="cuda")
B = torch.randn(n, nc, device="cuda")
op = xt.LinearOperator.m(A, is_hermitian=True)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(2):
        x = xt.linalg.solve(op, B, method=method, posdef=True, rtol=1e-30, atol=0.0, max_niter=niter)
torch.cuda.synchronize()
