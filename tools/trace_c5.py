import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle, xitorch_b200 as xt, time
n, neig = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 16
A = oracle.make_herm(n, neig, torch.float32, seed=7).cuda()
op = xt.LinearOperator.m(A, is_hermitian=True)
for i in range(3):
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev, _ = xt.linalg.symeig(op, neig=neig, mode="lowest", method="lanczos", min_eps=1e-4, info=info)
    torch.cuda.synchronize(); print("%.2f ms" % ((time.perf_counter() - t0) * 1e3), info)
