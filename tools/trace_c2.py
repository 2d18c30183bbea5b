import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle, xitorch_b200 as xt
n, neig = 16384, 8
A = oracle.make_herm(n, neig, torch.float32, seed=7).cuda()
op = xt.LinearOperator.m(A, is_hermitian=True)
for i in range(3):
    info = {}
    ev, _ = xt.linalg.symeig(op, neig=neig, mode="lowest", method="davidson", min_eps=1e-4, info=info)
    torch.cuda.synchronize()
    print(info, ev.cpu().numpy())
