"""host-side overhead of one symeig call on the GPU box (not a pytest file): cProfile over 40 solves."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import xitorch_b200 as xt

n, neig = 16384, 8
A = oracle.make_herm(n, neig, torch.float32, seed=7).cuda()
op = xt.LinearOperator.m(A, is_hermitian=True)


def solve():
    info = {}
    return xt.linalg.symeig(op, neig=neig, mode="lowest", method="davidson", min_eps=1e-4, info=info)


for _ in range(3):
    solve()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40):
    solve()
torch.cuda.synchronize()
print("wall per solve: %.1f us" % ((time.perf_counter() - t0) / 40 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    solve()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
