"""ATen operations and host synchronisations per function evaluation of the Broyden rootfinder (device-independent:
the same Python drives CPU and CUDA tensors).  python tools/count_rootfinder_ops.py [n]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.utils._python_dispatch import TorchDispatchMode
import oracle
from xitorch_b200.optimize import rootfinder
from xitorch_b200._impls import rootsolver

class Counter(TorchDispatchMode):
    def __init__(self):
        super().__init__(); self.c = collections.Counter()
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        self.c[str(func)] += 1
        return func(*args, **(kwargs or {}))

def fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A, _ = oracle.make_rootfinder_c4(n, dtype=torch.float64)
y0 = torch.zeros(n, 1, dtype=torch.float64)
nfev = [0]
def f2(y, A):
    nfev[0] += 1
    return fcn(y, A)
with Counter() as c:
    with torch.no_grad():
        y = rootsolver.broyden1(f2, y0, (A,))
tot = sum(c.c.values())
print("nfev", nfev[0], "aten ops", tot, "per eval %.1f" % (tot / nfev[0]))
for k, v in c.c.most_common(25): print("%6d %s" % (v, k))
