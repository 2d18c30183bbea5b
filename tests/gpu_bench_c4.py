"""BASELINE config 4 on the GPU box (not a pytest file): rootfinder broyden1 on tanh(A@y+0.1)+y/2, y in R^8192, with the
backward adjoint solve; prints timing next to the CPU oracle (reference algorithm) on a smaller sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import xitorch_b200 as xt
from xitorch_b200.optimize import rootfinder
from xitorch_b200 import _lib

def fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2.0

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
DT = torch.float64      # fp32 cannot reach the default f_tol = 1e-6 at this size (rounding floor of |f|)
A_cpu, y0_cpu = oracle.make_rootfinder_c4(n, dtype=DT)
A = A_cpu.cuda().requires_grad_()
y0 = y0_cpu.cuda()
nfev = [0]
def fcn_count(y, A):
    nfev[0] += 1
    return fcn(y, A)
for rep in range(3):
    nfev[0] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = rootfinder(fcn_count, y0, params=(A,), maxiter=1000)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    nf_fwd = nfev[0]
    _lib.profile_reset(False)
    (g,) = torch.autograd.grad(y.sum(), A)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    _, _, nlaunch = _lib.profile_read()
print("C4 n=%d fp64: forward %.1f ms (%d fcn evals, %.0f evals/s, A-read %.0f GB/s), backward %.1f ms (%d fcn evals incl. VJPs, %d of our launches); |f(y)| = %.2e"
      % (n, (t1 - t0) * 1e3, nf_fwd, nf_fwd / (t1 - t0), nf_fwd * 8 * n * n / (t1 - t0) / 1e9, (t2 - t1) * 1e3, nfev[0] - nf_fwd, nlaunch,
         fcn(y.detach(), A.detach()).norm().item()))
# CPU oracle on the same problem (reference algorithm, all host threads)
torch.set_num_threads(os.cpu_count())
t0 = time.perf_counter()
y_o, info = oracle.broyden1_root(fcn, y0_cpu, (A_cpu,), return_info=True, maxiter=1000)
t1 = time.perf_counter()
print("CPU oracle (%d threads): forward %.1f ms (%d fcn evals, %.0f evals/s); |y_gpu - y_cpu|/|y| = %.2e"
      % (os.cpu_count(), (t1 - t0) * 1e3, info["nfev"], info["nfev"] / (t1 - t0), ((y.detach().cpu() - y_o).norm() / y_o.norm()).item()))
