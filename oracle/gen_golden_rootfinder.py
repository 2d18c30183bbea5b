"""
Generate tests/golden/rootfinder_golden.pt by running the UNMODIFIED reference rootfinder (imported from
/root/reference, present only in the build container) and check the oracle restatement against it.

    python oracle/gen_golden_rootfinder.py

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import os
import sys
import warnings

import torch

REF = os.environ.get("XITORCH_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import xitorch                                            # noqa: E402  (the reference)
from xitorch.optimize import rootfinder as ref_rootfinder, equilibrium as ref_equilibrium   # noqa: E402

import oracle                                             # noqa: E402

warnings.simplefilter("ignore")


def fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2.0


cases = []
for (n, dtype, method) in [(2, torch.float32, "broyden1"), (40, torch.float64, "broyden1"),
                           (128, torch.float64, "broyden1"), (40, torch.float64, "broyden2"),
                           (40, torch.float64, "linearmixing")]:
    if n == 2:
        A = torch.tensor([[1.1, 0.4], [0.3, 0.8]], dtype=dtype)          # the doctest of rootfinder.py:86-93
    else:
        A, _ = oracle.make_rootfinder_c4(n, dtype=dtype)
    y0 = torch.zeros(n, 1, dtype=dtype)
    Ar = A.clone().requires_grad_()
    y_ref = ref_rootfinder(fcn, y0, params=(Ar,), method=method)
    (g_ref,) = torch.autograd.grad(y_ref.sum(), Ar)                        # reference backward (bicgstab for n > 5)
    (g_exact,) = oracle.implicit_grad_dense(fcn, y_ref, (A,), torch.ones_like(y_ref))
    rec = {"n": n, "dtype": dtype, "method": method, "A": A, "y": y_ref.detach(), "grad_A": g_ref,
           "grad_A_exact": g_exact}
    if method == "broyden1":
        y_o, info = oracle.broyden1_root(fcn, y0, (A,), return_info=True)
        err = (y_o - y_ref.detach()).abs().max().item()
        print("n=%d %s  max|oracle-ref| = %.3e  (oracle niter %d, nfev %d)  |grad_ref-grad_exact| = %.2e"
              % (n, method, err, info["niter"], info["nfev"], (g_ref - g_exact).abs().max().item()))
        assert err <= 1e-12 * max(1.0, y_ref.abs().max().item()) or dtype == torch.float32 and err <= 1e-6
        rec["oracle_niter"] = info["niter"]
    cases.append(rec)
A = torch.tensor([[1.1, 0.4], [0.3, 0.8]])
ye = ref_equilibrium(fcn, torch.zeros(2, 1), params=(A,))
out = {"rootfinder": cases, "equilibrium_doc": {"A": A, "y": ye.detach()}}
path = os.path.join(ROOT, "tests", "golden", "rootfinder_golden.pt")
torch.save(out, path)
print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024), "reference xitorch", xitorch.__version__)
