"""
Generate tests/golden/optimize_golden.pt: outputs of the UNMODIFIED reference (imported from /root/reference, present
only in the build container) for the callers around the rootfinder path -- `newton`, `anderson_acc`, `minimize`
("broyden1", "newton", "gd", "adam"), object methods as `fcn` (EditableModule / nn.Module: first and second
derivatives with respect to tensors hidden in the object) and the degenerate-spectrum gradient of `exacteig`.

    python oracle/gen_golden_optimize.py

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

import xitorch as xt                                                     # noqa: E402  (the reference)
from xitorch.optimize import rootfinder, equilibrium, minimize           # noqa: E402
from xitorch.linalg import symeig                                        # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from optimize_cases import (rf_fcn, min_fcn, make_inputs, make_module, METHOD_CASES, module_loss,  # noqa: E402
                            degenerate_loss, first_and_second)

warnings.simplefilter("ignore")
torch.set_default_dtype(torch.float64)

out = {"methods": [], "modules": [], "degenerate": []}
api = {"rootfinder": rootfinder, "equilibrium": equilibrium, "minimize": minimize}
for (kind, method, opts) in METHOD_CASES:
    A, y0 = make_inputs(kind)
    Ar = A.clone().requires_grad_()
    fcn = min_fcn if kind == "minimize" else rf_fcn
    y = api[kind](fcn, y0, params=(Ar,), method=method,
                      bck_options={"method": "exactsolve"}, **opts)
    (g,) = torch.autograd.grad(y.sum(), Ar)
    out["methods"].append({"kind": kind, "method": method, "y": y.detach(), "grad_A": g})
    print("%-11s %-12s |y| = %.12f  |grad| = %.12f" % (kind, method, y.norm().item(), g.norm().item()))

for base in ("editable", "nn", "nn_in_editable"):
    for kind in ("rootfinder", "equilibrium", "minimize"):
        tensors = make_module(xt, base, kind)[1]
        loss = module_loss(xt, api[kind], base, kind, *tensors)
        grads, grads2 = first_and_second(loss, tensors)
        gsum = sum((g ** 2).sum() for g in grads)
        out["modules"].append({"base": base, "kind": kind, "loss": loss.detach(), "grads": grads, "grads2": grads2})
        print("%-15s %-11s loss = %.12f  |g| = %.6e  |g2| = %.6e"
              % (base, kind, loss.item(), gsum.sqrt().item(), sum((g ** 2).sum() for g in grads2).sqrt().item()))

for offset in (0.0, -4.0):
    a, mat, P2 = degenerate_loss.inputs(offset)
    loss = degenerate_loss(xt, symeig, a, mat, P2)
    grads = torch.autograd.grad(loss, (a, mat, P2))
    out["degenerate"].append({"offset": offset, "loss": loss.detach(), "grads": list(grads)})
    print("degenerate offset %+.0f  loss = %.12f" % (offset, loss.item()))

path = os.path.join(ROOT, "tests", "golden", "optimize_golden.pt")
torch.save(out, path)
print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024), "reference xitorch", xt.__version__)
