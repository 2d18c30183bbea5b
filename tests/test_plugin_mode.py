"""Plug-in mode: our method callables handed to the UNMODIFIED upstream xitorch through its
`method=` hook.  Needs the reference checkout (present only in the build container) -- skipped elsewhere.
On CPU the callables must be reached (signature accepted by upstream) and then refuse loudly (no CPU path);
on a GPU box with the reference installed the same test runs the kernels under upstream's autograd Functions."""
import os
import sys

import pytest
import torch

REF = os.environ.get("XITORCH_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "xitorch")), reason="reference checkout not present")


def _upstream():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import xitorch
    import xitorch.linalg
    return xitorch


def test_upstream_accepts_our_callables():
    up = _upstream()
    import xitorch_b200._impls.solve as bs
    import xitorch_b200._impls.symeig as be
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(0)
    n = 64
    A = torch.randn(n, n, dtype=torch.float64)
    A = (A @ A.t() / n + torch.eye(n, dtype=torch.float64)).to(dev)
    B = torch.randn(n, 2, dtype=torch.float64, device=dev)
    op = up.LinearOperator.m(A, is_hermitian=True)              # upstream MatrixLinearOperator
    if dev == "cpu":
        with pytest.raises(RuntimeError, match="CUDA"):
            up.linalg.solve(op, B, method=bs.cg)
        with pytest.raises(RuntimeError, match="CUDA"):
            up.linalg.symeig(op, neig=2, method=be.davidson)
        return
    x = up.linalg.solve(op, B, method=bs.cg, rtol=1e-10, bck_options={"method": bs.cg, "rtol": 1e-10})
    assert torch.allclose(A @ x, B, atol=1e-7)
    ev, vec = up.linalg.symeig(op, neig=2, method=be.davidson, min_eps=1e-8)
    assert torch.allclose(A @ vec, vec * ev, atol=1e-6)


def test_upstream_autograd_around_our_methods_with_standin_engine(monkeypatch):
    """plug-in mode end to end on CPU: upstream's autograd Functions (forward, backward re-entering `solve`) around our
    method callables, the CUDA library replaced by tests/standin_engine.py -- upstream's LinearOperator objects reach
    our host layer by duck typing (`fullmatrix` / `mm`)."""
    import standin_engine
    up = _upstream()
    import xitorch_b200._impls.solve as bs
    import xitorch_b200._impls.symeig as be
    eng = standin_engine.install(monkeypatch)
    torch.manual_seed(1)
    n = 6
    A0 = torch.randn(n, n, dtype=torch.float64)
    A0 = (A0 @ A0.t() / n + torch.eye(n, dtype=torch.float64)).requires_grad_()
    B0 = torch.randn(n, 2, dtype=torch.float64).requires_grad_()

    def solve_fcn(A, B):
        As = (A + A.t()) * 0.5
        return up.linalg.solve(up.LinearOperator.m(As, is_hermitian=True), B, method=bs.cg,
                               bck_options={"method": bs.bicgstab})

    assert torch.autograd.gradcheck(solve_fcn, (A0, B0))
    assert {r["method"] for r in eng.log} == {"cg", "bicgstab"}

    def eig_fcn(A):
        As = (A + A.t()) * 0.5
        ev, vec = up.linalg.symeig(up.LinearOperator.m(As, is_hermitian=True), neig=2, method=be.davidson,
                                   bck_options={"method": bs.cg})
        return ev, vec ** 2

    assert torch.autograd.gradcheck(eig_fcn, (A0,))
    assert eng.log[-1]["method"] in ("cg", "symeig")
