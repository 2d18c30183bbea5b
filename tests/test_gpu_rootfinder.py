"""GPU parity of the rootfinder path (BASELINE config 4) and of the matrix-free Krylov solves behind its backward:
CUDA result vs the CPU oracle on the same seeded inputs, vs the committed reference outputs, and vs the exact implicit
gradient (dense Jacobian, direct solve)."""
import os
import warnings

import pytest
import torch

import oracle
import xitorch_b200 as xt
from xitorch_b200.optimize import rootfinder
from xitorch_b200._impls.rootsolver import LowRankMatrix

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2.0


class _UserOp(xt.LinearOperator):
    """matrix-free operator in the style of the reference's ALarge test operator (test_linop_fcns.py:129-176):
    diagonal + cyclic neighbours, optionally made non-symmetric"""

    def __init__(self, diag, off, asym=0.0):
        super().__init__(shape=(diag.numel(), diag.numel()), is_hermitian=(asym == 0.0), dtype=diag.dtype,
                         device=diag.device)
        self.diag, self.off, self.asym = diag, off, asym

    def _mv(self, x):
        return x * self.diag + self.off * ((1 + self.asym) * torch.roll(x, 1, -1) + (1 - self.asym) * torch.roll(x, -1, -1))

    def _rmv(self, x):
        return x * self.diag + self.off * ((1 - self.asym) * torch.roll(x, 1, -1) + (1 + self.asym) * torch.roll(x, -1, -1))

    def _getparamnames(self, prefix=""):
        return [prefix + "diag"]


@pytest.mark.parametrize("method", ["cg", "bicgstab", "gmres"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_matrix_free_solve(method, dtype):
    n = 300
    g = torch.Generator().manual_seed(5)
    diag = (2.0 + torch.rand(n, generator=g, dtype=torch.float64)).to(dtype).to(DEV)
    asym = 0.0 if method == "cg" else 0.3
    op = _UserOp(diag, 0.2, asym)
    B = torch.randn(n, 3, generator=g, dtype=torch.float64).to(dtype).to(DEV)
    info = {}
    rtol = 1e-10 if dtype == torch.float64 else 1e-5
    with torch.no_grad():
        x = xt.linalg.solve(op, B, method=method, rtol=rtol, atol=1e-14, posdef=True, info=info)
        dense = op.fullmatrix().double()
    assert info["converged"] and info["matrix_free"] and info["napply"] >= 2
    x_ref = torch.linalg.solve(dense, B.double())
    err = ((x.double() - x_ref).norm() / x_ref.norm()).item()
    assert err <= (1e-8 if dtype == torch.float64 else 2e-4), err
    # the CPU oracle (reference algorithm) on the same operator needs about as many operator applications
    if method != "gmres":
        xo, oinfo = getattr(oracle, method)(oracle.DenseOp(dense.cpu(), asym == 0.0), B.double().cpu(), posdef=True,
                                            rtol=rtol, atol=1e-14, return_info=True)
        assert abs(info["niter"] - oinfo["niter"]) <= max(3, oinfo["niter"] // 4)


def test_matrix_free_solve_with_E_and_M():
    n = 200
    g = torch.Generator().manual_seed(6)
    diag = (2.0 + torch.rand(n, generator=g, dtype=torch.float64)).to(DEV)
    op = _UserOp(diag, 0.2, 0.0)
    Mm = torch.rand(n, n, generator=g, dtype=torch.float64) * 0.02
    Mm = ((Mm + Mm.t()) * 0.5 + 0.5 * torch.eye(n, dtype=torch.float64)).to(DEV)
    B = torch.randn(n, 4, generator=g, dtype=torch.float64).to(DEV)
    E = (torch.rand(4, generator=g, dtype=torch.float64) * 0.1).to(DEV)
    with torch.no_grad():
        x = xt.linalg.solve(op, B, E=E, M=xt.LinearOperator.m(Mm, True), method="bicgstab", rtol=1e-11, atol=1e-14,
                            posdef=True)
        resid = op.mm(x) - (Mm @ x) * E - B
    assert resid.abs().max().item() <= 1e-8
    # normal equations (posdef=False) through the adjoint callback
    with torch.no_grad():
        x2 = xt.linalg.solve(_UserOp(diag, 0.2, 0.3), B, method="cg", rtol=1e-11, atol=1e-14)
        r2 = _UserOp(diag, 0.2, 0.3).mm(x2) - B
    assert r2.abs().max().item() <= 1e-7


def test_low_rank_matrix_on_gpu_matches_cpu():
    n = 1000
    g = torch.Generator().manual_seed(3)
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 2e-5)):
        Gc = LowRankMatrix(-0.5, n, dtype, torch.device("cpu"))
        Gg = LowRankMatrix(-0.5, n, dtype, torch.device(DEV))
        v = torch.randn(n, generator=g, dtype=dtype)
        for i in range(45):
            c, d = torch.randn(n, generator=g, dtype=dtype) * 0.1, torch.randn(n, generator=g, dtype=dtype) * 0.1
            Gc.append(c, d)
            Gg.append(c.to(DEV), d.to(DEV))
            if i in (0, 3, 31, 32, 44):
                for f in ("mv", "rmv"):
                    a, b = getattr(Gc, f)(v), getattr(Gg, f)(v.to(DEV)).cpu()
                    assert (a - b).abs().max().item() <= tol * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("n,dtype", [(128, torch.float64), (512, torch.float64), (2048, torch.float32)])
def test_rootfinder_c4_form(n, dtype):
    A_cpu, y0_cpu = oracle.make_rootfinder_c4(n, dtype=dtype)
    A = A_cpu.to(DEV).requires_grad_()
    bck = {"rtol": 1e-10, "atol": 1e-14} if dtype == torch.float64 else {}
    # fp32 cannot reach the default f_tol = 1e-6 at this size (rounding floor of |f| ~ 4e-6; the reference would spin
    # for 100 (n + 1) iterations): looser tolerances there, and the iteration count is always bounded
    fwd = {"maxiter": 600} if dtype == torch.float64 else {"maxiter": 600, "f_tol": 5e-5, "x_tol": 1e-4}
    y_o, oinfo = oracle.broyden1_root(fcn, y0_cpu, (A_cpu,), return_info=True, maxiter=600,
                                      f_tol=fwd.get("f_tol", 1e-6), x_tol=fwd.get("x_tol", 1e-6))
    with warnings.catch_warnings():
        warnings.simplefilter("error", xt.ConvergenceWarning)
        y = rootfinder(fcn, y0_cpu.to(DEV), params=(A,), bck_options=bck, **fwd)
    assert fcn(y.detach(), A.detach()).norm().item() <= 2 * fwd.get("f_tol", 1e-6)
    ftol = 1e-7 if dtype == torch.float64 else 5e-4
    assert ((y.detach().cpu() - y_o).norm() / y_o.norm()).item() <= ftol
    (gA,) = torch.autograd.grad(y.sum(), A)
    (g_exact,) = oracle.implicit_grad_dense(fcn, y_o.double(), (A_cpu.double(),), torch.ones_like(y_o.double()))
    gerr = ((gA.double().cpu() - g_exact).norm() / g_exact.norm()).item()
    assert gerr <= (1e-7 if dtype == torch.float64 else 2e-3), gerr


def test_rootfinder_matches_committed_reference_outputs():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "rootfinder_golden.pt"), weights_only=False)
    for c in gold["rootfinder"]:
        if c["n"] <= 5:
            continue
        A = c["A"].to(DEV).requires_grad_()
        y = rootfinder(fcn, torch.zeros(c["n"], 1, dtype=c["dtype"], device=DEV), params=(A,), method=c["method"],
                       maxiter=2000)
        assert (y.detach().cpu() - c["y"]).abs().max().item() <= 1e-7
        if c["method"] == "broyden1":
            (g,) = torch.autograd.grad(y.sum(), A)       # default backward: bicgstab on the matrix-free Jacobian
            assert (g.cpu() - c["grad_A"]).abs().max().item() <= 5e-6
            assert (g.cpu() - c["grad_A_exact"]).abs().max().item() <= 5e-6
