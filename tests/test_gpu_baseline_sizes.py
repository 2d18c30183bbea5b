"""Parity at the FULL sizes of BASELINE.json's configurations.  The checkers are fp64 library
calls on the device (`eigvalsh`, `solve`, autograd) -- test infrastructure, never on the product path."""
import warnings

import pytest
import torch

import oracle
import xitorch_b200 as xt

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_c2_eigenvalues_vs_fp64_eigvalsh():
    """configs[1]: davidson neig=8 on N=16384 fp32 -- eigenvalues against the TRUE spectrum (device-side fp64 `eigvalsh` of
    the same fp32-rounded matrix), 1e-5 relative (north_star), not only the Rayleigh self-check"""
    n, neig = 16384, 8
    A = oracle.make_herm(n, neig, torch.float32, seed=7).to(DEV)
    info = {}
    ev, vec = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="davidson", min_eps=1e-4, info=info)
    ref = torch.linalg.eigvalsh(A.double())[:neig]
    assert info["converged"] and info["niter"] <= 14
    assert ((ev.double() - ref).abs() / ref.abs()).max().item() <= 1e-5
    # tighter stop test: the error follows
    ev2, _ = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="davidson", min_eps=2e-5)
    assert ((ev2.double() - ref).abs() / ref.abs()).max().item() <= 3e-6


def test_c3_shard_bf16_vs_fp64_solve():
    """configs[2], one GPU's shard: bicgstab on 64 independent 4096 x 4096 bf16 systems (fp32 vectors) against the fp64
    direct solution of the SAME bf16-rounded matrices"""
    nb, n = 64, 4096
    g = torch.Generator(device=DEV)
    g.manual_seed(99)
    A = torch.empty(nb, n, n, dtype=torch.bfloat16, device=DEV)
    for b in range(nb):
        Ab = torch.randn(n, n, generator=g, device=DEV) * (0.3 / n ** 0.5)
        Ab.diagonal().add_(1.0)
        A[b] = Ab.to(torch.bfloat16)
    B = torch.randn(nb, n, 1, generator=g, device=DEV)
    info = {}
    with warnings.catch_warnings():
        warnings.simplefilter("error", xt.ConvergenceWarning)
        X = xt.linalg.solve(xt.LinearOperator.m(A, is_hermitian=False), B, method="bicgstab", rtol=1e-7, atol=1e-12,
                            posdef=True, info=info)
    worst = 0.0
    for b0 in range(0, nb, 8):                                   # fp64 checker in slices of 8 systems (1 GiB each)
        Xr = torch.linalg.solve(A[b0:b0 + 8].double(), B[b0:b0 + 8].double())
        worst = max(worst, ((X[b0:b0 + 8].double() - Xr).norm(dim=1) / Xr.norm(dim=1)).max().item())
    assert info["converged"]
    assert worst <= 1e-5, worst


def _fcn(y, A):
    return torch.tanh(A @ y + 0.1) + y / 2


def test_c4_rootfinder_8192_forward_and_adjoint():
    """configs[3]: Broyden rootfinder on tanh(A y + 0.1) + y/2, y in R^8192 (fp64), and the backward adjoint
    solve, against the exact implicit gradient  -(df/dA)^T (df/dy)^-T g  formed with a dense fp64 Jacobian
    (oracle.implicit_grad_dense's formula, evaluated on the device)"""
    from xitorch_b200.optimize import rootfinder
    n = 8192
    A_cpu, y0_cpu = oracle.make_rootfinder_c4(n, dtype=torch.float64)
    A = A_cpu.to(DEV).requires_grad_()
    with warnings.catch_warnings():
        warnings.simplefilter("error", xt.ConvergenceWarning)
        y = rootfinder(_fcn, y0_cpu.to(DEV), params=(A,), bck_options={"rtol": 1e-10, "atol": 1e-14}, maxiter=800)
    assert _fcn(y.detach(), A.detach()).norm().item() <= 2e-6
    (gA,) = torch.autograd.grad(y.sum(), A)
    # exact: J = diag(1 - tanh^2(A y + 0.1)) A + I/2;  J^T v = -1;  dL/dA = (diag(1 - tanh^2) v) y^T
    with torch.no_grad():
        yd, Ad = y.detach(), A.detach()
        s = 1.0 - torch.tanh(Ad @ yd + 0.1) ** 2                          # (n, 1)
        J = s * Ad + 0.5 * torch.eye(n, dtype=torch.float64, device=DEV)
        v = torch.linalg.solve(J.t(), -torch.ones_like(yd))
        g_exact = (s * v) @ yd.t()
    gerr = ((gA - g_exact).norm() / g_exact.norm()).item()
    assert gerr <= 1e-6, gerr


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_second_device_after_first_in_one_process():
    """kernel attributes (opt-in shared memory sizes) are per device: a process that used cuda:0 must be able to solve on
    cuda:1 (VERDICT round 1, weak point 7)"""
    n, neig = 4096, 8
    A = oracle.make_herm(n, neig, torch.float32, seed=3)
    ref = torch.linalg.eigvalsh(A.double())[:neig]
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        with torch.cuda.device(dev):
            Ad = A.to(dev)
            ev, _ = xt.linalg.symeig(xt.LinearOperator.m(Ad, True), neig=neig, method="davidson", min_eps=1e-4)
            assert ((ev.double().cpu() - ref).abs() / ref.abs()).max().item() <= 1e-5
            B = torch.randn(n, 2, device=dev)
            M = Ad + 30.0 * torch.eye(n, device=dev)
            X = xt.linalg.solve(xt.LinearOperator.m(M, True), B, method="cg", rtol=1e-7)
            assert ((M @ X - B).norm() / B.norm()).item() <= 1e-5
            X2 = xt.linalg.solve(xt.LinearOperator.m(M, False), B, method="gmres", rtol=1e-7)
            assert ((M @ X2 - B).norm() / B.norm()).item() <= 1e-5
