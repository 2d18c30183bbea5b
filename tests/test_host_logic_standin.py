"""CPU tests of the complete HOST logic of `linalg.solve` / `linalg.symeig` with the Krylov methods, the CUDA library
replaced by tests/standin_engine.py (numpy on the argument structs).  Every case checks the answer against dense linear
algebra, i.e. that what the host marshals -- pointers, strides, batch flattening / broadcasting, shifts, M, the
normal-equation rewrite, the real-equivalent form of complex systems, operator and preconditioner callbacks -- describes
the right problem, and that the autograd boundary around it differentiates correctly."""
import warnings

import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200.linalg import solve, symeig, svd

import standin_engine

DT = torch.float64


@pytest.fixture()
def eng(monkeypatch):
    return standin_engine.install(monkeypatch)


def _rand(*shape, seed=0, dtype=DT):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=dtype)


def _spd(n, *batch, seed=0, dtype=DT):
    a = _rand(*batch, n, n, seed=seed, dtype=dtype)
    return a @ a.transpose(-2, -1).conj() / n + torch.eye(n, dtype=dtype)


class MatrixFree(xt.LinearOperator):
    def __init__(self, mat, is_hermitian=False):
        super().__init__(shape=mat.shape, is_hermitian=is_hermitian, dtype=mat.dtype, device=mat.device)
        self.mat = mat

    def _mv(self, x):
        return torch.matmul(self.mat, x.unsqueeze(-1)).squeeze(-1)

    def _rmv(self, x):
        return torch.matmul(self.mat.transpose(-2, -1).conj(), x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


# ---------------------------------------------------------------------------------------------- solve: dense operands
@pytest.mark.parametrize("method", ["cg", "bicgstab", "gmres"])
def test_dense_plain(eng, method):
    n = 12
    A = _spd(n, seed=1)
    B = _rand(n, 3, seed=2)
    info = {}
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, method=method, info=info)
    assert torch.allclose(A @ X, B, atol=1e-10)
    assert eng.log[-1]["method"] == method and not eng.log[-1]["matrix_free"]
    assert info["converged"]


def test_batches_broadcast(eng):
    n = 7
    A = _spd(n, 2, 1, seed=3)                       # (2, 1, n, n)
    B = _rand(3, n, 2, seed=4)                      # (3, n, 2)
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, method="cg")
    assert X.shape == (2, 3, n, 2)
    assert torch.allclose(A @ X, B.expand(2, 3, n, 2), atol=1e-10)
    assert eng.log[-1]["nbatch"] == 6


def test_strided_operator_and_rhs(eng):
    n = 9
    big = _spd(2 * n, seed=5)
    A = big[:n, :n]                                 # leading dimension 2n
    Bt = _rand(4, n, seed=6).T                      # column-major right-hand side
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), Bt, method="cg")
    assert torch.allclose(A @ X, Bt, atol=1e-10)


@pytest.mark.parametrize("with_M", [False, True])
@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_shifts_and_metric(eng, method, with_M):
    n, nc = 10, 3
    A = _spd(n, seed=7) + 5 * torch.eye(n, dtype=DT)
    E = torch.tensor([0.1, -0.4, 0.25], dtype=DT)
    Mm = _spd(n, seed=8)
    B = _rand(n, nc, seed=9)
    M = xt.LinearOperator.m(Mm, is_hermitian=True) if with_M else None
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, E=E, M=M, method=method)
    MX = Mm @ X if with_M else X
    assert torch.allclose(A @ X - MX * E, B, atol=1e-9)
    assert eng.log[-1]["has_E"] and eng.log[-1]["has_M"] == with_M


def test_batched_shifts(eng):
    n, nc = 6, 2
    A = _spd(n, 3, seed=10) + 4 * torch.eye(n, dtype=DT)
    E = _rand(3, nc, seed=11) * 0.3
    B = _rand(3, n, nc, seed=12)
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, E=E, method="bicgstab")
    assert torch.allclose(A @ X - X * E.unsqueeze(-2), B, atol=1e-9)


@pytest.mark.parametrize("with_E", [False, True])
def test_cg_on_indefinite_or_nonsymmetric_uses_normal_equations(eng, with_E):
    n, nc = 8, 2
    A = _rand(n, n, seed=13) + 3 * torch.eye(n, dtype=DT)          # not symmetric
    B = _rand(n, nc, seed=14)
    E = torch.tensor([0.2, -0.1], dtype=DT) if with_E else None
    X = solve(xt.LinearOperator.m(A), B, E=E, method="cg")
    lhs = A @ X - (X * E if with_E else 0)
    assert torch.allclose(lhs, B, atol=1e-8)
    sysm = torch.from_numpy(eng.log[-1]["systems"])
    assert torch.allclose(sysm, sysm.transpose(-2, -1), atol=1e-12)     # what reached the engine is A^H A - like
    # explicitly requested for a Hermitian operator
    As = _spd(n, seed=15) - 1.2 * torch.eye(n, dtype=DT)             # indefinite
    X2 = solve(xt.LinearOperator.m(As, is_hermitian=True), B, method="cg", posdef=False)
    assert torch.allclose(As @ X2, B, atol=1e-8)


def test_float32_and_zero_rhs(eng):
    n = 8
    A = _spd(n, seed=16, dtype=torch.float32)
    B = _rand(n, 2, seed=17, dtype=torch.float32)
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, method="cg")
    assert X.dtype == torch.float32 and torch.allclose(A @ X, B, atol=1e-4)
    ncalls = len(eng.log)
    Z = solve(xt.LinearOperator.m(A, is_hermitian=True), torch.zeros(n, 2), method="cg")
    assert torch.equal(Z, torch.zeros(n, 2)) and len(eng.log) == ncalls        # answered on the host


# ---------------------------------------------------------------------------------------------- solve: callbacks
def test_gmres_refuses_shifts_like_the_reference(eng):
    A = xt.LinearOperator.m(_spd(6, seed=41), is_hermitian=True)
    with pytest.raises(RuntimeError, match="gmres does not support E"):
        solve(A, _rand(6, 2, seed=42), E=torch.tensor([0.1, 0.2], dtype=DT), method="gmres")


@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_matrix_free_operator_with_shift_and_metric(eng, method):
    n, nc = 9, 2
    A = _spd(n, seed=18) + 3 * torch.eye(n, dtype=DT)
    Mm = _spd(n, seed=19)
    E = torch.tensor([0.3, -0.2], dtype=DT)
    B = _rand(n, nc, seed=20)
    X = solve(MatrixFree(A, True), B, E=E, M=MatrixFree(Mm, True), method=method)
    assert eng.log[-1]["matrix_free"]
    assert torch.allclose(A @ X - Mm @ X * E, B, atol=1e-9)


def test_matrix_free_batched(eng):
    n = 6
    A = _spd(n, 2, seed=21)
    B = _rand(2, n, 3, seed=22)
    X = solve(MatrixFree(A, True), B, method="cg")
    assert torch.allclose(A @ X, B, atol=1e-9)


def test_preconditioner_callbacks(eng):
    n, nc = 8, 2
    A = _spd(n, seed=23)
    P = _spd(n, seed=24)
    Q = _spd(n, seed=25)
    B = _rand(n, nc, seed=26)
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, method="cg", precond=xt.LinearOperator.m(P, is_hermitian=True))
    rec = eng.log[-1]
    assert rec["precond_l"] and not rec["precond_r"] and rec["check_every"] == 1
    probe = torch.from_numpy(rec["precond_probe"])[0]
    assert torch.allclose(torch.from_numpy(rec["precond_l_out"])[0], P @ probe, atol=1e-12)
    assert torch.allclose(A @ X, B, atol=1e-9)
    solve(xt.LinearOperator.m(A), B, method="bicgstab", precond_l=xt.LinearOperator.m(P), precond_r=MatrixFree(Q))
    rec = eng.log[-1]
    probe = torch.from_numpy(rec["precond_probe"])[0]
    assert torch.allclose(torch.from_numpy(rec["precond_l_out"])[0], P @ probe, atol=1e-12)
    assert torch.allclose(torch.from_numpy(rec["precond_r_out"])[0], Q @ probe, atol=1e-12)
    with pytest.raises(TypeError):
        solve(xt.LinearOperator.m(A), B, method="bicgstab", precond_l=P)


def test_exception_in_operator_surfaces(eng):
    class Broken(MatrixFree):
        def _mv(self, x):
            raise ValueError("operator failed")

    with pytest.raises(ValueError, match="operator failed"):
        solve(Broken(_spd(5, seed=27), True), _rand(5, 1, seed=28), method="cg")


# ---------------------------------------------------------------------------------------------- solve: complex
@pytest.mark.parametrize("method", ["cg", "bicgstab", "gmres"])
@pytest.mark.parametrize("with_E", [False, True])
def test_complex_systems(eng, method, with_E):
    n, nc = 7, 2
    A = _spd(n, seed=29, dtype=torch.complex128) + 2 * torch.eye(n, dtype=torch.complex128)     # Hermitian pos. def.
    B = _rand(n, nc, seed=30, dtype=torch.complex128)
    E = torch.tensor([0.3 + 0.1j, -0.2j], dtype=torch.complex128) if with_E else None
    if with_E and method == "gmres":
        pytest.skip("gmres takes no E")
    X = solve(xt.LinearOperator.m(A, is_hermitian=True), B, E=E, method=method)
    assert X.dtype == torch.complex128
    lhs = A @ X - (X * E if with_E else 0)
    assert torch.allclose(lhs, B, atol=1e-9)
    assert eng.log[-1]["n"] == 2 * n                                   # the real-equivalent doubled system


# ---------------------------------------------------------------------------------------------- solve: autograd
@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_solve_gradients_through_the_boundary(eng, method):
    n, nc = 5, 2
    A0 = _spd(n, seed=31).requires_grad_()
    B0 = _rand(n, nc, seed=32).requires_grad_()
    E0 = (_rand(nc, seed=33) * 0.1).requires_grad_()

    def fcn(A, B, E):
        As = (A + A.T) * 0.5
        return solve(xt.LinearOperator.m(As, is_hermitian=True), B, E=E, method=method,
                     bck_options={"method": method})

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert torch.autograd.gradcheck(fcn, (A0, B0, E0))
        assert torch.autograd.gradgradcheck(fcn, (A0, B0, E0))
    assert all(r["method"] == method for r in eng.log)                  # backward re-entered the same method


# ---------------------------------------------------------------------------------------------- symeig
@pytest.mark.parametrize("method", ["davidson", "lanczos"])
@pytest.mark.parametrize("mode", ["lowest", "uppest"])
def test_symeig_dense_batched(eng, method, mode):
    n, k = 10, 3
    A = _spd(n, 2, 2, seed=34)
    evals, evecs = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=k, mode=mode, method=method)
    assert evals.shape == (2, 2, k) and evecs.shape == (2, 2, n, k)
    w = torch.linalg.eigvalsh(A)
    ref = w[..., :k] if mode == "lowest" else w[..., -k:]
    assert torch.allclose(evals, ref, atol=1e-10)
    assert torch.allclose(A @ evecs, evecs * evals.unsqueeze(-2), atol=1e-9)
    assert eng.log[-1]["nbatch"] == 4 and eng.log[-1]["mode"] == (0 if mode == "lowest" else 1)


def test_symeig_generalized_and_strided(eng):
    n, k = 9, 2
    big = _spd(2 * n, seed=35)
    A = big[n:, n:]
    Mm = _spd(n, seed=36)
    evals, evecs = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=k, M=xt.LinearOperator.m(Mm, is_hermitian=True),
                          method="davidson")
    assert torch.allclose(A @ evecs, Mm @ evecs * evals, atol=1e-9)
    assert torch.allclose(evecs.T @ Mm @ evecs, torch.eye(k, dtype=DT), atol=1e-10)


def test_symeig_options_and_errors(eng):
    A = xt.LinearOperator.m(_spd(8, seed=37), is_hermitian=True)
    with pytest.raises(RuntimeError, match="nguess"):
        symeig(A, neig=3, method="davidson", nguess=2)
    with pytest.raises(ValueError, match="v_init"):
        symeig(A, neig=2, method="davidson", v_init="nope")
    with pytest.raises(RuntimeError, match="expansion"):
        symeig(A, neig=2, method="davidson", expansion="nope")
    info = {}
    symeig(A, neig=2, method="davidson", v_init="eye", max_basis=6, expansion="residual", info=info)
    assert eng.log[-1]["max_basis"] == 6 and eng.log[-1]["expansion"] == 0 and info["converged"]
    with pytest.raises(RuntimeError, match="complex"):
        symeig(xt.LinearOperator.m(_spd(6, seed=38, dtype=torch.complex128), is_hermitian=True), neig=2,
               method="davidson")


def test_symeig_gradients_with_krylov_forward_and_backward(eng):
    n, k = 6, 2
    A0 = _spd(n, seed=39).requires_grad_()

    def fcn(A):
        As = (A + A.T) * 0.5
        evals, evecs = symeig(xt.LinearOperator.m(As, is_hermitian=True), neig=k, method="davidson",
                              bck_options={"method": "cg"})
        return evals, (evecs ** 2)                                   # sign-independent function of the vectors

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert torch.autograd.gradcheck(fcn, (A0,))
    assert {r["method"] for r in eng.log} >= {"symeig", "cg"}


def test_svd_through_davidson(eng):
    B = _rand(9, 6, seed=40)
    u, s, vh = svd(xt.LinearOperator.m(B), k=2, mode="uppest", method="davidson")
    sref = torch.linalg.svdvals(B)[:2]
    assert torch.allclose(s.sort(descending=True).values, sref, atol=1e-10)
    assert torch.allclose(B @ vh.transpose(-2, -1), u * s.unsqueeze(-2), atol=1e-9)


# ---------------------------------------------------------------------------------------------- object lifetimes
def test_callbacks_do_not_keep_tensors_alive(eng):
    """the operator / all-gather callbacks must not sit in a reference cycle: the workspace and the operator's tensors
    have to be released when the call returns, not when the cyclic collector happens to run (on the GPU that is HBM
    held by dead solves; the reference guards the same property in _tests/test_memleak.py).  `ctypes.cast(cb, c_void_p)`
    creates exactly such a cycle, hence `_lib.fn_address`."""
    import gc
    import weakref
    from xitorch_b200.optimize import rootfinder

    gc.collect()
    gc.disable()
    try:
        A = _spd(8, seed=50)
        wa = weakref.ref(A)
        op = MatrixFree(A, True)
        X = solve(op, _rand(8, 2, seed=51), method="cg", precond=MatrixFree(_spd(8, seed=52), True))
        ev, _ = symeig(op, neig=2, method="davidson", matrix_free=True)
        del op, A, X, ev, _
        assert wa() is None

        a = (torch.ones(12, dtype=DT) + 0.5).requires_grad_()
        wr = weakref.ref(a)

        def f(y, a):
            return y * y - 3.0 * a + torch.sigmoid(y)

        y = rootfinder(f, torch.ones(12, dtype=DT), params=(a,), f_tol=1e-9, alpha=-0.5,
                       bck_options={"method": "bicgstab"})
        (grad,) = torch.autograd.grad((y ** 2).sum(), a, create_graph=True)
        assert eng.log[-1]["method"] == "bicgstab" and eng.log[-1]["matrix_free"]
        del y, grad, a
        assert wr() is None
    finally:
        gc.enable()


def test_symeig_with_fewer_than_two_blocks(eng):
    # n < 2 neig: the reference's first expansion already spans the whole space -> exact pairs; no engine call
    A = _spd(5, seed=60)
    ncalls = len(eng.log)
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=3, method="davidson", info=info)
    assert len(eng.log) == ncalls and info["converged"] and info["completed_full_space"]
    assert torch.allclose(ev, torch.linalg.eigvalsh(A)[:3], atol=1e-12)
    assert torch.allclose(A @ vec, vec * ev, atol=1e-12)
    Mm = _spd(5, seed=61)
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, mode="uppest",
                     M=xt.LinearOperator.m(Mm, is_hermitian=True), method="lanczos")
    assert torch.allclose(A @ vec, Mm @ vec * ev, atol=1e-10)
