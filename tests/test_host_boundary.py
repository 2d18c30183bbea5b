"""CPU tests of the host-side mirror of the reference interface: the LinearOperator
contract (reference tests: xitorch/_tests/test_linop.py), the method plug-in dispatch, default-method rules,
error conventions (test_linop_fcns.py:16-49, 382-425) and the analytic backward of symeig / solve exercised
through the custom_exact* methods exactly as the reference tests do (test_linop_fcns.py:52-127, 427-629)."""
import warnings

import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200 import LinearOperator
from xitorch_b200._utils import get_method

dt = torch.float64


class DiagOp(LinearOperator):
    """matrix-free operator: only _mv (+ _getparamnames)"""

    def __init__(self, d, herm=True):
        super().__init__(shape=(*d.shape[:-1], d.shape[-1], d.shape[-1]), is_hermitian=herm, dtype=d.dtype,
                         device=d.device)
        self.d = d

    def _mv(self, x):
        return self.d * x

    def _getparamnames(self, prefix=""):
        return [prefix + "d"]


class ShiftOp(LinearOperator):
    """non-symmetric matrix-free operator with only _mv: rmv must come from the adjoint trick"""

    def __init__(self, mat):
        super().__init__(shape=mat.shape, is_hermitian=False, dtype=mat.dtype, device=mat.device)
        self.mat = mat

    def _mv(self, x):
        return torch.matmul(self.mat, x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


def test_mv_is_required():
    class Bad(LinearOperator):
        def __init__(self):
            super().__init__(shape=(2, 2))

    with pytest.raises(RuntimeError):
        Bad()


def test_missing_super_init_message():
    class NoInit(LinearOperator):
        def __init__(self):
            pass

        def _mv(self, x):
            return x

    with pytest.raises(RuntimeError, match="__init__"):
        NoInit().mv(torch.ones(2))


def test_mm_from_mv_rmv_adjoint_trick_fullmatrix():
    torch.manual_seed(0)
    mat = torch.randn(2, 4, 4, dtype=dt)
    op = ShiftOp(mat)
    x = torch.randn(4, 3, dtype=dt)
    assert torch.allclose(op.mm(x), mat @ x)
    z = torch.randn(4, dtype=dt)
    assert torch.allclose(op.rmv(z), (mat.transpose(-2, -1) @ z.unsqueeze(-1)).squeeze(-1))
    assert torch.allclose(op.rmm(x), mat.transpose(-2, -1) @ x)
    assert torch.allclose(op.fullmatrix(), mat)
    assert torch.allclose(op.H.fullmatrix(), mat.transpose(-2, -1))
    # no caching of fullmatrix
    op.mat = mat * 2
    assert torch.allclose(op.fullmatrix(), mat * 2)


def test_shape_errors_and_hermitian_checks():
    mat = torch.randn(3, 3, dtype=dt)
    op = LinearOperator.m(mat)
    assert not op.is_hermitian
    for fn, bad in ((op.mv, torch.ones(4, dtype=dt)), (op.mm, torch.ones(4, 2, dtype=dt)),
                    (op.rmv, torch.ones(4, dtype=dt)), (op.rmm, torch.ones(4, 2, dtype=dt))):
        with pytest.raises(RuntimeError):
            fn(bad)
    with pytest.raises(RuntimeError):
        LinearOperator.m(mat, is_hermitian=True)
    sym = mat + mat.t()
    assert LinearOperator.m(sym).is_hermitian
    with pytest.raises(RuntimeError):
        DiagOp(torch.ones(3, dtype=dt)).__class__.__init__(DiagOp.__new__(DiagOp), torch.ones(3, dtype=dt)) or \
            LinearOperator.__init__(DiagOp.__new__(DiagOp), shape=(2, 3), is_hermitian=True)
    assert "MatrixLinearOperator" in repr(op)
    op.check()
    LinearOperator.m(sym, True).check()


def test_algebra():
    torch.manual_seed(1)
    a = torch.randn(3, 3, dtype=dt)
    d = torch.rand(3, dtype=dt)
    A, D = LinearOperator.m(a), DiagOp(d)
    x = torch.randn(3, 2, dtype=dt)
    assert torch.allclose((A + D).mm(x), a @ x + d.unsqueeze(-1) * x)
    assert torch.allclose((A - D).mm(x), a @ x - d.unsqueeze(-1) * x)
    assert torch.allclose((A * 2.5).mm(x), 2.5 * (a @ x))
    assert torch.allclose((2 * D).mm(x), 2 * d.unsqueeze(-1) * x)
    assert torch.allclose(A.matmul(D).mm(x), a @ (d.unsqueeze(-1) * x))
    assert isinstance(A + A, xt.MatrixLinearOperator)
    assert torch.allclose((A + D).rmm(x), a.t() @ x + d.unsqueeze(-1) * x)
    with pytest.raises(TypeError):
        A * "x"
    with pytest.raises(RuntimeError):
        A + LinearOperator.m(torch.randn(4, 4, dtype=dt))


def test_uselinopparams_swaps_and_restores():
    d = torch.rand(3, dtype=dt)
    op = DiagOp(d)
    assert op.getlinopparams()[0] is d
    d2 = torch.ones(3, dtype=dt)
    with op.uselinopparams(d2):
        assert op.d is d2
    assert op.d is d
    comp = LinearOperator.m(torch.eye(3, dtype=dt)) + op
    names = comp._getparamnames()
    assert names == ["a.mat", "b.d"]
    p = comp.getlinopparams()
    with comp.uselinopparams(p[0] * 3, p[1]):
        assert torch.allclose(comp.mv(torch.ones(3, dtype=dt)), 3 + d)
    assert torch.allclose(comp.mv(torch.ones(3, dtype=dt)), 1 + d)


def test_get_method_rules():
    methods = {"a": lambda: 1}
    assert get_method("x", methods, "A")() == 1
    f = lambda: 2   # noqa: E731
    assert get_method("x", methods, f) is f
    with pytest.raises(RuntimeError, match="Unknown x method"):
        get_method("x", methods, "zzz")
    with pytest.raises(TypeError):
        get_method("x", methods, 3)


def test_symeig_errors_and_defaults():
    mat = torch.randn(4, 4, dtype=dt)
    with pytest.raises(RuntimeError):
        xt.linalg.lsymeig(LinearOperator.m(mat, is_hermitian=False))
    sym = LinearOperator.m(mat + mat.t(), True)
    with pytest.raises(RuntimeError):
        xt.linalg.lsymeig(sym, M=LinearOperator.m(torch.eye(5, dtype=dt), True))
    ev, vec = xt.linalg.symeig(sym)                      # method None -> exacteig, neig None -> all
    assert tuple(ev.shape) == (4,) and tuple(vec.shape) == (4, 4)
    ev2, _ = xt.linalg.symeig(sym, 2, "uppermost")
    assert torch.allclose(ev2, ev[-2:])
    with pytest.raises(RuntimeError, match="Unknown symeig method"):
        xt.linalg.symeig(sym, 2, method="nope")


def test_solve_errors_and_defaults():
    a = torch.randn(4, 4, dtype=dt) + 4 * torch.eye(4, dtype=dt)
    with pytest.raises(RuntimeError):
        xt.linalg.solve(LinearOperator.m(torch.randn(3, 4, dtype=dt)), torch.ones(3, 1, dtype=dt))
    with pytest.raises(RuntimeError):
        xt.linalg.solve(LinearOperator.m(a), torch.ones(5, 1, dtype=dt))
    x = xt.linalg.solve(LinearOperator.m(a), torch.ones(4, 2, dtype=dt))      # dense -> exactsolve
    assert torch.allclose(a @ x, torch.ones(4, 2, dtype=dt))
    E = torch.tensor([0.1, 0.2], dtype=dt)
    x = xt.linalg.solve(LinearOperator.m(a), torch.ones(4, 2, dtype=dt), E)
    assert torch.allclose(a @ x - x * E, torch.ones(4, 2, dtype=dt))
    with pytest.warns(UserWarning, match="ignored"):
        xt.linalg.solve(LinearOperator.m(a), torch.ones(4, 2, dtype=dt), M=LinearOperator.m(torch.eye(4, dtype=dt), True))
    # the plug-in point: a user callable receives (A, B, E, M, **opts) with grad disabled
    seen = {}

    def mymethod(A, B, E=None, M=None, tol=1e-3, **unused):
        seen["grad"] = torch.is_grad_enabled()
        seen["tol"] = tol
        return torch.linalg.solve(A.fullmatrix(), B)

    x = xt.linalg.solve(LinearOperator.m(a), torch.ones(4, 1, dtype=dt), method=mymethod, tol=5.0)
    assert seen == {"grad": False, "tol": 5.0} and torch.allclose(a @ x, torch.ones(4, 1, dtype=dt))


def test_zero_rhs_shortcut_in_boundary():
    a = torch.randn(4, 4, dtype=dt) + 4 * torch.eye(4, dtype=dt)

    def boom(*args, **kw):
        raise AssertionError("must not be called for an all-zero B")

    x = xt.linalg.solve(LinearOperator.m(a), torch.zeros(2, 4, 3, dtype=dt), method=boom)
    assert tuple(x.shape) == (2, 4, 3) and torch.count_nonzero(x) == 0


@pytest.mark.parametrize("with_M", [False, True])
def test_symeig_gradcheck_through_boundary(with_M):
    torch.manual_seed(3)
    n = 5
    A0 = torch.randn(n, n, dtype=dt)
    M0 = torch.randn(n, n, dtype=dt)
    M0 = M0 @ M0.t() + n * torch.eye(n, dtype=dt)

    def fcn(a, m):
        a = (a + a.t()) / 2
        m = (m + m.t()) / 2
        ev, vec = xt.linalg.lsymeig(LinearOperator.m(a, True), 2, M=LinearOperator.m(m, True) if with_M else None,
                                    method="custom_exacteig", bck_options={"method": "custom_exactsolve"})
        return ev, vec.abs()

    a = A0.clone().requires_grad_()
    m = M0.clone().requires_grad_()
    assert torch.autograd.gradcheck(fcn, (a, m))
    assert torch.autograd.gradgradcheck(fcn, (a, m))


@pytest.mark.parametrize("with_E,with_M", [(False, False), (True, False), (True, True)])
def test_solve_gradcheck_through_boundary(with_E, with_M):
    torch.manual_seed(4)
    n, nc = 4, 2
    A0 = torch.randn(n, n, dtype=dt) + 4 * torch.eye(n, dtype=dt)
    B0 = torch.randn(n, nc, dtype=dt)
    E0 = torch.rand(nc, dtype=dt) * 0.1
    M0 = torch.randn(n, n, dtype=dt)
    M0 = M0 @ M0.t() / n + torch.eye(n, dtype=dt)

    def fcn(a, b, e, m):
        m = (m + m.t()) / 2
        return xt.linalg.solve(LinearOperator.m(a), b, e if with_E else None,
                               LinearOperator.m(m, True) if with_M else None,
                               method="custom_exactsolve", bck_options={"method": "custom_exactsolve"})

    args = tuple(t.clone().requires_grad_() for t in (A0, B0, E0, M0))
    assert torch.autograd.gradcheck(fcn, args)
    assert torch.autograd.gradgradcheck(fcn, args)


def test_matrixfree_operator_through_boundary_with_user_method():
    d = torch.linspace(1, 2, 6, dtype=dt).requires_grad_()

    def mymethod(A, B, E=None, M=None, **unused):
        return B / A.d.unsqueeze(-1)

    x = xt.linalg.solve(DiagOp(d), torch.ones(6, 1, dtype=dt), method=mymethod,
                        bck_options={"method": mymethod})
    (g,) = torch.autograd.grad(x.sum(), (d,))
    assert torch.allclose(g, -1 / d.detach() ** 2)


def test_svd_exact():
    torch.manual_seed(5)
    a = torch.randn(6, 4, dtype=dt)
    u, s, vh = xt.linalg.svd(LinearOperator.m(a), k=2)
    sref = torch.linalg.svdvals(a)
    assert torch.allclose(s, sref[:2].flip(0)) or torch.allclose(s, sref[:2])
    assert torch.allclose(a @ vh.t(), u * s, atol=1e-10)


def test_debug_mode_runs_checks():
    sym = torch.eye(3, dtype=dt)
    with xt.enable_debug():
        assert xt.is_debug_enabled()
        xt.linalg.symeig(LinearOperator.m(sym, True), 1)
    assert not xt.is_debug_enabled()


def test_failing_hermiticity_kernel_is_reported_not_hidden(monkeypatch):
    """`LinearOperator.m` checks the Hermitian flag with the one-pass CUDA kernel; if that path FAILS the library test
    still decides, but the failure is reported (round-1 review: the fallback used to swallow every exception)"""
    import warnings
    from xitorch_b200 import linop, _dense

    class LooksLikeCuda(torch.Tensor):
        is_cuda = property(lambda self: True)

    a = torch.randn(5, 5, dtype=torch.float64)
    sym = (a + a.T).as_subclass(LooksLikeCuda)

    def broken(mat, *args, **kwargs):
        raise RuntimeError("kernel launch failed")

    monkeypatch.setattr(_dense, "hermitian_check", broken)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert linop._equals_adjoint(sym) is True                      # the library test still settles it
    assert any(issubclass(i.category, RuntimeWarning) and "Hermiticity check failed" in str(i.message) for i in w)
    monkeypatch.setattr(_dense, "hermitian_check", lambda mat, *a_, **k_: False)   # a "not Hermitian" verdict is silent
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert linop._equals_adjoint(sym) is True
    assert not [i for i in w if issubclass(i.category, RuntimeWarning)]
