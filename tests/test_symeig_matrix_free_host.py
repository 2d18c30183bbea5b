"""CPU test of the HOST side of the matrix-free eigensolver path (`xt_symeig_args.apply`): the CUDA engine is replaced by
a stand-in that drives the very same callback protocol (block pointers inside the workspace, one `apply` per
iteration) with a textbook block-Lanczos in numpy.  What is checked is everything Python owns -- pointer arithmetic of
the callback, dtype / shape handling, the whitening of a generalized problem, error propagation out of the callback,
the decision between materialising and calling back.  The engine itself is covered by the `-m gpu` tests.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200 import _lib
from xitorch_b200._impls import symeig as impl
from xitorch_b200.linalg import symeig, svd


class _StandInEngine(object):
    """same entry points as the shared library for the calls `_call_engine` makes"""

    def __init__(self):
        self.calls = 0
        self.used_apply = None

    def xt_symeig_workspace_bytes(self, dtype, n, neig, max_basis, world):
        return 4 * n * neig * 8 + 256

    def xt_last_error(self):
        return b"stand-in"

    def xt_symeig_krylov(self, g):
        n, k = g.n, g.neig
        npdt = np.float32 if g.dtype == _lib.XT_F32 else np.float64
        esz = np.dtype(npdt).itemsize
        self.used_apply = bool(g.apply)
        assert g.apply, "the stand-in only implements the matrix-free protocol"
        assert not g.A and g.nbatch == 1
        apply = C.cast(g.apply, _lib.APPLY_FN)
        xoff, yoff = 64, 64 + ((n * k * esz + 63) // 64) * 64           # two blocks somewhere inside the workspace

        def view(ptr, count):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if npdt == np.float32 else C.c_double)),
                                         shape=(count,))

        xblk = view(g.workspace + xoff, n * k).reshape(n, k)
        yblk = view(g.workspace + yoff, n * k).reshape(n, k)

        def A(x):
            xblk[...] = x.astype(npdt)
            apply(None, g.workspace + xoff, g.workspace + yoff, None)
            self.calls += 1
            return yblk.astype(np.float64).copy()

        q, _ = np.linalg.qr(view(g.V0, n * k).reshape(n, k).astype(np.float64))
        basis, images = [q], []
        nblocks = min(g.max_niter, n // k)                # up to the whole space: the Ritz pairs are then exact
        for it in range(nblocks):
            w = A(basis[it])
            images.append(w)
            if it == nblocks - 1:
                break
            V = np.concatenate(basis, axis=1)
            w = w - V @ (V.T @ w)
            w = w - V @ (V.T @ w)
            qn, _ = np.linalg.qr(w)
            basis.append(qn)
        V = np.concatenate(basis[:len(images)], axis=1)
        AV = np.concatenate(images, axis=1)
        T = V.T @ AV
        w, S = np.linalg.eigh(0.5 * (T + T.T))
        sel = slice(0, k) if g.mode == 0 else slice(len(w) - k, len(w))
        view(g.evals, k)[...] = w[sel].astype(npdt)
        view(g.evecs, n * k).reshape(n, k)[...] = (V @ S[:, sel]).astype(npdt)
        if g.niter_out:
            g.niter_out[0] = len(images)
        if g.converged_out:
            g.converged_out[0] = 1
        if g.napply_out:
            g.napply_out[0] = len(images)
        return 0


@pytest.fixture()
def engine(monkeypatch):
    eng = _StandInEngine()
    monkeypatch.setattr(_lib, "lib", lambda: eng)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda dev: 0)

    class _NoDevice(object):
        def __init__(self, dev):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    monkeypatch.setattr(torch.cuda, "device", _NoDevice)
    monkeypatch.setattr(impl, "_start_block",
                        lambda kind, nb, n, neig, dtype, dev: torch.randn(
                            nb, n, neig, dtype=dtype, generator=torch.Generator().manual_seed(12421)))
    return eng


def _sym(n, dtype, seed=3):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    a = (a + a.T) / (2 * n) ** 0.5 + torch.diag(torch.linspace(1, 20, n, dtype=torch.float64))
    return a.to(dtype)


class UserOperator(xt.LinearOperator):
    """known only through _mv, as a user-defined operator of the reference"""

    def __init__(self, mat):
        super().__init__(shape=mat.shape, is_hermitian=True, dtype=mat.dtype, device=mat.device)
        self.mat = mat
        self.nmv = 0

    def _mv(self, x):
        self.nmv += 1
        return torch.matmul(self.mat, x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("mode", ["lowest", "uppest"])
def test_user_operator_goes_through_the_callback(engine, dtype, mode):
    n, k = 48, 4
    mat = _sym(n, dtype)
    op = UserOperator(mat)
    info = {}
    evals, evecs = symeig(op, neig=k, mode=mode, method="davidson", matrix_free=True, info=info)
    assert engine.used_apply and engine.calls == info["napply"] > 0
    assert evals.shape == (k,) and evecs.shape == (n, k) and evals.dtype == dtype
    ref = torch.linalg.eigvalsh(mat.double())
    ref = ref[:k] if mode == "lowest" else ref[-k:]
    tol = 1e-9 if dtype == torch.float64 else 2e-5
    assert (evals.double() - ref).abs().max().item() <= tol * ref.abs().max().item()
    resid = mat.double() @ evecs.double() - evecs.double() * evals.double()
    # (fp32: the stand-in's plain Lanczos on fp32-rounded images, not the engine's accuracy)
    assert resid.abs().max().item() <= (1e-7 if dtype == torch.float64 else 2e-2)


def test_dense_operator_never_uses_the_callback(engine):
    mat = _sym(32, torch.float64)
    with pytest.raises(AssertionError, match="stand-in only implements"):
        symeig(xt.LinearOperator.m(mat, is_hermitian=True), neig=2, method="davidson", matrix_free=True)


def test_default_materialises_small_composites(engine, monkeypatch):
    mat = _sym(32, torch.float64)
    op = UserOperator(mat)
    seen = {}

    def fake_dense(A, *a, **kw):
        seen["materialised"] = True
        raise KeyboardInterrupt            # stop right there: only the decision is under test

    monkeypatch.setattr(impl, "_dense_of", fake_dense)
    with pytest.raises(KeyboardInterrupt):
        symeig(op, neig=2, method="davidson")
    assert seen.get("materialised")
    monkeypatch.setattr(impl, "_MATERIALISE_MAX_N", 16)     # "too large to materialise": automatic matrix-free
    evals, _ = symeig(op, neig=2, method="davidson")
    assert engine.used_apply
    assert torch.allclose(evals, torch.linalg.eigvalsh(mat)[:2], atol=1e-9)


def test_generalized_problem_with_a_matrix_free_operator(engine):
    n, k = 42, 3
    A = _sym(n, torch.float64, seed=5)
    Mm = _sym(n, torch.float64, seed=6)
    Mm = Mm @ Mm.T / 50 + torch.eye(n, dtype=torch.float64)
    evals, evecs = symeig(UserOperator(A), neig=k, M=xt.LinearOperator.m(Mm, is_hermitian=True), method="lanczos",
                          matrix_free=True)
    resid = A @ evecs - Mm @ evecs * evals
    assert resid.abs().max().item() <= 1e-8
    assert torch.allclose(evecs.T @ Mm @ evecs, torch.eye(k, dtype=torch.float64), atol=1e-9)   # M-orthonormal


def test_svd_of_rectangular_operator(engine):
    g = torch.Generator().manual_seed(9)
    B = torch.randn(60, 36, generator=g, dtype=torch.float64)

    class Rect(xt.LinearOperator):
        def __init__(self):
            super().__init__(shape=B.shape, dtype=B.dtype, device=B.device)

        def _mv(self, x):
            return torch.matmul(B, x.unsqueeze(-1)).squeeze(-1)

        def _rmv(self, y):
            return torch.matmul(B.T, y.unsqueeze(-1)).squeeze(-1)

        def _getparamnames(self, prefix=""):
            return []

    u, s, vh = svd(Rect(), k=3, mode="uppest", method="davidson", matrix_free=True)
    assert engine.used_apply                                    # A^H A is a composite operator
    sref = torch.linalg.svdvals(B)[:3]
    assert torch.allclose(s.sort(descending=True).values, sref, rtol=1e-8)
    assert (B @ vh.transpose(-2, -1) - u * s.unsqueeze(-2)).abs().max().item() <= 1e-7


def test_exception_in_user_operator_surfaces(engine):
    class Broken(UserOperator):
        def _mv(self, x):
            raise ValueError("user operator failed")

    with pytest.raises(ValueError, match="user operator failed"):
        symeig(Broken(_sym(24, torch.float64)), neig=2, method="davidson", matrix_free=True)


def test_batched_operator_cannot_be_matrix_free(engine):
    mats = torch.stack([_sym(16, torch.float64, seed=s) for s in (1, 2)])

    class Batched(xt.LinearOperator):
        def __init__(self):
            super().__init__(shape=mats.shape, is_hermitian=True, dtype=mats.dtype, device=mats.device)

        def _mv(self, x):
            return torch.matmul(mats, x.unsqueeze(-1)).squeeze(-1)

        def _getparamnames(self, prefix=""):
            return []

    with pytest.raises(RuntimeError, match="without batch dimensions"):
        symeig(Batched(), neig=2, method="davidson", matrix_free=True)
