"""CPU tests of the host-composed Davidson (`_impls/symeig.py:_davidson_host`): generalized problems, start blocks wider
than neig, preconditioned expansion.  The loop only talks to the operators through `A.mm` / `M.mm` (the block-matvec
kernel on the GPU, torch.matmul for the CPU tensors of these tests) and never forms anything of order n^3, so what is
under test here is the algorithm: against the reference's own generalized run (tests/golden), against dense
generalized `eigh`, and through the public `linalg.symeig` including gradients."""
import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200.linalg import symeig
from xitorch_b200._impls import symeig as simpl

import standin_engine

DT = torch.float64


@pytest.fixture()
def eng(monkeypatch):
    return standin_engine.install(monkeypatch)


def _rand(*shape, seed=0, dtype=DT):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=dtype)


def _spd(n, *batch, seed=0, dtype=DT):
    a = _rand(*batch, n, n, seed=seed, dtype=dtype)
    return a @ a.transpose(-2, -1) / n + torch.eye(n, dtype=dtype)


def _herm(n, *batch, seed=0, dtype=DT):
    a = _rand(*batch, n, n, seed=seed, dtype=dtype)
    return 0.5 * (a + a.transpose(-2, -1)) + torch.diag(torch.arange(n, dtype=dtype) * 0.5)


def _exact(A, M, k, mode):
    Li = torch.inverse(torch.linalg.cholesky(M.double()))
    w = torch.linalg.eigvalsh(Li @ A.double() @ Li.transpose(-2, -1))
    return w[..., :k] if mode == "lowest" else w[..., -k:]


class Counting(xt.LinearOperator):
    """matrix-free operator that records the widths of the blocks it is applied to"""

    def __init__(self, mat):
        super().__init__(shape=mat.shape, is_hermitian=True, dtype=mat.dtype, device=mat.device)
        self.mat = mat
        self.widths = []

    def _mv(self, x):
        return self._mm(x.unsqueeze(-1)).squeeze(-1)

    def _mm(self, x):
        self.widths.append(x.shape[-1])
        return torch.matmul(self.mat, x)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


def test_reference_generalized_case(eng, golden):
    """the reference's davidson on (A, M): same eigenvalues, same eigenvectors up to sign"""
    c = golden["davidson_M"]
    info = {}
    evals, evecs = symeig(xt.LinearOperator.m(c["A"], True), neig=c["neig"], M=xt.LinearOperator.m(c["M"], True),
                          method="davidson", min_eps=c["min_eps"], info=info)
    assert info["engine"] == "host-composed" and info["converged"]
    assert not eng.log, "generalized problems must not reach the standard-problem engine"
    assert ((evals - c["evals"]).abs() / c["evals"].abs()).max().item() <= 1e-10
    assert (evecs.abs() - c["evecs_abs"]).abs().max().item() <= 1e-6
    assert (c["A"] @ evecs - c["M"] @ evecs * evals).abs().max().item() <= c["min_eps"]


def test_reference_generalized_and_wide_start_fixture(eng, golden_generalized):
    """every case of tests/golden/generalized_golden.pt (outputs of the unmodified reference: generalized problems in
    both modes, batched, the full-space exit, nguess > neig with and without M): same eigenvalues, same eigenvectors up
    to sign, M-orthonormal, and no more operator applications than the reference needed iterations"""
    for c in golden_generalized:
        A, Mm = c["A"], c["M"]
        info = {}
        evals, evecs = symeig(xt.LinearOperator.m(A, True), neig=c["neig"], mode=c["mode"],
                              M=None if Mm is None else xt.LinearOperator.m(Mm, True), method="davidson",
                              nguess=c["nguess"], min_eps=c["min_eps"], info=info)
        assert info["engine"] == "host-composed" and info["converged"], c["tag"]
        assert ((evals - c["evals"]).abs() / c["evals"].abs()).max().item() <= 1e-9, c["tag"]
        assert (evecs.abs() - c["evecs_abs"]).abs().max().item() <= 1e-5, c["tag"]
        Md = torch.eye(A.shape[-1], dtype=DT) if Mm is None else Mm
        gram = evecs.transpose(-2, -1) @ Md @ evecs
        assert (gram - torch.eye(c["neig"], dtype=DT)).abs().max().item() <= 1e-9, c["tag"]
        assert (A @ evecs - Md @ evecs * evals.unsqueeze(-2)).abs().max().item() <= c["min_eps"], c["tag"]
        assert info["napply"] <= c["oracle_niter"] + 3, (c["tag"], info, c["oracle_niter"])


@pytest.mark.parametrize("mode", ["lowest", "uppest"])
@pytest.mark.parametrize("with_m", [True, False])
@pytest.mark.parametrize("n,neig,nguess", [(200, 4, None), (200, 4, 7), (300, 8, 12), (9, 2, None), (9, 2, 5)])
def test_against_dense_eigh(eng, n, neig, nguess, with_m, mode):
    if not with_m and nguess is None:
        pytest.skip("standard problem with nguess = neig is the engine's case")
    A, Mm = _herm(n, seed=3), (_spd(n, seed=4) if with_m else torch.eye(n, dtype=DT))
    info = {}
    evals, evecs = symeig(xt.LinearOperator.m(A, True), neig=neig, mode=mode,
                          M=xt.LinearOperator.m(Mm, True) if with_m else None, method="davidson", nguess=nguess,
                          min_eps=1e-8, info=info)
    ref = _exact(A, Mm, neig, mode)
    assert info["converged"] and info["engine"] == "host-composed"
    assert ((evals - ref).abs() / ref.abs()).max().item() <= 1e-10
    assert (A @ evecs - Mm @ evecs * evals.unsqueeze(-2)).abs().max().item() <= 1e-8
    assert (evecs.T @ Mm @ evecs - torch.eye(neig, dtype=DT)).abs().max().item() <= 1e-10     # M-orthonormal (:182-185)


def test_one_application_of_each_operator_per_iteration(eng):
    """the incremental basis applies A and M to the new block only (the reference applies M to the whole basis,
    symeig.py:212): widths of all applications are <= neig after the start block, counts are niter (+ the start)"""
    n, neig, nguess = 240, 4, 6
    A, Mm = Counting(_herm(n, seed=5)), Counting(_spd(n, seed=6))
    info = {}
    simpl._davidson_host(A, neig, "lowest", Mm, 1000, nguess, "randn", 1e-8, None, None, info, "davidson")
    assert info["converged"]
    assert A.widths[0] == nguess and Mm.widths[0] == nguess
    assert max(A.widths[1:]) <= neig and max(Mm.widths[1:]) <= neig
    assert len(A.widths) == info["napply"] and len(Mm.widths) == info["napply_M"]
    assert info["napply"] <= info["niter"] + 1 and info["napply_M"] <= info["niter"] + 1


def test_thick_restart_and_max_basis(eng):
    n, neig = 400, 4
    A, Mm = _herm(n, seed=7), _spd(n, seed=8)
    info = {}
    evals, _ = symeig(xt.LinearOperator.m(A, True), neig=neig, M=xt.LinearOperator.m(Mm, True), method="davidson",
                      max_basis=20, min_eps=1e-8, info=info)
    assert info["max_basis"] == 20 and info["niter"] > 20 // neig          # went through restarts
    assert ((evals - _exact(A, Mm, neig, "lowest")).abs()).max().item() <= 1e-9


def test_diagonal_preconditioner_cuts_iterations(eng):
    """Davidson's diagonal preconditioner on a diagonally dominant pair; same answer, fewer iterations"""
    n, neig = 600, 6
    A = _herm(n, seed=9) * 0.05 + torch.diag(torch.arange(n, dtype=DT) * 2.0 + 1.0)
    Mm = _spd(n, seed=10) * 0.02 + torch.eye(n, dtype=DT)
    Aop, Mop = xt.LinearOperator.m(A, True), xt.LinearOperator.m(Mm, True)
    plain, pre, usr = {}, {}, {}
    e0, _ = symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=1e-8, info=plain)
    e1, _ = symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=1e-8, precond="diag", info=pre)
    dA, dM = A.diagonal().unsqueeze(-1), Mm.diagonal().unsqueeze(-1)
    e2, _ = symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=1e-8, info=usr,
                   precond=lambda r, lam: r / (dA - lam.unsqueeze(-2) * dM).abs().clamp_min(1e-2))
    ref = _exact(A, Mm, neig, "lowest")
    for e in (e0, e1, e2):
        assert ((e - ref).abs() / ref.abs()).max().item() <= 1e-10
    assert pre["niter"] < plain["niter"] and usr["niter"] < plain["niter"], (plain, pre, usr)
    with pytest.raises(RuntimeError, match="precond"):
        symeig(Aop, neig=neig, method="davidson", precond="nope")
    with pytest.raises(RuntimeError, match="dense"):
        symeig(Counting(A), neig=neig, method="davidson", precond="diag")


def test_batched_and_broadcast(eng):
    n, neig = 120, 3
    A, Mm = _herm(n, 2, seed=11), _spd(n, seed=12)                 # A (2, n, n), M (n, n) broadcast
    evals, evecs = symeig(xt.LinearOperator.m(A, True), neig=neig, M=xt.LinearOperator.m(Mm, True),
                          method="lanczos", min_eps=1e-8)
    assert evals.shape == (2, neig) and evecs.shape == (2, n, neig)
    ref = torch.stack([_exact(A[i], Mm, neig, "lowest") for i in range(2)])
    assert (evals - ref).abs().max().item() <= 1e-9


def test_float32_and_unsupported_dtypes(eng):
    n, neig = 256, 4
    A, Mm = _herm(n, seed=13, dtype=torch.float32), _spd(n, seed=14, dtype=torch.float32)
    evals, evecs = symeig(xt.LinearOperator.m(A, True), neig=neig, M=xt.LinearOperator.m(Mm, True),
                          method="davidson", min_eps=1e-3)
    ref = _exact(A, Mm, neig, "lowest")
    assert evals.dtype == torch.float32
    assert ((evals.double() - ref).abs() / ref.abs()).max().item() <= 1e-4
    with pytest.raises(RuntimeError, match="float32 or float64"):
        symeig(xt.LinearOperator.m(A.to(torch.complex64), True), neig=neig, M=xt.LinearOperator.m(Mm, True),
               method="davidson")


def test_gradients_of_generalized_problem(eng):
    n, k = 8, 2
    A0, M0 = _herm(n, seed=15).requires_grad_(), _spd(n, seed=16).requires_grad_()

    def fcn(A, Mm):
        As, Ms = 0.5 * (A + A.T), 0.5 * (Mm + Mm.T)
        evals, evecs = symeig(xt.LinearOperator.m(As, True), neig=k, M=xt.LinearOperator.m(Ms, True),
                              method="davidson", min_eps=1e-11, bck_options={"method": "exactsolve"})
        return evals, evecs.abs()

    torch.autograd.gradcheck(fcn, (A0, M0), atol=1e-5, rtol=1e-4)


def test_cpu_tensors_still_raise_without_the_standin():
    A, Mm = _herm(16, seed=17), _spd(16, seed=18)
    with pytest.raises(RuntimeError, match="CUDA"):
        symeig(xt.LinearOperator.m(A, True), neig=2, M=xt.LinearOperator.m(Mm, True), method="davidson")


def test_start_block_kinds_iteration_cap_and_metric_batch(eng):
    n, neig = 80, 3
    A, Mm = _herm(n, seed=19), _spd(n, 2, seed=20)                  # M (2, n, n), A (n, n) broadcast
    Aop, Mop = xt.LinearOperator.m(A, True), xt.LinearOperator.m(Mm, True)
    ref = torch.stack([_exact(A, Mm[i], neig, "lowest") for i in range(2)])
    for v_init in ("randn", "rand", "eye"):
        evals, evecs = symeig(Aop, neig=neig, M=Mop, method="davidson", v_init=v_init, min_eps=1e-8)
        assert evals.shape == (2, neig) and evecs.shape == (2, n, neig)
        assert (evals - ref).abs().max().item() <= 1e-9, v_init
    with pytest.raises(ValueError, match="v_init"):
        symeig(Aop, neig=neig, M=Mop, method="davidson", v_init="nope")
    # out of iterations: the best pair so far comes back, flagged as not converged (the reference returns it silently)
    info = {}
    evals, evecs = symeig(Aop, neig=neig, M=Mop, method="davidson", max_niter=3, min_eps=1e-12, info=info)
    assert not info["converged"] and info["niter"] == 3 and info["best_resid"] > 1e-12
    assert evals.shape == (2, neig) and torch.isfinite(evecs).all()
    # info is optional
    symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=1e-6)
