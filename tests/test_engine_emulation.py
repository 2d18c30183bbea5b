"""The WHOLE eigensolver engine on the CPU: csrc/symeig.cu -- the host loop `run_symeig` (kernel sequencing, lagged Ritz
checks, run-ahead window, thick restart, operator callback) and every kernel it launches -- is rewritten textually for a
host compiler (tools/emu_engine: `<<<>>>` launches become host-thread launches, `__shared__` a per-CTA arena, PTX fences
std::atomic fences; the TMA block matvec is replaced by a plain loop) and driven through the real Python host wrapper
(`xitorch_b200.linalg.symeig` -> `_call_engine` -> `xt_symeig_krylov`).  What is compared is what the GPU tests compare:
the reference's committed outputs (tests/golden/krylov_golden.pt: eigenvalues, |eigenvectors|, ITERATION COUNTS), the
residual identity, orthonormality.  This is also the first execution of the matrix-free `apply` hook of the engine.

TEST INFRASTRUCTURE: the emulated library is built in a temporary directory and never shipped; the product on a GPU
box loads libxitorch_b200.so only."""
import os

import warnings

import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200 import _lib
from xitorch_b200.linalg import symeig, svd, solve

import emu_engine_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def engine(emu_lib, monkeypatch):
    emu_engine_lib.install(monkeypatch, emu_lib)
    return emu_lib


def _residual(A, ev, vec):
    return (A.double() @ vec.double() - vec.double() * ev.double().unsqueeze(-2)).abs().max().item()


# ---------------------------------------------------------------------------------------------- reference outputs
@pytest.mark.parametrize("idx", [0, 1, 2])
def test_golden_davidson_cases(engine, golden, idx):
    c = golden["davidson"][idx]
    A = c["A"]
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=c["neig"], mode=c["mode"], method="davidson",
                     min_eps=c["min_eps"], max_basis=c["n"], info=info)
    assert info["converged"]
    # north_star tolerance against the reference's own output and against fp64 eigvalsh
    assert ((ev - c["evals"]).abs() / c["evals"].abs()).max().item() <= 1e-5
    assert ((ev - c["evals_exact_f64"]).abs() / c["evals_exact_f64"].abs()).max().item() <= 1e-9
    assert _residual(A, ev, vec) <= 20 * c["min_eps"]
    assert (vec.abs() - c["evecs_abs"]).abs().max().item() <= 1e-5
    # same Krylov space as the reference: same number of iterations (the reference counts from 0; +-1 for the tie
    # between its residual test and ours on the last step)
    assert abs(info["niter"] - c["oracle_niter"]) <= 1, (info, c["oracle_niter"])


def test_golden_fp32_c2_shape(engine, golden):
    import oracle
    c = golden["davidson"][3]
    A = oracle.make_herm(c["n"], c["neig"], torch.float32, seed=c["seed"])
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=c["neig"], method="davidson",
                     min_eps=c["min_eps"], info=info)
    assert info["converged"] and abs(info["niter"] - c["oracle_niter"]) <= 1
    assert ((ev.double() - c["evals_exact_f64"]).abs() / c["evals_exact_f64"].abs()).max().item() <= 1e-5
    assert _residual(A, ev, vec) <= 20 * c["min_eps"]
    # the lagged Ritz check costs at most two applications beyond the reference's count
    assert info["napply"] <= info["niter"] + 3


# ---------------------------------------------------------------------------------------------- engine paths
def _herm(n, k, dtype=torch.float64, seed=3):
    import oracle
    return oracle.make_herm(n, k, dtype, seed=seed)


def test_lanczos_and_residual_expansion_agree(engine):
    A = _herm(96, 4)
    ref = torch.linalg.eigvalsh(A)[:4]
    infos = {}
    for name, kw in (("lanczos", dict(method="lanczos")), ("krylov", dict(method="davidson")),
                     ("residual", dict(method="davidson", expansion="residual"))):
        infos[name] = {}
        ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, min_eps=1e-8, info=infos[name], **kw)
        assert infos[name]["converged"], name
        assert ((ev - ref).abs() / ref.abs()).max().item() <= 1e-9, name
        assert _residual(A, ev, vec) <= 2e-7, name
        assert (vec.T @ vec - torch.eye(4, dtype=A.dtype)).abs().max().item() <= 1e-10
    # identical Krylov spaces: identical iteration counts, whichever way the subspace is expanded
    assert infos["lanczos"]["niter"] == infos["krylov"]["niter"]
    assert abs(infos["residual"]["niter"] - infos["krylov"]["niter"]) <= 1


def test_thick_restart(engine):
    # a basis cap far below what convergence needs: several restarts, still the right pairs
    g = torch.Generator().manual_seed(5)
    n = 128
    A = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (A + A.T) / (2 * n) ** 0.5 + torch.diag(torch.linspace(1, 3, n, dtype=torch.float64))
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, method="davidson", min_eps=1e-6,
                     max_basis=24, info=info)
    ref = torch.linalg.eigvalsh(A)[:4]
    assert info["converged"] and info["niter"] > 24 // 4          # it did have to restart
    assert ((ev - ref).abs() / ref.abs()).max().item() <= 1e-8
    assert _residual(A, ev, vec) <= 2e-5


def test_uppest_float32_odd_size(engine):
    A = _herm(102, 3, torch.float32, seed=9)                     # n not a multiple of the 64-row chunks
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=3, mode="uppest", method="davidson",
                     min_eps=2e-4, info=info)
    ref = torch.linalg.eigvalsh(A.double())[-3:]
    assert info["converged"]
    assert ((ev.double() - ref).abs() / ref.abs()).max().item() <= 1e-5
    assert _residual(A, ev, vec) <= 20 * 2e-4


# ---------------------------------------------------------------------------------------------- matrix-free hook
class UserOperator(xt.LinearOperator):
    def __init__(self, mat):
        super().__init__(shape=mat.shape, is_hermitian=True, dtype=mat.dtype, device=mat.device)
        self.mat = mat
        self.napply = 0

    def _mv(self, x):
        self.napply += 1
        return torch.matmul(self.mat, x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_matrix_free_hook_matches_dense_path(engine, dtype):
    A = _herm(96, 4, dtype)
    eps = 1e-4 if dtype == torch.float32 else 1e-8
    info_d, info_f = {}, {}
    ev_d, vec_d = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, method="davidson", min_eps=eps, info=info_d)
    op = UserOperator(A)
    ev_f, vec_f = symeig(op, neig=4, method="davidson", min_eps=eps, matrix_free=True, info=info_f)
    assert info_f["converged"] and info_f["niter"] == info_d["niter"]
    assert op.napply == info_f["napply"] > 0                          # one callback per application, none extra
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    assert ((ev_f - ev_d).abs() / ev_d.abs()).max().item() <= tol
    assert (vec_f.abs() - vec_d.abs()).abs().max().item() <= (1e-3 if dtype == torch.float32 else 1e-8)
    assert _residual(A, ev_f, vec_f) <= 20 * eps


def test_matrix_free_generalized_and_svd(engine):
    n, k = 96, 3
    A = _herm(n, k)
    g = torch.Generator().manual_seed(8)
    Mh = torch.randn(n, n, generator=g, dtype=torch.float64) / n ** 0.5
    Mm = Mh @ Mh.T * 0.1 + torch.eye(n, dtype=torch.float64)
    ev, X = symeig(UserOperator(A), neig=k, M=xt.LinearOperator.m(Mm, is_hermitian=True), method="davidson",
                   matrix_free=True, min_eps=1e-8)
    assert (A @ X - Mm @ X * ev).abs().max().item() <= 1e-6
    assert (X.T @ Mm @ X - torch.eye(k, dtype=torch.float64)).abs().max().item() <= 1e-9

    B = torch.randn(140, 64, generator=g, dtype=torch.float64) / 12
    B[:4, :4] += torch.diag(torch.tensor([6.0, 5.0, 4.0, 3.0], dtype=torch.float64))

    class Rect(xt.LinearOperator):
        def __init__(self):
            super().__init__(shape=B.shape, dtype=B.dtype, device=B.device)

        def _mv(self, x):
            return torch.matmul(B, x.unsqueeze(-1)).squeeze(-1)

        def _rmv(self, y):
            return torch.matmul(B.T, y.unsqueeze(-1)).squeeze(-1)

        def _getparamnames(self, prefix=""):
            return []

    u, s, vh = svd(Rect(), k=2, mode="uppest", method="davidson", matrix_free=True, min_eps=1e-9)
    sref = torch.linalg.svdvals(B)[:2]
    assert ((s.sort(descending=True).values - sref).abs() / sref).max().item() <= 1e-8
    assert (B @ vh.transpose(-2, -1) - u * s.unsqueeze(-2)).abs().max().item() <= 1e-6


# ---------------------------------------------------------------------------------------------- small dense eigh entry
def test_small_eigh_entry(engine):
    g = torch.Generator().manual_seed(2)
    m, nev = 40, 6
    T = torch.randn(m, m, generator=g, dtype=torch.float64)
    T = (T + T.T) / 2
    w = torch.zeros(nev, dtype=torch.float64)
    S = torch.zeros(m, nev, dtype=torch.float64)
    scratch = torch.zeros(m * (m | 1) + 32, dtype=torch.float64)
    rc = engine.xt_small_eigh(T.data_ptr(), m, nev, 0, w.data_ptr(), S.data_ptr(), scratch.data_ptr(), None)
    assert rc == 0
    assert torch.allclose(w, torch.linalg.eigvalsh(T)[:nev], atol=1e-12)
    assert (T @ S - S * w).abs().max().item() <= 1e-11


# ---------------------------------------------------------------------------------------------- linear solvers
def _golden_solve(c):
    A = xt.LinearOperator.m(c["A"], is_hermitian=c["herm"])
    M = xt.LinearOperator.m(c["M"], is_hermitian=True) if c.get("M") is not None else None
    info = {}
    x = solve(A, c["B"], E=c.get("E"), M=M, method=c["method"], info=info, **c["opts"])
    return x, info


def test_golden_solve_cases(engine, golden):
    """every committed output of the reference's cg / bicgstab / gmres (plain, batched, non-symmetric, normal equations,
    E and M): the same solution, and the same number of iterations (cg exactly; bicgstab / gmres count the step that
    detects convergence differently by one)"""
    for c in golden["solve"]:
        x, info = _golden_solve(c)
        assert info["converged"], c["tag"]
        rtol = c["opts"].get("rtol", 1e-6)
        assert ((x - c["x"]).norm() / c["x"].norm()).item() <= 100 * rtol, (c["tag"], c["method"])
        assert ((x - c["x_exact"]).norm() / c["x_exact"].norm()).item() <= 100 * rtol, (c["tag"], c["method"])
        slack = 0 if c["method"] == "cg" else 1
        assert abs(info["niter"] - c["oracle_niter"]) <= slack, (c["tag"], c["method"], info["niter"], c["oracle_niter"])


@pytest.mark.parametrize("slices", [3])
def test_golden_solve_cases_with_sliced_step_kernels(engine, golden, slices, monkeypatch):
    """the sliced step kernels of cg / bicgstab / gmres (rows of a system split over co-resident CTAs, deterministic
    cross-slice reductions through the arrival counter; csrc/solve_common.cuh `slice_allreduce`, csrc/gmres.cu
    `gm_step_sliced_kernel`) on the host build -- cooperative launches run all CTAs at once there: the reference's
    committed solutions and iteration counts again, and the same result on a second run"""
    monkeypatch.setenv("XT_EMU_SLICES", str(slices))
    seen = set()
    for c in golden["solve"]:
        x, info = _golden_solve(c)
        assert info["converged"], c["tag"]
        rtol = c["opts"].get("rtol", 1e-6)
        assert ((x - c["x"]).norm() / c["x"].norm()).item() <= 100 * rtol, (c["tag"], c["method"])
        slack = 0 if c["method"] == "cg" else 1
        assert abs(info["niter"] - c["oracle_niter"]) <= slack, (c["tag"], c["method"], info["niter"], c["oracle_niter"])
        if c["method"] not in seen:                 # once per method: slice-ordered sums give the same bits again
            seen.add(c["method"])
            x2, info2 = _golden_solve(c)
            assert torch.equal(x, x2) and info["niter"] == info2["niter"], c["tag"]


@pytest.mark.parametrize("method", ["cg", "bicgstab", "gmres"])
def test_sliced_step_kernels_against_one_cta(engine, method, monkeypatch):
    """rows split unevenly over the slices (n = 131, 6 slices of 22, the last one 21):
    dense solution, and the one-CTA kernels' iteration count; the sums are taken in another order, so the iterates
    agree to rounding but not to the bit -- which also shows that the sliced path ran"""
    n, nc = 131, 3
    g = torch.Generator().manual_seed(31)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = a @ a.T / n + 0.5 * torch.eye(n, dtype=torch.float64) if method != "gmres" else \
        torch.eye(n, dtype=torch.float64) * 2.0 + a / n ** 0.5
    B = torch.randn(n, nc, generator=g, dtype=torch.float64)
    op = xt.LinearOperator.m(A, is_hermitian=(method != "gmres"))
    kw = dict(rtol=1e-10, atol=1e-14)
    i1, i6 = {}, {}
    x1 = solve(op, B, method=method, info=i1, **kw)
    monkeypatch.setenv("XT_EMU_SLICES", "6")
    x6 = solve(op, B, method=method, info=i6, **kw)
    ref = torch.linalg.solve(A, B)
    assert i6["converged"] and ((x6 - ref).norm() / ref.norm()).item() <= 1e-8
    assert abs(i6["niter"] - i1["niter"]) <= 1, (i1, i6)
    assert ((x6 - x1).norm() / ref.norm()).item() <= 1e-8
    assert not torch.equal(x6, x1)


@pytest.mark.parametrize("method,rce,slices", [("cg", 10, 1), ("cg", 0, 4), ("bicgstab", 3, 1)])
def test_replayed_periods_number_their_iterations_relatively(engine, method, rce, slices, monkeypatch):
    """launch-bound solves replay whole periods of iterations (on the GPU from an instantiated CUDA graph): the step
    kernels of a replayed period get iteration numbers and reduction epochs RELATIVE to `SolveCtl::graph_base /
    epoch_base`, which a one-thread kernel advances after every period; the true-residual iteration closes a period; the
    stop flag is looked at one period late; the last iterations go out as plain launches again.  The host build replays
    a period by enqueueing it again (XT_EMU_GRAPH=1), everything else is the shipped code: same iterates, to the bit, as
    the plain loop -- iteration count, best iterate, solution -- with and without sliced step kernels."""
    n, nc = 90, 2
    g = torch.Generator().manual_seed(41)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = a @ a.T / n + 0.02 * torch.eye(n, dtype=torch.float64)              # ~70-110 iterations at rtol = 1e-10
    B = torch.randn(n, nc, generator=g, dtype=torch.float64)
    op = xt.LinearOperator.m(A, is_hermitian=True)
    kw = dict(rtol=1e-10, atol=1e-14, resid_calc_every=rce)
    monkeypatch.setenv("XT_EMU_SLICES", str(slices))
    i0, i1, i2 = {}, {}, {}
    x0 = solve(op, B, method=method, info=i0, **kw)
    monkeypatch.setenv("XT_EMU_GRAPH", "1")
    x1 = solve(op, B, method=method, info=i1, **kw)
    assert i0["converged"] and i0["niter"] > 40, i0                        # several replayed periods
    assert i1["niter"] == i0["niter"] and i1["converged"]
    assert torch.equal(x1, x0)
    # an iteration cap that leaves a tail of plain launches after the last whole period, without convergence
    cap = 47
    monkeypatch.delenv("XT_EMU_GRAPH")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xa = solve(op, B, method=method, info=i2, max_niter=cap, **kw)
        monkeypatch.setenv("XT_EMU_GRAPH", "1")
        i3 = {}
        xb = solve(op, B, method=method, info=i3, max_niter=cap, **kw)
    assert i2["niter"] == i3["niter"] == cap and not i3["converged"]
    assert torch.equal(xa, xb)


def test_golden_preconditioned_cases(engine, golden):
    for c in golden["solve_precond"]:
        A = xt.LinearOperator.m(c["A"], is_hermitian=c["herm"])
        pre = {k: xt.LinearOperator.m(v) for k, v in c["precond"].items()}
        info = {}
        x = solve(A, c["B"], method=c["method"], info=info, **pre, **c["opts"])
        assert ((x - c["x_exact"]).norm() / c["x_exact"].norm()).item() <= 1e-7, c["tag"]
        if c["oracle_niter"] < c["plain_niter"]:            # the preconditioner pays off exactly as in the reference
            assert abs(info["niter"] - c["oracle_niter"]) <= 1, (c["tag"], info["niter"], c["oracle_niter"])


def test_complex_and_float32_systems(engine):
    g = torch.Generator().manual_seed(12)
    n = 40
    a = torch.randn(n, n, generator=g, dtype=torch.complex128)
    A = a @ a.conj().T / n + 2 * torch.eye(n, dtype=torch.complex128)
    B = torch.randn(n, 2, generator=g, dtype=torch.complex128)
    x = solve(xt.LinearOperator.m(A, is_hermitian=True), B, method="cg", rtol=1e-10)
    assert (A @ x - B).abs().max().item() <= 1e-8
    Af = (A.real + torch.eye(n)).float()
    Bf = B.real.float()
    xf = solve(xt.LinearOperator.m(Af, is_hermitian=True), Bf, method="bicgstab", rtol=1e-6)
    assert xf.dtype == torch.float32 and (Af @ xf - Bf).abs().max().item() <= 1e-4


def test_rootfinder_backward_through_the_solver_engine(engine):
    """BASELINE config 4 in small: Broyden forward, then the adjoint solve of the backward pass with the matrix-free
    Jacobian driven by the bicgstab ENGINE through its operator callback; gradient against the exact implicit one"""
    import oracle
    from xitorch_b200.optimize import rootfinder

    def fcn(y, A):
        return torch.tanh(A @ y + 0.1) + y / 2.0

    n = 24
    A, _ = oracle.make_rootfinder_c4(n, dtype=torch.float64)
    Ar = A.clone().requires_grad_()
    y = rootfinder(fcn, torch.zeros(n, 1, dtype=torch.float64), params=(Ar,),
                   bck_options={"method": "bicgstab", "rtol": 1e-10})
    (g,) = torch.autograd.grad(y.sum(), Ar)
    (g_exact,) = oracle.implicit_grad_dense(fcn, y.detach(), (A,), torch.ones_like(y))
    assert (g - g_exact).abs().max().item() <= 1e-8 * max(1.0, g_exact.abs().max().item())


# ---------------------------------------------------------------------------------------------- tiny problems
@pytest.mark.parametrize("n,k,mode", [(9, 4, "lowest"), (70, 16, "uppest")])
def test_subspace_filling_the_whole_space(engine, n, k, mode):
    """n so small that the block-wise subspace runs out of room before the residual test passes: the reference then
    adds a partial block, its subspace becomes the whole space and the next Rayleigh-Ritz is exact (symeig.py:204-211).
    The engine grows in whole blocks only; the host wrapper completes the space the same way."""
    g = torch.Generator().manual_seed(n + k)
    A = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (A + A.T) / (2 * n) ** 0.5 + torch.diag(torch.linspace(1, 10, n, dtype=torch.float64))
    info = {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=k, mode=mode, method="davidson", min_eps=1e-8,
                     info=info)
    w = torch.linalg.eigvalsh(A)
    ref = w[:k] if mode == "lowest" else w[-k:]
    assert info["converged"] and info.get("completed_full_space")
    assert ((ev - ref).abs() / ref.abs()).max().item() <= 1e-12
    assert _residual(A, ev, vec) <= 1e-10


def test_batch_of_small_operators(engine):
    g = torch.Generator().manual_seed(21)
    A = torch.randn(3, 40, 40, generator=g, dtype=torch.float64)
    A = (A + A.transpose(-2, -1)) / 80 ** 0.5 + torch.diag(torch.linspace(1, 10, 40, dtype=torch.float64))
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=3, method="davidson", min_eps=1e-8)
    assert (ev - torch.linalg.eigvalsh(A)[..., :3]).abs().max().item() <= 1e-9
    assert _residual(A, ev, vec) <= 1e-7


# ---------------------------------------------------------------------------------------------- engine code paths
@pytest.mark.parametrize("env", [{}, {"XT_SYNC_CHECK": "1", "XT_LAG1_M": "128"}, {"XT_SYNC_CHECK": "1", "XT_LAG1_M": "0"},
                                 {"XT_NO_FUSE": "1"}, {"XT_PO_NOSTAGE": "1"}, {"XT_ROTATE_NAIVE": "1"},
                                 {"XT_SYNC_CHECK": "1", "XT_ROTATE_NAIVE": "1"}])
def test_alternative_engine_paths_give_the_same_answer(engine, monkeypatch, env):
    """the switches select code paths the default run at test sizes does not take: the in-stream Ritz checks (P0 of the
    fused kernel) lagging one or two iterations instead of the asynchronous checks at the end of the Rayleigh-Ritz
    kernels, the multi-kernel iteration without the fused cooperative kernel, the basis read from L2 instead of staged
    in shared memory, the untiled restart rotation (with thick restarts, i.e. the hand-over between asynchronous checks
    and the restart path).  Same eigenpairs, same iteration count.  On this host build every launch completes before the
    next one starts, so an asynchronous check stops the solve with exactly one application per iteration; an in-stream
    check lagging one (two) iterations costs exactly one (two) applications more."""
    for k_, v_ in env.items():
        monkeypatch.setenv(k_, v_)
    A = _herm(96, 4)
    info = {}
    kw = dict(max_basis=24) if "XT_ROTATE_NAIVE" in env else {}
    ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, method="davidson", min_eps=1e-8, info=info, **kw)
    ref = torch.linalg.eigvalsh(A)[:4]
    assert info["converged"]
    assert ((ev - ref).abs() / ref.abs()).max().item() <= 1e-9
    assert _residual(A, ev, vec) <= 2e-7
    if not env:
        assert info["napply"] == info["niter"]
    if env.get("XT_LAG1_M") == "128":
        assert info["napply"] == info["niter"] + 1
    if env.get("XT_LAG1_M") == "0":
        assert info["napply"] == info["niter"] + 2


@pytest.mark.parametrize("dtype,mode,kw", [
    (torch.float64, "lowest", dict(max_niter=5)),                 # stops unconverged: the best pair is returned
    (torch.float32, "uppest", dict(max_niter=6)),
    (torch.float64, "lowest", dict(max_basis=16)),                # thick restarts between asynchronous checks
    (torch.float32, "lowest", dict(max_basis=24, max_niter=9)),   # restarts AND no convergence
])
def test_async_checks_match_in_stream_checks(engine, monkeypatch, dtype, mode, kw):
    """The asynchronous Ritz check (Lanczos residual formula R = Q_{j+1} L^T S_last at the end of rr_kernel, best pair
    kept as coefficients and turned into Ritz vectors once) against the in-stream check that forms X = V S and
    R = AV S - X theta from the whole basis (XT_SYNC_CHECK=1): same iteration count, same stop decision, same residual
    maximum to rounding, same eigenpairs -- also when the iteration limit ends the solve (best pair bookkeeping) and
    across thick restarts (coefficients materialised before the basis is rotated)."""
    A = _herm(96, 4, dtype=dtype)
    out = {}
    for sync in ("0", "1"):
        if sync == "1":
            monkeypatch.setenv("XT_SYNC_CHECK", "1")
        else:
            monkeypatch.delenv("XT_SYNC_CHECK", raising=False)
        info = {}
        ev, vec = symeig(xt.LinearOperator.m(A, is_hermitian=True), neig=4, mode=mode, method="davidson",
                         min_eps=1e-9 if dtype == torch.float64 else 1e-5, info=info, **kw)
        out[sync] = (ev, vec, info)
    (ev0, vec0, i0), (ev1, vec1, i1) = out["0"], out["1"]
    assert i0["niter"] == i1["niter"] and i0["converged"] == i1["converged"], (i0, i1)
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    assert (ev0 - ev1).abs().max().item() <= tol
    assert (vec0.abs() - vec1.abs()).abs().max().item() <= (1e-8 if dtype == torch.float64 else 2e-4)
    assert abs(i0["best_resid"] - i1["best_resid"]) <= 1e-3 * i1["best_resid"] + (1e-13 if dtype == torch.float64 else 2e-6)
    # the reported residual maximum is the true one of the returned pair
    true = _residual(A, ev0, vec0)
    assert abs(true - i0["best_resid"]) <= 1e-3 * true + (1e-12 if dtype == torch.float64 else 1e-5)


# ---------------------------------------------------------------------------------------------- Hermiticity check
def test_hermitian_check_kernel_matches_allclose(engine):
    """`xt_hermitian_check` (csrc/linop.cu, what LinearOperator.m runs on CUDA matrices) against torch.allclose(A, A^T):
    symmetric matrices of awkward sizes, a single violating entry in every kind of tile position, the tolerance boundary
    in both directions, NaN, batches, padded leading dimension, both precisions"""
    from xitorch_b200 import _dense
    g = torch.Generator().manual_seed(0)

    def ref(m):
        return bool(torch.allclose(m, m.transpose(-2, -1)))

    for dtype in (torch.float32, torch.float64):
        for n in (1, 5, 32, 33, 64, 70, 97):
            a = torch.randn(n, n, generator=g, dtype=dtype)
            sym = a + a.T
            assert _dense.hermitian_check(sym) and ref(sym), (dtype, n)
            for (i, j) in {(0, n - 1), (n - 1, 0), (n // 2, n // 3), (min(31, n - 1), min(32, n - 1)), (n - 1, n - 2)}:
                if i == j or min(i, j) < 0:
                    continue
                bad = sym.clone()
                bad[i, j] += 1e-2 * (1 + bad[i, j].abs())
                assert _dense.hermitian_check(bad) == ref(bad) == False, (dtype, n, i, j)       # noqa: E712
            if n > 1:
                # inside the tolerance in both directions / outside in one
                near = sym.clone()
                near[0, 1] = near[1, 0] * (1 + 5e-6)
                assert _dense.hermitian_check(near) == ref(near), (dtype, n)
                tiny = sym.clone()
                tiny[0, 1], tiny[1, 0] = 3e-9, -3e-9                 # |d| = 6e-9 <= atol = 1e-8
                assert _dense.hermitian_check(tiny) == ref(tiny) == True      # noqa: E712
                tiny[0, 1] = 3e-8
                assert _dense.hermitian_check(tiny) == ref(tiny) == False     # noqa: E712
                nan = sym.clone()
                nan[1, 0] = nan[0, 1] = float("nan")
                assert _dense.hermitian_check(nan) == ref(nan) == False       # noqa: E712
    # batch: one bad item spoils the verdict; padded rows (lda > n)
    a = torch.randn(3, 40, 40, generator=g, dtype=torch.float64)
    sym = a + a.transpose(-2, -1)
    assert _dense.hermitian_check(sym)
    sym[2, 7, 30] += 1.0
    assert not _dense.hermitian_check(sym)
    big = torch.randn(64, 64, generator=g, dtype=torch.float32)
    big = big + big.T
    assert _dense.hermitian_check(big[:45, :45]) and not _dense.hermitian_check(big[:45, 1:46])
    # and through the public entry: LinearOperator.m trusts / rejects the flag accordingly (CPU tensors take torch's test)
    with pytest.raises(RuntimeError, match="hermitian"):
        xt.LinearOperator.m(sym[2], is_hermitian=True)
