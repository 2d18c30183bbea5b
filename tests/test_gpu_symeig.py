"""GPU parity of davidson / lanczos (public API -> C ABI `xt_symeig_krylov`) against the committed reference
outputs, the CPU oracle and fp64 eigvalsh.  Tolerance: eigenvalues within 1e-5 relative (north_star)."""
import ctypes

import pytest
import torch

import oracle
import xitorch_b200 as xt
from xitorch_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"
EIG_RTOL = 1e-5


def _check_pairs(A, evals, evecs, ref_evals, min_eps):
    A64 = A.double()
    ev = evals.double().cpu()
    X = evecs.double().cpu()
    rel = ((ev - ref_evals.double()).abs() / ref_evals.double().abs()).max().item()
    assert rel <= EIG_RTOL, (rel, ev, ref_evals)
    resid = (A64 @ X - X * ev.unsqueeze(-2)).abs().max().item()
    assert resid <= 20 * min_eps + 1e-12, resid
    gram = X.transpose(-2, -1) @ X
    assert (gram - torch.eye(gram.shape[-1], dtype=torch.float64)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("method", ["davidson", "davidson-residual", "lanczos"])
def test_golden_davidson_cases(golden, method):
    extra = {}
    if method == "davidson-residual":
        method, extra = "davidson", {"expansion": "residual"}
    for case in golden["davidson"]:
        n, neig, mode, dtype = case["n"], case["neig"], case["mode"], case["dtype"]
        if case["A"] is not None:
            A = case["A"]
        else:
            A = oracle.make_herm(n, neig, torch.float64, seed=case["seed"]).to(dtype)
        info = {}
        evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A.to(DEV), True), neig=neig, mode=mode,
                                        method=method, min_eps=case["min_eps"], info=info, **extra)
        assert evals.dtype == dtype and tuple(evals.shape) == (*case["batch"], neig)
        assert info["converged"], (case["n"], mode, info)
        _check_pairs(A, evals, evecs, case["evals"], case["min_eps"])                # vs the reference's davidson
        _check_pairs(A, evals, evecs, case["evals_exact_f64"], case["min_eps"])      # vs fp64 eigvalsh
        if case["evecs_abs"] is not None and dtype == torch.float64:
            assert (evecs.abs().cpu() - case["evecs_abs"]).abs().max().item() <= 1e-5
        # same Krylov space => about the same number of expansions as the reference algorithm (rounding at the
        # threshold shifts the stopping iteration by a few)
        assert abs(info["niter"] - case["oracle_niter"]) <= max(3, case["oracle_niter"] // 8), (info, case["oracle_niter"])


def test_generalized_problem(golden):
    c = golden["davidson_M"]
    evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(c["A"].to(DEV), True), neig=c["neig"],
                                    M=xt.LinearOperator.m(c["M"].to(DEV), True), method="davidson",
                                    min_eps=c["min_eps"])
    assert ((evals.cpu() - c["evals"]).abs() / c["evals"].abs()).max().item() <= EIG_RTOL
    X = evecs.cpu()
    assert (c["A"] @ X - c["M"] @ X * evals.cpu().unsqueeze(-2)).abs().max().item() <= 1e-6


def _gen_exact(A, M, k, mode="lowest"):
    Li = torch.inverse(torch.linalg.cholesky(M.double()))
    w = torch.linalg.eigvalsh(Li @ A.double() @ Li.transpose(-2, -1))
    return w[..., :k] if mode == "lowest" else w[..., -k:]


def _spd_metric(n, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, 64, generator=g, dtype=torch.float64)
    return (a @ a.T / 64 * 0.2 + torch.eye(n, dtype=torch.float64)).to(dtype)


@pytest.mark.parametrize("n,dtype,min_eps", [(4096, torch.float32, 1e-5), (1536, torch.float64, 1e-9)])
def test_generalized_problem_through_the_matvec_kernel(n, dtype, min_eps):
    """A x = lambda M x with dense A and M: both are applied with the block-matvec kernel, one launch each per iteration,
    and nothing of order n^3 (no Cholesky whitening, no n x n temporaries) is formed -- checked on the allocator's
    high-water mark -- against fp64 generalized eigh (reference: symeig.py:182-185, 212-214)."""
    neig = 8
    A = oracle.make_herm(n, neig, dtype)
    M = _spd_metric(n, dtype, 5)
    Ad, Md = A.to(DEV), M.to(DEV)
    Aop, Mop = xt.LinearOperator.m(Ad, True), xt.LinearOperator.m(Md, True)
    xt.linalg.symeig(Aop, neig=2, M=Mop, method="davidson", min_eps=1e-2)      # library handles etc. allocated
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    _lib.profile_reset(True)
    info = {}
    evals, evecs = xt.linalg.symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=min_eps, info=info)
    torch.cuda.synchronize()
    _, mv_launches, _ = _lib.profile_read()
    _lib.profile_reset(False)
    extra = torch.cuda.max_memory_allocated() - base
    assert info["converged"] and info["engine"] == "host-composed", info
    assert extra < 0.5 * n * n * A.element_size(), (extra, n * n * A.element_size())
    assert mv_launches >= info["napply"] + info["napply_M"] > 0, (mv_launches, info)   # A and M through OUR kernel
    ref = _gen_exact(A, M, neig)
    assert ((evals.double().cpu() - ref).abs() / ref.abs()).max().item() <= EIG_RTOL
    X = evecs.double().cpu()
    assert (A.double() @ X - M.double() @ X * evals.double().cpu().unsqueeze(-2)).abs().max().item() <= 20 * min_eps
    assert (X.T @ M.double() @ X - torch.eye(neig, dtype=torch.float64)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("with_m", [False, True])
def test_wide_start_block_and_preconditioner(with_m):
    """nguess > neig (reference symeig.py:137-138) and Davidson's diagonal preconditioner"""
    n, neig = 2048, 6
    g = torch.Generator().manual_seed(11)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    A = (0.5 * (a + a.T) * 0.05 + torch.diag(torch.arange(n, dtype=torch.float64) * 2.0 + 1.0))
    M = _spd_metric(n, torch.float64, 6) if with_m else None
    Aop = xt.LinearOperator.m(A.to(DEV), True)
    Mop = xt.LinearOperator.m(M.to(DEV), True) if with_m else None
    ref = _gen_exact(A, M, neig) if with_m else torch.linalg.eigvalsh(A)[:neig]
    runs = {}
    for tag, kw in (("wide", dict(nguess=10)), ("plain", dict(nguess=neig + 1)), ("diag", dict(nguess=neig + 1, precond="diag"))):
        info = {}
        evals, evecs = xt.linalg.symeig(Aop, neig=neig, M=Mop, method="davidson", min_eps=1e-8, info=info, **kw)
        assert info["converged"] and info["engine"] == "host-composed", (tag, info)
        assert ((evals.cpu() - ref).abs() / ref.abs()).max().item() <= 1e-9, tag
        runs[tag] = info["niter"]
    assert runs["diag"] < runs["plain"], runs


@pytest.mark.parametrize("method", ["davidson", "lanczos"])
def test_fp32_2048_vs_oracle_fp64(method):
    """fp32 operator at N=2048: the reference's own fp32 davidson only survives min_eps >= 1e-4
    (SURVEY.md 8a A3); ours is compared with the oracle run in fp64 on the same fp32 matrix."""
    n, neig = 2048, 8
    A = oracle.make_herm(n, neig, torch.float32)
    ev_o, _ = oracle.davidson(A.double(), neig, "lowest", min_eps=1e-9)
    info = {}
    evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A.to(DEV), True), neig=neig, method=method,
                                    min_eps=1e-4, info=info)
    assert info["converged"], info
    _check_pairs(A, evals, evecs, ev_o, 1e-4)


def test_thick_restart_slow_spectrum():
    """slow-converging shifted GOE matrix with a small restart cap: exercises the thick restart; the
    converged pairs must still match fp64 eigvalsh."""
    n, neig = 1024, 4
    A = oracle.make_slow_herm(n, torch.float64)
    ref = torch.linalg.eigvalsh(A)[:neig]
    info = {}
    for expansion in ("krylov", "residual"):
        info = {}
        evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A.to(DEV), True), neig=neig, method="davidson",
                                        min_eps=1e-7, max_basis=32, max_niter=2000, info=info, expansion=expansion)
        assert info["converged"], info
        _check_pairs(A, evals, evecs, ref, 1e-7)
        assert info["niter"] > 32 // neig          # i.e. at least one restart happened


def test_uppest_and_batch():
    A = torch.stack([oracle.make_herm(200, 4, torch.float64, seed=s) for s in (1, 2, 3)])
    ref = torch.linalg.eigvalsh(A)[..., -3:]
    evals, evecs = xt.linalg.usymeig(xt.LinearOperator.m(A.to(DEV), True), neig=3, method="davidson", min_eps=1e-8)
    assert tuple(evals.shape) == (3, 3) and tuple(evecs.shape) == (3, 200, 3)
    assert ((evals.cpu() - ref).abs() / ref.abs()).max().item() <= EIG_RTOL


def _small_eigh(T, nev, mode):
    L = _lib.lib()
    m = T.shape[0]
    Td = T.contiguous().to(DEV)
    w = torch.zeros(nev, dtype=torch.float64, device=DEV)
    S = torch.zeros(m, nev, dtype=torch.float64, device=DEV)
    scratch = torch.zeros(m * (m | 1) + 16, dtype=torch.float64, device=DEV)
    vp = ctypes.c_void_p
    rc = L.xt_small_eigh(vp(Td.data_ptr()), m, nev, mode, vp(w.data_ptr()), vp(S.data_ptr()), vp(scratch.data_ptr()),
                         vp(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.xt_last_error()
    torch.cuda.synchronize()
    return w.cpu(), S.cpu()


@pytest.mark.parametrize("m", [2, 3, 8, 33, 104, 128, 160, 200, 256])
@pytest.mark.parametrize("mode", [0, 1])
def test_small_eigh_kernel_random(m, mode):
    g = torch.Generator().manual_seed(m)
    T = torch.randn(m, m, generator=g, dtype=torch.float64)
    T = (T + T.t()) / 2
    nev = min(8, m)
    w, S = _small_eigh(T, nev, mode)
    wr = torch.linalg.eigvalsh(T)
    wr = wr[:nev] if mode == 0 else wr[-nev:]
    scale = max(1.0, wr.abs().max().item())
    assert (w - wr).abs().max().item() <= 1e-12 * scale * m
    assert (T @ S - S * w).abs().max().item() <= 1e-11 * scale * m
    assert (S.t() @ S - torch.eye(nev, dtype=torch.float64)).abs().max().item() <= 1e-12


def test_small_eigh_kernel_special_cases():
    g = torch.Generator().manual_seed(0)
    m = 96
    Q, _ = torch.linalg.qr(torch.randn(m, m, generator=g, dtype=torch.float64))
    # degenerate and tightly clustered spectra, a diagonal matrix, a block-tridiagonal matrix, many pairs (restart)
    spectra = [torch.cat([torch.tensor([1., 1., 1., 2., 2.], dtype=torch.float64), torch.linspace(3, 30, m - 5, dtype=torch.float64)]),
               torch.cat([1 + 1e-9 * torch.arange(4, dtype=torch.float64), torch.linspace(2, 5, m - 4, dtype=torch.float64)])]
    mats = [(Q * w) @ Q.t() for w in spectra]
    mats.append(torch.diag(torch.arange(m, 0, -1, dtype=torch.float64)))
    bt = torch.diag(torch.arange(1.0, m + 1, dtype=torch.float64))
    bt = bt + torch.diag(torch.full((m - 8,), 0.1, dtype=torch.float64), 8) + torch.diag(torch.full((m - 8,), 0.1, dtype=torch.float64), -8)
    mats.append(bt)
    for T in mats:
        T = (T + T.t()) / 2
        for nev, mode in ((8, 0), (8, 1), (48, 0), (48, 1), (96, 0)):
            w, S = _small_eigh(T, nev, mode)
            wr = torch.linalg.eigvalsh(T)
            wr = wr[:nev] if mode == 0 else wr[-nev:]
            assert (w - wr).abs().max().item() <= 1e-11 * m
            assert (T @ S - S * w).abs().max().item() <= 1e-10 * m
            assert (S.t() @ S - torch.eye(nev, dtype=torch.float64)).abs().max().item() <= 1e-11


def test_symeig_backward_with_krylov_adjoint():
    """symeig backward re-enters `solve(A, -B, E=evals)` with the CUDA cg (symeig.py:365-367)."""
    n, neig = 128, 3
    A0 = oracle.make_herm(n, neig, torch.float64, seed=4)
    A = A0.to(DEV).requires_grad_()
    bck = {"method": "cg", "rtol": 1e-11, "atol": 1e-14, "posdef": True}
    evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="davidson", min_eps=1e-9,
                                    bck_options=bck)
    loss = evals.sum() + (evecs.abs() ** 4).sum()
    (gA,) = torch.autograd.grad(loss, (A,))
    Ar = A0.clone().requires_grad_()
    ev, vec = torch.linalg.eigh((Ar + Ar.t()) / 2)
    lr = ev[:neig].sum() + (vec[:, :neig].abs() ** 4).sum()
    (gr,) = torch.autograd.grad(lr, (Ar,))
    gsym = (gA + gA.t()).cpu() / 2
    assert torch.allclose(gsym, gr, rtol=1e-5, atol=1e-7), (gsym - gr).abs().max()


def test_full_size_residual_property():
    """BASELINE configs[1] at full size (N=16384, neig=8, fp32): size-independent checks -- residual
    identity ||A x - lambda x||, orthonormality, and the known structure of make_herm (eigenvalues near 1..8)."""
    n, neig = 16384, 8
    A = oracle.make_herm(n, neig, torch.float32).to(DEV)
    info = {}
    evals, evecs = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="davidson", min_eps=1e-4,
                                    info=info)
    assert info["converged"], info
    R = A.double() @ evecs.double() - evecs.double() * evals.double()
    assert R.abs().max().item() <= 5e-4
    G = evecs.double().t() @ evecs.double()
    assert (G - torch.eye(neig, dtype=torch.float64, device=DEV)).abs().max().item() <= 1e-4
    # Rayleigh quotients in fp64 agree with the returned eigenvalues to 1e-5 relative
    rq = (evecs.double() * (A.double() @ evecs.double())).sum(0)
    assert ((rq - evals.double()).abs() / rq.abs()).max().item() <= EIG_RTOL
    assert ((evals.cpu() - (1 + torch.arange(neig))).abs() < 0.05).all()


@pytest.mark.parametrize("env", ["XT_PO_NOSTAGE", "XT_NO_FUSE"])
def test_expansion_step_variants(env, monkeypatch):
    """the fused expansion kernel with the basis read from L2 (what large n uses) and the multi-kernel path give the
    same eigenpairs as the default (basis slice staged in shared memory)"""
    n, neig = 2048, 8
    A = oracle.make_herm(n, neig, torch.float32, seed=5).to(DEV)
    op = xt.LinearOperator.m(A, True)
    ev0, _ = xt.linalg.symeig(op, neig=neig, method="davidson", min_eps=1e-4)
    monkeypatch.setenv(env, "1")
    info = {}
    ev1, vec1 = xt.linalg.symeig(op, neig=neig, method="davidson", min_eps=1e-4, info=info)
    monkeypatch.delenv(env)
    assert info["converged"]
    assert ((ev1 - ev0).abs() / ev0.abs()).max().item() <= 1e-6
    R = A.double() @ vec1.double() - vec1.double() * ev1.double()
    assert R.abs().max().item() <= 2e-3


def test_lanczos_k16_large_n_fused_path():
    """k = 16 block at a size whose basis slice does not fit in shared memory (the C5 shape on one GPU, scaled down)"""
    n, neig = 24576, 16
    A = oracle.make_herm(n, neig, torch.float32, seed=9).to(DEV)
    info = {}
    ev, vec = xt.linalg.symeig(xt.LinearOperator.m(A, True), neig=neig, method="lanczos", min_eps=1e-4, info=info)
    assert info["converged"]
    ref = torch.arange(1, neig + 1, dtype=torch.float64)
    assert ((ev.double().cpu() - ref).abs() / ref).max().item() <= 2e-3          # make_herm: eigenvalues near 1..16
    R = A.double() @ vec.double() - vec.double() * ev.double()
    assert R.abs().max().item() <= 2e-3
    G = vec.double().t() @ vec.double()
    assert (G - torch.eye(neig, dtype=torch.float64, device=DEV)).abs().max().item() <= 1e-4
