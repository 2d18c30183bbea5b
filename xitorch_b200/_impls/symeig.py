"""
Method implementations behind `xitorch_b200.linalg.symeig` (plug-in layer L3 of SURVEY.md 1).

`davidson` keeps the reference signature and option names
(/root/reference/xitorch/_impls/linalg/symeig.py:100-109); `lanczos` is the new block-Lanczos method of
BASELINE.json (config 5) entering through the same `method=` plug-in point.  Both run on the B200 through
`xt_symeig_krylov` (csrc/symeig.cu) and need a CUDA operator -- there is NO CPU fallback.
`exacteig` is the dense `eigh` path (row f2 of SURVEY.md 8f).
"""
import ctypes as C
import warnings
from typing import Optional

import torch

from xitorch_b200 import _lib
from xitorch_b200._utils import ConvergenceWarning, MathWarning, bcast_dims
from xitorch_b200.debug import is_debug_enabled
from xitorch_b200.linop import LinearOperator, MatrixLinearOperator
from xitorch_b200._impls.solve import _dense_of, _mat3

_MATERIALISE_MAX_N = 2048       # composite / matrix-free operators up to this size are materialised by default; beyond,
                                # the engine's `apply` hook runs them matrix-free (verified on hardware in round 2)

__all__ = ["exacteig", "custom_exacteig", "davidson", "lanczos"]


def _take(evals, evecs, neig, mode):
    if mode == "lowest":
        return evals[..., :neig], evecs[..., :neig]
    return evals[..., -neig:], evecs[..., -neig:]


class _EighDegenerate(torch.autograd.Function):
    """`torch.linalg.eigh` whose backward stays finite at repeated eigenvalues.

    The eigenvector cotangent enters the textbook formula through the gaps 1/(lambda_j - lambda_i); inside a
    degenerate cluster that factor is dropped (its contribution vanishes whenever the loss does not depend on the
    choice of basis inside the cluster, which is the only case where the derivative exists -- Kasim 2020,
    arXiv:2011.04366).  Same contract as the reference's workaround (symeig.py:47-98): clusters are pairs closer than
    eps**0.6, debug mode warns when the basis-independence requirement is violated, the result is Hermitian.
    """

    @staticmethod
    def forward(ctx, A):
        lam, U = torch.linalg.eigh(A)
        ctx.save_for_backward(lam, U)
        return lam, U

    @staticmethod
    def backward(ctx, glam, gU):
        lam, U = ctx.saved_tensors
        Uh = U.transpose(-2, -1).conj()
        inner = None
        if gU is not None:
            gap = lam.unsqueeze(-2) - lam.unsqueeze(-1)                    # gap[i, j] = lam_j - lam_i
            clustered = gap.abs() <= torch.finfo(lam.dtype).eps ** 0.6     # includes the diagonal
            proj = torch.matmul(Uh, gU)
            if is_debug_enabled():
                skew = (proj - proj.transpose(-2, -1).conj())[clustered]
                if not torch.allclose(skew, torch.zeros_like(skew)):
                    warnings.warn(MathWarning(
                        "Degeneracy appears but the loss function seem to depend strongly on the eigenvector. "
                        "The gradient might be incorrect.\nEigenvalues:\n%s\nDegenerate map:\n%s\n"
                        "Requirements (should be all 0s):\n%s" % (lam, clustered, skew)))
            weight = torch.where(clustered, torch.zeros_like(gap), 1.0 / torch.where(clustered, torch.ones_like(gap), gap))
            inner = weight * proj
        if glam is not None:
            d = torch.diag_embed(glam).to(U.dtype)
            inner = d if inner is None else inner + d
        if inner is None:
            return torch.zeros_like(U)
        gA = torch.matmul(U, torch.matmul(inner, Uh))
        return 0.5 * (gA + gA.transpose(-2, -1).conj())


def exacteig(A: LinearOperator, neig: int, mode: str, M: Optional[LinearOperator]):
    """full dense eigendecomposition then truncation (reference symeig.py:11-44); the generalized
    problem is whitened with the Cholesky factor of M.  Differentiable also at degenerate spectra."""
    Amat = A.fullmatrix()
    if M is None:
        evals, evecs = _EighDegenerate.apply(Amat)
        return _take(evals, evecs, neig, mode)
    L = torch.linalg.cholesky(M.fullmatrix())
    Linv = torch.inverse(L)
    LinvT = Linv.transpose(-2, -1).conj()
    evals, q = _EighDegenerate.apply(torch.matmul(Linv, torch.matmul(Amat, LinvT)))
    evals, q = _take(evals, q, neig, mode)
    return evals, torch.matmul(LinvT, q)


def custom_exacteig(A, neig, mode, M=None, **options):
    return exacteig(A, neig, mode, M)


def _default_max_basis(n: int, neig: int) -> int:
    # up to 128 the projected matrix stays in the shared memory of the one-CTA eigensolver
    mb = min(max(16 * neig, 64), 128)
    mb = max(mb, 4 * neig)
    mb = min(mb, 512, n)
    return max((mb // neig) * neig, 2 * neig)


_V0_CACHE = {}


def _start_block(kind, nb, n, neig, dtype, dev):
    """seeded start block; it only depends on (kind, shape, dtype, device), so it is generated once and kept
    (the engine never writes to it)."""
    key = (kind, nb, n, neig, dtype, str(dev))
    v = _V0_CACHE.get(key)
    if v is None:
        gen = torch.Generator(device=dev)
        gen.manual_seed(12421)
        fn = torch.randn if kind == "randn" else torch.rand
        v = fn((nb, n, neig), dtype=dtype, device=dev, generator=gen)
        if len(_V0_CACHE) > 8:
            _V0_CACHE.clear()
        _V0_CACHE[key] = v
    return v


def _use_callback(A: LinearOperator, M: Optional[LinearOperator], matrix_free: Optional[bool]) -> bool:
    """matrix-free operators: dense ones never; others when asked for, or when they are too large to materialise
    (None = automatic).  Batched matrix-free operators are always materialised (the engine solves batch items one
    after another, an operator callback cannot be split per item)."""
    if isinstance(A, MatrixLinearOperator) or matrix_free is False:
        return False
    batched = any(s != 1 for s in A.shape[:-2]) or (M is not None and any(s != 1 for s in M.shape[:-2]))
    if batched:
        if matrix_free:
            raise RuntimeError("xitorch_b200: matrix_free=True needs an operator without batch dimensions")
        return False
    return bool(matrix_free) or A.shape[-1] > _MATERIALISE_MAX_N


def _start(kind: str, nb: int, n: int, neig: int, vdt, dev):
    # start block (the reference reseeds the GLOBAL RNG with 12421, symeig.py:236; a local generator with
    # the same seed is used here so callers' random streams are left alone)
    kind = kind.lower()
    if kind == "eye":
        return torch.eye(n, neig, dtype=vdt, device=dev).expand(nb, n, neig).contiguous()
    if kind in ("randn", "rand", "random"):
        return _start_block(kind, nb, n, neig, vdt, dev)
    raise ValueError("Unknown v_init type: %s" % kind)


def _space_exhausted(run: dict, n: int, neig: int, max_niter: int) -> bool:
    """the engine grows the subspace in whole blocks of `neig` vectors and stops when the next block no longer fits
    (m + neig > n).  The reference adds a PARTIAL block at that point, which makes the subspace the whole space and the
    next Rayleigh-Ritz step exact (symeig.py:204-206, 209-211).  True when that is what happened: not converged, not
    out of iterations, and no room left -- only possible for n below max_basis + neig (a few dozen rows)."""
    if run.get("converged", True) or run.get("niter", 0) >= max_niter:
        return False
    mb = min(int(run.get("max_basis", n)), (n // neig) * neig)
    return mb + neig > n


def _full_space_pairs(Amat: torch.Tensor, neig: int, mode: str, run: dict):
    """what the reference's last step computes once its basis spans everything: the eigenpairs of the projected
    matrix, which then IS the operator in another basis (a dense `eigh` of an n x n matrix with n <= ~100)"""
    w, S = torch.linalg.eigh(0.5 * (Amat + Amat.transpose(-2, -1)))
    w, S = _take(w, S, neig, mode)
    resid = (torch.matmul(Amat, S) - S * w.unsqueeze(-2)).abs().max().item()
    run.update(converged=True, best_resid=resid, niter=run.get("niter", 0) + 1, completed_full_space=True)
    return w.contiguous(), S.contiguous()


def _krylov_matrix_free(A, neig, mode, expansion, max_niter, v_init, min_eps, max_basis, check_every, info, name):
    """Krylov eigensolver on an operator that is only known through `A.mm` (user `_mv`, composite operators such as
    ``A^H A`` of `svd`, autograd Hessians): the whole iteration stays in the engine's kernels, each block application
    ``Y = A X`` comes back to Python through the `apply` hook of `xt_symeig_args` (one call per iteration, on the current
    stream), as the reference calls `A.mm` on whatever operator it gets (symeig.py:155,165).  Generalized problems go
    through `_davidson_host`."""
    n = A.shape[-1]
    probe = torch.empty(0, dtype=A.dtype, device=A.device)
    _lib.require_cuda(probe, "linalg.symeig(method=%r)" % name)
    if A.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("xitorch_b200.%s: matrix-free operators must be float32 or float64 (got %s)" % (name, A.dtype))
    vdt, dev = A.dtype, A.device
    op = A.mm
    if n < 2 * neig:
        # not even two blocks fit: the reference's first expansion already fills the whole space (symeig.py:209-211);
        # same shortcut as the dense path (a tiny operator, e.g. the A^H A of `svd` of a thin matrix)
        run = {} if info is None else info
        run.update(niter=0, converged=False, napply=1, max_basis=n)
        with torch.no_grad():
            full = op(torch.eye(n, dtype=vdt, device=dev))
        evals, evecs = _full_space_pairs(full.reshape(1, n, n), neig, mode, run)
        batch = tuple(A.shape[:-2])
        return evals.reshape(*batch, neig), evecs.reshape(*batch, n, neig)
    V0 = _start(v_init, 1, n, neig, vdt, dev)
    failure = []
    abort = C.c_int32(0)

    def make_apply(ws):
        base, nbytes = ws.data_ptr(), n * neig * V0.element_size()

        def _cb(user, xptr, yptr, stream):
            try:
                xv = ws[xptr - base: xptr - base + nbytes].view(vdt).view(n, neig)
                yv = ws[yptr - base: yptr - base + nbytes].view(vdt).view(n, neig)
                with torch.no_grad():
                    yv.copy_(op(xv).reshape(n, neig))
            except BaseException as exc:
                if not failure:
                    failure.append(exc)
                abort.value = 1          # the engine returns right after this callback (xt_symeig_args.abort)

        return _lib.APPLY_FN(_cb)

    run = {} if info is None else info
    try:
        evals, evecs = _call_engine(None, 0, 0, n, 1, neig, mode, expansion, V0, max_niter, max_basis, check_every,
                                    min_eps, run, name, make_apply=make_apply, abort=abort)
    except RuntimeError:
        if failure:
            raise failure[0]
        raise
    if failure:
        raise failure[0]
    if _space_exhausted(run, n, neig, max_niter):
        with torch.no_grad():
            full = op(torch.eye(n, dtype=vdt, device=dev))
        evals, evecs = _full_space_pairs(full.reshape(1, n, n), neig, mode, run)
    batch = tuple(A.shape[:-2])
    evals = evals.reshape(*batch, neig)
    evecs = evecs.reshape(*batch, n, neig)
    return evals, evecs


def _small_eigh_device(Tsym: torch.Tensor, nev: int, mode: str):
    """`nev` extreme eigenpairs (ascending) of a symmetric m x m device matrix with the engine's one-CTA eigensolver
    (`xt_small_eigh`, csrc/symeig.cu; replaces torch.linalg.eigh at reference symeig.py:174).  fp64 in and out."""
    m = Tsym.shape[-1]
    dev = Tsym.device
    T64 = Tsym.to(torch.float64).contiguous()
    w = torch.empty(nev, dtype=torch.float64, device=dev)
    S = torch.empty((m, nev), dtype=torch.float64, device=dev)
    scratch = torch.empty(m * (m | 1) + 16, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().xt_small_eigh(T64.data_ptr(), m, nev, 0 if mode == "lowest" else 1, w.data_ptr(),
                                            S.data_ptr(), scratch.data_ptr(), _lib.stream_ptr(dev)), "small_eigh")
    return w, S


def _m_orthonormalise(W, MW, rtol):
    """M-orthonormal basis of span(W), given ``MW = M W`` (``MW is W`` for the standard problem): with the Gram matrix
    ``W^T M W = U diag(d) U^T`` the result is ``W U d^-1/2`` (and the same transformation of MW).  Directions with
    ``d <= rtol * max(d)`` are dropped -- the Ritz residuals of converged pairs are numerically dependent, which is
    where a Cholesky factor of the Gram matrix (the reference's `tallqr`, _utils/tensor.py:8-19) breaks down.  The number
    of columns kept is the smallest numerical rank over the batch."""
    G = torch.matmul(W.transpose(-2, -1), MW)
    G = 0.5 * (G + G.transpose(-2, -1))
    dev = G.device
    if dev.type == "cuda" and G.numel() <= 65536:
        G = G.cpu()            # k x k with k <= 16: LAPACK on the host, one round trip instead of the device library's
                               # launch chain plus a second synchronisation for `keep`
    d, U = torch.linalg.eigh(G)                                         # ascending
    dmax = d[..., -1:].clamp_min(torch.finfo(d.dtype).tiny)
    keep = int((d > rtol * dmax).sum(-1).min().item())
    if keep == 0:
        return W[..., :0], MW[..., :0]
    U = U[..., -keep:] * d[..., -keep:].clamp_min(torch.finfo(d.dtype).tiny).rsqrt().unsqueeze(-2)
    U = U.to(dev)
    Wn = torch.matmul(W, U)
    return Wn, (Wn if MW is W else torch.matmul(MW, U))


def _davidson_host(A, neig, mode, M, max_niter, nguess, v_init, min_eps, max_basis, precond, info, name):
    """Davidson for the cases the one-kernel engine does not take: a generalized problem ``A x = lambda M x``, a start
    block wider than ``neig`` (`nguess`), a preconditioned expansion.  Follows the reference loop
    (/root/reference/xitorch/_impls/linalg/symeig.py:164-223): Rayleigh-Ritz on an M-orthonormal basis, residual
    ``A V s - lambda M V s``, expansion with the residuals -- but incrementally: the basis keeps ``V``, ``A V`` and
    ``M V``, so one iteration applies A once and M once to the NEW block only (the reference re-applies M to the whole
    basis and once more in the residual, symeig.py:184,212), and nothing of order n^3 is ever formed (no Cholesky
    whitening of M).  The two O(n^2 k) applications go through `A.mm` / `M.mm`, i.e. the block-matvec kernel for dense
    operators; the O(n m k) tall-skinny algebra is library calls on the same stream, the m x m eigenproblem goes to the
    engine's one-CTA eigensolver (`xt_small_eigh`; torch.linalg.eigh for batches and m > 256).  Problems up to n = 512
    never restart (as the reference); beyond, thick restart on ``max(2 * neig, nguess)`` Ritz vectors.

    precond: None | "diag" | callable(resid (*B, n, k), eigvals (*B, k)) -> (*B, n, k).  "diag" is Davidson's
    ``t = r / (diag(A) - lambda diag(M))`` for dense operators."""
    n = A.shape[-1]
    vdt, dev = A.dtype, A.device
    _lib.require_cuda(torch.empty(0, dtype=vdt, device=dev), "linalg.symeig(method=%r)" % name)
    if vdt not in (torch.float32, torch.float64):
        raise RuntimeError("xitorch_b200.%s: generalized / wide-start / preconditioned problems need float32 or "
                           "float64 operators (got %s)" % (name, vdt))
    batch = tuple(A.shape[:-2]) if M is None else tuple(bcast_dims(A.shape[:-2], M.shape[:-2]))
    nb = 1
    for s_ in batch:
        nb *= s_
    nguess = neig if nguess is None else int(nguess)
    if nguess < neig:
        raise RuntimeError("xitorch_b200.%s: nguess must be at least neig (got %d vs %d)" % (name, nguess, neig))
    nguess = min(nguess, n)
    if max_basis is None:
        # small problems never restart, like the reference (its basis grows until it is the whole space, symeig.py:203);
        # beyond, the engine's cap
        max_basis = n if n <= 512 else _default_max_basis(n, neig)
    nkeep = max(2 * neig, nguess)
    max_basis = min(max(int(max_basis), nkeep + neig), n)
    if max_basis + neig >= n:
        max_basis = n                  # small problems run to the full space, where the pairs are exact (symeig.py:203)
    nkeep = min(nkeep, max_basis)
    eps = torch.finfo(vdt).eps
    on_chip_eigh = dev.type == "cuda" and nb == 1 and max_basis <= 512 and max(nkeep, neig) <= 128

    dA = dM = None
    if precond == "diag":
        if not isinstance(A, MatrixLinearOperator) or (M is not None and not isinstance(M, MatrixLinearOperator)):
            raise RuntimeError("xitorch_b200.%s: precond='diag' needs dense operators" % name)
        dA = _dense_of(A, "A").diagonal(dim1=-2, dim2=-1).to(vdt).unsqueeze(-1)
        dM = None if M is None else _dense_of(M, "M").diagonal(dim1=-2, dim2=-1).to(vdt).unsqueeze(-1)
    elif precond is not None and not callable(precond):
        raise RuntimeError("Unknown precond: %s" % (precond,))

    def a_mm(X):
        return A.mm(X.contiguous())

    def m_mm(X):
        return X if M is None else M.mm(X.contiguous())

    with torch.no_grad():
        Vb = torch.empty((*batch, n, max_basis), dtype=vdt, device=dev)
        AVb = torch.empty_like(Vb)
        MVb = Vb if M is None else torch.empty_like(Vb)
        T = torch.zeros((*batch, max_basis, max_basis), dtype=vdt, device=dev)

        W = _start(v_init, nb, n, nguess, vdt, dev).reshape(*batch, n, nguess)
        MW = m_mm(W)
        for _ in range(2):
            W, MW = _m_orthonormalise(W, MW, 64 * eps)
        m = W.shape[-1]
        if m < neig:
            raise RuntimeError("xitorch_b200.%s: the start block has numerical rank %d < neig" % (name, m))
        AW = a_mm(W)
        napply, napply_m = 1, (0 if M is None else 1)
        Vb[..., :m], AVb[..., :m] = W, AW
        if M is not None:
            MVb[..., :m] = MW
        T[..., :m, :m] = torch.matmul(W.transpose(-2, -1), AW)

        best = (float("inf"), None, None, None, None)
        converged = False
        niter = 0
        for niter in range(1, max_niter + 1):
            V, AV, MV = Vb[..., :m], AVb[..., :m], MVb[..., :m]
            Tm = T[..., :m, :m]
            Tsym = 0.5 * (Tm + Tm.transpose(-2, -1))
            theta = S = None
            if on_chip_eigh and m <= 256:          # sizes the kernel's hardware tests cover
                # the engine's one-CTA eigensolver (`xt_small_eigh`: 0.2 ms at m = 88, no synchronisation) for the neig
                # wanted pairs; the device library's `eigh` is a chain of ~ms launches for a matrix this small
                lam, Sk = _small_eigh_device(Tsym.reshape(m, m), neig, mode)
                lam, Sk = lam.to(vdt).reshape(*batch, neig), Sk.to(vdt).reshape(*batch, m, neig)
            else:
                theta, S = torch.linalg.eigh(Tsym)
                lam, Sk = _take(theta, S, neig, mode)
            X = torch.matmul(V, Sk)
            AX = torch.matmul(AV, Sk)
            MX = X if M is None else torch.matmul(MV, Sk)
            R = AX - MX * lam.unsqueeze(-2)
            resid = R.abs().max().item()                   # the reference's stop test, global over the batch (:188,201)
            if resid < best[0]:
                best = (resid, lam, X, AX, MX)
            if resid < min_eps:
                converged = True
                break
            if m == n:                                     # the basis is the whole space: the pairs are exact (:203)
                converged = True
                break
            if niter == max_niter:
                break
            if m + 1 > max_basis:
                # thick restart: the basis becomes the nkeep extreme Ritz vectors (still M-orthonormal, T diagonal)
                if theta is None:
                    thk, Skp = _small_eigh_device(Tsym.reshape(m, m), nkeep, mode)
                    thk, Skp = thk.to(vdt).reshape(*batch, nkeep), Skp.to(vdt).reshape(*batch, m, nkeep)
                else:
                    thk, Skp = _take(theta, S, nkeep, mode)
                Vk, AVk = torch.matmul(V, Skp), torch.matmul(AV, Skp)
                MVk = None if M is None else torch.matmul(MV, Skp)
                m = nkeep
                Vb[..., :m], AVb[..., :m] = Vk, AVk
                if M is not None:
                    MVb[..., :m] = MVk
                T.zero_()
                T[..., :m, :m] = torch.diag_embed(thk)
                V, AV, MV = Vb[..., :m], AVb[..., :m], MVb[..., :m]
            # expansion block: the (preconditioned) residuals, at most what still fits
            if dA is not None:
                den = dA - lam.unsqueeze(-2) * (1.0 if dM is None else dM)
                floor = 1e-3 * dA.abs().amax(dim=(-2, -1), keepdim=True).clamp_min(torch.finfo(vdt).tiny)
                den = torch.where(den.abs() < floor, torch.where(den < 0, -floor, floor), den)
                W = R / den
            elif precond is not None:
                W = precond(R, lam)
            else:
                W = R
            nadd = min(neig, n - m, max_basis - m)
            W = W[..., :nadd]
            W = W / W.norm(dim=-2, keepdim=True).clamp_min(torch.finfo(vdt).tiny)
            MVt = MV.transpose(-2, -1)
            for _ in range(2):                             # block Gram-Schmidt against the basis in the M inner product
                W = W - torch.matmul(V, torch.matmul(MVt, W))
            MW = m_mm(W)
            napply_m += 0 if M is None else 1
            W, MW = _m_orthonormalise(W, MW, 64 * eps)
            if W.shape[-1] > 0:
                c = torch.matmul(MVt, W)                   # what normalising amplified
                W = W - torch.matmul(V, c)
                MW = W if M is None else MW - torch.matmul(MV, c)
                W, MW = _m_orthonormalise(W, MW, 64 * eps)
            k = W.shape[-1]
            if k == 0:
                break                                      # the residuals add nothing to the basis: stagnation
            AW = a_mm(W)
            napply += 1
            Vb[..., m:m + k], AVb[..., m:m + k] = W, AW
            if M is not None:
                MVb[..., m:m + k] = MW
            C1 = torch.matmul(V.transpose(-2, -1), AW)
            T[..., :m, m:m + k] = C1
            T[..., m:m + k, :m] = C1.transpose(-2, -1)
            T[..., m:m + k, m:m + k] = torch.matmul(W.transpose(-2, -1), AW)
            m += k
        resid, evals, evecs, AX, MX = best
        if vdt == torch.float32:
            # final Rayleigh-Ritz on the neig Ritz vectors with fp64 accumulation: removes what fp32 sums of length n
            # put into T and into the M-orthonormality of the basis (a few 1e-6 relative at n = 4096)
            Xd = evecs.double()
            AXd, MXd = AX.double(), (Xd if M is None else MX.double())
            Xt = Xd.transpose(-2, -1)
            Tk, Gk = torch.matmul(Xt, AXd), torch.matmul(Xt, MXd)
            Li = torch.inverse(torch.linalg.cholesky(0.5 * (Gk + Gk.transpose(-2, -1))))
            Tw = torch.matmul(Li, torch.matmul(0.5 * (Tk + Tk.transpose(-2, -1)), Li.transpose(-2, -1)))
            w, Q = torch.linalg.eigh(Tw)
            evals = w.to(vdt)
            evecs = torch.matmul(Xd, torch.matmul(Li.transpose(-2, -1), Q)).to(vdt)
    if info is not None:
        info.update(niter=niter, converged=converged, best_resid=resid, napply=napply, napply_M=napply_m,
                    max_basis=int(max_basis), engine="host-composed")
    return evals.contiguous(), evecs.contiguous()


def _krylov(A: LinearOperator, neig: int, mode: str, M: Optional[LinearOperator], expansion: int,
            max_niter: int, nguess: Optional[int], v_init: str, min_eps: float, max_basis: Optional[int],
            check_every: Optional[int], info: Optional[dict], name: str, matrix_free: Optional[bool] = None,
            precond=None):
    if mode not in ("lowest", "uppest"):
        raise RuntimeError("Unknown mode: %s" % mode)
    n = A.shape[-1]
    if M is not None or precond is not None or (nguess is not None and nguess != neig):
        # generalized problem, wide start block or preconditioned expansion: host-composed loop over the block-matvec
        # kernel (the one-kernel engine is the standard problem with a start block of neig vectors)
        return _davidson_host(A, neig, mode, M, max_niter, nguess, v_init, min_eps, max_basis, precond, info, name)
    if _use_callback(A, M, matrix_free):
        return _krylov_matrix_free(A, neig, mode, expansion, max_niter, v_init, min_eps, max_basis, check_every,
                                   info, name)
    Amat = _dense_of(A, "A")
    _lib.require_cuda(Amat, "linalg.symeig(method=%r)" % name)
    if Amat.is_complex():
        raise RuntimeError("xitorch_b200.%s: complex operators are not supported (the reference davidson is "
                           "real-only as well, SURVEY.md 3.6)" % name)
    if Amat.dtype == torch.bfloat16:
        raise RuntimeError("xitorch_b200.%s: bf16 operators are not supported for eigenproblems" % name)
    dev = Amat.device
    batch = tuple(Amat.shape[:-2])
    nb = 1
    for s in batch:
        nb *= s
    vdt = Amat.dtype

    V0 = _start(v_init, nb, n, neig, vdt, dev)

    A3, a_bs, lda = _mat3(Amat, batch)
    run = {} if info is None else info
    if n < 2 * neig:
        # not even two blocks fit: the reference's first expansion already fills the whole space (symeig.py:209-211)
        run.update(niter=0, converged=False, napply=0, max_basis=n)
        evals, evecs = _full_space_pairs(Amat.reshape(nb, n, n).to(vdt), neig, mode, run)
        return evals.reshape(*batch, neig), evecs.reshape(*batch, n, neig)
    evals, evecs = _call_engine(A3, lda, (a_bs if nb > 1 else 0), n, nb, neig, mode, expansion, V0, max_niter,
                                max_basis, check_every, min_eps, run, name, out_batch=batch)
    if _space_exhausted(run, n, neig, max_niter):
        evals, evecs = _full_space_pairs(Amat.reshape(nb, n, n), neig, mode, run)
        evals = evals.reshape(*batch, neig)
        evecs = evecs.reshape(*batch, n, neig)
    return evals, evecs


_ENGINE_CACHE = {}          # prepared argument blocks of the plain dense path, see _call_engine
_ENGINE_CACHE_MAX = 4
_WS_CACHE_MAX_BYTES = 64 << 20


def _call_engine(A3, lda, a_bs, n, nb, neig, mode, expansion, V0, max_niter, max_basis, check_every, min_eps,
                 info, name, dist_ctx=None, make_apply=None, out_batch=None, abort=None):
    """fill `xt_symeig_args` and run `xt_symeig_krylov`.  dist_ctx = (world, rank, group) for the row-partitioned
    operator: A3 is then this rank's (n/world, n) row block and the per-iteration all-gather hook is installed.
    make_apply(workspace) -> APPLY_FN for a matrix-free operator (A3 is then None).
    out_batch: batch dimensions of the results (None: the flat (nb, ...) shapes).

    Host time matters for small and medium problems (a C2 solve is 2.6 ms): on the plain dense path the filled
    argument block, its output cells and the workspace tensor are kept per (operator storage, problem, stream) and only
    the result pointers change from call to call."""
    vdt, dev = V0.dtype, V0.device
    if max_basis is None:
        max_basis = _default_max_basis(n, neig)
    eshape = (nb, neig) if out_batch is None else (*out_batch, neig)
    vshape = (nb, n, neig) if out_batch is None else (*out_batch, n, neig)
    evals = torch.empty(eshape, dtype=vdt, device=dev)
    evecs = torch.empty(vshape, dtype=vdt, device=dev)
    L_ = _lib.lib()
    plain = dist_ctx is None and make_apply is None and A3 is not None
    stream = _lib.stream_ptr(dev)
    key = None
    if plain:
        key = (A3.data_ptr(), lda, a_bs, n, nb, neig, mode, expansion, V0.data_ptr(), int(max_niter), int(max_basis),
               float(min_eps), vdt, dev, stream, check_every)
        ent = _ENGINE_CACHE.get(key)
        if ent is not None:
            g, niter, conv, best, napply, ws, _keep = ent
            g.evals, g.evecs = evals.data_ptr(), evecs.data_ptr()
            if ws is None:
                ws = torch.empty(g.workspace_bytes, dtype=torch.uint8, device=dev)
                g.workspace = ws.data_ptr()
            with torch.cuda.device(dev):
                _lib.check(L_.xt_symeig_krylov(g), name)
            if info is not None:
                info.update(niter=niter.value, converged=bool(conv.value), best_resid=best.value, napply=napply.value,
                            max_basis=int(max_basis))
            return evals, evecs
    g = _lib.SymeigArgs()
    g.dtype = _lib.dtype_code(vdt)
    g.n, g.nbatch, g.neig = n, nb, neig
    g.mode = 0 if mode == "lowest" else 1
    g.expansion = expansion
    g.A, g.lda, g.a_bstride = (A3.data_ptr() if A3 is not None else None), lda, a_bs
    g.V0, g.ldv0, g.v0_bstride = V0.data_ptr(), neig, n * neig
    g.evals, g.evals_bstride = evals.data_ptr(), neig
    g.evecs, g.ldv, g.evecs_bstride = evecs.data_ptr(), neig, n * neig
    g.max_niter, g.max_basis = int(max_niter), int(max_basis)
    world = dist_ctx[0] if dist_ctx is not None else 1
    ce = check_every
    if ce is None:
        t_iter = max((n // world) * n * V0.element_size() / 6.0e12, 3e-5)
        ce = max(4, min(16, int(1e-3 / t_iter)))
    g.check_every = int(ce)
    g.min_eps = float(min_eps)
    niter, conv, best, napply = C.c_int32(0), C.c_int32(0), C.c_double(0.0), C.c_int64(0)
    g.niter_out, g.converged_out = C.pointer(niter), C.pointer(conv)
    g.best_resid_out, g.napply_out = C.pointer(best), C.pointer(napply)
    wsb = L_.xt_symeig_workspace_bytes(g.dtype, n, neig, g.max_basis, world)
    if wsb == 0:
        raise RuntimeError("xitorch_b200.%s: n=%d is too small for neig=%d (need n >= 2*neig)" % (name, n, neig))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    g.workspace, g.workspace_bytes = ws.data_ptr(), wsb
    g.stream = stream
    keep = [ws]
    if dist_ctx is not None and world > 1:
        import torch.distributed as dist
        _, rank, group = dist_ctx
        base = ws.data_ptr()

        def _gather(user, buf, count, esize, stream):
            # in-place all-gather of `world` chunks of `count` elements inside the workspace tensor
            off = buf - base
            whole = ws[off: off + world * count * esize].view(vdt)
            mine = whole[rank * count: (rank + 1) * count]
            dist.all_gather_into_tensor(whole, mine, group=group)

        cb = _lib.ALLGATHER_FN(_gather)
        keep.append(cb)
        g.world, g.rank = world, rank
        g.allgather = _lib.fn_address(cb)
    if make_apply is not None:
        acb = make_apply(ws)
        keep.append(acb)
        g.apply = _lib.fn_address(acb)
        if abort is not None:
            g.abort = C.pointer(abort)
    with torch.cuda.device(dev):
        _lib.check(L_.xt_symeig_krylov(g), name)
    if info is not None:
        info.update(niter=niter.value, converged=bool(conv.value), best_resid=best.value, napply=napply.value,
                    max_basis=int(max_basis))
    if plain:
        if len(_ENGINE_CACHE) >= _ENGINE_CACHE_MAX:
            _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE)))
        # the block only holds ADDRESSES: the operator's storage is not kept alive (a new tensor at the same address with
        # the same geometry is the same call as far as the engine is concerned); the small start block is, because its
        # address is part of the key; a large workspace is not held between calls
        _ENGINE_CACHE[key] = (g, niter, conv, best, napply, ws if wsb <= _WS_CACHE_MAX_BYTES else None, (V0,))
    del keep
    return evals, evecs


def _krylov_row_partitioned(A_local, n, neig, mode, expansion, group, min_eps, max_niter, max_basis, check_every,
                            info):
    """row-partitioned operator: every rank holds rows [rank*n/world, (rank+1)*n/world)."""
    import torch.distributed as dist
    _lib.require_cuda(A_local, "symeig on a row-partitioned operator")
    if A_local.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("row-partitioned symeig supports float32 / float64 operators")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if tuple(A_local.shape) != (n // world, n) or n % world != 0:
        raise RuntimeError("expected a local row block of shape %s, got %s" % ((n // world, n), tuple(A_local.shape)))
    A_local = A_local.contiguous()
    V0 = _start_block("randn", 1, n, neig, A_local.dtype, A_local.device)      # same start block on every rank
    evals, evecs = _call_engine(A_local, A_local.stride(0), 0, n, 1, neig, mode, expansion, V0, max_niter, max_basis,
                                check_every, min_eps, info, "lanczos" if expansion == 1 else "davidson",
                                dist_ctx=(world, rank, group))
    return evals[0], evecs[0]


def _sharded_applies(A_local, n, neig, max_basis, group) -> bool:
    """the row-sharded engine takes fp32 / fp64 CUDA blocks, n divisible by the world size and at least twice the basis cap"""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if not A_local.is_cuda or A_local.dtype not in (torch.float32, torch.float64) or world > 8 or n % world != 0:
        return False
    mb = _default_max_basis(n, neig) if max_basis is None else max_basis
    return _lib.lib().xt_symeig_peer_bytes(_lib.dtype_code(A_local.dtype), n, neig, int(mb), world) > 0


def _krylov_row_sharded(A_local, n, neig, mode, group, min_eps, max_niter, max_basis, info, regions, restart_keep,
                        gather=True):
    """row-sharded block Lanczos (BASELINE config 5): `xt_symeig_krylov` with `peers` set."""
    import torch.distributed as dist
    _lib.require_cuda(A_local, "symeig on a row-partitioned operator")
    if A_local.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("row-partitioned symeig supports float32 / float64 operators")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if n % world != 0 or tuple(A_local.shape) != (n // world, n):
        raise RuntimeError("expected a local row block of shape %s, got %s" % ((n // world, n), tuple(A_local.shape)))
    if mode not in ("lowest", "uppest"):
        raise RuntimeError("Unknown mode: %s" % mode)
    A_local = A_local.contiguous()
    vdt, dev = A_local.dtype, A_local.device
    n_loc = n // world
    if max_basis is None:
        max_basis = _default_max_basis(n, neig)
    L_ = _lib.lib()
    dcode = _lib.dtype_code(vdt)
    pbytes = L_.xt_symeig_peer_bytes(dcode, n, neig, int(max_basis), world)
    wsb = L_.xt_symeig_sharded_workspace_bytes(dcode, n, neig, int(max_basis), world)
    if pbytes == 0 or wsb == 0:
        raise RuntimeError("xitorch_b200.lanczos (sharded): unsupported sizes n=%d neig=%d max_basis=%d world=%d "
                           "(need n %% world == 0, max_basis >= 4*neig, n >= 2*max_basis, world <= 8)"
                           % (n, neig, max_basis, world))
    if regions is None:
        from xitorch_b200.dist import _regions_for
        regions = _regions_for(pbytes, group, signature=(str(vdt), n, neig, int(max_basis)))
    if regions.nbytes < pbytes or regions.world != world:
        raise RuntimeError("exchange regions too small or made for another world size")
    regions.epoch += 1
    V0 = _start_block("randn", 1, n, neig, vdt, dev)                 # same start block on every rank
    evals = torch.empty((neig,), dtype=vdt, device=dev)
    evecs_loc = torch.empty((n_loc, neig), dtype=vdt, device=dev)
    g = _lib.SymeigArgs()
    g.dtype = dcode
    g.n, g.nbatch, g.neig = n, 1, neig
    g.mode = 0 if mode == "lowest" else 1
    g.expansion = 1
    g.A, g.lda, g.a_bstride = A_local.data_ptr(), A_local.stride(0), 0
    g.V0, g.ldv0, g.v0_bstride = V0.data_ptr(), neig, n * neig
    g.evals, g.evals_bstride = evals.data_ptr(), neig
    g.evecs, g.ldv, g.evecs_bstride = evecs_loc.data_ptr(), neig, n_loc * neig
    g.max_niter, g.max_basis, g.check_every = int(max_niter), int(max_basis), 1
    g.min_eps = float(min_eps)
    niter, conv, best, napply = C.c_int32(0), C.c_int32(0), C.c_double(0.0), C.c_int64(0)
    g.niter_out, g.converged_out = C.pointer(niter), C.pointer(conv)
    g.best_resid_out, g.napply_out = C.pointer(best), C.pointer(napply)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    g.workspace, g.workspace_bytes = ws.data_ptr(), wsb
    g.stream = _lib.stream_ptr(dev)
    g.world, g.rank = world, rank
    parr = (C.c_void_p * world)(*regions.ptrs)
    g.peers = C.cast(parr, C.POINTER(C.c_void_p))
    g.epoch = regions.epoch
    g.restart_keep = int(restart_keep or 0)
    with torch.cuda.device(dev):
        _lib.check(L_.xt_symeig_krylov(g), "lanczos (sharded)")
    if info is not None:
        info.update(niter=niter.value, converged=bool(conv.value), best_resid=best.value, napply=napply.value,
                    max_basis=int(max_basis), engine="sharded")
    del parr
    if not gather or world == 1:
        return evals, evecs_loc
    evecs = torch.empty((n, neig), dtype=vdt, device=dev)
    dist.all_gather_into_tensor(evecs, evecs_loc, group=group)      # once per solve, outside the iteration
    return evals, evecs


def davidson(A: LinearOperator, neig: int, mode: str, M: Optional[LinearOperator] = None,
             max_niter: int = 1000, nguess: Optional[int] = None, v_init: str = "randn",
             max_addition: Optional[int] = None, min_eps: float = 1e-6, verbose: bool = False,
             max_basis: Optional[int] = None, check_every: Optional[int] = None,
             info: Optional[dict] = None, expansion: str = "krylov", matrix_free: Optional[bool] = None,
             precond=None, **unused):
    """
    Block Davidson (Rayleigh-Ritz on span{V0, r0, r1, ...}) on the B200.

    Arguments
    ---------
    max_niter: int
        Maximum number of iterations (subspace expansions)
    nguess: int or None
        Size of the start block (None: ``neig``).  A block wider than ``neig`` runs the host-composed loop
        (`_davidson_host`), as do generalized problems (``M``) and ``precond``
    v_init: str
        Mode of the initial guess (``"randn"``, ``"rand"``, ``"eye"``)
    max_addition: int or None
        Accepted for compatibility and unused, exactly as in the reference (always ``neig``)
    min_eps: float
        Stop when the largest entry of the residual ``|A X - X E|`` is below this value
    verbose: bool
        Ignored (convergence control lives on the device)
    max_basis: int or None
        Thick-restart cap on the subspace dimension (None: ``max(16 * neig, 64)``, at most 512 and n)
    check_every: int or None
        Host polls the device convergence flag every this many iterations
    info: dict or None
        If given, receives ``niter``, ``converged``, ``best_resid``, ``napply``, ``max_basis``
    expansion: str
        ``"krylov"`` (default): the subspace is expanded with the orthonormalised ``A @ (last block)``.  Without a
        preconditioner the reference's expansion block ``-R`` (the Ritz residuals) spans exactly the same space
        (``R = Q_next @ (k x k)``), so the Ritz pairs and iteration counts are the same, but the expansion no
        longer waits for the Rayleigh-Ritz step, which then overlaps with the next matvec on the GPU.
        ``"residual"``: append the orthonormalised Ritz residuals literally as the reference does
        (symeig.py:207-220); Rayleigh-Ritz is then on the critical path of every iteration.
    matrix_free: bool or None
        How an operator that is not a dense matrix (user ``_mv``, ``A.H.matmul(A)``, Jacobians ...) is applied.
        ``True``: through ``A.mm`` once per iteration, nothing is materialised.  ``False``: ``A.fullmatrix()`` is built
        once and the dense kernels are used.  ``None``: materialise up to n = 2048, matrix-free beyond.
    precond: None, "diag" or callable
        Preconditioner of the expansion block (the reference has the hook but no preconditioner, symeig.py:206-207).
        ``"diag"``: Davidson's ``r / (diag(A) - lambda diag(M))`` for dense operators; a callable gets
        ``(resid, eigvals)`` and returns the block to add.
    """
    if expansion not in ("krylov", "residual"):
        raise RuntimeError("Unknown expansion: %s" % expansion)
    return _krylov(A, neig, mode, M, 1 if expansion == "krylov" else 0, max_niter, nguess, v_init, min_eps,
                   max_basis, check_every, info, "davidson", matrix_free, precond)


def lanczos(A: LinearOperator, neig: int, mode: str, M: Optional[LinearOperator] = None,
            max_niter: int = 1000, nguess: Optional[int] = None, v_init: str = "randn",
            min_eps: float = 1e-6, verbose: bool = False, max_basis: Optional[int] = None,
            check_every: Optional[int] = None, info: Optional[dict] = None,
            matrix_free: Optional[bool] = None, **unused):
    """
    Block Lanczos with full reorthogonalisation and Rayleigh-Ritz extraction on the B200: the same
    Krylov space as ``davidson`` (which has no preconditioner), expanded with the orthonormalised
    ``A @ (last block)`` instead of the Ritz residuals.

    Arguments
    ---------
    max_niter: int
        Maximum number of iterations (subspace expansions)
    nguess: int or None
        Size of the start block (None: ``neig``); as in :func:`davidson`.  Generalized problems (``M``) and wide start
        blocks are expanded with the Ritz residuals (there is no ``M^-1`` to build a Krylov space of ``M^-1 A`` with)
    v_init: str
        Mode of the initial guess (``"randn"``, ``"rand"``, ``"eye"``)
    min_eps: float
        Stop when the largest entry of the residual ``|A X - X E|`` is below this value
    verbose: bool
        Ignored
    max_basis, check_every, info, matrix_free:
        As in :func:`davidson`
    """
    return _krylov(A, neig, mode, M, 1, max_niter, nguess, v_init, min_eps, max_basis, check_every, info,
                   "lanczos", matrix_free)
