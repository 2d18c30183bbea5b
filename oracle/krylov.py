"""
CPU oracle for the xitorch Krylov hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function restates the algorithm of the cited reference lines with plain
torch-CPU calls (the reference's own arithmetic backend is ATen: torch.matmul,
torch.linalg.eigh/cholesky/inverse/lstsq, einsum, cat).  Operators are dense
tensors wrapped in `DenseOp`; there is no LinearOperator plumbing here.

Pinned against the reference by tests/golden/*.pt (see oracle/gen_golden.py).
"""
import warnings
from typing import Optional, Tuple

import torch


class OracleConvergenceWarning(UserWarning):
    pass


class DenseOp:
    """Dense operator, the restatement of MatrixLinearOperator
    (/root/reference/xitorch/_core/linop.py:676-708): mm = mat @ x, rmm = mat^H @ x,
    Hermitian operators reuse mm for rmm (linop.py:326-327)."""

    def __init__(self, mat: torch.Tensor, is_hermitian: Optional[bool] = None):
        self.mat = mat
        if is_hermitian is None:
            is_hermitian = bool(torch.allclose(mat, mat.transpose(-2, -1).conj()))
        self.is_hermitian = is_hermitian
        self.shape = mat.shape
        self.dtype = mat.dtype
        self.device = mat.device
        self.napply = 0  # number of operator applications (for bench accounting)

    def mm(self, x: torch.Tensor) -> torch.Tensor:
        self.napply += 1
        return torch.matmul(self.mat, x)

    def rmm(self, x: torch.Tensor) -> torch.Tensor:
        if self.is_hermitian:
            return self.mm(x)
        self.napply += 1
        return torch.matmul(self.mat.transpose(-2, -1).conj(), x)


def _as_op(A) -> Optional[DenseOp]:
    if A is None or isinstance(A, DenseOp):
        return A
    return DenseOp(A)


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _bcast_dims(*shapes):
    """broadcast of batch shapes (/root/reference/xitorch/_utils/bcast.py:4-10)."""
    return list(torch.broadcast_shapes(*[tuple(s) for s in shapes]))


def tallqr(V: torch.Tensor, MV: Optional[torch.Tensor] = None):
    """Cholesky-QR of a tall matrix: G = V^T (M) V, R = chol(G^H)^H, Q = V R^-1
    (/root/reference/xitorch/_utils/tensor.py:8-19)."""
    if MV is None:
        MV = V
    gram = torch.matmul(V.transpose(-2, -1), MV)
    R = torch.linalg.cholesky(gram.transpose(-2, -1).conj()).transpose(-2, -1).conj()
    Q = torch.matmul(V, torch.inverse(R))
    return Q, R


def to_fortran_order(V: torch.Tensor) -> torch.Tensor:
    """make the last two dims column-major (/root/reference/xitorch/_utils/tensor.py:21-32)."""
    if V.is_contiguous():
        return V.transpose(-2, -1).contiguous().transpose(-2, -1)
    if V.transpose(-2, -1).is_contiguous():
        return V
    raise RuntimeError("Only the last two dimensions can be made Fortran order.")


def _dot(r, z):
    """column-wise <r, z> keeping a singleton row dim (/root/reference/xitorch/_impls/linalg/solve.py:441-445)."""
    return torch.einsum("...rc,...rc->...c", r.conj(), z).unsqueeze(-2)


def _safedenom(r: torch.Tensor, eps: float) -> torch.Tensor:
    """replace exact zeros IN PLACE (/root/reference/xitorch/_impls/linalg/solve.py:437-439)."""
    r[r == 0] = eps
    return r


# --------------------------------------------------------------------------
# symeig: davidson
# --------------------------------------------------------------------------
def _initial_v(kind, dtype, device, batch, n, nguess, M):
    """/root/reference/xitorch/_impls/linalg/symeig.py:229-253 (reseeds the GLOBAL RNG)."""
    torch.manual_seed(12421)
    if kind == "eye":
        nb = 1
        for b in batch:
            nb *= b
        V = torch.eye(n, nguess, dtype=dtype, device=device).unsqueeze(0).repeat(nb, 1, 1)
        V = V.reshape(*batch, n, nguess)
    elif kind == "randn":
        V = torch.randn((*batch, n, nguess), dtype=dtype, device=device)
    elif kind in ("rand", "random"):
        V = torch.rand((*batch, n, nguess), dtype=dtype, device=device)
    else:
        raise ValueError("Unknown v_init type: %s" % kind)
    if M is not None:
        V, _ = tallqr(V, MV=M.mm(V))
    else:
        V, _ = tallqr(V)
    return V


def davidson(A, neig: int, mode: str = "lowest", M=None, max_niter: int = 1000,
             nguess: Optional[int] = None, v_init: str = "randn",
             min_eps: float = 1e-6, return_info: bool = False, **unused):
    """Unpreconditioned, unrestarted block Davidson
    (/root/reference/xitorch/_impls/linalg/symeig.py:100-227).

    Per iteration: T = V^T AV (:170), eigh(T) (:174), keep lowest/uppest neig (:175,255-264),
    X = V S (:178), R = AV S - (M) X Lambda (:181-185), stop on max|R| < min_eps (:188,200),
    best-so-far bookkeeping (:196-199), exit when the basis is square (:202-203),
    append -R, re-orthonormalise the whole basis with tallqr (:207-220),
    AV <- [AV, A Vnew] (:221-223).
    """
    A = _as_op(A)
    M = _as_op(M)
    if nguess is None:
        nguess = neig
    n = A.shape[-1]
    batch = list(A.shape[:-2]) if M is None else _bcast_dims(A.shape[:-2], M.shape[:-2])
    V = _initial_v(v_init.lower(), A.dtype, A.device, batch, n, nguess, M)

    best_resid = float("inf")
    best_vals = best_vecs = None
    AV = A.mm(V)
    niter = 0
    for it in range(max_niter):
        niter = it + 1
        T = torch.matmul(V.transpose(-2, -1), AV)
        tvals, tvecs = torch.linalg.eigh(T)
        if mode == "lowest":
            tvals, tvecs = tvals[..., :neig], tvecs[..., :neig]
        else:
            tvals, tvecs = tvals[..., -neig:], tvecs[..., -neig:]
        X = torch.matmul(V, tvecs)
        AX = torch.matmul(AV, tvecs)
        LX = tvals.unsqueeze(-2) * X
        if M is not None:
            LX = M.mm(LX)
        resid = AX - LX
        max_resid = resid.abs().max()
        if max_resid < best_resid:
            best_resid, best_vals, best_vecs = max_resid, tvals, X
        if max_resid < min_eps:
            break
        if AV.shape[-1] == AV.shape[-2]:
            break
        t = to_fortran_order(-resid)
        Vnew = torch.cat((V, t), dim=-1)
        if Vnew.shape[-1] > Vnew.shape[-2]:
            Vnew = Vnew[..., :Vnew.shape[-2]]
        nadd = Vnew.shape[-1] - V.shape[-1]
        if M is not None:
            V, _ = tallqr(Vnew, MV=M.mm(Vnew))
        else:
            V, _ = tallqr(Vnew)
        AVnew = to_fortran_order(A.mm(V[..., -nadd:]))
        AV = torch.cat((AV, AVnew), dim=-1)
    if return_info:
        return best_vals, best_vecs, {"niter": niter, "best_resid": float(best_resid),
                                      "napply": A.napply}
    return best_vals, best_vecs


def exacteig(A, neig: int, mode: str = "lowest", M=None):
    """full eigh then truncate (/root/reference/xitorch/_impls/linalg/symeig.py:11-44),
    generalized problem via Cholesky of M."""
    Amat = A.mat if isinstance(A, DenseOp) else A
    if M is None:
        vals, vecs = torch.linalg.eigh(Amat)
    else:
        Mmat = M.mat if isinstance(M, DenseOp) else M
        L = torch.linalg.cholesky(Mmat)
        Linv = torch.inverse(L)
        LinvT = Linv.transpose(-2, -1).conj()
        vals, q = torch.linalg.eigh(torch.matmul(Linv, torch.matmul(Amat, LinvT)))
        vecs = torch.matmul(LinvT, q)
    if mode == "lowest":
        return vals[..., :neig], vecs[..., :neig]
    return vals[..., -neig:], vecs[..., -neig:]


# --------------------------------------------------------------------------
# solve: problem setup
# --------------------------------------------------------------------------
def _batchdims(A, B, E, M):
    """/root/reference/xitorch/_impls/linalg/solve.py:540-549"""
    dims = [A.shape[:-2], B.shape[:-2]]
    if E is not None:
        dims.append(E.shape[:-1])
        if M is not None:
            dims.append(M.shape[:-2])
    return _bcast_dims(*dims)


def _normalize_bcast(*shapes):
    """left-pad batch shapes with 1s to equal rank (/root/reference/xitorch/_utils/bcast.py:12-18)."""
    nd = max(len(s) for s in shapes)
    return [[1] * (nd - len(s)) + list(s) for s in shapes]


def largest_eival_power(Afcn, x):
    """<=10-step power iteration returning the norm of the last iterate
    (/root/reference/xitorch/_impls/linalg/solve.py:645-663)."""
    niter, rtol, atol = 10, 1e-3, 1e-6
    prev = None
    xnorm = None
    for i in range(niter):
        x = Afcn(x)
        xnorm = x.norm(dim=-2, keepdim=True)
        if i > 0:
            if torch.all(torch.abs(prev - xnorm) <= rtol * xnorm + atol):
                break
        prev = xnorm
        if i < niter - 1:
            x = x / xnorm
    return xnorm


def setup_linear_problem(A: DenseOp, B, E, M: Optional[DenseOp], batchdims, posdef, need_hermit):
    """Builds A_fcn / AT_fcn / B2 / col_swapped
    (/root/reference/xitorch/_impls/linalg/solve.py:560-643):
    plain operator when E is None; otherwise the column-swapped layout x:(ncols,*B,nr,1)
    with x -> A x - (M x) E (:585-603); posdef probe by power iteration with an
    (unseeded) randn start (:617-634); normal equations when not posdef (:637-643)."""
    if E is None:
        A_fcn = lambda x: A.mm(x)          # noqa: E731
        AT_fcn = lambda x: A.rmm(x)        # noqa: E731
        B_new = B
        swapped = False
    else:
        if M is None:
            BAs, BBs, BEs = _normalize_bcast(A.shape[:-2], B.shape[:-2], E.shape[:-1])
        else:
            BAs, BBs, BEs, BMs = _normalize_bcast(A.shape[:-2], B.shape[:-2], E.shape[:-1], M.shape[:-2])
        E = E.reshape(*BEs, *E.shape[-1:])
        E_new = E.unsqueeze(0).transpose(-1, 0).unsqueeze(-1)     # (ncols, *BE, 1, 1)
        B = B.reshape(*BBs, *B.shape[-2:])
        B_new = B.unsqueeze(0).transpose(-1, 0)                    # (ncols, *BB, nr, 1)

        def A_fcn(x):
            Mx = M.mm(x) if M is not None else x
            return A.mm(x) - Mx * E_new

        def AT_fcn(x):
            MTx = M.rmm(x) if M is not None else x
            return A.rmm(x) - MTx * E_new

        swapped = True

    if need_hermit:
        if not (A.is_hermitian and (M is None or M.is_hermitian)):
            posdef = False

    if posdef is None:
        nr, ncols = B.shape[-2:]
        x0shape = (ncols, *batchdims, nr, 1) if swapped else (*batchdims, nr, ncols)
        x0 = torch.randn(x0shape, dtype=A.dtype, device=A.device)
        x0 = x0 / x0.norm(dim=-2, keepdim=True)
        big = largest_eival_power(A_fcn, x0)
        neg = big <= 0
        if torch.all(neg):
            posdef = False
        else:
            offset = torch.clamp(big, min=0.0)
            shifted = lambda x: A_fcn(x) - offset * x   # noqa: E731
            mostneg = largest_eival_power(shifted, x0)
            posdef = bool(torch.all(torch.logical_or(-mostneg <= offset, neg)).item())

    if posdef:
        return A_fcn, AT_fcn, B_new, swapped
    normal = lambda x: AT_fcn(A_fcn(x))   # noqa: E731
    return normal, normal, AT_fcn(B_new), swapped


def _unswap(x, swapped):
    return x.transpose(0, -1).squeeze(0) if swapped else x


# --------------------------------------------------------------------------
# solve: cg / bicgstab / gmres
# --------------------------------------------------------------------------
def cg(A, B, E=None, M=None, posdef=None, precond=None, max_niter=None, rtol=1e-6, atol=1e-8, eps=1e-12,
       resid_calc_every=10, return_info=False, **unused):
    """(Preconditioned) conjugate gradient, all columns/batches in lock-step with a GLOBAL
    stop test and best-iterate bookkeeping (/root/reference/xitorch/_impls/linalg/solve.py:69-190;
    `precond`: z = precond.mm(r), :122,136,171)."""
    A, M = _as_op(A), _as_op(M)
    precond = _as_op(precond)
    pfcn = (lambda x: x) if precond is None else precond.mm
    nr, ncols = A.shape[-1], B.shape[-1]
    if max_niter is None:
        max_niter = int(1.5 * nr)
    batchdims = _batchdims(A, B, E, M)
    if torch.allclose(B, B * 0, rtol=rtol, atol=atol):                       # :116-119
        x0 = torch.zeros((*batchdims, nr, ncols), dtype=A.dtype, device=A.device)
        return (x0, {"niter": 0, "converged": True, "napply": A.napply}) if return_info else x0
    A_fcn, _, B2, swapped = setup_linear_problem(A, B, E, M, batchdims, posdef, True)
    stop = torch.max(rtol * B2.norm(dim=-2, keepdim=True),
                     atol * torch.ones_like(B2.norm(dim=-2, keepdim=True)))  # :128-129
    shape = (ncols, *batchdims, nr, 1) if swapped else (*batchdims, nr, ncols)
    xk = torch.zeros(shape, dtype=A.dtype, device=A.device)
    rk = B2 - A_fcn(xk)                                                      # :135 (spends a matvec)
    zk = pfcn(rk)
    pk = zk
    rkzk = _dot(rk, zk)
    converged = False
    best_resid = rk.norm(dim=-2).max().item()
    best_x = xk
    k = 0
    for k in range(1, max_niter + 1):
        Apk = A_fcn(pk)
        alpha = rkzk / _safedenom(_dot(pk, Apk), eps)
        xk1 = xk + alpha * pk
        if resid_calc_every != 0 and k % resid_calc_every == 0:
            rk1 = B2 - A_fcn(xk1)                                            # :148-149
        else:
            rk1 = rk - alpha * Apk
        rnorm = rk1.norm(dim=-2, keepdim=True)
        mx = rnorm.max().item()
        if mx < best_resid:
            best_resid, best_x = mx, xk1
        if torch.all(rnorm < stop):
            converged = True
            break
        zk1 = pfcn(rk1)
        rkzk1 = _dot(rk1, zk1)
        beta = rkzk1 / _safedenom(rkzk, eps)
        pk = zk1 + beta * pk
        xk, rk, rkzk = xk1, rk1, rkzk1
    if not converged:
        warnings.warn(OracleConvergenceWarning(
            "Convergence is not achieved after %d iterations. Max norm of best resid: %.3e"
            % (max_niter, best_resid)))
    x = _unswap(best_x, swapped)
    if return_info:
        return x, {"niter": k, "converged": converged, "best_resid": best_resid, "napply": A.napply}
    return x


def bicgstab(A, B, E=None, M=None, posdef=None, precond_l=None, precond_r=None, max_niter=None, rtol=1e-6,
             atol=1e-8, eps=1e-12, resid_calc_every=10, return_info=False, **unused):
    """(Preconditioned) BiCGSTAB with r0hat = r0
    (/root/reference/xitorch/_impls/linalg/solve.py:192-324); first step uses the scalar
    initial values alpha=1, omega=1, v=p=0 (:265-268); y = precond_r(p), z = precond_r(s),
    omega = <K t, K s> / <K t, K t> with K = precond_l (:276-287)."""
    A, M = _as_op(A), _as_op(M)
    precond_l, precond_r = _as_op(precond_l), _as_op(precond_r)
    pl = (lambda x: x) if precond_l is None else precond_l.mm
    pr = (lambda x: x) if precond_r is None else precond_r.mm
    nr, ncols = B.shape[-2:]
    if max_niter is None:
        max_niter = int(1.5 * nr)
    batchdims = _batchdims(A, B, E, M)
    if torch.allclose(B, B * 0, rtol=rtol, atol=atol):
        x0 = torch.zeros((*batchdims, nr, ncols), dtype=A.dtype, device=A.device)
        return (x0, {"niter": 0, "converged": True, "napply": A.napply}) if return_info else x0
    A_fcn, _, B2, swapped = setup_linear_problem(A, B, E, M, batchdims, posdef, False)
    bn = B2.norm(dim=-2, keepdim=True)
    stop = torch.max(rtol * bn, atol * torch.ones_like(bn))
    shape = (ncols, *batchdims, nr, 1) if swapped else (*batchdims, nr, ncols)
    xk = torch.zeros(shape, dtype=A.dtype, device=A.device)
    rk = B2 - A_fcn(xk)
    r0hat = rk
    rho_k = _dot(r0hat, rk)
    omega_k = torch.tensor(1.0, dtype=A.dtype, device=A.device)
    alpha = 1.0
    vk = 0.0
    pk = 0.0
    converged = False
    best_resid = rk.norm(dim=-2).max()
    best_x = xk
    k = 0
    for k in range(1, max_niter + 1):
        rho_new = _dot(r0hat, rk)
        omega_den = _safedenom(omega_k, eps)
        beta = rho_new / _safedenom(rho_k, eps) * (alpha / omega_den)
        pk = rk + beta * (pk - omega_k * vk)
        y = pr(pk)
        vk = A_fcn(y)
        alpha = rho_new / _safedenom(_dot(r0hat, vk), eps)
        h = xk + alpha * y
        s = rk - alpha * vk
        z = pr(s)
        t = A_fcn(z)
        Kt = pl(t)
        omega_k = _dot(Kt, pl(s)) / _safedenom(_dot(Kt, Kt), eps)
        xk = h + omega_k * z
        if resid_calc_every != 0 and k % resid_calc_every == 0:
            rk = B2 - A_fcn(xk)
        else:
            rk = s - omega_k * t
        rnorm = rk.norm(dim=-2, keepdim=True)
        mx = rnorm.max().item()
        if mx < best_resid:
            best_resid, best_x = mx, xk
        if torch.all(rnorm < stop):
            converged = True
            break
        rho_k = rho_new
    if not converged:
        warnings.warn(OracleConvergenceWarning(
            "Convergence is not achieved after %d iterations. Max norm of resid: %.3e"
            % (max_niter, float(best_resid))))
    x = _unswap(best_x, swapped)
    if return_info:
        return x, {"niter": k, "converged": converged, "best_resid": float(best_resid),
                   "napply": A.napply}
    return x


def gmres(A, B, E=None, M=None, posdef=None, max_niter=None, rtol=1e-6, atol=1e-8, eps=1e-12,
          return_info=False, **unused):
    """Unrestarted GMRES: MGS Arnoldi, dense lstsq on the Hessenberg every step using the
    first k basis vectors, explicit residual each step
    (/root/reference/xitorch/_impls/linalg/solve.py:326-433).  E is not supported by the
    reference method either (its Hessenberg indexing assumes the un-swapped layout)."""
    A, M = _as_op(A), _as_op(M)
    nr, ncols = A.shape[-1], B.shape[-1]
    if max_niter is None:
        max_niter = int(nr)
    batchdims = _batchdims(A, B, E, M)
    if torch.allclose(B, B * 0, rtol=rtol, atol=atol):
        x0 = torch.zeros((*batchdims, nr, ncols), dtype=A.dtype, device=A.device)
        return (x0, {"niter": 0, "converged": True, "napply": A.napply}) if return_info else x0
    A_fcn, _, B2, swapped = setup_linear_problem(A, B, E, M, batchdims, posdef, False)
    bn = B2.norm(dim=-2, keepdim=True)
    stop = torch.max(rtol * bn, atol * torch.ones_like(bn))
    shape = (ncols, *batchdims, nr, 1) if swapped else (*batchdims, nr, ncols)
    x0 = torch.zeros(shape, dtype=A.dtype, device=A.device)
    r = B2 - A_fcn(x0)
    best_resid = r.norm(dim=-2, keepdim=True).max().item()
    best = x0
    q = torch.empty([max_niter] + list(r.shape), dtype=A.dtype, device=A.device)
    q[0] = r / _safedenom(r.norm(dim=-2, keepdim=True), eps)
    h = torch.zeros((*batchdims, ncols, max_niter + 1, max_niter), dtype=A.dtype, device=A.device)
    h = h.reshape((-1, ncols, max_niter + 1, max_niter))
    converged = False
    k = 0
    for k in range(min(nr, max_niter)):
        y = A_fcn(q[k])
        for j in range(k + 1):
            h[..., j, k] = _dot(q[j], y).reshape(-1, ncols)
            y = y - h[..., j, k].reshape(*batchdims, 1, ncols) * q[j]
        h[..., k + 1, k] = torch.linalg.norm(y, dim=-2)
        if torch.any(h[..., k + 1, k]) != 0 and k != max_niter - 1:
            q[k + 1] = (y.reshape(-1, nr, ncols) / h[..., k + 1, k].reshape(-1, 1, ncols)
                        ).reshape(*batchdims, nr, ncols)
        b = torch.zeros((*batchdims, ncols, k + 1), dtype=A.dtype, device=A.device).reshape(-1, ncols, k + 1)
        b[..., 0] = torch.linalg.norm(r, dim=-2)
        coef = torch.linalg.lstsq(h[..., :k + 1, :k], b)[0]
        res = None
        for i in range(k):
            term = q[i] * coef[..., i].reshape(*batchdims, 1, ncols) + x0
            res = term if res is None else res + term
        if res is not None:
            resid = B2 - A_fcn(res)
            rnorm = resid.norm(dim=-2, keepdim=True)
            mx = rnorm.max().item()
            if mx < best_resid:
                best_resid, best = mx, res
            if torch.all(rnorm < stop):
                converged = True
                break
    if not converged:
        warnings.warn(OracleConvergenceWarning(
            "Convergence is not achieved after %d iterations. Max norm of resid: %.3e"
            % (max_niter, best_resid)))
    if return_info:
        return best, {"niter": k + 1, "converged": converged, "best_resid": best_resid,
                      "napply": A.napply}
    return best


def exactsolve(A, B, E=None, M=None):
    """dense direct solve of A X - M X diag(E) = B, one shifted matrix per column when E is given
    (/root/reference/xitorch/_impls/linalg/solve.py:481-537)."""
    Amat = A.mat if isinstance(A, DenseOp) else A
    if E is None:
        return torch.linalg.solve(Amat, B)
    Mmat = None if M is None else (M.mat if isinstance(M, DenseOp) else M)
    n = Amat.shape[-1]
    eye = torch.eye(n, dtype=Amat.dtype, device=Amat.device) if Mmat is None else Mmat
    cols = []
    for c in range(B.shape[-1]):
        shifted = Amat - E[..., c].unsqueeze(-1).unsqueeze(-1) * eye
        cols.append(torch.linalg.solve(shifted, B[..., c:c + 1]))
    return torch.cat(cols, dim=-1)
