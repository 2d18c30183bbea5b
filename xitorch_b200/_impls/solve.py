"""
Method implementations behind `xitorch_b200.linalg.solve` (plug-in layer L3 of SURVEY.md 1).

`cg`, `bicgstab`, `gmres` keep the reference signatures and option names
(/root/reference/xitorch/_impls/linalg/solve.py:69-80, 192-204, 326-334) but run on the B200 through the
C ABI (`xt_cg`, `xt_bicgstab`, `xt_gmres`): one fused block-matvec kernel per operator application with
the iteration's dot products in its epilogue, device-side convergence control.  They need a CUDA
operator; there is NO CPU fallback (a CPU operator raises RuntimeError).

Host logic kept here (the part of `_setup_linear_problem`, solve.py:560-643, that is not arithmetic):
batch broadcasting, the zero-RHS shortcut, the Hermitian / posdef decision and the normal-equation
switch.  The reference's posdef *probe* (power iterations, solve.py:617-634) compares two vector norms
(`-mostneg_eival <= offset` with both sides >= 0) and therefore always yields posdef=True for finite
operators; it is elided (saves up to 20 operator applications, same decision).
"""
import ctypes as C
import math
import warnings
from typing import Optional, Sequence

import torch

from xitorch_b200 import _lib
from xitorch_b200._utils import ConvergenceWarning, bcast_dims, normalize_bcast_dims
from xitorch_b200.linop import LinearOperator, MatrixLinearOperator

__all__ = ["exactsolve", "custom_exactsolve", "cg", "bicgstab", "gmres", "broyden1_solve", "get_batchdims"]


def get_batchdims(A, B, E, M):
    """broadcast batch shape of the solution (solve.py:540-549)."""
    dims = [A.shape[:-2], B.shape[:-2]]
    if E is not None:
        dims.append(E.shape[:-1])
        if M is not None:
            dims.append(M.shape[:-2])
    return bcast_dims(*dims)


# ----------------------------------------------------------------------------- exact (dense) path
def exactsolve(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor], M: Optional[LinearOperator]):
    """Direct dense solve (row f2 of SURVEY.md 8f; reference solve.py:481-537).  Without E: LU of the full matrix.  With
    E the reference materialises one shifted N x N matrix PER COLUMN (`_solve_ABE`, solve.py:514-537: ncols x N x N --
    unusable at N = 16384); here a Hermitian operator is diagonalised ONCE and every column is a diagonal
    solve in its eigenbasis, a non-Hermitian one is solved column chunk by column chunk within a fixed memory budget.
    Cholesky whitening when M is given."""
    Amat = A.fullmatrix()
    if E is None:
        return torch.linalg.solve(Amat, B)
    hermit = bool(A.is_hermitian) and (M is None or bool(M.is_hermitian))
    if M is None:
        return _solve_shifted(Amat, B, E, hermit)
    L = torch.linalg.cholesky(M.fullmatrix())
    Linv = torch.inverse(L)
    LinvT = Linv.transpose(-2, -1).conj()
    A2 = torch.matmul(Linv, A.mm(LinvT))
    X2 = _solve_shifted(A2, torch.matmul(Linv, B), E, hermit)
    return torch.matmul(LinvT, X2)


_SHIFT_CHUNK_BYTES = 1 << 30        # memory budget of the shifted matrices formed at once (non-Hermitian path)


def _solve_shifted(Amat, B, E, hermitian: bool = False):
    """columns x_j of (A - e_j I) x_j = b_j"""
    n = Amat.shape[-1]
    BA, BB, BE = normalize_bcast_dims(Amat.shape[:-2], B.shape[:-2], E.shape[:-1])
    eps = torch.finfo(Amat.dtype).eps
    if hermitian:
        # A = Q diag(lam) Q^H once:  x_j = Q ((Q^H b_j) / (lam - e_j)).  An (almost) singular shift -- the symeig
        # backward solves at its own eigenvalues -- gets the same remedy as the reference's fallback (a diagonal bump
        # of 10 eps max|A|, solve.py:531-536), applied to the offending denominators only.
        As = 0.5 * (Amat + Amat.transpose(-2, -1).conj())
        lam, Q = torch.linalg.eigh(As.reshape(*BA, n, n))
        Qh = Q.transpose(-2, -1).conj()
        Y = torch.matmul(Qh, B.reshape(*BB, *B.shape[-2:]).to(Q.dtype))              # (*B, n, ncols)
        den = lam.unsqueeze(-1) - E.reshape(*BE, 1, E.shape[-1]).to(lam.dtype)       # (*B, n, ncols)
        bump = 10 * eps * Amat.abs().reshape(*BA, -1).max(dim=-1)[0][..., None, None]
        small = den.abs() < bump
        den = torch.where(small, torch.where(den < 0, -bump, bump).expand_as(den), den)
        return torch.matmul(Q, Y / den)
    ncols = B.shape[-1]
    nbatch = 1
    for d in bcast_dims(BA, BB, BE):
        nbatch *= d
    per_col = max(1, nbatch) * n * n * Amat.element_size()
    chunk = max(1, min(ncols, _SHIFT_CHUNK_BYTES // max(per_col, 1)))
    eye = torch.eye(n, dtype=Amat.dtype, device=Amat.device)
    outs = []
    for c0 in range(0, ncols, chunk):
        c1 = min(ncols, c0 + chunk)
        Ec = E[..., c0:c1].reshape(1, *BE, c1 - c0).transpose(0, -1)               # (cols, *BE, 1)
        Bc = B[..., c0:c1].reshape(1, *BB, n, c1 - c0).transpose(0, -1)            # (cols, *BB, n, 1)
        AE = Amat.reshape(*BA, n, n) - Ec.unsqueeze(-1) * eye                       # (cols, *BAE, n, n)
        try:
            r = torch.linalg.solve(AE, Bc)
        except torch._C._LinAlgError:
            bump = 10 * eps * AE.reshape(*AE.shape[:-2], -1).max(dim=-1)[0][..., None, None]
            r = torch.linalg.solve(AE + eye * bump, Bc)
        outs.append(r.transpose(0, -1).squeeze(0))
        del AE
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)


def custom_exactsolve(A, B, E=None, M=None, **options):
    return exactsolve(A, B, E, M)


# ----------------------------------------------------------------------------- Krylov methods (CUDA only)
def _dense_of(op: LinearOperator, what: str) -> torch.Tensor:
    if isinstance(op, MatrixLinearOperator):
        return op.mat
    # composite / matrix-free operators: materialise (rows f3 of SURVEY.md 8f are not fused yet)
    if op.shape[-1] > 16384:
        raise RuntimeError("xitorch_b200: %s is matrix-free with n=%d; only dense (or materialisable) operators "
                           "are supported by the fused Krylov kernels" % (what, op.shape[-1]))
    with torch.no_grad():
        return op.fullmatrix()


def _default_check_every(n: int, nbatch: int, esize: int) -> int:
    # every poll drains the launch pipeline (~30 us of idle GPU), an iteration enqueued after convergence costs a few
    # microseconds (its kernels exit at once): never poll more often than every 8 iterations
    t_iter = max(nbatch * n * n * esize / 6.0e12, 8e-6)     # one pass over A at ~6 TB/s, >= launch floor
    return max(8, min(64, int(4e-4 / t_iter)))


def _flatten(t: torch.Tensor, batch: Sequence[int], tail: Sequence[int], dtype) -> torch.Tensor:
    nb = 1
    for s in batch:
        nb *= s
    return t.to(dtype).expand(*batch, *tail).reshape(nb, *tail).contiguous()


def _mat3(mat: torch.Tensor, batch: Sequence[int]):
    if mat.dim() == 2 and mat.stride(1) == 1 and mat.data_ptr() % 16 == 0:
        return mat, 0, mat.stride(0)               # the common case without the general broadcast bookkeeping
    from xitorch_b200._dense import flatten_batch
    m3, bstride, ld = flatten_batch(mat, tuple(batch))
    if m3.data_ptr() % 16 != 0:
        m3 = m3.clone()
    return m3, bstride, ld


def _block_callback(fn, ws, nbytes, vdt, shape, opdt, failure, abort=None):
    """ctypes callback `(user, X, Y, stream)` computing ``Y = fn(X)`` on (nbatch, n, ncols) blocks that live inside the
    workspace tensor `ws`; exceptions cannot cross the C frame: they are collected in `failure` and raise the `abort`
    cell (`xt_solve_args.abort`), on which the engine returns at once instead of iterating on garbage."""
    base = ws.data_ptr()

    def _cb(user, xptr, yptr, stream):
        try:
            xv = ws[xptr - base: xptr - base + nbytes].view(vdt).view(*shape)
            yv = ws[yptr - base: yptr - base + nbytes].view(vdt).view(*shape)
            with torch.no_grad():
                yv.copy_(fn(xv.to(opdt)))
        except BaseException as exc:
            if not failure:
                failure.append(exc)
            if abort is not None:
                abort.value = 1

    return _lib.APPLY_FN(_cb)


def _rhs_is_zero(B: torch.Tensor, atol: float) -> bool:
    """``allclose(B, 0, atol=atol)`` of the reference (solve.py:116-119) as one reduction and one synchronisation"""
    return bool(B.abs().amax() <= atol)


def _check_precond(**kw):
    for k, v in kw.items():
        if v is not None and not isinstance(v, LinearOperator):
            raise TypeError("%s can only be LinearOperator or None" % k)


def _is_dense(op: Optional[LinearOperator]) -> bool:
    return op is None or isinstance(op, MatrixLinearOperator)


def _run_matrix_free(name: str, A: LinearOperator, B: torch.Tensor, E, M, posdef, need_hermit: bool,
                     max_niter: int, rtol: float, atol: float, eps: float, resid_calc_every: int,
                     info: Optional[dict], precond_l=None, precond_r=None):
    """Krylov solve with a matrix-free operator (user `_mv`, autograd Jacobians, composite operators).

    The solver loop -- every vector update, dot product, norm, the stop test and the best-iterate bookkeeping --
    runs in the library's CUDA kernels exactly as for a dense operator; only the operator application is handed
    back to Python through the `apply` callback of the C ABI (one call per application, on the current stream).
    The operator built here is the reference's `A_fcn` / `AT_fcn` of `_setup_linear_problem`
    (/root/reference/xitorch/_impls/linalg/solve.py:560-643) in the un-swapped layout:
    ``X -> A.mm(X) - M.mm(X) * E`` and, for the normal equations, its adjoint composed with it."""
    n, ncols = A.shape[-1], B.shape[-1]
    batch = get_batchdims(A, B, E, M)
    if B.is_complex() or A.dtype.is_complex:
        raise RuntimeError("xitorch_b200: complex operators are not supported by the fused Krylov kernels")
    vdt = torch.float64 if A.dtype == torch.float64 else torch.float32
    out_dtype = B.dtype if B.dtype in (torch.float32, torch.float64) else vdt
    if _rhs_is_zero(B, atol):
        return torch.zeros((*batch, n, ncols), dtype=out_dtype, device=B.device)
    opdt = A.dtype
    Er = None if E is None else E.to(opdt).unsqueeze(-2)            # (*BE, 1, ncols)

    def a_fcn(x):
        y = A.mm(x)
        if Er is not None:
            y = y - (M.mm(x) if M is not None else x) * Er
        return y

    def at_fcn(x):
        y = A.rmm(x)
        if Er is not None:
            y = y - (M.rmm(x) if M is not None else x) * Er
        return y

    hermit = A.is_hermitian and (M is None or E is None or M.is_hermitian)
    if need_hermit and not hermit:
        posdef = False
    if posdef is None:
        posdef = True          # what the reference's probe always concludes (see module docstring)
    Bv = B.to(opdt)
    op = a_fcn
    if not posdef:             # normal equations (solve.py:637-643)
        with torch.no_grad():
            Bv = at_fcn(Bv.expand(*batch, n, ncols))
        op = lambda x: at_fcn(a_fcn(x))

    L = _lib.lib()
    nb = 1
    for s_ in batch:
        nb *= s_
    Bf = _flatten(Bv, batch, (n, ncols), vdt)
    X = torch.empty((nb, n, ncols), dtype=vdt, device=Bf.device)
    g = _lib.SolveArgs()
    g.dtype = _lib.dtype_code(vdt)
    g.n, g.nbatch, g.ncols = n, nb, ncols
    g.B, g.ldb, g.b_bstride = Bf.data_ptr(), ncols, n * ncols
    g.X, g.ldx, g.x_bstride = X.data_ptr(), ncols, n * ncols
    g.rtol, g.atol, g.eps = float(rtol), float(atol), float(eps)
    g.max_niter = int(max_niter)
    g.resid_calc_every = int(resid_calc_every)
    g.check_every = 1          # the operator is user code: never run it past convergence
    niter, conv, best, napply = C.c_int32(0), C.c_int32(0), C.c_double(0.0), C.c_int64(0)
    g.niter_out, g.converged_out = C.pointer(niter), C.pointer(conv)
    g.best_resid_out, g.napply_out = C.pointer(best), C.pointer(napply)
    wsb = L.xt_solve_workspace_bytes(name.encode(), g.dtype, n, nb, ncols, g.max_niter, 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device=Bf.device)
    g.workspace, g.workspace_bytes = ws.data_ptr(), wsb
    g.stream = _lib.stream_ptr(Bf.device)
    nbytes = nb * n * ncols * ws.new_empty(0, dtype=vdt).element_size()
    failure = []
    abort = C.c_int32(0)
    g.abort = C.pointer(abort)
    shape = (*batch, n, ncols)
    cb = _block_callback(op, ws, nbytes, vdt, shape, opdt, failure, abort)
    g.apply = _lib.fn_address(cb)
    keep = [cb]
    for field, pc in (("precond_l", precond_l), ("precond_r", precond_r)):
        if pc is not None:
            pcb = _block_callback(pc.mm, ws, nbytes, vdt, shape, opdt, failure, abort)
            setattr(g, field, _lib.fn_address(pcb))
            keep.append(pcb)
    with torch.cuda.device(Bf.device):
        rc = getattr(L, "xt_" + name)(g)
    if failure:
        raise failure[0]
    _lib.check(rc, name)
    if info is not None:
        info.update(niter=niter.value, converged=bool(conv.value), best_resid=best.value, napply=napply.value,
                    matrix_free=True)
    if not conv.value:
        warnings.warn(ConvergenceWarning(
            "Convergence is not achieved after %d iterations. Max norm of best resid: %.3e"
            % (max_niter, best.value)))
    del keep, cb
    return X.reshape(*batch, n, ncols).to(out_dtype)


# ----------------------------------------------------------------------------- complex systems
def _to_real(x: torch.Tensor) -> torch.Tensor:
    """(..., n, c) complex -> (..., 2n, c) real: real parts stacked over imaginary parts."""
    return torch.cat([x.real, x.imag], dim=-2)


def _to_cplx(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[-2] // 2
    return torch.complex(x[..., :n, :].contiguous(), x[..., n:, :].contiguous())


class _RealEquivalent(LinearOperator):
    """the real 2n x 2n form ``[[Re, -Im], [Im, Re]]`` of a complex operator given by callables for the operator and
    its adjoint; Hermitian complex <=> symmetric real, and the real inner product of stacked vectors is ``Re(x^H y)``,
    which is all CG needs (the quantities it divides are real for Hermitian operators)."""

    def __init__(self, fcn, fcn_adj, shape, is_hermitian, rdtype, device):
        super().__init__(shape=shape, is_hermitian=is_hermitian, dtype=rdtype, device=device,
                         _suppress_hermit_warning=True)
        self._fcn, self._fcn_adj = fcn, fcn_adj

    def _mv(self, x):
        return self._mm(x.unsqueeze(-1)).squeeze(-1)

    def _mm(self, x):
        return _to_real(self._fcn(_to_cplx(x)))

    def _rmv(self, x):
        return self._rmm(x.unsqueeze(-1)).squeeze(-1)

    def _rmm(self, x):
        return _to_real(self._fcn_adj(_to_cplx(x)))

    def _getparamnames(self, prefix=""):
        return []


def _run_complex(name, A, B, E, M, posdef, need_hermit, max_niter, rtol, atol, eps, resid_calc_every, check_every,
                 info, precond_l, precond_r):
    """Complex systems (the reference's cg / bicgstab are complex-correct through `_dot`'s conjugate,
    solve.py:441-445) are solved in their real-equivalent form with the same real kernels: for cg the recurrences are
    identical to the complex ones; bicgstab / gmres become the real-arithmetic method on the doubled system (same
    solution at convergence, different iteration path)."""
    cdt = A.dtype if A.dtype.is_complex else (torch.complex128 if B.dtype in (torch.float64, torch.complex128)
                                              else torch.complex64)
    rdt = torch.float64 if cdt == torch.complex128 else torch.float32
    n = A.shape[-1]
    Bc = B.to(cdt)
    e_real = E is None or not E.is_complex()
    if isinstance(A, MatrixLinearOperator) and A.mat.is_cuda and M is None and e_real and precond_l is None and \
            precond_r is None:
        # dense fast path: the doubled matrix goes through the block-matvec kernels like any real operator
        Am = A.mat.to(cdt)
        Ar = torch.cat([torch.cat([Am.real, -Am.imag], dim=-1), torch.cat([Am.imag, Am.real], dim=-1)], dim=-2)
        opr = MatrixLinearOperator(Ar.contiguous(), is_hermitian=A.is_hermitian)
        Er = None if E is None else E.to(rdt)
        xr = _run_krylov(name, opr, _to_real(Bc), Er, None, posdef, need_hermit, max_niter, rtol, atol, eps,
                         resid_calc_every, check_every, info)
        return _to_cplx(xr)
    Ec = None if E is None else E.to(cdt).unsqueeze(-2)

    def fcn(x):
        y = A.mm(x)
        if Ec is not None:
            y = y - (M.mm(x) if M is not None else x) * Ec
        return y

    def fcn_adj(x):
        y = A.rmm(x)
        if Ec is not None:
            y = y - (M.rmm(x) if M is not None else x) * Ec.conj()
        return y

    hermit = A.is_hermitian and (E is None or (e_real and (M is None or M.is_hermitian)))
    batch = get_batchdims(A, B, E, M)
    opr = _RealEquivalent(fcn, fcn_adj, (*batch, 2 * n, 2 * n), hermit, rdt, B.device)
    pcs = []
    for pc in (precond_l, precond_r):
        pcs.append(None if pc is None else
                   _RealEquivalent(pc.mm, pc.rmm, (*pc.shape[:-2], 2 * n, 2 * n), pc.is_hermitian, rdt, B.device))
    xr = _run_matrix_free(name, opr, _to_real(Bc), None, None, posdef, need_hermit, max_niter, rtol, atol, eps,
                          resid_calc_every, info, pcs[0], pcs[1])
    return _to_cplx(xr)


def _run_krylov(name: str, A: LinearOperator, B: torch.Tensor, E, M, posdef, need_hermit: bool,
                max_niter: int, rtol: float, atol: float, eps: float, resid_calc_every: int,
                check_every: Optional[int], info: Optional[dict], precond_l=None, precond_r=None):
    _lib.require_cuda(B, "linalg.solve(method=%r)" % name)
    _check_precond(precond_l=precond_l, precond_r=precond_r)
    if A.dtype.is_complex or B.is_complex() or (E is not None and E.is_complex()):
        return _run_complex(name, A, B, E, M, posdef, need_hermit, max_niter, rtol, atol, eps, resid_calc_every,
                            check_every, info, precond_l, precond_r)
    if not _is_dense(A) or (E is not None and not _is_dense(M)):
        return _run_matrix_free(name, A, B, E, M, posdef, need_hermit, max_niter, rtol, atol, eps,
                                resid_calc_every, info, precond_l, precond_r)
    n, ncols = A.shape[-1], B.shape[-1]
    batch = get_batchdims(A, B, E, M)
    Amat = _dense_of(A, "A")
    _lib.require_cuda(Amat, "linalg.solve(method=%r)" % name)
    if Amat.is_complex() or B.is_complex():
        raise RuntimeError("xitorch_b200: complex operators are not supported by the fused Krylov kernels")
    vdt = _lib.vec_dtype(Amat.dtype)
    out_dtype = B.dtype if B.dtype in (torch.float32, torch.float64) else vdt

    # zero right-hand side -> zeros (solve.py:116-119)
    if _rhs_is_zero(B, atol):
        return torch.zeros((*batch, n, ncols), dtype=out_dtype, device=B.device)

    Mmat = None
    if E is not None and M is not None:
        Mmat = _dense_of(M, "M").to(Amat.dtype)

    hermit = A.is_hermitian and (M is None or E is None or M.is_hermitian)
    if need_hermit and not hermit:
        posdef = False
    if posdef is None:
        posdef = True          # what the reference's probe always concludes (see module docstring)

    Bv = B
    if not posdef:
        # normal equations (solve.py:637-643), MATRIX-FREE as in the reference: the operator is x -> F^H (F x) with
        # F x = A x - (M x) E, two (three with M) dense passes per application -- the forward block matvec and the
        # transposed-access one (`xt_block_matvec`, trans = 1), which reads A once without materialising A^T.  No
        # N x N product is formed: nothing is squared in storage precision and the E case needs no matrix per column.
        return _run_matrix_free(name, A, B, E, M, False, need_hermit, max_niter, rtol, atol, eps, resid_calc_every,
                                info, precond_l, precond_r)
    x = _call(name, Amat, Mmat, E, Bv, batch, n, ncols, vdt, max_niter, rtol, atol, eps, resid_calc_every,
              check_every, info, precond_l, precond_r)
    return x.to(out_dtype)


def _call(name, Amat, Mmat, E, B, batch, n, ncols, vdt, max_niter, rtol, atol, eps, resid_calc_every,
          check_every, info, precond_l=None, precond_r=None):
    L = _lib.lib()
    nb = 1
    for s in batch:
        nb *= s
    A3, a_bs, lda = _mat3(Amat, batch)
    Bf = _flatten(B, batch, (n, ncols), vdt)
    X = torch.empty((nb, n, ncols), dtype=vdt, device=Bf.device)
    keep = [A3, Bf, X]
    g = _lib.SolveArgs()
    g.dtype = _lib.dtype_code(Amat.dtype)
    g.n, g.nbatch, g.ncols = n, nb, ncols
    g.A, g.lda, g.a_bstride = A3.data_ptr(), lda, a_bs
    if Mmat is not None:
        M3, m_bs, ldm = _mat3(Mmat, batch)
        keep.append(M3)
        g.M, g.ldm, g.m_bstride = M3.data_ptr(), ldm, m_bs
    if E is not None:
        Ef = _flatten(E, batch, (ncols,), vdt)
        keep.append(Ef)
        g.E, g.e_bstride = Ef.data_ptr(), ncols
    g.B, g.ldb, g.b_bstride = Bf.data_ptr(), ncols, n * ncols
    g.X, g.ldx, g.x_bstride = X.data_ptr(), ncols, n * ncols
    g.rtol, g.atol, g.eps = float(rtol), float(atol), float(eps)
    g.max_niter = int(max_niter)
    g.resid_calc_every = int(resid_calc_every)
    g.check_every = int(check_every) if check_every else _default_check_every(n, nb, Amat.element_size())
    niter, conv, best, napply = C.c_int32(0), C.c_int32(0), C.c_double(0.0), C.c_int64(0)
    g.niter_out, g.converged_out = C.pointer(niter), C.pointer(conv)
    g.best_resid_out, g.napply_out = C.pointer(best), C.pointer(napply)
    wsb = L.xt_solve_workspace_bytes(name.encode(), g.dtype, n, nb, ncols, g.max_niter, 1 if Mmat is not None else 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device=Bf.device)
    g.workspace, g.workspace_bytes = ws.data_ptr(), wsb
    g.stream = _lib.stream_ptr(Bf.device)
    failure = []
    if precond_l is not None or precond_r is not None:
        # preconditioners reach the library as block callbacks (a dense one is a single xt_block_matvec per call)
        nbytes = nb * n * ncols * X.element_size()
        abort = C.c_int32(0)
        g.abort = C.pointer(abort)
        keep.append(abort)
        for field, pc in (("precond_l", precond_l), ("precond_r", precond_r)):
            if pc is not None:
                pcb = _block_callback(pc.mm, ws, nbytes, vdt, (*batch, n, ncols), pc.dtype, failure, abort)
                setattr(g, field, _lib.fn_address(pcb))
                keep.append(pcb)
        g.check_every = 1
    with torch.cuda.device(Bf.device):
        rc = getattr(L, "xt_" + name)(g)
    if failure:
        raise failure[0]
    _lib.check(rc, name)
    if info is not None:
        info.update(niter=niter.value, converged=bool(conv.value), best_resid=best.value, napply=napply.value)
    if not conv.value:
        warnings.warn(ConvergenceWarning(
            "Convergence is not achieved after %d iterations. Max norm of best resid: %.3e"
            % (max_niter, best.value)))
    return X.reshape(*batch, n, ncols)


def cg(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor] = None, M: Optional[LinearOperator] = None,
       posdef: Optional[bool] = None, precond: Optional[LinearOperator] = None,
       max_niter: Optional[int] = None, rtol: float = 1e-6, atol: float = 1e-8, eps: float = 1e-12,
       resid_calc_every: int = 10, verbose: bool = False, check_every: Optional[int] = None,
       info: Optional[dict] = None, **unused) -> torch.Tensor:
    r"""
    Conjugate gradient on the B200 (fused matvec + p.Ap kernel per iteration).

    Keyword arguments
    -----------------
    posdef: bool or None
        Whether :math:`\mathbf{AX-MXE}` is positive definite for all columns and batches.
        ``False`` (or a non-Hermitian operator) switches to the normal equations. ``None`` means True.
    precond: LinearOperator or None
        Preconditioner ``z = precond.mm(r)`` (one more operator application per iteration).
    max_niter: int or None
        Maximum number of iterations. If None, ``int(1.5 * A.shape[-1])``.
    rtol, atol: float
        Stop when every column satisfies ``||r|| < max(rtol * ||b||, atol)``.
    eps: float
        Replacement for exactly-zero denominators.
    resid_calc_every: int
        Recompute the true residual ``B - A x`` every this many iterations (0: never).
    verbose: bool
        Ignored (convergence control lives on the device).
    check_every: int or None
        Host polls the device convergence flag every this many iterations (None: chosen from the size).
    info: dict or None
        If given, receives ``niter``, ``converged``, ``best_resid``, ``napply``.
    """
    if max_niter is None:
        max_niter = int(1.5 * A.shape[-1])
    return _run_krylov("cg", A, B, E, M, posdef, True, max_niter, rtol, atol, eps, resid_calc_every,
                       check_every, info, precond_l=precond)


def bicgstab(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor] = None,
             M: Optional[LinearOperator] = None, posdef: Optional[bool] = None,
             precond_l: Optional[LinearOperator] = None, precond_r: Optional[LinearOperator] = None,
             max_niter: Optional[int] = None, rtol: float = 1e-6, atol: float = 1e-8, eps: float = 1e-12,
             verbose: bool = False, resid_calc_every: int = 10, check_every: Optional[int] = None,
             info: Optional[dict] = None, **unused) -> torch.Tensor:
    r"""
    Stabilised bi-conjugate gradient on the B200 (two fused matvec kernels per iteration carrying
    ``r0hat.v`` and ``t.s, t.t``).

    Keyword arguments
    -----------------
    posdef: bool or None
        ``False`` switches to the normal equations; ``None`` means True.
    precond_l, precond_r: LinearOperator or None
        Left / right preconditioners, used exactly as the reference does: ``y = precond_r.mm(p)``,
        ``z = precond_r.mm(s)`` and ``omega = <K t, K s> / <K t, K t>`` with ``K = precond_l``.
    max_niter: int or None
        Maximum number of iterations. If None, ``int(1.5 * A.shape[-1])``.
    rtol, atol: float
        Stop when every column satisfies ``||r|| < max(rtol * ||b||, atol)``.
    eps: float
        Replacement for exactly-zero denominators.
    verbose: bool
        Ignored.
    resid_calc_every: int
        Recompute the true residual every this many iterations (0: never).
    check_every, info:
        As in :func:`cg`.
    """
    if max_niter is None:
        max_niter = int(1.5 * A.shape[-1])
    return _run_krylov("bicgstab", A, B, E, M, posdef, False, max_niter, rtol, atol, eps, resid_calc_every,
                       check_every, info, precond_l=precond_l, precond_r=precond_r)


def gmres(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor] = None, M: Optional[LinearOperator] = None,
          posdef: Optional[bool] = None, max_niter: Optional[int] = None, rtol: float = 1e-6,
          atol: float = 1e-8, eps: float = 1e-12, check_every: Optional[int] = None,
          info: Optional[dict] = None, **unused) -> torch.Tensor:
    r"""
    Unrestarted GMRES on the B200 (fused matvec, block classical Gram-Schmidt twice, Givens
    least squares on the device).

    Keyword arguments
    -----------------
    posdef: bool or None
        ``False`` switches to the normal equations; ``None`` means True.
    max_niter: int or None
        Maximum number of iterations. If None, ``A.shape[-1]`` (as the reference, solve.py:357).  One Krylov cycle holds
        at most 256 vectors (the Hessenberg matrix lives on chip); beyond that the method restarts from the current
        iterate (GMRES(256)) until ``max_niter`` iterations have been spent.
    rtol, atol: float
        Stop when every column satisfies ``||r|| < max(rtol * ||b||, atol)``.
    eps: float
        Replacement for exactly-zero denominators.
    check_every, info:
        As in :func:`cg`.
    """
    if E is not None:
        raise RuntimeError("gmres does not support E (neither does the reference method: "
                           "xitorch/_impls/linalg/solve.py:386-398)")
    if max_niter is None:
        max_niter = int(A.shape[-1])
    cycle = 256
    if max_niter <= cycle:
        return _run_krylov("gmres", A, B, E, M, posdef, False, max_niter, rtol, atol, eps, 0, check_every, info)
    # restarted cycles: solve A dx = r for the current residual, with the ABSOLUTE target of the original problem
    run = {}
    X = _run_krylov("gmres", A, B, E, M, posdef, False, cycle, rtol, atol, eps, 0, check_every, run)
    spent = int(run.get("niter", cycle))
    if not run.get("converged", True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", ConvergenceWarning)
            bnorm = B.norm(dim=-2, keepdim=True)
            target = torch.clamp(rtol * bnorm, min=atol)
            while spent < max_niter:
                R = B - A.mm(X)
                if bool((R.norm(dim=-2, keepdim=True) < target).all()):
                    run["converged"] = True
                    break
                run = {}
                dX = _run_krylov("gmres", A, R, None, M, posdef, False, min(cycle, max_niter - spent), 0.0,
                                 float(target.min().item()), eps, 0, check_every, run)
                X = X + dX
                spent += int(run.get("niter", cycle))
        if not run.get("converged", True):
            R = B - A.mm(X)
            if not bool((R.norm(dim=-2, keepdim=True) < target).all()):
                warnings.warn(ConvergenceWarning("Convergence is not achieved after %d iterations. Max norm of best "
                                                 "resid: %.3e" % (spent, R.norm(dim=-2).max().item())))
            else:
                run["converged"] = True
    if info is not None:
        info.update(run)
        info["niter"] = spent
    return X


def broyden1_solve(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor] = None,
                   M: Optional[LinearOperator] = None, **options) -> torch.Tensor:
    """Solve the linear system as the root of ``A X - M X E - B`` with the first Broyden method
    (reference: _rootfinder_solve, /root/reference/xitorch/_impls/linalg/solve.py:448-478); options of
    :func:`xitorch_b200._impls.rootsolver.broyden1`."""
    from xitorch_b200._impls.rootsolver import broyden1
    nr, ncols = A.shape[-1], B.shape[-1]

    def resid(xi):
        x = xi.reshape(*xi.shape[:-1], nr, ncols)
        y = A.mm(x) - B
        if E is not None:
            y = y - (M.mm(x) if M is not None else x) * E.unsqueeze(-2)
        return y.reshape(*xi.shape[:-1], -1)

    batch = get_batchdims(A, B, E, M)
    x0 = torch.zeros((*batch, nr * ncols), dtype=A.dtype, device=A.device)
    x = broyden1(resid, x0, **options)
    return x.reshape(*x.shape[:-1], nr, ncols)


for _m in (cg, bicgstab, gmres):
    _m._checks_zero_rhs = True          # see linalg/solve.py::_forward
