"""
Quasi-Newton root solvers behind `xitorch_b200.optimize.rootfinder` (row f1 of SURVEY.md 8f; reference:
/root/reference/xitorch/_impls/optimize/root/rootsolver.py:15-149 and _jacobian.py:51-199, themselves after
scipy.optimize.nonlin).

Same algorithm and options as the reference -- inexact-Newton outer loop with a Broyden approximation of the inverse
Jacobian ``G = -alpha I + sum_i c_i d_i^T``, Armijo backtracking on ``|f|^2``, best-iterate bookkeeping and the
four-norm termination test -- with one structural change for the GPU: the rank-r matrix is not a Python list of
vectors applied by r dot + r axpy launches (_jacobian.py:172-182) but four growing device buffers, so that
``G v`` / ``G^T v`` are two launches of the dense block-matvec kernel (`xt_block_matvec`) whatever the rank:
``coef = D_rows v`` (r x n) and ``alpha v + C_cols coef`` (n x r).
"""
import functools
import warnings
from typing import Optional, Tuple, Union

import numpy as np
import torch

from xitorch_b200._utils import ConvergenceWarning

__all__ = ["newton", "broyden1", "broyden2", "linearmixing"]


# ----------------------------------------------------------------------------- rank-r inverse Jacobian
class LowRankMatrix(object):
    """``alpha I + sum_i c_i d_i^T`` (reference _jacobian.py:156-199).  c_i / d_i are kept both as rows of
    (cap, n) buffers and as columns of (n, cap) buffers so that both products of `mv` and of `rmv` read a
    row-major matrix."""

    def __init__(self, alpha, n: int, dtype, device, uv0=None):
        self.alpha = alpha
        self.n = n
        self.r = 0
        self.cap = 0
        self.dtype, self.device = dtype, device
        self.c_rows = self.d_rows = self.c_cols = self.d_cols = None
        if uv0 is not None:
            self.append(uv0[0].reshape(-1), uv0[1].reshape(-1))

    def _grow(self):
        cap = max(32, 2 * self.cap)
        def rows():
            return torch.zeros((cap, self.n), dtype=self.dtype, device=self.device)
        def cols():
            return torch.zeros((self.n, cap), dtype=self.dtype, device=self.device)
        new = [rows(), rows(), cols(), cols()]
        if self.r > 0:
            new[0][:self.r] = self.c_rows[:self.r]
            new[1][:self.r] = self.d_rows[:self.r]
            new[2][:, :self.r] = self.c_cols[:, :self.r]
            new[3][:, :self.r] = self.d_cols[:, :self.r]
        self.c_rows, self.d_rows, self.c_cols, self.d_cols = new
        self.cap = cap

    @staticmethod
    def _matvec(mat: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
        if mat.is_cuda and mat.dtype in (torch.float32, torch.float64):
            from xitorch_b200 import _dense
            return _dense.block_matvec(mat, v.unsqueeze(-1)).squeeze(-1)
        return torch.matmul(mat, v)

    def _apply(self, left_rows, right_cols, v):
        res = self.alpha * v
        if self.r == 0:
            return res
        coef = self._matvec(left_rows[:self.r], v)                 # (r,)   = [d_i . v]
        return res + self._matvec(right_cols[:, :self.r], coef)    # (n,)  += sum_i c_i coef_i

    def mv(self, v):
        return self._apply(self.d_rows, self.c_cols, v)

    def rmv(self, v):
        return self._apply(self.c_rows, self.d_cols, v)

    def append(self, c, d):
        if self.r == self.cap:
            self._grow()
        i = self.r
        self.c_rows[i] = c
        self.d_rows[i] = d
        self.c_cols[:, i] = c
        self.d_cols[:, i] = d
        self.r += 1
        return self

    def reduce(self, max_rank):
        if self.r > max_rank:          # "restart" policy of the reference (_jacobian.py:186-190)
            self.r = 0


class _Broyden(object):
    """Broyden's first ("good") method: inverse-Jacobian update (_jacobian.py:51-125)."""
    second = False

    def __init__(self, alpha=None, uv0=None, max_rank=None):
        self.alpha, self.uv0, self.max_rank = alpha, uv0, max_rank

    def setup(self, x0, y0, func):
        self.x_prev, self.y_prev = x0, y0
        if self.max_rank is None:
            self.max_rank = float("inf")
        if self.alpha is None:
            normy0 = torch.norm(y0)
            ones = torch.ones_like(normy0)
            self.alpha = 0.5 * torch.max(torch.norm(x0), ones) / normy0 if normy0 else ones
        if isinstance(self.uv0, str):
            if self.uv0 != "svd":
                raise RuntimeError("Unknown uv0: %s" % self.uv0)
            self.uv0 = _svd_uv0(func, x0)
        self.Gm = LowRankMatrix(-self.alpha, x0.numel(), x0.dtype, x0.device, self.uv0)

    def solve(self, v, tol=0):
        return self.Gm.mv(v)

    def update(self, x, y):
        dy = y - self.y_prev
        dx = x - self.x_prev
        self.Gm.reduce(self.max_rank)
        if self.second:
            v = dy
            c = dx - self.Gm.mv(dy)
            d = v / torch.dot(dy, dy)
        else:
            v = self.Gm.rmv(dx)
            c = dx - self.Gm.mv(dy)
            d = v / torch.dot(dy, v)
        self.Gm.append(c, d)
        self.y_prev, self.x_prev = y, x


class _Broyden2(_Broyden):
    second = True


class _Newton(object):
    """exact Jacobian: each step solves ``J(x) s = f(x)`` with `linalg.solve` (_jacobian.py:26-49).  With a Krylov
    `solver_method` on CUDA tensors this is Newton-Krylov: the linear solve runs in the CUDA solver kernels with the
    matrix-free Jacobian applied through the operator callback."""

    def __init__(self, solver_method="exactsolve", solver_kwargs=None):
        self.solver_method = solver_method
        self.solver_kwargs = {} if solver_kwargs is None else solver_kwargs

    def setup(self, x0, y0, func):
        self.x, self.func = x0, func

    def solve(self, v, tol=0):
        from xitorch_b200.grad import jac
        from xitorch_b200.linalg import solve
        J = jac(self.func, (self.x.clone().requires_grad_(),), idxs=0)
        return solve(J, v.unsqueeze(-1), method=self.solver_method, **self.solver_kwargs).squeeze(-1)

    def update(self, x, y):
        self.x = x


class _LinearMixing(object):
    def __init__(self, alpha=None):
        self.alpha = -1.0 if alpha is None else alpha

    def setup(self, x0, y0, func):
        pass

    def solve(self, v, tol=0):
        return -v * self.alpha

    def update(self, x, y):
        pass


def _svd_uv0(func, x0):
    # rank-1 start from the smallest singular triplet of the Jacobian (_jacobian.py:224-233)
    from xitorch_b200.grad import jac
    from xitorch_b200.linalg import svd
    fjac = jac(func, (x0.clone().requires_grad_(),), idxs=[0])[0]
    u, s, vh = svd(fjac, k=1, mode="lowest", method="davidson", min_eps=1e-3)
    sinv_sqrt = 1.0 / torch.sqrt(torch.clamp(s, min=0.1))
    return (sinv_sqrt * vh.squeeze(-2), sinv_sqrt * u.squeeze(-1))


# ----------------------------------------------------------------------------- outer loop
def _host_scalars(*tensors):
    """the values of several 0-dim tensors with ONE device -> host transfer, as numpy scalars of the tensors' dtype: the
    control arithmetic done with them on the host (line-search interpolation, forcing terms, stop tests) rounds exactly
    like the 0-dim tensor arithmetic of the reference, without a kernel launch per scalar operation."""
    items = [t if t.dim() == 0 else t.reshape(()) for t in tensors]
    dt = items[0].dtype
    if any(t.dtype != dt for t in items):
        items = [t.double() for t in items]                    # mixed precisions: exact in fp64
    elif dt not in (torch.float32, torch.float64):
        items = [t.float() for t in items]                     # half / bfloat16 have no numpy scalar type
    stacked = torch.stack(items)
    arr = stacked.cpu().numpy()
    return [arr[i] for i in range(len(tensors))]


class TerminationCondition(object):
    def __init__(self, f_tol, f_rtol, f0_norm, x_tol, x_rtol):
        self.f_tol = 1e-6 if f_tol is None else f_tol
        self.f_rtol = float("inf") if f_rtol is None else f_rtol
        self.x_tol = 1e-6 if x_tol is None else x_tol
        self.x_rtol = float("inf") if x_rtol is None else x_rtol
        self.f0_norm = _host_scalars(f0_norm)[0] if isinstance(f0_norm, torch.Tensor) else f0_norm

    def check_norms(self, xnorm, ynorm, dxnorm) -> bool:
        with np.errstate(all="ignore"):
            return bool((dxnorm < self.x_tol) and (dxnorm < self.x_rtol * xnorm) and
                        (ynorm < self.f_tol) and (ynorm < self.f_rtol * self.f0_norm))

    def check(self, x, y, dx) -> bool:
        return self.check_norms(*_host_scalars(x.norm(), y.norm(), dx.norm()))


def _nonlin_solver(fcn, x0, params, jacobian, maxiter=None, f_tol=None, f_rtol=None, x_tol=None, x_rtol=None,
                   line_search=True, verbose=False, custom_terminator=None, **unused):
    """
    Keyword arguments
    -----------------
    maxiter: int or None
        Maximum number of iterations, or ``100 * (numel + 1)`` if None.
    f_tol: float or None
        The absolute tolerance of the norm of the output ``f``.
    f_rtol: float or None
        The relative tolerance of the norm of the output ``f``.
    x_tol: float or None
        The absolute tolerance of the norm of the input ``x``.
    x_rtol: float or None
        The relative tolerance of the norm of the input ``x``.
    line_search: bool or str
        Options to perform line search. If ``True``, it is set to ``"armijo"``.
    verbose: bool
        Options for verbosity
    """
    if maxiter is None:
        maxiter = 100 * (torch.numel(x0) + 1)
    if line_search is True:
        line_search = "armijo"
    elif line_search is False:
        line_search = None
    xshape = x0.shape
    is_cplx = torch.is_complex(x0)

    if is_cplx:                      # complex unknowns: real and imaginary parts stacked into one real vector
        def ravel(x):
            return torch.cat((x.real, x.imag), dim=0).reshape(-1)

        def pack(x):
            h = len(x) // 2
            return (x[:h] + 1j * x[h:]).reshape(xshape)
    else:
        def ravel(x):
            return x.reshape(-1)

        def pack(x):
            return x.reshape(xshape)

    def func(x):
        return ravel(fcn(pack(x), *params))

    x = ravel(x0)
    y = func(x)
    y_norm, x_norm = _host_scalars(y.norm(), x.norm())
    stop_cond = custom_terminator if custom_terminator is not None else \
        TerminationCondition(f_tol, f_rtol, y_norm, x_tol, x_rtol)
    if y_norm == 0:
        return x.reshape(xshape)
    jacobian.setup(x, y, func)

    # All control decisions below are taken on host copies of the norms, fetched with one transfer per function
    # evaluation (the reference syncs on every comparison of 0-dim tensors: ~4 per evaluation plus ~40 scalar kernels
    # per line-search interpolation).
    gamma, eta_max, eta_threshold, eta = 0.9, 0.9999, 0.1, 1e-3     # forcing terms of the inexact Newton step
    converged = False
    best_ynorm, best_x, best_dxnorm, best_iter = y_norm, x, x_norm, 0
    for i in range(maxiter):
        tol = float(min(eta, eta * y_norm))
        dx = -jacobian.solve(y, tol=tol)
        if line_search:
            s, xnew, ynew, (y_norm_new, x_norm_new, dx_norm) = _line_search(func, x, y, y_norm, dx,
                                                                           search_type=line_search)
        else:
            xnew = x + dx
            ynew = func(xnew)
            y_norm_new, x_norm_new, dx_norm = _host_scalars(ynew.norm(), xnew.norm(), dx.norm())
        if dx_norm == 0:
            raise ValueError("Jacobian inversion yielded zero vector. This indicates a bug in the Jacobian "
                             "approximation.")
        if y_norm_new < best_ynorm:
            best_x, best_dxnorm, best_ynorm, best_iter = xnew, dx_norm, y_norm_new, i + 1
        jacobian.update(xnew, ynew)
        if custom_terminator is not None:
            to_stop = stop_cond.check(xnew, ynew, dx)
        else:
            to_stop = stop_cond.check_norms(x_norm_new, y_norm_new, dx_norm)
        if verbose and (i < 10 or i % 10 == 0 or to_stop):
            print("%6d: |dx|=%.3e, |f|=%.3e" % (i, dx_norm, y_norm))
        if to_stop:
            # as in the reference (rootsolver.py:129-143) the iterate returned on convergence is the one BEFORE this
            # last step (|dx| < x_tol apart from xnew)
            converged = True
            break
        with np.errstate(all="ignore"):
            ratio = y_norm_new / y_norm
            eta_A = float(gamma * (ratio * ratio))
        gamma_eta2 = gamma * eta * eta
        eta = min(eta_max, eta_A) if gamma_eta2 < eta_threshold else min(eta_max, max(eta_A, gamma_eta2))
        y_norm, x, y = y_norm_new, xnew, ynew
    if not converged:
        warnings.warn(ConvergenceWarning(
            "The rootfinder does not converge after %d iterations. Best |dx|=%.3e, |f|=%.3e at iter %d"
            % (maxiter, best_dxnorm, best_ynorm, best_iter)))
        x = best_x
    return pack(x)


def _line_search(func, x, y, y_norm, dx, search_type="armijo", rdiff=1e-8, smin=1e-2):
    """step length by backtracking on phi(s) = |f(x + s dx)|^2.  Returns ``(s, xnew, ynew, (|ynew|, |xnew|, |dx|))``, the
    norms on the host: every evaluation of phi brings them along in its one transfer, so an accepted step needs no
    further synchronisation."""
    dx_norm_dev = dx.norm()
    cache = {"s": 0, "phi": y_norm * y_norm, "x": x, "y": y, "norms": None}

    def phi(s):
        if s == cache["s"]:
            return cache["phi"]
        xs = x + float(s) * dx
        v = func(xs)
        vf = v.reshape(-1)
        p, vn, xn, dn = _host_scalars(torch.dot(vf, vf), v.norm(), xs.norm(), dx_norm_dev)
        cache.update(s=s, phi=p, x=xs, y=v, norms=(vn, xn, dn))
        return p

    s = None
    if search_type == "armijo":
        with np.errstate(all="ignore"):
            s, _ = _armijo(phi, cache["phi"], -cache["phi"], amin=smin)
    if s is None:
        s = 1.0                      # no acceptable step length: take the full step and hope for the best
    if s == cache["s"] and cache["norms"] is not None:
        return s, cache["x"], cache["y"], cache["norms"]
    xnew = x + float(s) * dx
    ynew = func(xnew)
    return s, xnew, ynew, tuple(_host_scalars(ynew.norm(), xnew.norm(), dx_norm_dev))


def _armijo(phi, phi0, derphi0, c1=1e-4, alpha0=1, amin=0, max_niter=20):
    # backtracking with quadratic, then cubic interpolation (scipy's scalar_search_armijo) on host scalars of the
    # problem's dtype; powers are written as products, which is what the tensor `**` of the reference evaluates
    def sq(a):
        return a * a

    def cube(a):
        return a * a * a

    phi_a0 = phi(alpha0)
    if phi_a0 <= phi0 + c1 * alpha0 * derphi0:
        return alpha0, phi_a0
    alpha1 = -derphi0 * sq(alpha0) / 2.0 / (phi_a0 - phi0 - derphi0 * alpha0)
    phi_a1 = phi(alpha1)
    if phi_a1 <= phi0 + c1 * alpha1 * derphi0:
        return alpha1, phi_a1
    niter = 0
    alpha2, phi_a2 = alpha1, phi_a1
    while alpha1 > amin and niter < max_niter:
        factor = sq(alpha0) * sq(alpha1) * (alpha1 - alpha0)
        a = sq(alpha0) * (phi_a1 - phi0 - derphi0 * alpha1) - sq(alpha1) * (phi_a0 - phi0 - derphi0 * alpha0)
        a = a / factor
        b = -cube(alpha0) * (phi_a1 - phi0 - derphi0 * alpha1) + cube(alpha1) * (phi_a0 - phi0 - derphi0 * alpha0)
        b = b / factor
        alpha2 = (-b + np.sqrt(np.abs(sq(b) - 3 * a * derphi0))) / (3.0 * a)
        phi_a2 = phi(alpha2)
        if phi_a2 <= phi0 + c1 * alpha2 * derphi0:
            return alpha2, phi_a2
        if (alpha1 - alpha2) > alpha1 / 2.0 or (1 - alpha2 / alpha1) < 0.96:
            alpha2 = alpha1 / 2.0
        alpha0, alpha1, phi_a0, phi_a1 = alpha1, alpha2, phi_a1, phi_a2
        niter += 1
    if niter == max_niter:
        return alpha2, phi_a2
    return None, phi_a1


# ----------------------------------------------------------------------------- methods
def newton(fcn, x0, params=(), *, solver_method: str = "exactsolve", solver_kwargs: Optional[dict] = None, **kwargs):
    """
    Solve the root finder using the Newton method: ``x <- x - J(x)^{-1} f(x)`` with ``J`` the Jacobian of ``f``.

    Keyword arguments
    -----------------
    solver_method: str
        The `xitorch_b200.linalg.solve` method that solves with the Jacobian.
    solver_kwargs: dict or None
        The keyword arguments of that solve method.
    """
    return _nonlin_solver(fcn, x0, params, jacobian=_Newton(solver_method, solver_kwargs), **kwargs)


def broyden1(fcn, x0, params=(), *, alpha: Optional[float] = None,
             uv0: Optional[Union[str, Tuple[torch.Tensor, torch.Tensor]]] = None,
             max_rank: Optional[int] = None, **kwargs):
    """
    Solve the root finder or linear equation using the first Broyden method.

    Keyword arguments
    -----------------
    alpha: float or None
        The initial guess of inverse Jacobian is ``- alpha * I + u v^T``.
    uv0: tuple of tensors or str or None
        The initial guess of inverse Jacobian is ``- alpha * I + u v^T``.
        If ``"svd"``, then it uses 1-rank svd to obtain ``u`` and ``v``.
        If None, then ``u`` and ``v`` are zeros.
    max_rank: int or None
        The maximum rank of inverse Jacobian approximation. If ``None``, it is ``inf``.
    """
    return _nonlin_solver(fcn, x0, params, jacobian=_Broyden(alpha=alpha, uv0=uv0, max_rank=max_rank), **kwargs)


@functools.wraps(_nonlin_solver, assigned=("__annotations__",))
def broyden2(fcn, x0, params=(), *, alpha: Optional[float] = None,
             uv0: Optional[Union[str, Tuple[torch.Tensor, torch.Tensor]]] = None,
             max_rank: Optional[int] = None, **kwargs):
    """
    Solve the root finder or linear equation using the second Broyden method.

    Keyword arguments
    -----------------
    alpha: float or None
        The initial guess of inverse Jacobian is ``- alpha * I + u v^T``.
    uv0: tuple of tensors or str or None
        The initial guess of inverse Jacobian is ``- alpha * I + u v^T``.
        If ``"svd"``, then it uses 1-rank svd to obtain ``u`` and ``v``.
        If None, then ``u`` and ``v`` are zeros.
    max_rank: int or None
        The maximum rank of inverse Jacobian approximation. If ``None``, it is ``inf``.
    """
    return _nonlin_solver(fcn, x0, params, jacobian=_Broyden2(alpha=alpha, uv0=uv0, max_rank=max_rank), **kwargs)


def linearmixing(fcn, x0, params=(), *, alpha: Optional[float] = None, **kwargs):
    """
    Solve the root finding problem by approximating the inverse of Jacobian to be a constant scalar.

    Keyword arguments
    -----------------
    alpha: float or None
        The initial guess of inverse Jacobian is ``-alpha * I``.
    """
    return _nonlin_solver(fcn, x0, params, jacobian=_LinearMixing(alpha=alpha), **kwargs)


for _f in (newton, broyden1, broyden2, linearmixing):
    _f.__doc__ += _nonlin_solver.__doc__
