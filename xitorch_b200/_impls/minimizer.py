"""
First-order minimizers behind `xitorch_b200.optimize.minimize(method="gd" | "adam")` (contract:
/root/reference/xitorch/_impls/optimize/minimizer.py:5-211).  Not on the Krylov hot path -- elementwise updates
around the user's ``fcn`` -- provided so that `minimize` is a complete drop-in; only its backward (the adjoint solve
with the Hessian operator) reaches the CUDA solvers.

``fcn(x, *params)`` returns ``(f, df/dx)``.  Both methods stop as soon as ANY of the four criteria holds (after the
first iteration) and, when none ever held, warn and return the iterate with the lowest ``f`` seen.
"""
import warnings
from typing import Callable, List, Optional

import torch

from xitorch_b200._impls.rootsolver import _host_scalars

__all__ = ["gd", "adam"]


class _Progress(object):
    """OR-combined stopping test plus best-iterate bookkeeping shared by the minimizers"""

    def __init__(self, f_tol: float, f_rtol: float, x_tol: float, x_rtol: float, verbose: bool):
        self.f_tol, self.f_rtol, self.x_tol, self.x_rtol = f_tol, f_rtol, x_tol, x_rtol
        self.verbose = verbose
        self.converged = False
        self.last_iter = -1
        self.best = dict(f=float("inf"), x=None, dx=float("inf"), df=float("inf"))

    def step(self, i: int, x_new: torch.Tensor, x_old: torch.Tensor, f: torch.Tensor, f_old: torch.Tensor) -> bool:
        # one device -> host transfer per iteration; the differences are formed in the tensors' dtype as before
        dx, xnorm, df, fval = (float(v) for v in _host_scalars(
            (x_old - x_new).detach().norm(), x_old.detach().norm(), (f_old - f).detach().abs(), f.detach()))
        hit = (dx < self.x_tol) or (dx < self.x_rtol * xnorm) or \
              (df < self.f_tol) or (df < self.f_rtol * abs(fval))
        if self.verbose:
            if i == 0:
                print("   #:             f |        dx,        df")
            if hit:
                print("Finish with convergence")
            if i == 0 or (i + 1) % 10 == 0 or hit:
                print("%4d: %.6e | %.3e, %.3e" % (i + 1, fval, dx, df))
        stop = hit and i > 0
        self.converged = self.converged or stop
        self.last_iter = max(self.last_iter, i)
        if fval < self.best["f"]:
            self.best = dict(f=fval, x=x_old, dx=dx, df=df)
        return stop

    def result(self, x: torch.Tensor) -> torch.Tensor:
        if self.converged or self.last_iter < 0:        # maxiter == 0 is used to wrap only the backward
            return x
        warnings.warn("The minimizer does not converge after %d iterations. Best |dx|=%.4e, |df|=%.4e, f=%.4e"
                      % (self.last_iter, self.best["dx"], self.best["df"], self.best["f"]))
        return self.best["x"]


def gd(fcn: Callable[..., torch.Tensor], x0: torch.Tensor, params: List, step: float = 1e-3, gamma: float = 0.9,
       maxiter: int = 1000, f_tol: float = 0.0, f_rtol: float = 1e-8, x_tol: float = 0.0, x_rtol: float = 1e-8,
       verbose=False, **unused):
    r"""
    Gradient descent with momentum: :math:`v \leftarrow \gamma v - \eta \nabla f(x)`, :math:`x \leftarrow x + v`.

    Keyword arguments
    -----------------
    step: float
        The step size :math:`\eta`.
    gamma: float
        The momentum factor :math:`\gamma`.
    maxiter: int
        Maximum number of iterations.
    f_tol, f_rtol: float
        Absolute / relative tolerance on the change of ``f``.
    x_tol, x_rtol: float
        Absolute / relative tolerance on the norm of the change of ``x``.
    """
    x = x0.clone()
    prog = _Progress(f_tol, f_rtol, x_tol, x_rtol, verbose)
    f_old = torch.zeros((), dtype=x0.dtype, device=x0.device)
    vel = torch.zeros_like(x)
    for i in range(maxiter):
        f, g = fcn(x, *params)
        vel = (gamma * vel - step * g).detach()
        x_old = x.detach()
        x = (x_old + vel).detach()
        if prog.step(i, x, x_old, f, f_old):
            break
        f_old = f
    return prog.result(x)


def adam(fcn: Callable[..., torch.Tensor], x0: torch.Tensor, params: List, step: float = 1e-3, beta1: float = 0.9,
         beta2: float = 0.999, eps: float = 1e-8, maxiter: int = 1000, f_tol: float = 0.0, f_rtol: float = 1e-8,
         x_tol: float = 0.0, x_rtol: float = 1e-8, verbose=False, **unused):
    r"""
    Adam (Kingma & Ba 2015) with bias-corrected moments.

    Keyword arguments
    -----------------
    step: float
        The step size.
    beta1, beta2: float
        Exponential decay rates of the first / second moment estimates.
    eps: float
        Small number to prevent division by 0.
    maxiter: int
        Maximum number of iterations.
    f_tol, f_rtol: float
        Absolute / relative tolerance on the change of ``f``.
    x_tol, x_rtol: float
        Absolute / relative tolerance on the norm of the change of ``x``.
    """
    x = x0.clone()
    prog = _Progress(f_tol, f_rtol, x_tol, x_rtol, verbose)
    f_old = torch.zeros((), dtype=x0.dtype, device=x0.device)
    m1 = torch.zeros_like(x)
    m2 = torch.zeros_like(x)
    b1t, b2t = beta1, beta2
    for i in range(maxiter):
        f, g = fcn(x, *params)
        f, g = f.detach(), g.detach()
        m1 = beta1 * m1 + (1 - beta1) * g
        m2 = beta2 * m2 + (1 - beta2) * g ** 2
        upd = (m1 / (1 - b1t)) / ((m2 / (1 - b2t)) ** 0.5 + eps)
        b1t, b2t = b1t * beta1, b2t * beta2
        x_old = x.detach()
        x = (x_old - step * upd).detach()
        if prog.step(i, x, x_old, f, f_old):
            break
        f_old = f
    return prog.result(x)
