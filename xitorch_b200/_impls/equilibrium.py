"""
Anderson acceleration behind `xitorch_b200.optimize.equilibrium(method="anderson_acc")` (contract:
/root/reference/xitorch/_impls/optimize/equilibrium.py:9-135; Walker & Ni, SIAM J. Numer. Anal. 49, 1715).
Not on the Krylov hot path (a ring of ``msize`` iterates and one tiny bordered solve per step); provided so that
`equilibrium` is a complete drop-in.  Its backward is the rootfinder adjoint solve, which is on the hot path.
"""
import math
import warnings
from typing import Callable, List

import torch

from xitorch_b200._utils import ConvergenceWarning
from xitorch_b200._impls.rootsolver import TerminationCondition

__all__ = ["anderson_acc"]


def anderson_acc(fcn: Callable[..., torch.Tensor], x0: torch.Tensor, params: List, feat_ndims: int = 1,
                 msize: int = 5, beta: float = 1.0, lmbda: float = 1e-4,
                 maxiter=None, f_tol=None, f_rtol=None, x_tol=None, x_rtol=None, custom_terminator=None,
                 verbose: bool = False) -> torch.Tensor:
    """
    Solve the fixed-point problem ``x = fcn(x, *params)`` with Anderson acceleration.

    Keyword arguments
    -----------------
    feat_ndims: int
        The number of trailing dimensions that are features (the leading ones are batch dimensions).
    msize: int
        How many previous iterates enter the mixing.
    beta: float
        Damping (< 1) or over-relaxation (> 1) parameter.
    lmbda: float
        Tikhonov shift that keeps the small Gram matrix invertible.
    maxiter: int or None
        Maximum number of iterations (``100 * (nfeat + 1)`` if None).
    f_tol, f_rtol: float or None
        Absolute / relative tolerance of ``|f(x) - x|``.
    x_tol, x_rtol: float or None
        Absolute / relative tolerance of the step norm.
    verbose: bool
        Options for verbosity
    """
    batch = x0.shape[:x0.ndim - feat_ndims]
    feat = x0.shape[x0.ndim - feat_ndims:]
    nfeat = math.prod(feat)
    if maxiter is None:
        maxiter = 100 * (nfeat + 1)
    opts = dict(dtype=x0.dtype, device=x0.device)

    def apply(flat):
        return fcn(flat.reshape(*batch, *feat), *params).reshape(*batch, nfeat)

    # two plain fixed-point steps fill the first two slots of the history
    xs = torch.zeros((*batch, msize, nfeat), **opts)
    fs = torch.zeros((*batch, msize, nfeat), **opts)
    xk = x0.reshape(*batch, nfeat)
    fk = apply(xk)
    xs[..., 0, :], fs[..., 0, :] = xk, fk
    xk = fk
    fk = apply(xk)
    xs[..., 1, :], fs[..., 1, :] = xk, fk

    dev0 = (fk - xk).norm()
    term = custom_terminator if custom_terminator is not None else TerminationCondition(f_tol, f_rtol, dev0, x_tol, x_rtol)
    if dev0 == 0:
        return x0

    # bordered system  [[0, 1^T], [1, G G^T + lmbda I]] [mu; alpha] = [1; 0]   (sum(alpha) = 1)
    bord = torch.zeros((*batch, msize + 1, msize + 1), **opts)
    bord[..., 0, 1:] = 1.0
    bord[..., 1:, 0] = 1.0
    rhs = torch.zeros((*batch, msize + 1, 1), **opts)
    rhs[..., 0, :] = 1.0

    done = False
    for k in range(2, maxiter):
        h = min(k, msize)
        resid = fs[..., :h, :] - xs[..., :h, :]
        bord[..., 1:h + 1, 1:h + 1] = torch.einsum("...nf,...mf->...nm", resid, resid) + lmbda * torch.eye(h, **opts)
        alpha = torch.linalg.solve(bord[..., :h + 1, :h + 1], rhs[..., :h + 1, :])[..., 1:h + 1, 0]
        xnew = beta * torch.einsum("...n,...nf->...f", alpha, fs[..., :h, :]) + \
            (1 - beta) * torch.einsum("...n,...nf->...f", alpha, xs[..., :h, :])
        fnew = apply(xnew)
        xs[..., k % msize, :], fs[..., k % msize, :] = xnew, fnew
        done = term.check(xnew, fnew - xnew, xnew - xk)
        if verbose and (k < 10 or k % 10 == 0 or done):
            print("%6d: |dx|=%.3e, |f-x|=%.3e" % (k, (xnew - xk).norm(), (fnew - xnew).norm()))
        xk = xnew
        if done:
            break
    if not done:
        warnings.warn(ConvergenceWarning("The rootfinder does not converge after %d iterations." % maxiter))
    return xk.reshape(*batch, *feat)
