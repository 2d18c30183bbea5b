"""
TEST INFRASTRUCTURE (see oracle/__init__.py) -- CPU restatement of the reference's Broyden rootfinder and of the
gradient its backward produces.

  * broyden1 / _nonlin_solver   /root/reference/xitorch/_impls/optimize/root/rootsolver.py:15-149, 176-204
  * BroydenFirst, LowRankMatrix /root/reference/xitorch/_impls/optimize/root/_jacobian.py:51-199 (the inverse Jacobian
                                 as a Python list of rank-1 terms, exactly the reference's data structure)
  * line search                 rootsolver.py:272-357
  * implicit gradient           /root/reference/xitorch/optimize/rootfinder.py:331-366, evaluated here with a DENSE
                                 Jacobian and a direct solve (the exact value the reference's Krylov backward approximates)
"""
import torch


def broyden1_root(fcn, x0, params=(), alpha=None, maxiter=None, f_tol=1e-6, x_tol=1e-6, return_info=False):
    xshape = x0.shape
    func = lambda x: fcn(x.reshape(xshape), *params).reshape(-1)      # noqa: E731
    x = x0.reshape(-1)
    y = func(x)
    y_norm = y.norm()
    nfev = 1
    if maxiter is None:
        maxiter = 100 * (x.numel() + 1)
    if y_norm == 0:
        return (x.reshape(xshape), {"nfev": nfev, "niter": 0}) if return_info else x.reshape(xshape)
    if alpha is None:
        alpha = 0.5 * torch.max(x.norm(), torch.ones_like(y_norm)) / y_norm
    cns, dns = [], []
    g_alpha = -alpha

    def gm_mv(v):
        res = g_alpha * v
        for c, d in zip(cns, dns):
            res = res + c * torch.dot(d, v)
        return res

    def gm_rmv(v):
        res = g_alpha * v
        for c, d in zip(cns, dns):
            res = res + d * torch.dot(c, v)
        return res

    x_prev, y_prev = x, y
    converged = False
    best = (y_norm, x)
    niter = 0
    for i in range(maxiter):
        niter = i + 1
        dx = -gm_mv(y)
        # Armijo backtracking on phi(s) = |f(x + s dx)|^2 (rootsolver.py:272-357)
        cache = {"s": 0, "y": y, "phi": y.norm() ** 2}

        def phi(s, cache=cache, x=x, dx=dx):
            if s == cache["s"]:
                return cache["phi"]
            v = func(x + s * dx)
            cache.update(s=s, y=v, phi=torch.dot(v, v))
            cache["n"] = cache.get("n", 0) + 1
            return cache["phi"]

        phi0 = cache["phi"]
        s = _armijo(phi, phi0, -phi0, amin=1e-2)
        if s is None:
            s = 1.0
        xnew = x + s * dx
        ynew = cache["y"] if s == cache["s"] else func(xnew)
        nfev += cache.get("n", 0) + (0 if s == cache["s"] else 1)
        y_norm_new = ynew.norm()
        if y_norm_new < best[0]:
            best = (y_norm_new, xnew)
        # Broyden's first update of the inverse Jacobian (_jacobian.py:103-119)
        dy_, dx_ = ynew - y_prev, xnew - x_prev
        v = gm_rmv(dx_)
        c = dx_ - gm_mv(dy_)
        d = v / torch.dot(dy_, v)
        cns.append(c)
        dns.append(d)
        x_prev, y_prev = xnew, ynew
        if dx.norm() < x_tol and y_norm_new < f_tol:
            converged = True
            break                         # the reference returns the iterate before this last step
        x, y, y_norm = xnew, ynew, y_norm_new
    if not converged:
        x = best[1]
    out = x.reshape(xshape)
    return (out, {"nfev": nfev, "niter": niter, "converged": converged}) if return_info else out


def _armijo(phi, phi0, derphi0, c1=1e-4, alpha0=1, amin=0, max_niter=20):
    phi_a0 = phi(alpha0)
    if phi_a0 <= phi0 + c1 * alpha0 * derphi0:
        return alpha0
    alpha1 = -derphi0 * alpha0 ** 2 / 2.0 / (phi_a0 - phi0 - derphi0 * alpha0)
    phi_a1 = phi(alpha1)
    if phi_a1 <= phi0 + c1 * alpha1 * derphi0:
        return alpha1
    niter = 0
    alpha2 = alpha1
    while alpha1 > amin and niter < max_niter:
        factor = alpha0 ** 2 * alpha1 ** 2 * (alpha1 - alpha0)
        a = (alpha0 ** 2 * (phi_a1 - phi0 - derphi0 * alpha1) - alpha1 ** 2 * (phi_a0 - phi0 - derphi0 * alpha0)) / factor
        b = (-alpha0 ** 3 * (phi_a1 - phi0 - derphi0 * alpha1) + alpha1 ** 3 * (phi_a0 - phi0 - derphi0 * alpha0)) / factor
        alpha2 = (-b + torch.sqrt(torch.abs(b ** 2 - 3 * a * derphi0))) / (3.0 * a)
        phi_a2 = phi(alpha2)
        if phi_a2 <= phi0 + c1 * alpha2 * derphi0:
            return alpha2
        if (alpha1 - alpha2) > alpha1 / 2.0 or (1 - alpha2 / alpha1) < 0.96:
            alpha2 = alpha1 / 2.0
        alpha0, alpha1, phi_a0, phi_a1 = alpha1, alpha2, phi_a1, phi_a2
        niter += 1
    return alpha2 if niter == max_niter else None


def implicit_grad_dense(fcn, y, params, grad_y):
    """exact parameter gradients of L(y*(theta)) with dL/dy* = grad_y:  -(df/dtheta)^T (df/dy)^-T grad_y."""
    y = y.detach()
    J = torch.autograd.functional.jacobian(lambda yy: fcn(yy, *params).reshape(-1), y).reshape(y.numel(), y.numel())
    g = torch.linalg.solve(J.t(), -grad_y.reshape(-1, 1)).reshape(y.shape)
    prm = [p.detach().clone().requires_grad_() for p in params]
    out = fcn(y, *prm)
    return torch.autograd.grad(out, prm, grad_outputs=g)
