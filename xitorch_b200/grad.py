"""
`jac` / `hess` -- matrix-free Jacobian / Hessian LinearOperators built on autograd (boundary B2 of SURVEY.md 8b;
reference: /root/reference/xitorch/grad/jachess.py:9-208).

These are the operators the rootfinder backward hands to `linalg.solve` (rootfinder.py:346-348): `mv` is the
double-backward trick (one JVP per operator application), `rmv` one VJP.  With `xitorch_b200` the Krylov loop around
them runs in the CUDA solver kernels and only these products are autograd calls (`xt_solve_args.apply`).

`fcn` may be a function, a `torch.jit` script function, or a method of an `EditableModule` / `torch.nn.Module`: it is
wrapped by `get_pure_function`, and the tensors hidden in the object (`objparams`) are operator parameters like the
explicit ones, so derivatives with respect to them flow through `solve` / `symeig` backward.
"""
from typing import Any, Callable, List, Sequence, Union

import torch

from xitorch_b200.linop import LinearOperator
from xitorch_b200.pure_function import get_pure_function, make_sibling

__all__ = ["jac", "hess"]


def _tensor_params(params: Sequence[Any]) -> List[torch.Tensor]:
    return [p for p in params if isinstance(p, torch.Tensor)]


def _resolve_idxs(idxs, params) -> List[int]:
    if idxs is None:
        idxs = [i for i, t in enumerate(params) if isinstance(t, torch.Tensor) and t.requires_grad]
    elif isinstance(idxs, int):
        idxs = [idxs]
    for i in idxs:
        if not (isinstance(params[i], torch.Tensor) and params[i].requires_grad):
            raise TypeError("The %d-th element (0-based) must be a tensor which requires grad" % i)
    return list(idxs)


def jac(fcn: Callable[..., torch.Tensor], params: Sequence[Any],
        idxs: Union[None, int, Sequence[int]] = None) -> Union[LinearOperator, List[LinearOperator]]:
    """LinearOperator(s) of shape ``(nout, nin)`` acting as the Jacobian of ``fcn`` w.r.t. ``params[idx]``."""
    lst = _resolve_idxs(idxs, params)
    pfcn = get_pure_function(fcn)
    res = [_Jac(pfcn, params, i) for i in lst]
    return res[0] if isinstance(idxs, int) else res


def hess(fcn: Callable[..., torch.Tensor], params: Sequence[Any],
         idxs: Union[None, int, Sequence[int]] = None) -> Union[LinearOperator, List[LinearOperator]]:
    """LinearOperator(s) of shape ``(nin, nin)`` acting as the Hessian of the scalar ``fcn`` w.r.t. ``params[idx]``."""
    lst = _resolve_idxs(idxs, params)
    pfcn = get_pure_function(fcn)

    def grad_of(idx):
        @make_sibling(pfcn)
        def gfcn(*prm):
            with torch.enable_grad():
                z = pfcn(*prm)
            (g,) = torch.autograd.grad(z, (prm[idx],), retain_graph=True, create_graph=torch.is_grad_enabled())
            return g
        return gfcn

    res = [_Jac(grad_of(i), params, i, is_hermitian=True) for i in lst]
    return res[0] if isinstance(idxs, int) else res


def _tie(out: torch.Tensor, tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    # keeps `out` attached to the graph of every parameter (needed by create_graph consumers); without a graph
    # (the operator callback of the CUDA solvers runs under no_grad) it would only be launches that add zero
    if not torch.is_grad_enabled():
        return out
    for t in tensors:
        out = out + t.reshape(-1)[0] * 0
    return out


class _Jac(LinearOperator):
    """d fcn(*params) / d params[idx] as an operator; the linearisation point is rebuilt when the operator's
    parameters are swapped (`uselinopparams`)."""

    def __init__(self, fcn, params: Sequence[Any], idx: int, is_hermitian: bool = False) -> None:
        self.fcn = get_pure_function(fcn)
        self.objparams = list(self.fcn.objparams())
        self.idx = idx
        self.params = list(params)
        self._tensor_pos = [i for i, p in enumerate(self.params) if isinstance(p, torch.Tensor)]
        self.params_tensor = [self.params[i] for i in self._tensor_pos]
        self._linearise()
        yparam = self.params[idx]
        super().__init__(shape=(self.yout.numel(), yparam.numel()), is_hermitian=is_hermitian,
                         dtype=yparam.dtype, device=yparam.device, _suppress_hermit_warning=True)
        self.inshape, self.outshape = yparam.shape, self.yout.shape
        self.nin, self.nout = yparam.numel(), self.yout.numel()

    def _linearise(self):
        for pos, t in zip(self._tensor_pos, self.params_tensor):
            self.params[pos] = t
        self.yparam = self.params[self.idx]
        with torch.enable_grad(), self.fcn.useobjparams(self.objparams):
            self.yout = self.fcn(*self.params)
            self.v = torch.ones_like(self.yout).requires_grad_()
            (self.dfdy,) = torch.autograd.grad(self.yout, (self.yparam,), grad_outputs=self.v, create_graph=True)
        self._ids = [id(t) for t in self.params_tensor] + [id(t) for t in self.objparams]

    def _refresh(self):
        if [id(t) for t in self.params_tensor] + [id(t) for t in self.objparams] != self._ids:
            self._linearise()

    def _getparamnames(self, prefix: str = "") -> List[str]:
        return [prefix + ("params_tensor[%d]" % i) for i in range(len(self.params_tensor))] + \
               [prefix + ("objparams[%d]" % i) for i in range(len(self.objparams))]

    def _mv(self, gy: torch.Tensor) -> torch.Tensor:
        # J g = d/dv <dfdy(v), g>   (dfdy is linear in the dummy cotangent v)
        self._refresh()
        g2 = gy.reshape(-1, self.nin)
        rows = []
        for i in range(g2.shape[0]):
            (r,) = torch.autograd.grad(self.dfdy, (self.v,), grad_outputs=g2[i].reshape(self.inshape),
                                       retain_graph=True, create_graph=torch.is_grad_enabled())
            rows.append(r.reshape(1, self.nout))
        res = torch.cat(rows, dim=0).reshape(*gy.shape[:-1], self.nout)
        return _tie(_tie(res, self.params_tensor), self.objparams)

    def _rmv(self, gout: torch.Tensor) -> torch.Tensor:
        # J^T g: one vector-Jacobian product per vector
        self._refresh()
        g2 = gout.reshape(-1, self.nout)
        rows = []
        for i in range(g2.shape[0]):
            (r,) = torch.autograd.grad(self.yout, (self.yparam,), grad_outputs=g2[i].reshape(self.outshape),
                                       retain_graph=True, create_graph=torch.is_grad_enabled())
            rows.append(r.reshape(1, self.nin))
        res = torch.cat(rows, dim=0).reshape(*gout.shape[:-1], self.nin)
        return _tie(_tie(res, self.params_tensor), self.objparams)
