"""
`rootfinder` / `equilibrium` -- public functionals + autograd boundary (BASELINE config 4; reference:
/root/reference/xitorch/optimize/rootfinder.py:35-366).

Same signatures, default method ("broyden1") and analytic backward as the reference: the backward builds the
matrix-free Jacobian ``df/dy`` at the root (`xitorch_b200.grad.jac`), solves the adjoint system
``(df/dy)^H g = -grad_y`` by re-entering `linalg.solve` with `bck_options` (rootfinder.py:346-348) and pulls the
parameter gradients through one more evaluation of ``fcn`` (:352-362).  With CUDA tensors that adjoint solve runs in
the fused Krylov kernels (default for n > 5: "bicgstab") with the Jacobian applied through the operator callback of
the C ABI.

`fcn` may be a function or a method of an `EditableModule` / `torch.nn.Module` (`get_pure_function`): the tensors hidden
in the object travel through `torch.autograd.Function.apply` next to the explicit parameters, exactly as in the
reference, so gradients reach them.

`minimize` is `rootfinder` on the gradient (methods "newton" / "broyden*" / "linearmixing") or a first-order descent
("gd", "adam"); its backward is the same adjoint solve with the Hessian operator.
"""
import inspect
from typing import Any, Callable, Mapping, Sequence, Union

import torch

from xitorch_b200._utils import get_method
from xitorch_b200._impls.rootsolver import newton, broyden1, broyden2, linearmixing
from xitorch_b200._impls.equilibrium import anderson_acc
from xitorch_b200._impls.minimizer import gd, adam
from xitorch_b200.debug import is_debug_enabled
from xitorch_b200.editable_module import EditableModule
from xitorch_b200.grad import jac
from xitorch_b200.pure_function import get_pure_function, make_sibling
from xitorch_b200.linalg.solve import solve

__all__ = ["rootfinder", "equilibrium", "minimize"]

_RF_METHODS = {"newton": newton, "broyden1": broyden1, "broyden2": broyden2, "linearmixing": linearmixing}
_EQUIL_METHODS = {"anderson_acc": anderson_acc}     # solvers that take y = f(y) itself rather than the residual
_OPT_METHODS = {"gd": gd, "adam": adam}
_METHOD_TABLES = {"rootfinder": _RF_METHODS, "equilibrium": _EQUIL_METHODS, "minimizer": _OPT_METHODS}


def _entry_checks(fcn, y0, params):
    if is_debug_enabled() and inspect.ismethod(fcn) and isinstance(fcn.__self__, EditableModule):
        fcn.__self__.assertparams(fcn, y0, *params)


def rootfinder(fcn: Callable[..., torch.Tensor], y0: torch.Tensor, params: Sequence[Any] = [],
               bck_options: Mapping[str, Any] = {}, method: Union[str, Callable, None] = None,
               **fwd_options) -> torch.Tensor:
    r"""Solve :math:`\mathbf{0} = \mathbf{f}(\mathbf{y}, \theta)` for ``y`` (shape of ``y0``).

    ``fcn(y, *params)`` returns a tensor of the shape of ``y``; ``method``: "broyden1" (default) | "broyden2" |
    "linearmixing" | callable ``fcn_method(fcn, y0, params, **opts)``; ``bck_options``: options of the adjoint
    `linalg.solve` in backward (``method`` among them); ``**fwd_options``: options of the method.
    """
    _entry_checks(fcn, y0, params)
    pfunc = get_pure_function(fcn)
    fwd_options["method"] = "broyden1" if method is None else method
    return _RootFinder.apply(pfunc, y0, pfunc, "rootfinder", fwd_options, bck_options, len(params), *params,
                             *pfunc.objparams())


def equilibrium(fcn: Callable[..., torch.Tensor], y0: torch.Tensor, params: Sequence[Any] = [],
                bck_options: Mapping[str, Any] = {}, method: Union[str, Callable, None] = None,
                **fwd_options) -> torch.Tensor:
    r"""Solve :math:`\mathbf{y} = \mathbf{f}(\mathbf{y}, \theta)` for ``y``: a rootfinder method on ``y - f(y)``, or
    "anderson_acc" on the fixed-point map itself."""
    _entry_checks(fcn, y0, params)
    pfunc = get_pure_function(fcn)

    @make_sibling(pfunc)
    def resid(y, *prm):
        return y - pfunc(y, *prm)

    method = "broyden1" if method is None else method
    fwd_options["method"] = method
    fixed_point = isinstance(method, str) and method.lower() in _EQUIL_METHODS
    return _RootFinder.apply(resid, y0, pfunc if fixed_point else resid, "equilibrium" if fixed_point else "rootfinder",
                             fwd_options, bck_options, len(params), *params, *pfunc.objparams())


def minimize(fcn: Callable[..., torch.Tensor], y0: torch.Tensor, params: Sequence[Any] = [],
             bck_options: Mapping[str, Any] = {}, method: Union[str, Callable, None] = None,
             **fwd_options) -> torch.Tensor:
    r"""Solve :math:`\mathbf{y^*} = \arg\min_\mathbf{y} f(\mathbf{y}, \theta)` for the one-element output ``fcn``.

    ``method``: a rootfinder method applied to the gradient ("broyden1" default, "broyden2", "newton",
    "linearmixing"), a minimizer ("gd", "adam") or a callable ``fcn_method(fcn, y0, params, **opts)`` receiving a
    function that returns ``(f, df/dy)``.
    """
    assert not torch.is_complex(y0), "complex number is not supported on xitorch.optimize.rootfinder at the moment"
    _entry_checks(fcn, y0, params)
    pfunc = get_pure_function(fcn)
    method = "broyden1" if method is None else method
    fwd_options["method"] = method

    @make_sibling(pfunc)
    def value_and_grad(y, *prm):
        with torch.enable_grad():
            y1 = y.clone().requires_grad_()
            z = pfunc(y1, *prm)
        (g,) = torch.autograd.grad(z, (y1,), retain_graph=True, create_graph=torch.is_grad_enabled())
        return z, g

    @make_sibling(value_and_grad)
    def grad_only(y, *prm):
        return value_and_grad(y, *prm)[1]

    # the root solvers walk against the function's output, i.e. they take the gradient alone
    as_root = isinstance(method, str) and method.lower() in _RF_METHODS
    return _RootFinder.apply(grad_only, y0, grad_only if as_root else value_and_grad,
                             "rootfinder" if as_root else "minimizer", fwd_options, bck_options, len(params),
                             *params, *pfunc.objparams())


class _RootFinder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fcn, y0, fwd_fcn, alg_type, options, bck_options, nparams, *allparams):
        # fcn: what has to vanish at the solution (used by backward); fwd_fcn: what the forward method iterates on
        config = dict(options)
        ctx.bck_options = dict(bck_options)
        params, objparams = allparams[:nparams], allparams[nparams:]
        method = config.pop("method")
        method_fcn = get_method(alg_type, _METHOD_TABLES[alg_type], method)
        with fwd_fcn.useobjparams(objparams):
            y = method_fcn(fwd_fcn, y0, params, **config)
        ctx.fcn = fcn
        ctx.nparams = nparams
        ctx.is_tensor = [isinstance(p, torch.Tensor) for p in allparams]
        ctx.nontensors = [p for p in allparams if not isinstance(p, torch.Tensor)]
        ctx.save_for_backward(y, *[p for p in allparams if isinstance(p, torch.Tensor)])
        return y

    @staticmethod
    def backward(ctx, grad_yout):
        yout = ctx.saved_tensors[0]
        tensors = list(ctx.saved_tensors[1:])
        fcn, nparams = ctx.fcn, ctx.nparams

        def rebuild(tens):
            it, jt = iter(tens), iter(ctx.nontensors)
            return [next(it) if flag else next(jt) for flag in ctx.is_tensor]

        allparams = rebuild(tensors)
        params, objparams = allparams[:nparams], allparams[nparams:]
        with fcn.useobjparams(objparams):
            # dL/df: adjoint solve with the matrix-free Jacobian at the root
            # (the saved output carries the graph of this very function: second derivatives flow through it)
            y_lin = yout if yout.requires_grad else yout.detach().requires_grad_()
            jac_dfdy = jac(fcn, params=(y_lin, *params), idxs=[0])[0]
            gyfcn = solve(A=jac_dfdy.H, B=-grad_yout.reshape(-1, 1), bck_options=ctx.bck_options, **ctx.bck_options)
            gyfcn = gyfcn.reshape(grad_yout.shape)
            # gradients of the (explicit and object) parameters through one more evaluation of fcn
            with torch.enable_grad():
                copies = [p.clone().requires_grad_() for p in tensors]
                allcopy = rebuild(copies)
                with fcn.useobjparams(allcopy[nparams:]):
                    yfcn = fcn(yout, *allcopy[:nparams])
            grads = torch.autograd.grad(yfcn, copies, grad_outputs=gyfcn, create_graph=torch.is_grad_enabled(),
                                        allow_unused=True) if copies else ()
        it = iter(grads)
        grad_params = [next(it) if flag else None for flag in ctx.is_tensor]
        return (None, None, None, None, None, None, None, *grad_params)
