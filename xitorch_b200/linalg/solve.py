"""
`solve` -- public functional + autograd boundary (boundary B1 of SURVEY.md 8b).

Same call signature, default-method rules, error behaviour and analytic backward as the
reference (/root/reference/xitorch/linalg/solve.py:13-222):

  * asserts on shapes / Hermitian M / `_getparamnames` when grad is enabled (:70-88)
  * `method=None` -> "exactsolve" for dense operators or n<=5, else "cg" (Hermitian) /
    "bicgstab" (:96-104)
  * forward short-circuits an all-zero B (:139-141), dispatches `method` (str or callable)
    inside `torch.autograd.Function.forward` with the operator parameters swapped in (:143-153)
  * backward solves the adjoint system `(A - E M)^H v = grad_x` by re-entering `solve` with
    `bck_options` (:178-184), then obtains parameter gradients from `-A.mm(x)` (:188-195),
    `grad_E` (:198-205) and the M-parameter gradients (:208-219).

The Krylov methods ("cg", "bicgstab", "gmres") are the B200-native fused-kernel
implementations in `xitorch_b200._impls.solve`.
"""
import warnings
from typing import Any, Callable, Mapping, Optional, Union

import torch

from xitorch_b200._utils import assert_runtime, get_method, merged_options, null_context, bcast_dims
from xitorch_b200.debug import is_debug_enabled
from xitorch_b200.linop import LinearOperator, MatrixLinearOperator
from xitorch_b200._impls import solve as _impl

__all__ = ["solve"]


def _solve_methods():
    return {
        "custom_exactsolve": _impl.custom_exactsolve,
        "cg": _impl.cg,
        "bicgstab": _impl.bicgstab,
        "gmres": _impl.gmres,
        "broyden1": _impl.broyden1_solve,
    }


def solve(A: LinearOperator, B: torch.Tensor, E: Optional[torch.Tensor] = None,
          M: Optional[LinearOperator] = None,
          bck_options: Mapping[str, Any] = {},
          method: Union[str, Callable, None] = None,
          **fwd_options) -> torch.Tensor:
    r"""Solve :math:`\mathbf{AX - MXE = B}` for ``X``.

    A: LinearOperator ``(*BA, nr, nr)``; B: ``(*BB, nr, ncols)``; E: ``(*BE, ncols)`` or None;
    M: Hermitian LinearOperator ``(*BM, nr, nr)`` or None (ignored, with a warning, when E is None).
    ``method``: "cg" | "bicgstab" | "gmres" | "exactsolve" | callable ``fcn(A, B, E, M, **opts)``.
    ``bck_options``: ``method`` + options of the adjoint solve used in backward.
    Returns ``X`` of shape ``(*BABEM, nr, ncols)``.
    """
    assert_runtime(A.shape[-1] == A.shape[-2], "The linear operator A must have a square shape")
    assert_runtime(A.shape[-1] == B.shape[-2],
                   "Mismatch shape of A & B (A: %s, B: %s)" % (tuple(A.shape), tuple(B.shape)))
    assert_runtime(not torch.is_grad_enabled() or A.is_getparamnames_implemented,
                   "The _getparamnames(self, prefix) of linear operator A must be "
                   "implemented if using solve with grad enabled")
    if M is not None:
        assert_runtime(M.shape[-1] == M.shape[-2], "The linear operator M must have a square shape")
        assert_runtime(M.shape[-1] == A.shape[-1],
                       "The shape of A & M must match (A: %s, M: %s)" % (tuple(A.shape), tuple(M.shape)))
        assert_runtime(M.is_hermitian, "The linear operator M must be a Hermitian matrix")
        assert_runtime(not torch.is_grad_enabled() or M.is_getparamnames_implemented,
                       "The _getparamnames(self, prefix) of linear operator M must be "
                       "implemented if using solve with grad enabled")
    if E is not None:
        assert_runtime(E.shape[-1] == B.shape[-1],
                       "The last dimension of E & B must match (E: %s, B: %s)" % (tuple(E.shape), tuple(B.shape)))
    if E is None and M is not None:
        warnings.warn("M is supplied but will be ignored because E is not supplied")

    if is_debug_enabled():
        A.check()
        if M is not None:
            M.check()

    if method is None:
        dense = isinstance(A, MatrixLinearOperator) and (M is None or isinstance(M, MatrixLinearOperator))
        if dense or A.shape[-1] <= 5:
            method = "exactsolve"
        else:
            hermit = A.is_hermitian and (M is None or M.is_hermitian)
            method = "cg" if hermit else "bicgstab"

    if isinstance(method, str) and method.lower() == "exactsolve":
        return _impl.exactsolve(A, B, E, M)

    params = A.getlinopparams()
    mparams = M.getlinopparams() if M is not None else []
    needs_grad = torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad for t in (B, E, *params, *mparams))
    if not needs_grad:
        # nothing to differentiate: same result without the autograd-function round trip (host time per call)
        with torch.no_grad():            # methods always run without grad, as inside the autograd function
            return _forward(A, B, E, M, method, merged_options({}, fwd_options))
    return _SolveFunction.apply(A, B, E, M, method, fwd_options, bck_options,
                                len(params), *params, *mparams)


def _forward(A, B, E, M, method, config):
    fcn = get_method("solve", _solve_methods(), method)
    # a zero right-hand side is answered without calling the method (reference solve.py:176-180).  The library's own
    # Krylov methods make that test themselves (|B| <= atol, one reduction) -- no second pass and synchronisation here
    if getattr(fcn, "_checks_zero_rhs", False) or not torch.all(B == 0):
        return fcn(A, B, E, M, **config)
    dims = (*_impl.get_batchdims(A, B, E, M), *B.shape[-2:])
    return torch.zeros(dims, dtype=B.dtype, device=B.device)


class _SolveFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, E, M, method, fwd_options, bck_options, na, *all_params):
        params, mparams = all_params[:na], all_params[na:]
        config = merged_options({}, fwd_options)
        ctx.bck_config = merged_options({}, bck_options)

        with A.uselinopparams(*params), (M.uselinopparams(*mparams) if M is not None else null_context()):
            x = _forward(A, B, E, M, method, config)

        ctx.e_is_none = E is None
        ctx.A, ctx.M, ctx.na = A, M, na
        if ctx.e_is_none:
            ctx.save_for_backward(x, *all_params)
        else:
            ctx.save_for_backward(x, E, *all_params)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        x = ctx.saved_tensors[0]
        first = 1 if ctx.e_is_none else 2
        all_params = ctx.saved_tensors[first:]
        params, mparams = all_params[:ctx.na], all_params[ctx.na:]
        E = None if ctx.e_is_none else ctx.saved_tensors[1]
        A, M = ctx.A, ctx.M

        # adjoint solve: (A - E M)^H v = grad_x   (this is also grad_B)
        with A.uselinopparams(*params), (M.uselinopparams(*mparams) if M is not None else null_context()):
            AT = A.H
            MT = M.H if M is not None else None
            Ec = E.conj() if E is not None else None
            v = solve(AT, grad_x, Ec, MT, bck_options=ctx.bck_config, **ctx.bck_config)
        grad_B = v

        # parameter gradients of A through  -A x
        with torch.enable_grad():
            params = [p.clone().requires_grad_() for p in params]
            with A.uselinopparams(*params):
                loss = -A.mm(x)
        grad_params = torch.autograd.grad((loss,), params, grad_outputs=(v,),
                                          create_graph=torch.is_grad_enabled(), allow_unused=True)

        grad_E = None
        if E is not None:
            if M is None:
                Mx = x
            else:
                with M.uselinopparams(*mparams):
                    Mx = M.mm(x)
            grad_E = torch.einsum("...rc,...rc->...c", v, Mx.conj())

        grad_mparams = []
        if M is not None and E is not None:
            with torch.enable_grad():
                mparams = [p.clone().requires_grad_() for p in mparams]
                lx = x * E.unsqueeze(-2)
                with M.uselinopparams(*mparams):
                    mloss = M.mm(lx)
            grad_mparams = torch.autograd.grad((mloss,), mparams, grad_outputs=(v,),
                                               create_graph=torch.is_grad_enabled(), allow_unused=True)

        return (None, grad_B, grad_E, None, None, None, None, None, *grad_params, *grad_mparams)
