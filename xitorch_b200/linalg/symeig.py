"""
`symeig` / `lsymeig` / `usymeig` -- public functional + autograd boundary (B1 of SURVEY.md 8b).

Same signature, argument meaning, default-method rule and analytic backward as the reference
(/root/reference/xitorch/linalg/symeig.py:17-144, 252-448):

  * asserts: A (and M) Hermitian, shapes match, `_getparamnames` implemented when grad is on (:103-114)
  * mode "uppermost" -> "uppest"; `method=None` -> "exacteig"; `neig=None` -> all (:115-126)
  * forward dispatches `method` (str or callable) inside `torch.autograd.Function.forward`
    with the operator parameters swapped in (:254-289)
  * backward (:291-402): degeneracy map (:404-414), eigenvalue contribution `grad_evals * evecs`,
    eigenvector contribution through the shifted multi-RHS adjoint solve
    `solve(A, -B, E=evals, M, **bck_options)` (:365-367) with projections `_ortho` (:416-448),
    parameter gradients through `A.mm(evecs)` (and `M.mm(evecs)`).

Methods: "davidson" and "lanczos" are the B200-native implementations
(`xitorch_b200._impls.symeig`); "exacteig"/"custom_exacteig" are the dense `eigh` path.
"""
import warnings
from typing import Any, Callable, Mapping, Optional, Tuple, Union

import torch

from xitorch_b200._utils import (MathWarning, assert_runtime, get_method, merged_options,
                                 null_context, pop_keys)
from xitorch_b200.debug import is_debug_enabled
from xitorch_b200.linop import LinearOperator
from xitorch_b200.linalg.solve import solve
from xitorch_b200._impls import symeig as _impl

__all__ = ["lsymeig", "usymeig", "symeig", "svd"]


def _symeig_methods():
    return {
        "davidson": _impl.davidson,
        "lanczos": _impl.lanczos,
        "custom_exacteig": _impl.custom_exacteig,
    }


def lsymeig(A, neig=None, M=None, bck_options: Mapping[str, Any] = {}, method=None, **fwd_options):
    return symeig(A, neig, "lowest", M, method=method, bck_options=bck_options, **fwd_options)


def usymeig(A, neig=None, M=None, bck_options: Mapping[str, Any] = {}, method=None, **fwd_options):
    return symeig(A, neig, "uppest", M, method=method, bck_options=bck_options, **fwd_options)


def symeig(A: LinearOperator, neig: Optional[int] = None, mode: str = "lowest",
           M: Optional[LinearOperator] = None, bck_options: Mapping[str, Any] = {},
           method: Union[str, Callable, None] = None,
           **fwd_options) -> Tuple[torch.Tensor, torch.Tensor]:
    r"""``neig`` lowest (or uppermost) eigenpairs of :math:`\mathbf{AX = MXE}`.

    Returns ``(evals (*BAM, neig), evecs (*BAM, na, neig))``.
    ``method``: "davidson" | "lanczos" | "exacteig" | callable ``fcn(A, neig, mode, M, **opts)``.
    ``bck_options``: options of the adjoint `solve` plus ``degen_atol`` / ``degen_rtol``.
    """
    assert_runtime(A.is_hermitian, "The linear operator A must be Hermitian")
    assert_runtime(not torch.is_grad_enabled() or A.is_getparamnames_implemented,
                   "The _getparamnames(self, prefix) of linear operator A must be "
                   "implemented if using symeig with grad enabled")
    if M is not None:
        assert_runtime(M.is_hermitian, "The linear operator M must be Hermitian")
        assert_runtime(M.shape[-1] == A.shape[-1],
                       "The shape of A & M must match (A: %s, M: %s)" % (tuple(A.shape), tuple(M.shape)))
        assert_runtime(not torch.is_grad_enabled() or M.is_getparamnames_implemented,
                       "The _getparamnames(self, prefix) of linear operator M must be "
                       "implemented if using symeig with grad enabled")
    mode = mode.lower()
    if mode == "uppermost":
        mode = "uppest"
    if method is None:
        method = "exacteig"
    if neig is None:
        neig = A.shape[-1]

    if is_debug_enabled():
        A.check()
        if M is not None:
            M.check()

    if isinstance(method, str) and method.lower() == "exacteig":
        return _impl.exacteig(A, neig, mode, M)

    grad_on = torch.is_grad_enabled()
    if not grad_on:
        # nothing to differentiate and no mode to switch: straight to the method (host time per call matters: a C2 solve
        # is 2.6 ms on the device)
        return get_method("symeig", _symeig_methods(), method)(A, neig, mode, M, **fwd_options)
    params = A.getlinopparams()
    mparams = M.getlinopparams() if M is not None else []
    if not any(p.requires_grad for p in (*params, *mparams)):
        # nothing to differentiate: same result without the autograd-function round trip (host time per call)
        with torch.no_grad():            # methods always run without grad, as inside the autograd function
            return get_method("symeig", _symeig_methods(), method)(A, neig, mode, M, **fwd_options)
    fwd_options = dict(fwd_options)
    fwd_options["method"] = method
    return _SymeigFunction.apply(A, neig, mode, M, fwd_options, bck_options,
                                 len(params), *params, *mparams)


class _SymeigFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, neig, mode, M, fwd_options, bck_options, na, *amparams):
        params, mparams = amparams[:na], amparams[na:]
        config = merged_options({}, fwd_options)
        ctx.bck_config = merged_options({"degen_atol": None, "degen_rtol": None}, bck_options)
        ctx.bck_alg_config = pop_keys(ctx.bck_config, ["degen_atol", "degen_rtol"])

        method = config.pop("method")
        with A.uselinopparams(*params), (M.uselinopparams(*mparams) if M is not None else null_context()):
            fcn = get_method("symeig", _symeig_methods(), method)
            evals, evecs = fcn(A, neig, mode, M, **config)

        ctx.save_for_backward(evals, evecs, *amparams)
        ctx.na, ctx.A, ctx.M = na, A, M
        return evals, evecs

    @staticmethod
    def backward(ctx, grad_evals, grad_evecs):
        evals, evecs = ctx.saved_tensors[:2]
        amparams = ctx.saved_tensors[2:]
        params, mparams = amparams[:ctx.na], amparams[ctx.na:]
        A, M = ctx.A, ctx.M

        degen_atol = ctx.bck_alg_config["degen_atol"]
        degen_rtol = ctx.bck_alg_config["degen_rtol"]
        eps = torch.finfo(evals.dtype).eps
        if degen_atol is None:
            degen_atol = eps ** 0.6
        if degen_rtol is None:
            degen_rtol = eps ** 0.4

        idx_degen, isdegenerate = None, False
        if degen_atol > 0 or degen_rtol > 0:
            idx_degen, isdegenerate = _check_degen(evals, degen_atol, degen_rtol)
        if not isdegenerate:
            idx_degen = None

        with torch.enable_grad():
            params = [p.clone().requires_grad_() for p in params]
            with A.uselinopparams(*params):
                loss = A.mm(evecs)

        if is_debug_enabled() and isdegenerate:
            xtg = torch.matmul(evecs.transpose(-2, -1).conj(), grad_evecs)
            req1 = idx_degen * (xtg - xtg.transpose(-2, -1).conj())
            reqtol = xtg.abs().max() * grad_evecs.shape[-2] * torch.finfo(grad_evecs.dtype).eps
            if not torch.all(torch.abs(req1) <= reqtol):
                warnings.warn(MathWarning(
                    "Degeneracy appears but the loss function seem to depend strongly on the "
                    "eigenvector. The gradient might be incorrect.\nEigenvalues:\n%s\n"
                    "Degenerate map:\n%s\nRequirements (should be all 0s):\n%s"
                    % (str(evals), str(idx_degen), str(req1))))

        # eigenvalue contribution
        gevalsA = grad_evals.unsqueeze(-2) * evecs

        # eigenvector contribution: shifted multi-RHS adjoint solve
        with (M.uselinopparams(*mparams) if M is not None else null_context()):
            Bmat = _ortho(grad_evecs, evecs, D=idx_degen, M=M, mright=False)
            evals_shift = evals + 1e-14 if torch.is_complex(Bmat) else evals
            with A.uselinopparams(*params):
                gevecs = solve(A, -Bmat, evals_shift, M, bck_options=ctx.bck_config, **ctx.bck_config)
            gevecsA = _ortho(gevecs, evecs, D=None, M=M, mright=True)

        gaccumA = gevalsA + gevecsA
        grad_params = torch.autograd.grad((loss,), params, grad_outputs=(gaccumA,),
                                          create_graph=torch.is_grad_enabled())

        grad_mparams = []
        if M is not None:
            with torch.enable_grad():
                mparams = [p.clone().requires_grad_() for p in mparams]
                with M.uselinopparams(*mparams):
                    mloss = M.mm(evecs)
            gevalsM = -gevalsA * evals.unsqueeze(-2)
            gevecsM = -gevecsA * evals.unsqueeze(-2)
            par = (-0.5 * torch.einsum("...ae,...ae->...e", grad_evecs, evecs.conj())).unsqueeze(-2) * evecs
            grad_mparams = torch.autograd.grad((mloss,), mparams, grad_outputs=(gevalsM + gevecsM + par,),
                                               create_graph=torch.is_grad_enabled())

        return (None, None, None, None, None, None, None, *grad_params, *grad_mparams)


def _check_degen(evals: torch.Tensor, degen_atol: float, degen_rtol: float):
    diff = torch.abs(evals.unsqueeze(-2) - evals.unsqueeze(-1))
    thresh = degen_atol + degen_rtol * torch.abs(evals).unsqueeze(-1)
    idx = (diff < thresh).to(evals.dtype)
    return idx, bool(torch.sum(idx) > torch.numel(evals))


def _ortho(A: torch.Tensor, B: torch.Tensor, *, D: Optional[torch.Tensor] = None,
           M: Optional[LinearOperator] = None, mright: bool = False) -> torch.Tensor:
    """remove from the columns of A their (M-)components along the columns of B; D (if given) is the
    degeneracy map that widens the projection to degenerate partners."""
    if D is None:
        def coef(a):
            return torch.einsum("...rc,...rc->...c", a, B.conj()).unsqueeze(-2)
        if M is None:
            return A - coef(A) * B
        if mright:
            return A - coef(M.mm(A)) * B
        return A - M.mm(coef(A) * B)
    BH = B.transpose(-2, -1).conj()
    if M is None:
        return A - torch.matmul(B, D * torch.matmul(BH, A))
    if mright:
        return A - torch.matmul(B, D * torch.matmul(BH, M.mm(A)))
    return A - M.mm(torch.matmul(B, D * torch.matmul(BH, A)))


def svd(A: LinearOperator, k: Optional[int] = None, mode: str = "uppest",
        bck_options: Mapping[str, Any] = {}, method: Union[str, Callable, None] = None,
        **fwd_options):
    r"""``k`` extreme singular triplets via `symeig` on :math:`A^H A` (or :math:`A A^H` when
    p < q), the construction of /root/reference/xitorch/linalg/symeig.py:146-250."""
    if is_debug_enabled():
        A.check()
    m, n = A.shape[-2], A.shape[-1]
    if m < n:
        AAsym = A.matmul(A.H, is_hermitian=True)
    else:
        AAsym = A.H.matmul(A, is_hermitian=True)
    eivals, eivecs = symeig(AAsym, k, mode, bck_options=bck_options, method=method, **fwd_options)
    s = torch.sqrt(torch.clamp(eivals, min=0.0))
    sdiv = torch.clamp(s, min=1e-12).unsqueeze(-2)
    if m < n:
        u = eivecs
        v = A.rmm(u) / sdiv
    else:
        v = eivecs
        u = A.mm(v) / sdiv
    return u, s, v.transpose(-2, -1).conj()
