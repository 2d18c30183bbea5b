"""
`install_as_xitorch()` -- registers this package under the reference's module names, so that code written against
xitorch (DQC, xitorch's own benchmarks and tests) imports the B200 implementation without a source change:

    import xitorch_b200
    xitorch_b200.install_as_xitorch()
    import xitorch                       # -> xitorch_b200
    from xitorch.linalg import symeig    # the CUDA Krylov path
    from xitorch._core.editable_module import EditableModule

Only module *paths* are aliased (`sys.modules`); nothing of the reference is imported or needed.  The aliased paths
are the ones the hot path's callers use: `xitorch`, `.linalg[.symeig|.solve]`, `.optimize[.rootfinder]`,
`.grad[.jachess]`, `.debug[.modes]`, `._core.{editable_module,pure_function,linop}`,
`._utils.{exceptions,bcast,misc,attr,assertfuncs}`.
"""
import sys
import types

__all__ = ["install_as_xitorch"]


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install_as_xitorch(force: bool = False) -> None:
    if "xitorch" in sys.modules and not force:
        existing = sys.modules["xitorch"]
        if getattr(existing, "__xitorch_b200__", False):
            return
        raise RuntimeError("a module named 'xitorch' is already imported (%s); pass force=True to replace it"
                           % getattr(existing, "__file__", "?"))
    import xitorch_b200 as xb
    from xitorch_b200 import _utils, debug, editable_module, grad, linalg, linop, optimize, pure_function
    # (the packages re-export functions named like their submodules, so the modules are taken from sys.modules)
    solve_mod, symeig_mod = sys.modules["xitorch_b200.linalg.solve"], sys.modules["xitorch_b200.linalg.symeig"]
    rootfinder_mod = sys.modules["xitorch_b200.optimize.rootfinder"]

    top = _module("xitorch", __xitorch_b200__=True, __version__=xb.__version__, __path__=[])
    for nm in ("LinearOperator", "MatrixLinearOperator", "EditableModule", "get_pure_function", "make_sibling",
               "ConvergenceWarning", "MathWarning", "is_debug_enabled", "set_debug_mode", "enable_debug",
               "disable_debug"):
        setattr(top, nm, getattr(xb, nm))
    core = _module("xitorch._core", __path__=[], editable_module=editable_module, pure_function=pure_function,
                   linop=linop)
    utils = _module("xitorch._utils", __path__=[])
    exceptions = _module("xitorch._utils.exceptions", ConvergenceWarning=_utils.ConvergenceWarning,
                         MathWarning=_utils.MathWarning, GetSetParamsError=_utils.GetSetParamsError)
    bcast = _module("xitorch._utils.bcast", get_bcasted_dims=_utils.bcast_dims,
                    normalize_bcast_dims=_utils.normalize_bcast_dims)
    misc = _module("xitorch._utils.misc", get_method=_utils.get_method, set_default_option=_utils.merged_options,
                   get_and_pop_keys=_utils.pop_keys, dummy_context_manager=_utils.null_context)
    attr = _module("xitorch._utils.attr", get_attr=_utils.get_attr, set_attr=_utils.set_attr, del_attr=_utils.del_attr)
    asserts = _module("xitorch._utils.assertfuncs", assert_runtime=_utils.assert_runtime)
    for nm, sub in (("exceptions", exceptions), ("bcast", bcast), ("misc", misc), ("attr", attr),
                    ("assertfuncs", asserts)):
        setattr(utils, nm, sub)
    debug_pkg = _module("xitorch.debug", __path__=[], modes=debug, **{k: getattr(debug, k) for k in debug.__all__})
    grad_pkg = _module("xitorch.grad", __path__=[], jachess=grad, jac=grad.jac, hess=grad.hess)
    top.linalg, top.optimize, top.grad, top.debug, top._core, top._utils = linalg, optimize, grad_pkg, debug_pkg, core, utils
    table = {
        "xitorch": top,
        "xitorch.linalg": linalg, "xitorch.linalg.symeig": symeig_mod, "xitorch.linalg.solve": solve_mod,
        "xitorch.optimize": optimize, "xitorch.optimize.rootfinder": rootfinder_mod,
        "xitorch.grad": grad_pkg, "xitorch.grad.jachess": grad,
        "xitorch.debug": debug_pkg, "xitorch.debug.modes": debug,
        "xitorch._core": core, "xitorch._core.editable_module": editable_module,
        "xitorch._core.pure_function": pure_function, "xitorch._core.linop": linop,
        "xitorch._utils": utils, "xitorch._utils.exceptions": exceptions, "xitorch._utils.bcast": bcast,
        "xitorch._utils.misc": misc, "xitorch._utils.attr": attr, "xitorch._utils.assertfuncs": asserts,
    }
    sys.modules.update(table)
