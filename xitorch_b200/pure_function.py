"""
Pure-function view of functions and bound methods (layer L1 of SURVEY.md 1; contract =
/root/reference/xitorch/_core/pure_function.py:11-219 and _tests/test_pure_function.py).

`get_pure_function(fcn)` wraps a function, a `torch.jit` script function, a method of an `EditableModule` or of a
`torch.nn.Module` (or a callable object of those kinds).  The wrapper is called like `fcn`, and additionally exposes
the tensors hidden in the object the method is bound to: `objparams()` lists them (de-duplicated), `useobjparams(new)`
substitutes them for the duration of a `with` block (nested blocks restore in LIFO order).  `make_sibling(*fcns)` marks a
new function as sharing the state of existing ones, so that substituting its parameters substitutes theirs.  The
autograd boundaries (`rootfinder`, `jac`) pass `objparams()` through `torch.autograd.Function.apply` next to the explicit
parameters, which is how gradients reach object state.

Written from the contract: one wrapper class whose state comes from a list of *sources* (none / EditableModule method /
nn.Module / the sources of sibling functions) instead of the reference's class per kind.
"""
import inspect
from contextlib import contextmanager
from typing import Callable, List, Sequence

import torch

from xitorch_b200._utils import set_attr, del_attr
from xitorch_b200.editable_module import EditableModule

__all__ = ["get_pure_function", "make_sibling", "PureFunction"]


class _EditableSource(object):
    def __init__(self, obj: EditableModule, methodname: str):
        self.obj, self.methodname = obj, methodname

    def read(self) -> List:
        return list(self.obj.getparams(self.methodname))

    def write(self, values: Sequence) -> None:
        self.obj.setparams(self.methodname, *values)


class _ModuleSource(object):
    def __init__(self, module: torch.nn.Module):
        self.module = module
        self.names = [nm for nm, _ in module.named_parameters()]

    def read(self) -> List:
        return [p for _, p in self.module.named_parameters()]

    def write(self, values: Sequence) -> None:
        for nm, val in zip(self.names, values):
            del_attr(self.module, nm)       # needed when the new value is not an nn.Parameter
            set_attr(self.module, nm, val)


class PureFunction(object):
    """callable + the (substitutable) tensors its output silently depends on"""

    def __init__(self, fcn: Callable, sources: Sequence):
        self._fcn = fcn
        self._sources = list(sources)
        self._counts = []
        everything: List = []
        for src in self._sources:
            vals = src.read()
            self._counts.append(len(vals))
            everything.extend(vals)
        # positions of the first occurrence of every distinct object, and for each position its unique slot
        slot_of, self._first, self._slot = {}, [], []
        for pos, val in enumerate(everything):
            key = id(val)
            if key not in slot_of:
                slot_of[key] = len(self._first)
                self._first.append(pos)
            self._slot.append(slot_of[key])
        self._current = [everything[pos] for pos in self._first]
        self._stack: List = []
        self._frozen = False

    def __call__(self, *params):
        return self._fcn(*params)

    # ------------------------------------------------------------------ state
    def objparams(self) -> List:
        return self._current

    def _install(self, unique_values: Sequence) -> None:
        everything = [unique_values[s] for s in self._slot]
        start = 0
        for src, cnt in zip(self._sources, self._counts):
            src.write(everything[start:start + cnt])
            start += cnt

    def set_objparams(self, objparams: Sequence) -> None:
        same = len(objparams) == len(self._current) and all(a is b for a, b in zip(objparams, self._current))
        self._stack.append((self._current, same))
        if not same:
            if len(objparams) != len(self._first):
                raise RuntimeError("The uniqueobjs must have %d elements" % len(self._first))
            self._install(objparams)
            self._current = list(objparams)

    def restore_objparams(self) -> None:
        previous, same = self._stack.pop()
        if not same:
            self._install(previous)
            self._current = previous

    @contextmanager
    def useobjparams(self, objparams: Sequence):
        if self._frozen:
            raise RuntimeError("The state change is disabled")
        self.set_objparams(objparams)
        try:
            yield
        finally:
            self.restore_objparams()

    @contextmanager
    def disable_state_change(self):
        prev, self._frozen = self._frozen, True
        try:
            yield
        finally:
            self._frozen = prev


_ERRMSG = ("The input function must be a function, a method of torch.nn.Module, a method of "
           "xitorch.EditableModule, or a sibling method")


def get_pure_function(fcn) -> PureFunction:
    """the pure-function wrapper of a function / script function / method / callable object (idempotent)"""
    if isinstance(fcn, PureFunction):
        return fcn
    if inspect.isfunction(fcn) or isinstance(fcn, torch.jit.ScriptFunction):
        return PureFunction(fcn, [])
    if inspect.ismethod(fcn) or hasattr(fcn, "__call__"):
        if inspect.ismethod(fcn):
            owner = fcn.__self__
        else:
            owner, fcn = fcn, fcn.__call__
        if isinstance(owner, EditableModule):
            return PureFunction(fcn, [_EditableSource(owner, fcn.__name__)])
        if isinstance(owner, torch.nn.Module):
            return PureFunction(fcn, [_ModuleSource(owner)])
    raise RuntimeError(_ERRMSG)


def make_sibling(*fcns) -> Callable[[Callable], PureFunction]:
    """decorator: the decorated function shares (and can substitute) the state of `fcns`"""
    if len(fcns) == 0:
        raise TypeError("At least 1 function is required as the argument")
    sources: List = []
    for f in fcns:
        sources.extend(get_pure_function(f)._sources)
    return lambda fcn: PureFunction(fcn, sources)
