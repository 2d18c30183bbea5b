"""
`EditableModule` -- parameter plumbing the autograd boundaries need (layer L1 of SURVEY.md 1; contract =
/root/reference/xitorch/_core/editable_module.py:14-362 and its tests, _tests/test_editable_module.py).

An object exposes, per method, the *names* (attribute paths such as ``a.b[0]["k"]``) of the tensors that method's
output depends on (`getparamnames`).  The functionals (`solve`, `symeig`, `rootfinder`, `jac`) use the names to read the
tensors (`getparams` / `getuniqueparams`) and to temporarily substitute them (`setparams` / `setuniqueparams`), which is
how gradients and second derivatives reach tensors hidden inside objects.  `assertparams` is the debugging aid that
finds wrong name lists by differentiating through the method.

Written from the contract, not from the reference code: per method one small table (names + the groups of positions
that hold the same tensor) replaces the reference's three parallel dictionaries.
"""
import copy
import inspect
import warnings
from typing import Dict, List, Sequence, Tuple

import torch

from xitorch_b200._utils import get_attr, set_attr, del_attr, GetSetParamsError

__all__ = ["EditableModule"]

_FLOAT_TYPES = (torch.float16, torch.float32, torch.float64)


class _ParamTable(object):
    """names of one method's parameters and which positions alias the same tensor"""

    def __init__(self, names: List[str]):
        self.names = list(names)
        self.groups = None          # List[List[int]]: positions per unique tensor, in order of first appearance

    def resolve_groups(self, tensors: Sequence[torch.Tensor]) -> List[List[int]]:
        if self.groups is None:
            first: Dict[int, int] = {}
            groups: List[List[int]] = []
            for pos, t in enumerate(tensors):
                key = id(t)
                if key in first:
                    groups[first[key]].append(pos)
                else:
                    first[key] = len(groups)
                    groups.append([pos])
            self.groups = groups
        return self.groups


class EditableModule(object):
    """Base class of objects whose methods can be turned into pure functions of their tensor state."""

    # ------------------------------------------------------------------ to be implemented by subclasses
    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        """names of the tensors that affect the output of method `methodname` (KeyError for unknown methods)."""
        raise NotImplementedError("getparamnames(self, methodname, prefix) must be implemented by %s"
                                  % self.__class__.__name__)

    # ------------------------------------------------------------------ name table
    def _param_table(self, methodname: str) -> _ParamTable:
        tables = self.__dict__.setdefault("_xt_param_tables", {})
        tab = tables.get(methodname)
        if tab is None:
            tab = _ParamTable(self.getparamnames(methodname))
            tables[methodname] = tab
        return tab

    def cached_getparamnames(self, methodname: str, refresh: bool = False) -> List[str]:
        if refresh:
            self.__dict__.setdefault("_xt_param_tables", {}).pop(methodname, None)
        return self._param_table(methodname).names

    # ------------------------------------------------------------------ all parameters, in name order
    def getparams(self, methodname: str) -> List[torch.Tensor]:
        return [get_attr(self, nm) for nm in self._param_table(methodname).names]

    def setparams(self, methodname: str, *params) -> int:
        """substitute the method's parameters (more values than names may be given; the number consumed is returned
        as the reference does: the length of what was passed)."""
        for nm, val in zip(self._param_table(methodname).names, params):
            try:
                set_attr(self, nm, val)
            except TypeError:            # e.g. a plain tensor where an nn.Parameter is registered
                del_attr(self, nm)
                set_attr(self, nm, val)
        return len(params)

    # ------------------------------------------------------------------ de-duplicated view
    def getuniqueparams(self, methodname: str, onlyleaves: bool = False) -> List[torch.Tensor]:
        tensors = self.getparams(methodname)
        groups = self._param_table(methodname).resolve_groups(tensors)
        uniq = [tensors[g[0]] for g in groups]
        return [t for t in uniq if t.is_leaf] if onlyleaves else uniq

    def setuniqueparams(self, methodname: str, *uniqueparams) -> int:
        tab = self._param_table(methodname)
        if tab.groups is None:
            tab.resolve_groups(self.getparams(methodname))
        full = [None] * len(tab.names)
        for val, positions in zip(uniqueparams, tab.groups):
            for pos in positions:
                full[pos] = val
        return self.setparams(methodname, *full)

    # ------------------------------------------------------------------ debugging aid
    def assertparams(self, method, *args, **kwargs):
        """Check `getparamnames` of `method` (a bound method of this object) by running it on `args`: the method must
        leave the object's tensors unchanged (else GetSetParamsError), and the names it lists are compared with the
        tensors its output really depends on (warnings for missing / excess names)."""
        if not inspect.ismethod(method):
            raise TypeError("The input method must be a method")
        if method.__self__ is not self:
            raise RuntimeError("The method does not belong to the same instance")
        clsname, mname = self.__class__.__name__, method.__name__

        # 1. state preservation
        before, names = _collect_float_tensors(self)
        snapshot = [t.clone() for t in before]
        method(*args, **kwargs)
        after, _ = _collect_float_tensors(self)
        head = "The method %s.%s does not preserve the object's float tensors: \n" % (clsname, mname)
        if len(after) != len(snapshot):
            raise GetSetParamsError(head + "The number of parameters changed:\n"
                                    "* number of object's parameters before: %d\n"
                                    "* number of object's parameters after : %d\n" % (len(snapshot), len(after)))
        for nm, t0, t1 in zip(names, snapshot, after):
            if t0.shape != t1.shape:
                raise GetSetParamsError(head + "The shape of %s changed\n* (before) %s.shape: %s\n* (after ) %s.shape: %s\n"
                                        % (nm, nm, t0.shape, nm, t1.shape))
            if not torch.allclose(t0, t1):
                raise GetSetParamsError(head + "The value of %s changed\n* (before) %s: %s\n* (after ) %s: %s\n"
                                        % (nm, nm, t0, nm, t1))

        # 2. which tensors does the output depend on?  swap in fresh leaves, differentiate, swap back
        originals, onames = _collect_float_tensors(self)
        leaves = [t.clone().detach().requires_grad_() for t in originals]
        _replace_float_tensors(self, list(leaves))
        try:
            out = method(*args, **kwargs)
            if not isinstance(out, torch.Tensor):
                raise RuntimeError("The method to be asserted must have a tensor output")
            grads = torch.autograd.grad(out.sum(), leaves, retain_graph=True, allow_unused=True)
        finally:
            _replace_float_tensors(self, list(originals))
        used = [(nm, t) for nm, t, g in zip(onames, originals, grads) if g is not None]

        listed_names = self.getparamnames(mname)
        listed = [get_attr(self, nm) for nm in listed_names]
        for nm, t in zip(listed_names, listed):
            if not (isinstance(t, torch.Tensor) and t.dtype in _FLOAT_TYPES):
                raise GetSetParamsError("Parameter %s is a non-floating point tensor" % nm)
        listed_ids = {id(t) for t in listed}
        used_ids = {id(t) for _, t in used}
        missing = [nm for nm, t in used if id(t) not in listed_ids]
        if missing:
            warnings.warn("getparams for %s.%s does not include: %s" % (clsname, mname, ", ".join(missing)),
                          stacklevel=2)
        excess = [nm for nm, t in zip(listed_names, listed) if id(t) not in used_ids]
        if excess:
            warnings.warn("getparams for %s.%s has excess parameters: %s" % (clsname, mname, ", ".join(excess)),
                          stacklevel=2)
        print('"%s" method check done' % mname)


# ---------------------------------------------------------------------- object traversal
def _children(obj):
    """(key, value, container, display-name-format) of everything directly reachable from `obj`"""
    if isinstance(obj, torch.nn.Module):
        for store in (obj._parameters, obj._modules):
            for k, v in list(store.items()):
                yield k, v, store, "{p}{k}"
    elif hasattr(obj, "__dict__"):
        for k, v in list(obj.__dict__.items()):
            yield k, v, obj.__dict__, "{p}{k}"
    elif hasattr(obj, "__iter__"):
        items = obj.items() if isinstance(obj, dict) else enumerate(obj)
        for k, v in list(items):
            yield k, v, obj, "{p}[{k}]"
    else:
        raise RuntimeError("The object must be iterable or keyable")


def _is_float_tensor(x) -> bool:
    return isinstance(x, torch.Tensor) and x.dtype in _FLOAT_TYPES


def _walk(obj, prefix, visit, depth, seen):
    for key, val, store, fmt in _children(obj):
        name = fmt.format(p=prefix, k=key)
        if _is_float_tensor(val):
            visit(name, val, store, key)
            continue
        has_dict, has_iter = hasattr(val, "__dict__"), hasattr(val, "__iter__")
        if not (has_dict or has_iter) or isinstance(val, (str, bytes)) or id(val) in seen:
            continue
        seen.add(id(val))
        if depth <= 0:
            raise RecursionError("Maximum number of recursion reached")
        _walk(val, name + "." if has_dict else name, visit, depth - 1, seen)


def _collect_float_tensors(obj, prefix: str = "", max_depth: int = 20) -> Tuple[List[torch.Tensor], List[str]]:
    tensors, names = [], []

    def visit(name, val, store, key):
        tensors.append(val)
        names.append(name)

    _walk(obj, prefix, visit, max_depth, set())
    return tensors, names


def _replace_float_tensors(obj, new_values: List[torch.Tensor], max_depth: int = 20) -> None:
    def visit(name, val, store, key):
        store[key] = new_values.pop(0)

    _walk(obj, "", visit, max_depth, set())
