"""
Small host-side helpers shared by the boundary layer.

Mirrors the behaviour (not the code) of the reference helpers the hot path's
boundary relies on:
  * method dispatch by name/callable  (/root/reference/xitorch/_utils/misc.py:21-39)
  * option merging                    (/root/reference/xitorch/_utils/misc.py:6-19)
  * runtime asserts -> RuntimeError   (/root/reference/xitorch/_utils/assertfuncs.py)
  * warning classes                   (/root/reference/xitorch/_utils/exceptions.py)
  * batch-shape broadcasting          (/root/reference/xitorch/_utils/bcast.py:4-18)
  * dotted/indexed attribute paths    (/root/reference/xitorch/_utils/attr.py:7-62)
"""
import contextlib
import functools
import re
from typing import Any, Callable, Dict, List, Mapping, Sequence, Union

import torch


class ConvergenceWarning(Warning):
    """Emitted when an iterative method returns its best iterate without meeting the tolerance."""


class MathWarning(Warning):
    """Emitted when a mathematical requirement (e.g. degenerate-eigenvector gradients) is violated."""


class GetSetParamsError(Exception):
    """Raised by EditableModule.assertparams when a method changes the object's state or lists a non-float parameter."""


def assert_runtime(cond: bool, msg: str = "") -> None:
    if not cond:
        raise RuntimeError(msg)


def merged_options(defaults: Mapping[str, Any], given: Mapping[str, Any]) -> Dict[str, Any]:
    out = dict(defaults)
    out.update(given)
    return out


def pop_keys(dct: Dict[str, Any], keys: Sequence[str]) -> Dict[str, Any]:
    return {k: dct.pop(k) for k in keys}


def get_method(algname: str, methods: Mapping[str, Callable], method: Union[str, Callable]) -> Callable:
    """str -> lower-cased lookup (unknown -> RuntimeError); callable -> itself; else TypeError."""
    if isinstance(method, str):
        key = method.lower()
        if key not in methods:
            raise RuntimeError("Unknown %s method: %s" % (algname, method))
        return methods[key]
    if callable(method):
        return method
    raise TypeError("Invalid method type: %s. Only str and callable are accepted." % type(method))


@contextlib.contextmanager
def null_context():
    yield None


def bcast_dims(*shapes) -> List[int]:
    """broadcast of batch shapes (plain Python: torch.broadcast_shapes costs ~40 us per call on the host)"""
    nd = max((len(s) for s in shapes), default=0)
    out = [1] * nd
    for s in shapes:
        off = nd - len(s)
        for i, d in enumerate(s):
            d = int(d)
            cur = out[off + i]
            if cur == 1:
                out[off + i] = d
            elif d != 1 and d != cur:
                raise RuntimeError("Shape mismatch: objects cannot be broadcast to a single shape: %s"
                                   % (", ".join(str(tuple(x)) for x in shapes)))
    return out


def normalize_bcast_dims(*shapes) -> List[List[int]]:
    nd = max(len(s) for s in shapes)
    return [[1] * (nd - len(s)) + list(s) for s in shapes]


# ---- attribute paths such as "a.b[0].c" / 'd["key"]' -------------------------------------
_TOKEN = re.compile(r"\.?([A-Za-z_][A-Za-z_0-9]*)|\[([^\]]+)\]")


@functools.lru_cache(maxsize=4096)
def _parse_path(path: str):
    pos, toks = 0, []
    while pos < len(path):
        m = _TOKEN.match(path, pos)
        if m is None:
            raise AttributeError("cannot parse attribute path %r" % path)
        if m.group(1) is not None:
            toks.append(("attr", m.group(1)))
        else:
            raw = m.group(2).strip()
            if (raw[0] == raw[-1]) and raw[0] in "\"'":
                toks.append(("item", raw[1:-1]))
            else:
                toks.append(("item", int(raw)))
        pos = m.end()
    return tuple(toks)


def _step(obj, tok):
    return getattr(obj, tok[1]) if tok[0] == "attr" else obj[tok[1]]


def get_attr(obj, path: str):
    for tok in _parse_path(path):
        obj = _step(obj, tok)
    return obj


def set_attr(obj, path: str, val) -> None:
    toks = _parse_path(path)
    for tok in toks[:-1]:
        obj = _step(obj, tok)
    kind, key = toks[-1]
    if kind == "attr":
        # bypass nn.Module/Parameter type checks the same way a plain object would
        try:
            object.__setattr__(obj, key, val) if not isinstance(obj, torch.nn.Module) else _set_module_attr(obj, key, val)
        except AttributeError:
            setattr(obj, key, val)
    else:
        obj[key] = val


def _set_module_attr(mod: torch.nn.Module, key: str, val) -> None:
    if isinstance(val, (torch.nn.Parameter, torch.nn.Module)):
        mod.__dict__.pop(key, None)          # a plain tensor put there by an earlier substitution
        setattr(mod, key, val)               # registers the parameter / submodule properly
    elif key in mod._parameters:
        del mod._parameters[key]
        object.__setattr__(mod, key, val)
    elif key in mod._buffers:
        mod._buffers[key] = val
    else:
        object.__setattr__(mod, key, val)


def del_attr(obj, path: str) -> None:
    """delete the attribute / item at `path` (a list item is replaced by None so that the length is preserved)"""
    toks = _parse_path(path)
    for tok in toks[:-1]:
        obj = _step(obj, tok)
    kind, key = toks[-1]
    if kind == "attr":
        delattr(obj, key)
    elif isinstance(obj, list):
        obj[key] = None
    else:
        del obj[key]
