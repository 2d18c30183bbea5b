"""
Global debug-mode switch (the boundary calls `is_debug_enabled()` -> `A.check()`).
Behavioural mirror of /root/reference/xitorch/debug/modes.py:5-56.
"""
from contextlib import contextmanager

__all__ = ["is_debug_enabled", "set_debug_mode", "enable_debug", "disable_debug"]

_state = {"on": False}


def is_debug_enabled() -> bool:
    return _state["on"]


def set_debug_mode(mode: bool) -> None:
    _state["on"] = bool(mode)


@contextmanager
def enable_debug():
    prev = _state["on"]
    _state["on"] = True
    try:
        yield
    finally:
        _state["on"] = prev


@contextmanager
def disable_debug():
    prev = _state["on"]
    _state["on"] = False
    try:
        yield
    finally:
        _state["on"] = prev
