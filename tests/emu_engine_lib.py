"""
Builds and installs the HOST BUILD of the engines (TEST INFRASTRUCTURE): csrc/{symeig,solve,gmres,linop}.cu rewritten textually
for a host compiler by tools/emu_engine (host threads for CUDA threads, a loop for the block matvec) into one shared
library with the C ABI of include/xitorch_b200.h, loaded with ctypes and put in place of the CUDA library for the
duration of a test.  Everything above the C ABI -- and everything below it except the matvec kernel -- is then the
shipped code, running on CPU tensors.
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import torch

from xitorch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ("symeig", "solve", "gmres", "linop")


def build(workdir: str):
    if shutil.which("g++") is None:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tools", "emu_engine"))
    try:
        import preprocess
    finally:
        sys.path.pop(0)
    csrc = os.path.join(ROOT, "xitorch_b200", "csrc")
    cpps = []
    for name in SOURCES:
        out = os.path.join(workdir, name + "_host.cpp")
        open(out, "w").write(preprocess.transform(open(os.path.join(csrc, name + ".cu")).read(), csrc))
        cpps.append(out)
    so = os.path.join(workdir, "libxt_emu.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                           "-I", os.path.join(ROOT, "tools", "emu_engine"), "-I", os.path.join(ROOT, "include"),
                           "-o", so] + cpps)
    return load(so)


def load(so: str):
    lib = C.CDLL(so)
    lib.path = so
    lib.xt_symeig_workspace_bytes.argtypes = [C.c_int32] * 5
    lib.xt_symeig_workspace_bytes.restype = C.c_size_t
    lib.xt_symeig_krylov.argtypes = [C.POINTER(_lib.SymeigArgs)]
    lib.xt_symeig_krylov.restype = C.c_int
    lib.xt_symeig_sharded_workspace_bytes.argtypes = [C.c_int32] * 5
    lib.xt_symeig_sharded_workspace_bytes.restype = C.c_size_t
    lib.xt_symeig_peer_bytes.argtypes = [C.c_int32] * 5
    lib.xt_symeig_peer_bytes.restype = C.c_size_t
    lib.xt_small_eigh.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    lib.xt_small_eigh.restype = C.c_int
    lib.xt_solve_workspace_bytes.argtypes = [C.c_char_p] + [C.c_int32] * 6
    lib.xt_solve_workspace_bytes.restype = C.c_size_t
    for nm in ("xt_cg", "xt_bicgstab", "xt_gmres"):
        getattr(lib, nm).argtypes = [C.POINTER(_lib.SolveArgs)]
        getattr(lib, nm).restype = C.c_int
    lib.xt_hermitian_check.argtypes = [C.POINTER(_lib.HermCheckArgs)]
    lib.xt_hermitian_check.restype = C.c_int
    return lib


class _Hybrid(object):
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def xt_last_error(self):
        return b"emulated engine"


class _NoDevice(object):
    def __init__(self, dev):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install(monkeypatch, lib):
    """route the host layer to the emulated library (CPU tensors allowed, no CUDA stream / device handling)"""
    from xitorch_b200._impls import symeig as impl
    hyb = _Hybrid(lib)
    monkeypatch.setattr(_lib, "lib", lambda: hyb)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda dev: 0)
    monkeypatch.setattr(torch.cuda, "device", _NoDevice)
    # the start block the CUDA path draws (seed 12421, reference symeig.py:236), from the CPU generator
    monkeypatch.setattr(impl, "_start_block",
                        lambda kind, nb, n, neig, dtype, dev: (torch.randn if kind == "randn" else torch.rand)(
                            (nb, n, neig), dtype=dtype, generator=torch.Generator().manual_seed(12421)))
    return hyb
