"""
ctypes binding of the C ABI in include/xitorch_b200.h (csrc/libxitorch_b200.so).

The shared library is built in-tree by `python -m xitorch_b200.csrc.build`
(`__graft_entry__.build()` does that).  Loading never falls back to anything else:
if the library is missing, `lib()` raises.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libxitorch_b200.so")

XT_F32, XT_BF16, XT_F64 = 0, 1, 2

_DTYPES = {torch.float32: XT_F32, torch.bfloat16: XT_BF16, torch.float64: XT_F64}

# every symbol include/xitorch_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "xt_version", "xt_last_error", "xt_profile_reset", "xt_profile_read", "xt_block_matvec",
    "xt_solve_workspace_bytes", "xt_cg", "xt_bicgstab", "xt_gmres",
    "xt_symeig_workspace_bytes", "xt_symeig_krylov", "xt_small_eigh", "xt_hermitian_check",
    "xt_symeig_sharded_workspace_bytes", "xt_symeig_peer_bytes",
    "xt_peer_alloc", "xt_peer_open", "xt_peer_close", "xt_peer_free",
]


class MatvecArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("nbatch", C.c_int32), ("nrows", C.c_int32), ("ncolsA", C.c_int32), ("k", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_bstride", C.c_int64),
        ("X", C.c_void_p), ("ldx", C.c_int64), ("x_bstride", C.c_int64),
        ("Y", C.c_void_p), ("ldy", C.c_int64), ("y_bstride", C.c_int64),
        ("E", C.c_void_p), ("e_bstride", C.c_int64),
        ("Z", C.c_void_p), ("ldz", C.c_int64), ("z_bstride", C.c_int64),
        ("impl", C.c_int32),
        ("stream", C.c_void_p),
        ("trans", C.c_int32),
    ]


class SolveArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("n", C.c_int32), ("nbatch", C.c_int32), ("ncols", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_bstride", C.c_int64),
        ("M", C.c_void_p), ("ldm", C.c_int64), ("m_bstride", C.c_int64),
        ("E", C.c_void_p), ("e_bstride", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_bstride", C.c_int64),
        ("X", C.c_void_p), ("ldx", C.c_int64), ("x_bstride", C.c_int64),
        ("rtol", C.c_double), ("atol", C.c_double), ("eps", C.c_double),
        ("max_niter", C.c_int32), ("resid_calc_every", C.c_int32), ("check_every", C.c_int32),
        ("niter_out", C.POINTER(C.c_int32)), ("converged_out", C.POINTER(C.c_int32)),
        ("best_resid_out", C.POINTER(C.c_double)), ("napply_out", C.POINTER(C.c_int64)),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
        ("apply", C.c_void_p), ("apply_user", C.c_void_p),
        ("precond_l", C.c_void_p), ("precond_r", C.c_void_p), ("precond_user", C.c_void_p),
        ("abort", C.POINTER(C.c_int32)),
    ]


# matrix-free operator callback of xt_solve_args: apply(user, X, Y, stream)
APPLY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class SymeigArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("n", C.c_int32), ("nbatch", C.c_int32), ("neig", C.c_int32),
        ("mode", C.c_int32), ("expansion", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_bstride", C.c_int64),
        ("V0", C.c_void_p), ("ldv0", C.c_int64), ("v0_bstride", C.c_int64),
        ("evals", C.c_void_p), ("evals_bstride", C.c_int64),
        ("evecs", C.c_void_p), ("ldv", C.c_int64), ("evecs_bstride", C.c_int64),
        ("max_niter", C.c_int32), ("max_basis", C.c_int32), ("check_every", C.c_int32),
        ("min_eps", C.c_double),
        ("niter_out", C.POINTER(C.c_int32)), ("converged_out", C.POINTER(C.c_int32)),
        ("best_resid_out", C.POINTER(C.c_double)), ("napply_out", C.POINTER(C.c_int64)),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
        ("world", C.c_int32), ("rank", C.c_int32),
        ("allgather", C.c_void_p), ("allgather_user", C.c_void_p),
        ("apply", C.c_void_p), ("apply_user", C.c_void_p),
        ("peers", C.POINTER(C.c_void_p)), ("epoch", C.c_uint32), ("restart_keep", C.c_int32),
        ("abort", C.POINTER(C.c_int32)),
    ]


class HermCheckArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("n", C.c_int32), ("nbatch", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_bstride", C.c_int64),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("mismatch", C.c_void_p),
        ("stream", C.c_void_p),
    ]


ALLGATHER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p)


def fn_address(callback) -> int:
    """address of a CFUNCTYPE callback for a `void*` struct field.  Not `ctypes.cast(callback, c_void_p)`: cast() makes
    its result share the source's `_objects` dict and stores the source in it, i.e. the callback object then references
    itself -- and everything its closure holds (the workspace tensor, the operator and its tensors) would stay alive
    until the cyclic garbage collector runs.  The caller keeps `callback` alive for the duration of the library call."""
    return C.c_void_p.from_address(C.addressof(callback)).value

_lock = threading.Lock()
_lib = None


def lib():
    """the loaded shared library (raises RuntimeError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "xitorch_b200: the CUDA extension %s is missing. Build it with "
                "`python -m xitorch_b200.csrc.build` (there is no CPU / PyTorch fallback for the "
                "Krylov hot path)." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.xt_version.restype = C.c_int
        L.xt_last_error.restype = C.c_char_p
        L.xt_profile_reset.argtypes = [C.c_int]
        L.xt_profile_reset.restype = None
        L.xt_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.xt_profile_read.restype = C.c_int
        L.xt_block_matvec.argtypes = [C.POINTER(MatvecArgs)]
        L.xt_block_matvec.restype = C.c_int
        for name in ("xt_cg", "xt_bicgstab", "xt_gmres"):
            fn = getattr(L, name)
            fn.argtypes = [C.POINTER(SolveArgs)]
            fn.restype = C.c_int
        L.xt_solve_workspace_bytes.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32]
        L.xt_solve_workspace_bytes.restype = C.c_size_t
        L.xt_symeig_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.xt_symeig_workspace_bytes.restype = C.c_size_t
        L.xt_hermitian_check.argtypes = [C.POINTER(HermCheckArgs)]
        L.xt_hermitian_check.restype = C.c_int
        L.xt_symeig_krylov.argtypes = [C.POINTER(SymeigArgs)]
        L.xt_symeig_krylov.restype = C.c_int
        L.xt_symeig_sharded_workspace_bytes.argtypes = [C.c_int32] * 5
        L.xt_symeig_sharded_workspace_bytes.restype = C.c_size_t
        L.xt_symeig_peer_bytes.argtypes = [C.c_int32] * 5
        L.xt_symeig_peer_bytes.restype = C.c_size_t
        L.xt_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
        L.xt_peer_alloc.restype = C.c_int
        L.xt_peer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.xt_peer_open.restype = C.c_int
        L.xt_peer_close.argtypes = [C.c_void_p]
        L.xt_peer_close.restype = C.c_int
        L.xt_peer_free.argtypes = [C.c_void_p]
        L.xt_peer_free.restype = C.c_int
        L.xt_small_eigh.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
        L.xt_small_eigh.restype = C.c_int
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().xt_last_error()
        raise RuntimeError("xitorch_b200 %s failed (status %d): %s" % (what, rc, (msg or b"").decode()))


def dtype_code(dt: torch.dtype) -> int:
    if dt not in _DTYPES:
        raise RuntimeError("xitorch_b200: unsupported dtype %s (float32, bfloat16, float64 only)" % dt)
    return _DTYPES[dt]


def vec_dtype(dt: torch.dtype) -> torch.dtype:
    """dtype of vectors / results for an operator stored as `dt`."""
    return torch.float64 if dt == torch.float64 else torch.float32


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "xitorch_b200: %s needs CUDA tensors; this package has no CPU path for the Krylov "
            "methods (got a tensor on %s)" % (what, t.device))


def profile_reset(enable: bool = True) -> None:
    lib().xt_profile_reset(1 if enable else 0)


def profile_read():
    """(matvec_ms, matvec_launches, total_launches) since the last profile_reset."""
    ms, nmv, ntot = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
    check(lib().xt_profile_read(C.byref(ms), C.byref(nmv), C.byref(ntot)), "profile_read")
    return ms.value, nmv.value, ntot.value
