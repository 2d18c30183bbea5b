"""
Stand-in for the CUDA shared library (TEST INFRASTRUCTURE, CPU only): the entry points the Python host layer calls
(`xt_cg`, `xt_bicgstab`, `xt_gmres`, `xt_symeig_krylov`, the workspace queries), implemented with numpy on HOST memory
from nothing but the argument structs of include/xitorch_b200.h -- pointers, leading dimensions, batch strides,
dtype codes, shifts, callbacks.  With it the `-m "not gpu"` tier drives the complete host logic of
`xitorch_b200.linalg.solve / symeig` (marshalling of batched / broadcast / strided operands, E and M, normal
equations, complex real-equivalent form, matrix-free and preconditioner callbacks, autograd boundary) and checks the
answers against dense linear algebra.  It says nothing about the kernels: those are the `-m gpu` tests.

The numerics are deliberately direct (LU / eigh per system): what is under test is what the host hands over.
"""
import ctypes as C

import numpy as np
import torch

from xitorch_b200 import _lib

_NP = {_lib.XT_F32: np.float32, _lib.XT_F64: np.float64}


def _arr(ptr, count, npdt):
    ct = C.c_float if npdt == np.float32 else C.c_double
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(count,))


def _matrix(ptr, rows, cols, ld, npdt):
    """(rows, cols) view of a row-major matrix with leading dimension ld"""
    flat = _arr(ptr, (rows - 1) * ld + cols, npdt)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(ld * flat.itemsize, flat.itemsize))


class StandInLibrary(object):
    def __init__(self):
        self.log = []                   # one dict per engine call: what the host handed over
        self.probe_precond = True

    # ------------------------------------------------------------------ queries
    def xt_last_error(self):
        return b"stand-in engine"

    def xt_solve_workspace_bytes(self, method, dtype, n, nbatch, ncols, max_niter, has_M):
        return 8 * nbatch * n * ncols * 8 + 1024

    def xt_symeig_workspace_bytes(self, dtype, n, neig, max_basis, world):
        return 4 * n * neig * 8 + 256

    # ------------------------------------------------------------------ linear solvers
    def _solve(self, name, g):
        npdt = _NP[g.dtype]
        n, nb, nc = g.n, g.nbatch, g.ncols
        vdt = np.float64 if g.dtype == _lib.XT_F64 else np.float32
        esz = np.dtype(vdt).itemsize
        rec = dict(method=name, n=n, nbatch=nb, ncols=nc, dtype=g.dtype, has_E=bool(g.E), has_M=bool(g.M),
                   matrix_free=bool(g.apply), precond_l=bool(g.precond_l), precond_r=bool(g.precond_r),
                   rtol=g.rtol, atol=g.atol, max_niter=g.max_niter, check_every=g.check_every)
        self.log.append(rec)
        assert g.B and g.X and g.workspace and (g.A or g.apply)
        B = [_matrix(g.B + b * g.b_bstride * esz, n, nc, g.ldb, vdt).astype(np.float64) for b in range(nb)]
        X = [_matrix(g.X + b * g.x_bstride * esz, n, nc, g.ldx, vdt) for b in range(nb)]
        blk = nb * n * nc
        woff = [64, 64 + ((blk * esz + 63) // 64) * 64]

        def wview(i):
            return _arr(g.workspace + woff[i], blk, vdt).reshape(nb, n, nc)

        def call(fnptr, x):
            wview(0)[...] = x.astype(vdt)
            C.cast(fnptr, _lib.APPLY_FN)(None, g.workspace + woff[0], g.workspace + woff[1], None)
            return wview(1).astype(np.float64).copy()

        if g.apply:
            # the operator of every (batch, column) system, column by column of the identity
            mats = np.zeros((nb, nc, n, n))
            for i in range(n):
                x = np.zeros((nb, n, nc))
                x[:, i, :] = 1.0
                y = call(g.apply, x)                                  # (nb, n, nc)
                mats[:, :, :, i] = y.transpose(0, 2, 1)
        else:
            asz = np.dtype(npdt).itemsize
            mats = np.zeros((nb, nc, n, n))
            for b in range(nb):
                A = _matrix(g.A + b * g.a_bstride * asz, n, n, g.lda, npdt).astype(np.float64)
                Mm = np.eye(n)
                if g.M:
                    Mm = _matrix(g.M + b * g.m_bstride * asz, n, n, g.ldm, npdt).astype(np.float64)
                for j in range(nc):
                    e = float(_arr(g.E + b * g.e_bstride * esz, nc, vdt)[j]) if g.E else 0.0
                    mats[b, j] = A - e * Mm
        if self.probe_precond and (g.precond_l or g.precond_r):
            rng = np.random.default_rng(5)
            probe = rng.standard_normal((nb, n, nc))
            rec["precond_probe"] = probe
            for key, ptr in (("precond_l_out", g.precond_l), ("precond_r_out", g.precond_r)):
                if ptr:
                    rec[key] = call(ptr, probe)
        for b in range(nb):
            for j in range(nc):
                try:
                    X[b][:, j] = np.linalg.solve(mats[b, j], B[b][:, j]).astype(vdt)
                except np.linalg.LinAlgError:           # e.g. after a failed callback: the real engine returns too
                    X[b][:, j] = np.nan
        rec["systems"] = mats
        if g.niter_out:
            g.niter_out[0] = 1
        if g.converged_out:
            g.converged_out[0] = 1
        if g.best_resid_out:
            g.best_resid_out[0] = 0.0
        if g.napply_out:
            g.napply_out[0] = 1
        return 0

    def xt_cg(self, g):
        return self._solve("cg", g)

    def xt_bicgstab(self, g):
        return self._solve("bicgstab", g)

    def xt_gmres(self, g):
        return self._solve("gmres", g)

    # ------------------------------------------------------------------ eigensolver
    def xt_symeig_krylov(self, g):
        npdt = _NP[g.dtype]
        esz = np.dtype(npdt).itemsize
        n, k, nb = g.n, g.neig, g.nbatch
        rec = dict(method="symeig", n=n, nbatch=nb, neig=k, mode=g.mode, expansion=g.expansion,
                   matrix_free=bool(g.apply), max_basis=g.max_basis, min_eps=g.min_eps)
        self.log.append(rec)
        assert g.V0 and g.evals and g.evecs and g.workspace and (g.A or g.apply)
        napply = 0
        world = g.world if g.world > 1 else 1
        rec["world"] = world
        for b in range(nb):
            if world > 1:
                # row-partitioned operator: the local row block times identity columns, one all-gather per application
                # in the staging layout of the engine (chunk r = rank r's rows, plus one row carrying its stop flag)
                assert g.allgather and nb == 1 and not g.apply and n % world == 0
                n_local = n // world
                per = (n_local + 1) * k
                buf = _arr(g.workspace + 64, world * per, npdt)
                Aloc = _matrix(g.A, n_local, n, g.lda, npdt).astype(np.float64)
                A = np.zeros((n, n))
                for c0 in range(0, n, k):
                    X = np.zeros((n, k))
                    for j in range(min(k, n - c0)):
                        X[c0 + j, j] = 1
                    buf[...] = np.nan
                    mine = buf[g.rank * per:(g.rank + 1) * per]
                    mine[:n_local * k] = (Aloc @ X).astype(npdt).ravel()
                    mine[n_local * k:] = 0
                    C.cast(g.allgather, _lib.ALLGATHER_FN)(None, g.workspace + 64, per, esz, None)
                    napply += 1
                    for r in range(world):
                        blk = buf[r * per:r * per + n_local * k].reshape(n_local, k).astype(np.float64)
                        A[r * n_local:(r + 1) * n_local, c0:c0 + k] = blk[:, :min(k, n - c0)]
                        assert np.all(buf[r * per + n_local * k:(r + 1) * per] == 0)       # every rank's flag row arrived
            elif g.apply:
                assert nb == 1
                woff = [64, 64 + ((n * k * esz + 63) // 64) * 64]
                xblk = _arr(g.workspace + woff[0], n * k, npdt).reshape(n, k)
                yblk = _arr(g.workspace + woff[1], n * k, npdt).reshape(n, k)
                A = np.zeros((n, n))
                for c0 in range(0, n, k):                        # the operator, k identity columns per application
                    xblk[...] = 0
                    for j in range(min(k, n - c0)):
                        xblk[c0 + j, j] = 1
                    C.cast(g.apply, _lib.APPLY_FN)(None, g.workspace + woff[0], g.workspace + woff[1], None)
                    napply += 1
                    A[:, c0:c0 + k] = yblk.astype(np.float64)[:, :min(k, n - c0)]
            else:
                A = _matrix(g.A + b * g.a_bstride * esz, n, n, g.lda, npdt).astype(np.float64)
            V0 = _matrix(g.V0 + b * g.v0_bstride * esz, n, k, g.ldv0, npdt)
            assert np.linalg.matrix_rank(V0.astype(np.float64)) == k, "start block must have full rank"
            w, S = np.linalg.eigh(0.5 * (A + A.T))
            sel = slice(0, k) if g.mode == 0 else slice(n - k, n)
            _arr(g.evals + b * g.evals_bstride * esz, k, npdt)[...] = w[sel].astype(npdt)
            _matrix(g.evecs + b * g.evecs_bstride * esz, n, k, g.ldv, npdt)[...] = S[:, sel].astype(npdt)
        if g.niter_out:
            g.niter_out[0] = 1
        if g.converged_out:
            g.converged_out[0] = 1
        if g.best_resid_out:
            g.best_resid_out[0] = 0.0
        if g.napply_out:
            g.napply_out[0] = napply
        return 0


class _NoDevice(object):
    def __init__(self, dev):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install(monkeypatch):
    """route the host layer to the stand-in (CPU tensors allowed, no CUDA stream / device handling)"""
    eng = StandInLibrary()
    monkeypatch.setattr(_lib, "lib", lambda: eng)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda dev: 0)
    monkeypatch.setattr(torch.cuda, "device", _NoDevice)
    from xitorch_b200._impls import symeig as simpl
    monkeypatch.setattr(simpl, "_start_block",
                        lambda kind, nb, n, neig, dtype, dev: torch.randn(
                            nb, n, neig, dtype=dtype, generator=torch.Generator().manual_seed(12421)))
    return eng
