"""The fused expansion kernel of the eigensolver (`expand_fused_kernel`, csrc/symeig.cu: lagged Ritz check, two rounds of
Gram-Schmidt against the whole basis with cross-CTA fp64 atomics and the kernel's own grid barrier, Cholesky-QR of the
new block, the new block column of the projected matrix) checked WITHOUT a GPU: the device code is cut out of the .cu
file and compiled for the host, G CTAs x 512 real threads (tools/emu_expand_fused.cpp).  Checked against numpy: the new
block is orthonormal, orthogonal to the basis and spans the projected W; T's new block column is V^T W; the Ritz
vectors, the residual maximum, the best-pair / stop bookkeeping; the accumulators are left clean for the next launch
(the launch is repeated on the same inputs and must give the same answer)."""
import os
import re
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SE_MAXK, PO_NCOPY = 16, 8


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("emu_fused")
    src = open(os.path.join(ROOT, "xitorch_b200", "csrc", "symeig.cu")).read()

    def cut(start, end):
        i0 = src.index(start)
        return src[i0:src.index(end, i0)]

    body = cut("struct EigCtl {", "__device__ __forceinline__ unsigned long long gtimer() {")
    body += cut("__device__ int chol_inverse_warp(", "// Zp = Z - V Cin  (Cin may be null)")
    body += cut("constexpr int PO_THREADS = 512;", "// Out[:, 0..p) = In(:, 0..m) * Sr")
    # the constructs a host compiler cannot take
    body = body.replace('asm volatile("fence.acq_rel.gpu;" ::: "memory");', "emu_fence();")
    body = body.replace("extern __shared__ __align__(16) unsigned char po_raw[];",
                        "unsigned char* po_raw = emu_dyn_smem();")
    body = re.sub(r"__shared__ (\w+) (\w+)\[(\d+)\];", r"\1* \2 = emu_shared<\1>(__COUNTER__, \3);", body)
    body = re.sub(r"__shared__ (\w+) (\w+);", r"\1& \2 = *emu_shared<\1>(__COUNTER__, 1);", body)
    assert "asm" not in body and "__shared__" not in body
    open(os.path.join(d, "fused_body.inc"), "w").write(body)
    shutil.copy(os.path.join(ROOT, "tools", "emu_expand_fused.cpp"), os.path.join(d, "emu.cpp"))
    exe = os.path.join(d, "emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-o", exe,
                           os.path.join(d, "emu.cpp")], cwd=d)
    return exe, str(d)


def _problem(n, k, m, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    A = (A + A.T) / np.sqrt(2 * n) + np.diag(np.linspace(1, 10, n))
    V, _ = np.linalg.qr(rng.standard_normal((n, m)))
    return A, V


def _blocks(M, k):
    """(n, m) -> block layout [block][n][k] flattened"""
    n, m = M.shape
    return np.concatenate([M[:, b * k:(b + 1) * k].reshape(-1) for b in range(m // k)])


def _run(emu, n, k, m, G, tvbytes, stage_v, with_ritz, seed=0, launches=2, min_eps=1e-30, zero_w=False):
    exe, d = emu
    A, V = _problem(n, k, m, seed)
    dt = np.float32 if tvbytes == 4 else np.float64
    Vt = V.astype(dt).astype(np.float64)                       # what the kernel sees
    AV = (A @ Vt).astype(dt).astype(np.float64)
    if zero_w:
        AV[:, m - k:] = 0.0
    rz_m = m - k if with_ritz else 0                           # the lagged check concerns the basis one block ago
    nev = k
    if rz_m:
        Tz = Vt[:, :rz_m].T @ AV[:, :rz_m]
        w, S = np.linalg.eigh(0.5 * (Tz + Tz.T))
        S, theta = S[:, :nev], w[:nev]
    else:
        S, theta = np.zeros((1, 1)), np.zeros(1)
    fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("11i", n, k, m, G, tvbytes, stage_v, rz_m, nev if rz_m else 0, 0, 5, launches))
        f.write(struct.pack("d", min_eps))
        np.concatenate([_blocks(Vt, k), np.zeros(n * k)]).astype(np.float64).tofile(f)
        _blocks(AV, k).astype(np.float64).tofile(f)
        S.astype(np.float64).reshape(-1).tofile(f)
        theta.astype(np.float64).tofile(f)
    subprocess.run([exe, fin, fout], check=True, timeout=900)
    raw = np.fromfile(fout, dtype=np.float64)
    acc_stride = (m + SE_MAXK) * k
    sizes = [n * k, m * m, 2 * PO_NCOPY * acc_stride, 2 * n * k, 2 * SE_MAXK, 8]
    assert raw.size == sum(sizes)
    parts, off = [], 0
    for s in sizes:
        parts.append(raw[off:off + s])
        off += s
    Q, T, acc, X, evs, ctl = parts
    return dict(A=A, V=Vt, AV=AV, W=AV[:, m - k:], Q=Q.reshape(n, k), T=T.reshape(m, m),
                acc=acc.reshape(2, PO_NCOPY, acc_stride), X=X.reshape(2, n, k), evals=evs.reshape(2, SE_MAXK),
                ctl=dict(zip(("done", "converged", "niter", "best_slot", "best_resid", "breakdown", "bar_count",
                              "resmax_bits"), ctl)),
                S=S, theta=theta, rz_m=rz_m, k=k, m=m, eps=(6e-8 if tvbytes == 4 else 1.2e-16))


def _check(r):
    k, m, V, W, Q, eps = r["k"], r["m"], r["V"], r["W"], r["Q"], r["eps"]
    n = V.shape[0]
    tol = 50 * eps * max(1.0, np.abs(W).max()) * np.sqrt(n)
    assert r["ctl"]["breakdown"] == 0 and r["ctl"]["bar_count"] > 0 and r["ctl"]["bar_count"] % 2 == 0   # monotone: 2 barriers x CTAs
    # new block column of the projected matrix (diagonal block symmetrised)
    C = V.T @ W
    c0 = m - k
    Cref = C.copy()
    Cref[c0:] = 0.5 * (C[c0:] + C[c0:].T)
    assert np.abs(r["T"][:, c0:] - Cref).max() <= tol
    assert np.abs(r["T"][c0:, :] - Cref.T).max() <= tol
    # Q: orthonormal, orthogonal to V, same range as the twice-projected W
    assert np.abs(Q.T @ Q - np.eye(k)).max() <= max(tol, 200 * eps)
    assert np.abs(V.T @ Q).max() <= max(tol, 200 * eps)
    Wp = W - V @ (V.T @ W)
    Wp = Wp - V @ (V.T @ Wp)
    assert np.abs(Q @ (Q.T @ Wp) - Wp).max() <= 100 * tol
    # upper-triangular relation  W' = Q R  with positive diagonal (Cholesky-QR)
    Rm = Q.T @ Wp
    assert np.abs(np.tril(Rm, -1)).max() <= 100 * tol and np.all(np.diag(Rm) > 0)
    # set 0 of the accumulators is clean for the next launch
    assert np.all(r["acc"][0] == 0.0)
    if r["rz_m"]:
        rz_m, S, theta = r["rz_m"], r["S"], r["theta"]
        Xref = V[:, :rz_m] @ S
        res = r["AV"][:, :rz_m] @ S - Xref * theta
        slot = 1                                              # best_slot starts at 0: the candidate goes to slot 1
        assert np.abs(r["X"][slot] - Xref).max() <= tol
        assert r["ctl"]["best_slot"] == slot and r["ctl"]["niter"] == 4
        assert abs(r["ctl"]["best_resid"] - np.abs(res).max()) <= 1e-5 * np.abs(res).max() + tol
        assert np.allclose(r["evals"][slot, :k], theta)
        assert r["ctl"]["resmax_bits"] == 0


CASES = [  # n, k, m, G, bytes, stage_v
    (150, 4, 12, 3, 8, 1), (150, 4, 12, 3, 4, 1), (130, 8, 24, 2, 4, 1), (130, 8, 24, 2, 8, 0),
    (97, 3, 9, 3, 8, 1), (97, 6, 18, 2, 4, 0), (160, 16, 48, 2, 4, 1), (100, 8, 8, 3, 8, 1),
]


@pytest.mark.parametrize("n,k,m,G,tv,stage", CASES)
@pytest.mark.parametrize("with_ritz", [False, True])
def test_fused_expansion_emulated(emu, n, k, m, G, tv, stage, with_ritz):
    if with_ritz and m == k:
        pytest.skip("no earlier basis to check")
    _check(_run(emu, n, k, m, G, tv, stage, with_ritz, seed=n + k))


def test_stop_flag_and_breakdown(emu):
    # a huge tolerance: the lagged Ritz check must raise the stop flag
    r = _run(emu, 120, 4, 12, 2, 8, 1, True, seed=3, launches=1, min_eps=1e6)
    assert r["ctl"]["converged"] == 1 and r["ctl"]["done"] == 1
    # W = 0 exactly: the Gram matrix of the projected block vanishes -> breakdown flag, zero block, stop
    # (the pivot test is relative to the largest diagonal entry, so rounding-level blocks are still normalised)
    r2 = _run(emu, 60, 4, 12, 2, 8, 1, False, seed=4, launches=1, zero_w=True)
    assert r2["ctl"]["breakdown"] == 1 and r2["ctl"]["done"] == 1 and np.all(r2["Q"] == 0.0)
