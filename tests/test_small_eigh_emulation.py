"""The WHOLE one-CTA eigensolver of the Rayleigh-Ritz step (`eig_extreme_device`, csrc/symeig.cu: tridiagonalisation
in registers or shared memory, Sturm multisection, inverse iteration with pivoting, cluster Gram-Schmidt, Rayleigh
quotient refinement, paired back-transformation) checked WITHOUT a GPU: the device code is cut out of the .cu file and
compiled for the host, 256 real threads standing in for the CUDA threads (tools/emu_small_eigh.cpp).  Checks against
numpy's eigh of the same matrix: eigenvalues, residuals and orthonormality of the returned vectors -- for every
register-tile instantiation, both ends of the spectrum, and clustered / exactly degenerate / graded spectra."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("emu_eigh")
    src = open(os.path.join(ROOT, "xitorch_b200", "csrc", "symeig.cu")).read()
    a0 = src.index("struct EigPlan {")
    a1 = src.index("// ---------------------------------------------------------------------------- register-resident "
                   "tridiagonalisation")
    b0 = src.index("template <int MR, int MC>\n__device__ __noinline__ void tridiag_regs(")
    b1 = src.index("// T[:, new block] = C (and its transpose); then the nev extreme eigenpairs of T:")
    assert a0 < a1 < b0 < b1
    open(os.path.join(d, "eig_body.inc"), "w").write(src[a0:a1] + "\n" + src[b0:b1])
    shutil.copy(os.path.join(ROOT, "tools", "emu_small_eigh.cpp"), os.path.join(d, "emu.cpp"))
    exe = os.path.join(d, "emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-o", exe,
                           os.path.join(d, "emu.cpp")], cwd=d)
    return exe


def _run(emu, m, nev, mode, kind, debug=0):
    out = subprocess.run([emu, str(m), str(nev), str(mode), str(kind), str(debug)], capture_output=True, text=True,
                         timeout=600, check=True).stdout.split("\n")
    if out[0].startswith("noplan"):
        return None
    mm, nn, as_in_smem, slots = (int(x) for x in out[0].split())
    assert (mm, nn) == (m, nev)
    A = np.array(out[1].split(), dtype=np.float64).reshape(m, m)
    lam = np.array(out[2].split(), dtype=np.float64)
    Y = np.array(out[3].split(), dtype=np.float64).reshape(m, nev)
    return A, lam, Y


def _check(A, lam, Y, mode, kind):
    m, nev = Y.shape
    w = np.linalg.eigvalsh(A)
    ref = w[:nev] if mode == 0 else w[m - nev:]
    scale = max(1.0, np.abs(w).max())
    assert np.all(np.diff(lam) >= -1e-12 * scale)                       # ascending
    assert np.abs(lam - ref).max() <= 2e-13 * scale * max(1, m // 8)
    assert np.abs(Y.T @ Y - np.eye(nev)).max() <= 1e-10                 # orthonormal, also inside clusters
    resid = np.abs(A @ Y - Y * lam).max()
    # clustered eigenvalues (1e-9 apart): any orthonormal basis of the cluster is acceptable, residual ~ cluster width
    assert resid <= (5e-9 if kind == 1 else 1e-11) * scale * max(1, m // 8)


CASES = [(4, 2), (8, 8), (16, 8), (32, 8), (33, 8), (48, 16), (64, 8), (65, 8), (88, 8), (96, 16), (104, 8),
         (105, 8), (128, 8), (128, 16)]


@pytest.mark.parametrize("m,nev", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_random_spectrum(emu, m, nev, mode):
    A, lam, Y = _run(emu, m, nev, mode, 0)
    _check(A, lam, Y, mode, 0)


@pytest.mark.parametrize("kind", [1, 2, 3])
@pytest.mark.parametrize("m,nev", [(24, 8), (64, 8), (96, 12), (104, 16)])
def test_clustered_degenerate_graded(emu, m, nev, kind):
    A, lam, Y = _run(emu, m, nev, 0, kind)
    _check(A, lam, Y, 0, kind)


@pytest.mark.parametrize("m,nev", [(40, 8), (104, 8)])
def test_shared_memory_tridiagonalisation_and_unpaired_backtransform(emu, m, nev):
    # the debug switches select the alternative code paths (XT_EIG_DEBUG = 1 | 2): same answers
    A, lam, Y = _run(emu, m, nev, 0, 0, debug=3)
    _check(A, lam, Y, 0, 0)


def test_larger_than_register_tiles(emu):
    # m > 128: shared-memory tridiagonalisation, generic back-transformation
    res = _run(emu, 144, 8, 0, 0)
    assert res is not None
    _check(*res, 0, 0)
