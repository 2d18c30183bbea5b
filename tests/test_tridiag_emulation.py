"""The register-resident Householder tridiagonalisation of the eigensolver (`tridiag_regs`, csrc/symeig.cu) checked WITHOUT a
GPU: its body is cut out of the .cu file and compiled for the host with 256 real threads standing in for the CUDA threads
(std::barrier = __syncthreads; tools/emu_tridiag.cpp).  Checks: Q^T A Q = tridiag(d, e) with Q rebuilt from the stored
reflectors, Q orthogonal, and the spectrum of (d, e) against numpy's eigvalsh -- for every template instantiation and the
edge sizes around them."""
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
    d = tmp_path_factory.mktemp("emu")
    src = open(os.path.join(ROOT, "xitorch_b200", "csrc", "symeig.cu")).read()
    i0 = src.index("template <int MR, int MC>\n__device__ __noinline__ void tridiag_regs(")
    i1 = src.index("// As: work matrix (m x lds, destroyed).  Outputs: lam[nev] ascending")
    open(os.path.join(d, "tridiag_body.inc"), "w").write(src[i0:i1])
    shutil.copy(os.path.join(ROOT, "tools", "emu_tridiag.cpp"), os.path.join(d, "emu.cpp"))
    exe = os.path.join(d, "emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-o", exe, os.path.join(d, "emu.cpp")], cwd=d)
    return exe


@pytest.mark.parametrize("m", [1, 2, 3, 8, 32, 33, 64, 65, 96, 97, 104, 105, 128])
def test_tridiag_regs_emulated(emu, m):
    out = subprocess.run([emu, str(m)], capture_output=True, text=True, timeout=300, check=True).stdout.split("\n")
    mm, lds = (int(x) for x in out[0].split())
    assert mm == m
    A = np.array(out[1].split(), dtype=np.float64).reshape(m, m)
    d = np.array(out[2].split(), dtype=np.float64)
    e = np.array(out[3].split(), dtype=np.float64)
    tau = np.array(out[4].split(), dtype=np.float64)
    As = np.array(out[5].split(), dtype=np.float64).reshape(m, lds)
    T = np.diag(d) + np.diag(e[:m - 1], 1) + np.diag(e[:m - 1], -1)
    # Q = H_0 H_1 ... H_{m-3},  H_j = I - tau_j v_j v_j^T,  v_j = (0, ..., 0, 1 at j+1, As[j+2:, j])
    Q = np.eye(m)
    for j in range(max(m - 2, 0)):
        v = np.zeros(m)
        v[j + 1] = 1.0
        v[j + 2:] = As[j + 2:, j]
        assert As[j + 1, j] == 1.0
        Q = Q @ (np.eye(m) - tau[j] * np.outer(v, v))
    scale = max(1.0, np.abs(A).max() * m)
    assert np.abs(Q.T @ Q - np.eye(m)).max() <= 1e-13 * m
    assert np.abs(Q.T @ A @ Q - T).max() <= 1e-14 * scale
    assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A)).max() <= 1e-13 * scale
