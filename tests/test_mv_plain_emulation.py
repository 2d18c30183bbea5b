"""The plain-load block matvec (`mv_plain_kernel`, csrc/matvec.cu -- what runs when the TMA kernels cannot take the
operands) checked WITHOUT a GPU: kernel and argument struct are cut out of the .cu file and run on host threads
(tools/emu_engine).  Against numpy: Y = A X - Z diag(E) for fp32 / fp64 / bf16 operators, batches with shared or
per-item A, padded leading dimensions, k = 1 .. 16, ragged last tile, grid smaller than the tile count, and the fused
per-tile partial dot products the solvers consume."""
import os
import re
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MV_MAXK = 16


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("emu_mv")
    src = open(os.path.join(ROOT, "xitorch_b200", "csrc", "matvec.cu")).read()
    i0 = src.index("struct MvDev {")
    body = src[i0:src.index("};", i0) + 2] + "\n"
    k0 = src.index("template <typename TA, typename TV>\n__global__ void __launch_bounds__(256)\nmv_plain_kernel(")
    body += src[k0:src.index("// ============================================================================ launch", k0)]
    body = body.replace("__shared__ double dscr[8][2][MV_MAXK];",
                        "double (*dscr)[2][MV_MAXK] = reinterpret_cast<double (*)[2][MV_MAXK]>("
                        "emu_shared<double>(__COUNTER__, 8 * 2 * MV_MAXK));")
    assert "__shared__" not in body and "asm" not in re.sub(r"//.*", "", body)
    open(os.path.join(d, "mv_plain_body.inc"), "w").write(body)
    exe = os.path.join(d, "emu")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas",
                           "-I", os.path.join(ROOT, "tools", "emu_engine"), "-I", os.path.join(ROOT, "include"), "-I", str(d),
                           "-o", exe, os.path.join(ROOT, "tools", "emu_engine", "emu_mv_plain.cpp")])
    return exe, str(d)


def _bf16(a):
    u = a.astype(np.float32).view(np.uint32) & 0xFFFF0000
    return u.view(np.float32).astype(np.float64)


CASES = [  # dtype code, nb, nrows, ncols, k, lda pad, ldx pad, ldy pad, tile_rows, E, Z, U, A batched, grid
    (0, 1, 70, 50, 8, 0, 0, 0, 32, 0, 0, 0, 0, 3), (2, 1, 50, 50, 8, 3, 0, 0, 32, 1, 0, 1, 0, 2),
    (2, 2, 45, 45, 3, 0, 2, 1, 16, 1, 1, 1, 1, 4), (0, 3, 33, 40, 16, 1, 0, 0, 33, 0, 0, 1, 0, 1),
    (1, 2, 64, 64, 1, 0, 0, 0, 24, 1, 0, 0, 1, 5), (2, 1, 9, 130, 5, 0, 0, 0, 128, 0, 0, 0, 0, 7),
]


@pytest.mark.parametrize("case", CASES)
def test_mv_plain_emulated(emu, case):
    exe, d = emu
    code, nb, nrows, ncols, k, pa, px, py, tile_rows, has_e, has_z, has_u, a_batched, grid = case
    rng = np.random.default_rng(sum(case))
    lda, ldx, ldy = ncols + pa, k + px, k + py
    cast = {0: lambda a: a.astype(np.float32).astype(np.float64), 1: _bf16, 2: lambda a: a}[code]
    vcast = (lambda a: a) if code == 2 else (lambda a: a.astype(np.float32).astype(np.float64))
    A = cast(rng.standard_normal(((nb if a_batched else 1), nrows, lda)))
    X = vcast(rng.standard_normal((nb, ncols, ldx)))
    E = vcast(rng.standard_normal((nb, k)))
    Z = vcast(rng.standard_normal((nb, nrows, k)))
    U = vcast(rng.standard_normal((nb, nrows, k)))
    fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("14i", code, nb, nrows, ncols, k, lda, ldx, ldy, tile_rows, has_e, has_z, has_u, a_batched,
                            grid))
        A.tofile(f)
        X.tofile(f)
        if has_e:
            E.tofile(f)
        if has_z:
            Z.tofile(f)
        if has_u:
            U.tofile(f)
    subprocess.run([exe, fin, fout], check=True, timeout=600)
    raw = np.fromfile(fout, dtype=np.float64)
    tpb = (nrows + tile_rows - 1) // tile_rows
    Y = raw[:nb * nrows * ldy].reshape(nb, nrows, ldy)
    dots = raw[nb * nrows * ldy:].reshape(nb * tpb, 2, MV_MAXK)
    eps = 1.2e-16 if code == 2 else 6e-8
    for b in range(nb):
        Ab = A[b if a_batched else 0][:, :ncols]
        ref = Ab @ X[b][:, :k]
        if has_e:
            shift = Z[b] if has_z else X[b][:nrows, :k]          # Z defaults to X (square operators)
            ref = ref - shift * E[b]
        scale = np.abs(Ab) @ np.abs(X[b][:, :k]) + 1
        assert np.all(np.abs(Y[b][:, :k] - ref) <= 8 * eps * np.sqrt(ncols) * scale), case
        assert np.all(Y[b][:, k:] == -7.0)                        # padding columns untouched
        for t in range(tpb):
            rows = slice(t * tile_rows, min(nrows, (t + 1) * tile_rows))
            y = Y[b][rows, :k]
            assert np.allclose(dots[b * tpb + t, 1, :k], (y * y).sum(0), rtol=1e-10, atol=1e-12)
            if has_u:
                assert np.allclose(dots[b * tpb + t, 0, :k], (U[b][rows] * y).sum(0), rtol=1e-10, atol=1e-12)
