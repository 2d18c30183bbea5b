"""
Synthetic problem generators for the BASELINE.json configs.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything is seeded on the CPU
generator so that the CPU oracle and the GPU path see bit-identical inputs.
"""
import math
import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_herm(n: int, neig: int = 8, dtype=torch.float32, seed: int = 123) -> torch.Tensor:
    """C2/C5 matrix: A = (G+G^T) * 0.05/sqrt(2n) + diag(d), d = 20 + 10*linspace(0,1,n),
    d[:2*neig] = 1 + arange(2*neig)  (well separated lowest eigenvalues ~1..2*neig)."""
    G = torch.randn(n, n, generator=_gen(seed), dtype=torch.float32)
    A = (G + G.t()) * (0.05 / math.sqrt(2.0 * n))
    d = 20.0 + 10.0 * torch.linspace(0, 1, n, dtype=torch.float32)
    d[:2 * neig] = 1.0 + torch.arange(2 * neig, dtype=torch.float32)
    A.diagonal().add_(d)
    return A.to(dtype)


def make_slow_herm(n: int, dtype=torch.float32, seed: int = 123) -> torch.Tensor:
    """slow-converging shifted GOE matrix A = (G+G^T)/sqrt(2n) + 3 I (SURVEY.md 8d, C2 sustained)."""
    G = torch.randn(n, n, generator=_gen(seed), dtype=torch.float32)
    A = (G + G.t()) / math.sqrt(2.0 * n)
    A.diagonal().add_(3.0)
    return A.to(dtype)


def make_spd_c1(n: int = 256, min_eival: float = 0.2, max_eival: float = 1.0, seed: int = 123):
    """C1 matrix, the asv generator create_random_square_matrix(is_hermitian=True)
    (/root/reference/xitorch/_utils/tensor.py:46-76): Q^T diag(linspace) Q symmetrised, fp64,
    Q from QR of a seeded randn."""
    dtype = torch.float64
    eivals = torch.diag_embed(torch.linspace(min_eival, max_eival, n, dtype=dtype))
    torch.manual_seed(seed)      # create_random_square_matrix seeds (tensor.py:60-61)
    torch.manual_seed(seed)      # ... and create_random_ortho_matrix seeds again (tensor.py:72-73)
    a = torch.randn((n, n), dtype=dtype)
    q, _ = torch.linalg.qr(a)
    mat = torch.matmul(torch.matmul(q.transpose(-2, -1), eivals), q)
    return (mat + mat.transpose(-2, -1)) * 0.5


def make_nonsym_c3(nbatch: int, n: int = 4096, seed0: int = 0, dtype=torch.bfloat16):
    """C3 systems: A_b = I + 0.3*randn(n,n,seed=b)/sqrt(n) rounded to `dtype`, b_b = randn(n,1)."""
    As, Bs = [], []
    for b in range(nbatch):
        g = _gen(seed0 + b)
        A = torch.randn(n, n, generator=g, dtype=torch.float32) * (0.3 / math.sqrt(n))
        A.diagonal().add_(1.0)
        As.append(A.to(dtype))
        Bs.append(torch.randn(n, 1, generator=g, dtype=torch.float32))
    return torch.stack(As), torch.stack(Bs)


def make_rootfinder_c4(n: int = 8192, seed: int = 0, dtype=torch.float32):
    """C4: f(y, A) = tanh(A @ y + 0.1) + y / 2 with A = 0.1*randn(n,n)/sqrt(n), y0 = zeros(n,1)."""
    A = 0.1 * torch.randn(n, n, generator=_gen(seed), dtype=torch.float32) / math.sqrt(n)
    return A.to(dtype), torch.zeros(n, 1, dtype=dtype)


def make_herm_row_block(n: int, neig: int, lo: int, hi: int, device, bs: int = 2048) -> torch.Tensor:
    """rows [lo, hi) of a make_herm-like matrix of order n (C5: 16 GiB at n = 65536, never assembled in one place),
    generated on `device` from per-block-pair seeds so that every rank of a row partition draws a consistent,
    symmetric-by-construction matrix:  G = U + U^T with the (bi, bj) block (bi <= bj) of U from seed 1000003*bi + bj."""
    dev = torch.device(device)
    out = torch.empty(hi - lo, n, dtype=torch.float32, device=dev)
    for bi in range(lo // bs, (hi + bs - 1) // bs):
        r0, r1 = max(lo, bi * bs), min(hi, (bi + 1) * bs)
        for bj in range((n + bs - 1) // bs):
            c0, c1 = bj * bs, min(n, (bj + 1) * bs)
            a, b = min(bi, bj), max(bi, bj)
            g = torch.Generator(device=dev)
            g.manual_seed(1000003 * a + b)
            blk = torch.randn(bs, bs, generator=g, device=dev)
            if bi == bj:
                blk = blk + blk.t()
            elif bi > bj:
                blk = blk.t()
            sub = blk[r0 - bi * bs: r1 - bi * bs, : c1 - c0]
            out[r0 - lo: r1 - lo, c0:c1] = sub * (0.05 / (2.0 * n) ** 0.5) * (1.0 if bi == bj else 2.0 ** 0.5)
    d = 20.0 + 10.0 * torch.linspace(0, 1, n, device=dev)
    d[:2 * neig] = 1.0 + torch.arange(2 * neig, device=dev, dtype=torch.float32)
    idx = torch.arange(lo, hi, device=dev)
    out[idx - lo, idx] += d[lo:hi]
    return out
