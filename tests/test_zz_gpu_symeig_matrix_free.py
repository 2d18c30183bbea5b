"""GPU parity of the matrix-free eigensolver path (`xt_symeig_args.apply`, row f3 of SURVEY.md 8f): an operator known
only through `_mv` must give the eigenpairs the dense kernels give for the same matrix (same engine, same start block:
only the operator application differs) and satisfy the residual identity in fp64.

Its host side is also tested on CPU (tests/test_symeig_matrix_free_host.py) and the engine branch by running the engine's
source as a host build (tests/test_engine_emulation.py).  Green on hardware since round 2 (the round-1 xfail mark is gone).
"""
import pytest
import torch

import xitorch_b200 as xt
from xitorch_b200.linalg import symeig, svd

pytestmark = [pytest.mark.gpu]


def _herm(n, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, n, generator=g, dtype=torch.float64)
    a = (a + a.T) * (0.05 / (2 * n) ** 0.5) + torch.diag(20 + 10 * torch.linspace(0, 1, n, dtype=torch.float64))
    idx = torch.arange(16)
    a[idx, idx] = 1.0 + idx.double()
    return a.to(dtype).cuda()


class UserOperator(xt.LinearOperator):
    def __init__(self, mat):
        super().__init__(shape=mat.shape, is_hermitian=True, dtype=mat.dtype, device=mat.device)
        self.mat = mat

    def _mv(self, x):
        return torch.matmul(self.mat, x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix=""):
        return [prefix + "mat"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("method", ["davidson", "lanczos"])
def test_matrix_free_matches_dense_engine(dtype, method):
    n, k = 2048, 8
    mat = _herm(n, dtype)
    info_d, info_f = {}, {}
    with torch.no_grad():
        ev_d, _ = symeig(xt.LinearOperator.m(mat, is_hermitian=True), neig=k, method=method, min_eps=1e-4, info=info_d)
        ev_f, X = symeig(UserOperator(mat), neig=k, method=method, min_eps=1e-4, matrix_free=True, info=info_f)
    assert info_f["converged"] and info_f["napply"] > 0
    ref = torch.linalg.eigvalsh(mat.double())[:k]
    assert ((ev_f.double() - ref).abs() / ref.abs()).max().item() <= 1e-5          # north_star tolerance
    assert ((ev_f.double() - ev_d.double()).abs() / ref.abs()).max().item() <= 1e-5
    assert abs(info_f["niter"] - info_d["niter"]) <= 1                             # same Krylov space
    resid = mat.double() @ X.double() - X.double() * ev_f.double()
    assert resid.abs().max().item() <= 20 * 1e-4
    assert (X.double().T @ X.double() - torch.eye(k, device="cuda", dtype=torch.float64)).abs().max().item() <= 1e-4


def test_automatic_beyond_materialisation_limit():
    # n > 16384: materialising a composite operator used to be refused; now it is applied through the hook
    n, k = 16640, 4
    d = 20 + 10 * torch.linspace(0, 1, n, dtype=torch.float32, device="cuda")
    d[:8] = 1.0 + torch.arange(8, dtype=torch.float32, device="cuda")              # the make_herm spectrum

    class DiagPlusRank1(xt.LinearOperator):
        def __init__(self):
            super().__init__(shape=(n, n), is_hermitian=True, dtype=torch.float32, device=d.device)
            self.u = torch.ones(n, device="cuda") / n ** 0.5

        def _mv(self, x):
            return d * x - 0.5 * self.u * (x * self.u).sum(-1, keepdim=True)

        def _getparamnames(self, prefix=""):
            return []

    op = DiagPlusRank1()
    with torch.no_grad():
        ev, X = symeig(op, neig=k, method="davidson", min_eps=1e-4)
        resid = op.mm(X) - X * ev
    assert resid.abs().max().item() <= 20 * 1e-4
    assert ev[0].item() < d[0].item() and (ev[1:] > ev[:-1]).all()                 # interlacing of a rank-1 downdate


def test_svd_of_matrix_free_rectangular_operator():
    g = torch.Generator().manual_seed(4)
    B = (torch.randn(3000, 1024, generator=g, dtype=torch.float64) / 55).cuda()
    B[:8, :8] += torch.diag(torch.arange(10, 2, -1, dtype=torch.float64, device="cuda"))

    class Rect(xt.LinearOperator):
        def __init__(self):
            super().__init__(shape=B.shape, dtype=B.dtype, device=B.device)

        def _mv(self, x):
            return torch.matmul(B, x.unsqueeze(-1)).squeeze(-1)

        def _rmv(self, y):
            return torch.matmul(B.T, y.unsqueeze(-1)).squeeze(-1)

        def _getparamnames(self, prefix=""):
            return []

    with torch.no_grad():
        u, s, vh = svd(Rect(), k=4, mode="uppest", method="davidson", matrix_free=True, min_eps=1e-8)
    sref = torch.linalg.svdvals(B)[:4]
    assert ((s.sort(descending=True).values - sref).abs() / sref).max().item() <= 1e-5
    assert (B @ vh.transpose(-2, -1) - u * s.unsqueeze(-2)).abs().max().item() <= 1e-5


def test_generalized_matrix_free():
    n, k = 1024, 4
    A = _herm(n, torch.float64, seed=2)
    g = torch.Generator().manual_seed(3)
    Mh = torch.randn(n, n, generator=g, dtype=torch.float64).cuda() / n ** 0.5
    Mm = Mh @ Mh.T * 0.1 + torch.eye(n, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        ev, X = symeig(UserOperator(A), neig=k, M=xt.LinearOperator.m(Mm, is_hermitian=True), method="davidson",
                       matrix_free=True, min_eps=1e-7)
    assert (A @ X - Mm @ X * ev).abs().max().item() <= 1e-5
    assert (X.T @ Mm @ X - torch.eye(k, dtype=torch.float64, device="cuda")).abs().max().item() <= 1e-8


def test_hermitian_check_kernel_on_hardware():
    """`xt_hermitian_check` (csrc/linop.cu; logic verified on the host build, tests/test_engine_emulation.py) against
    torch.allclose on the GPU, and LinearOperator.m through it"""
    from xitorch_b200 import _dense
    g = torch.Generator().manual_seed(1)
    for dtype in (torch.float32, torch.float64):
        for n in (1, 33, 1000, 4096):
            a = torch.randn(n, n, generator=g, dtype=dtype).cuda()
            sym = a + a.T
            assert _dense.hermitian_check(sym)
            if n > 1:
                bad = sym.clone()
                bad[n - 1, 0] += 1.0
                assert not _dense.hermitian_check(bad)
                assert _dense.hermitian_check(bad) == bool(torch.allclose(bad, bad.T))
    batch = torch.randn(3, 200, 200, generator=g).cuda()
    batch = batch + batch.transpose(-2, -1)
    assert _dense.hermitian_check(batch)
    batch[1, 5, 150] += 0.5
    assert not _dense.hermitian_check(batch)
    with pytest.raises(RuntimeError, match="hermitian"):
        xt.LinearOperator.m(batch[1], is_hermitian=True)
    assert xt.LinearOperator.m(batch[0]).is_hermitian
