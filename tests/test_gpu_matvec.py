"""GPU parity of the dense block matvec (C ABI `xt_block_matvec`) against torch fp64 matmul."""
import pytest
import torch

from xitorch_b200 import _dense
import xitorch_b200 as xt

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(A, X, E=None, Z=None):
    y = A.double() @ X.double()
    if E is not None:
        y = y - (X.double() if Z is None else Z.double()) * E.double().unsqueeze(-2)
    return y


def _tol(dtype):
    return {torch.float32: 2e-6, torch.bfloat16: 2e-6, torch.float64: 1e-13}[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("n,k", [(64, 1), (100, 3), (256, 8), (1000, 2), (1024, 16), (2048, 5), (4100, 8)])
@pytest.mark.parametrize("impl", [1, 2, 3, 4, 5, 256 + 1, 256 + 4, (32 << 16) + 256 + 3])
def test_matvec_square(dtype, n, k, impl):
    g = torch.Generator().manual_seed(n * 31 + k)
    A = torch.randn(n, n, generator=g).to(dtype)
    X = torch.randn(n, k, generator=g)
    vdt = torch.float64 if dtype == torch.float64 else torch.float32
    es = A.element_size()
    if (impl & 0xff) in (1, 3, 4, 5) and (n * es) % 16 != 0:
        pytest.skip("row stride not 16-byte aligned: TMA path not applicable")
    y = _dense.block_matvec(A.to(DEV), X.to(vdt).to(DEV), impl=impl)
    ref = _ref(A, X.to(vdt))
    scale = (A.double().abs() @ X.double().abs()).max().item()
    err = (y.double().cpu() - ref).abs().max().item()
    assert y.dtype == vdt and tuple(y.shape) == (n, k)
    assert err <= _tol(dtype) * scale, (err, scale)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_matvec_rect_batched_shift(dtype):
    g = torch.Generator().manual_seed(5)
    A = torch.randn(3, 200, 200, generator=g).to(dtype)
    X = torch.randn(3, 200, 4, generator=g).to(dtype)
    E = torch.randn(3, 4, generator=g).to(dtype)
    Z = torch.randn(3, 200, 4, generator=g).to(dtype)
    for impl in (1, 2, 3):
        y = _dense.block_matvec(A.to(DEV), X.to(DEV), E=E.to(DEV), Z=Z.to(DEV), impl=impl)
        assert (y.double().cpu() - _ref(A, X, E, Z)).abs().max().item() <= _tol(dtype) * 200
        y = _dense.block_matvec(A.to(DEV), X.to(DEV), E=E.to(DEV), impl=impl)
        assert (y.double().cpu() - _ref(A, X, E)).abs().max().item() <= _tol(dtype) * 200
    # broadcast A over a batch of X, rectangular A, adjoint
    A2 = torch.randn(96, 160, generator=g).to(dtype)
    X2 = torch.randn(2, 5, 160, 3, generator=g).to(dtype)
    y = _dense.block_matvec(A2.to(DEV), X2.to(DEV))
    assert tuple(y.shape) == (2, 5, 96, 3)
    assert (y.double().cpu() - A2.double() @ X2.double()).abs().max().item() <= _tol(dtype) * 200
    X3 = torch.randn(96, 2, generator=g).to(dtype)
    y = _dense.block_matvec(A2.to(DEV), X3.to(DEV), adjoint=True)
    assert (y.double().cpu() - A2.double().t() @ X3.double()).abs().max().item() <= _tol(dtype) * 200


def test_matvec_colslice_shapes():
    # the column-slice kernel (fp32, k in 5..16): ragged tiles / chunks, many tiles per CTA, batched
    g = torch.Generator().manual_seed(11)
    for (nb, rows, cols, k) in [(1, 1000, 1000, 8), (1, 20000, 600, 16), (3, 200, 200, 7), (1, 72, 4100, 12)]:
        A = torch.randn(nb, rows, cols, generator=g)
        X = torch.randn(nb, cols, k, generator=g)
        y = _dense.block_matvec(A.to(DEV), X.to(DEV), impl=4)
        ref = A.double() @ X.double()
        scale = (A.abs().double() @ X.abs().double()).max().item()
        assert (y.double().cpu() - ref).abs().max().item() <= 2e-6 * scale


def test_matvec_many_columns_and_tiles():
    # k > 16 (column groups), more tiles than SMs (persistent loop), ragged last tile
    g = torch.Generator().manual_seed(9)
    A = torch.randn(20000, 520, generator=g)
    X = torch.randn(520, 37, generator=g)
    y = _dense.block_matvec(A.to(DEV), X.to(DEV))
    ref = A.double() @ X.double()
    assert (y.double().cpu() - ref).abs().max().item() <= 2e-6 * (A.abs().double() @ X.abs().double()).max().item()


def test_linop_mm_uses_kernel_and_matches():
    g = torch.Generator().manual_seed(3)
    A = torch.randn(512, 512, generator=g, dtype=torch.float64)
    A = (A + A.t()) / 2
    op = xt.LinearOperator.m(A.to(DEV), is_hermitian=True)
    x = torch.randn(512, 7, generator=g, dtype=torch.float64).to(DEV)
    with torch.no_grad():
        y = op.mm(x)
        yv = op.mv(x[:, 0])
        yr = op.rmm(x)
    assert torch.allclose(y.cpu(), A @ x.cpu(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(yv.cpu(), A @ x[:, 0].cpu(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(yr.cpu(), y.cpu())
    # with autograd on, the differentiable library path is taken and gives the same numbers
    Ag = A.to(DEV).requires_grad_()
    y2 = xt.LinearOperator.m(Ag, is_hermitian=True).mm(x)
    assert y2.requires_grad and torch.allclose(y2.detach(), y, rtol=1e-12, atol=1e-12)


def test_full_size_linearity_property():
    # BASELINE size (N=16384, k=8): size-independent checks -- linearity and a row-sum identity
    n, k = 16384, 8
    g = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn(n, n, device=DEV, generator=g)
    X = torch.randn(n, k, device=DEV, generator=g)
    W = torch.randn(n, k, device=DEV, generator=g)
    y1 = _dense.block_matvec(A, X)
    y2 = _dense.block_matvec(A, W)
    y12 = _dense.block_matvec(A, X + 2 * W)
    assert (y12 - (y1 + 2 * y2)).abs().max().item() <= 5e-4 * y12.abs().max().item()
    ones = torch.ones(n, 1, device=DEV)
    rs = _dense.block_matvec(A, ones)[:, 0]
    assert torch.allclose(rs.double(), A.double().sum(dim=1), rtol=1e-4, atol=1e-2)
    # and the two kernels agree on a row sample
    yp = _dense.block_matvec(A[:256].contiguous(), X, impl=2)
    assert (yp - y1[:256]).abs().max().item() <= 2e-4 * y1.abs().max().item()


@pytest.mark.parametrize("n,rows,k", [(1024, 1024, 16), (1000, 700, 12), (4096, 300, 9), (2080, 2080, 16), (72, 200, 16)])
def test_matvec_tensor_core_layout(n, rows, k):
    """fp32 wide blocks on the tensor cores (error-compensated TF32, impl = 6): same tolerance as the SIMT kernels,
    ragged tiles / chunks, strided and batched X"""
    g = torch.Generator().manual_seed(n + rows + k)
    A = torch.randn(2, rows, n, generator=g)
    X = torch.randn(2, n, k, generator=g)
    y = _dense.block_matvec(A.to(DEV), X.to(DEV), impl=6)
    ref = A.double() @ X.double()
    scale = (A.double().abs() @ X.double().abs()).max().item()
    err = (y.double().cpu() - ref).abs().max().item()
    assert tuple(y.shape) == (2, rows, k)
    assert err <= 2e-6 * scale, (err, scale, err / scale)
    # agrees with the SIMT kernel to rounding level
    y3 = _dense.block_matvec(A.to(DEV), X.to(DEV), impl=3)
    assert (y - y3).abs().max().item() <= 4e-6 * scale
    # values spanning many orders of magnitude (hi/lo split under scaling)
    Xs = X * torch.logspace(-12, 12, k).reshape(1, 1, k)
    ys = _dense.block_matvec(A.to(DEV), Xs.to(DEV), impl=6)
    refs = A.double() @ Xs.double()
    rel = ((ys.double().cpu() - refs).abs() / (A.double().abs() @ Xs.double().abs())).max().item()
    assert rel <= 2e-6, rel


@pytest.mark.parametrize("nr,nc", [(128, 64), (1000, 1000), (4096, 4096), (148 * 128 + 5, 2048 + 40), (8192, 16384)])
def test_matvec_k16_tcgen05(nr, nc):
    """fp32 k = 16 on the fifth-generation tensor cores (impl = 7: tcgen05.mma kind::tf32, operands and accumulators in
    tensor memory, error-compensated 3xTF32): the SIMT kernels' tolerance, ragged tiles and ragged last chunk, shift and
    fused dots through the common row epilogue."""
    g = torch.Generator().manual_seed(nr + nc)
    A = torch.randn(nr, nc, generator=g)
    X = torch.randn(nc, 16, generator=g)
    y = _dense.block_matvec(A.to(DEV), X.to(DEV), impl=7)
    ref = A.double() @ X.double()
    bound = A.double().abs() @ X.double().abs()
    assert ((y.cpu().double() - ref).abs() / bound).max().item() <= 2e-6
    y3 = _dense.block_matvec(A.to(DEV), X.to(DEV), impl=3)
    assert ((y.cpu().double() - y3.cpu().double()).abs() / bound).max().item() <= 2e-6
    if nr == nc:
        E = torch.randn(16, generator=g)
        ye = _dense.block_matvec(A.to(DEV), X.to(DEV), E=E.to(DEV), impl=7)
        refe = ref - X.double() * E.double()
        assert ((ye.cpu().double() - refe).abs() / (bound + (X.double() * E.double()).abs())).max().item() <= 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("nr,nc,k", [(128, 64, 1), (300, 200, 3), (1000, 1000, 8), (4096, 2048, 16), (777, 4100, 5),
                                     (2048, 4096, 20)])
def test_matvec_transposed_access(dtype, nr, nc, k):
    """Y = A^T X by column strips (trans = 1): no transposed copy of A (reference linop.py:698-702 materialises mat^H)"""
    g = torch.Generator().manual_seed(nr * 7 + nc + k)
    A = torch.randn(nr, nc, generator=g, dtype=dtype)
    X = torch.randn(nr, k, generator=g, dtype=dtype)
    if (nc * A.element_size()) % 16 != 0:
        pytest.skip("row stride not a 16-byte multiple: the wrapper materialises the transpose")
    Ad, Xd = A.to(DEV), X.to(DEV)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    y = _dense.block_matvec(Ad, Xd, adjoint=True)
    torch.cuda.synchronize()
    assert torch.cuda.max_memory_allocated() - base < nr * nc * A.element_size() // 2      # no copy of A
    ref = A.double().t() @ X.double()
    bound = A.double().abs().t() @ X.double().abs()
    tol = 2e-6 if dtype == torch.float32 else 1e-13
    assert ((y.cpu().double() - ref).abs() / bound).max().item() <= tol


def test_matvec_transposed_batched_and_rmm():
    g = torch.Generator().manual_seed(3)
    A = torch.randn(3, 512, 384, generator=g)
    X = torch.randn(3, 512, 4, generator=g)
    y = _dense.block_matvec(A.to(DEV), X.to(DEV), adjoint=True)
    assert torch.allclose(y.cpu(), A.transpose(-2, -1) @ X, rtol=1e-4, atol=1e-4)
    import xitorch_b200 as xt
    M = torch.randn(640, 640, generator=g)
    op = xt.LinearOperator.m(M.to(DEV), is_hermitian=False)
    V = torch.randn(640, 7, generator=g)
    assert torch.allclose(op.rmm(V.to(DEV)).cpu(), M.t() @ V, rtol=1e-4, atol=1e-4)
    assert torch.allclose(op.H.mm(V.to(DEV)).cpu(), M.t() @ V, rtol=1e-4, atol=1e-4)
    assert torch.allclose(op.rmv(V[:, 0].to(DEV)).cpu(), M.t() @ V[:, 0], rtol=1e-4, atol=1e-4)
