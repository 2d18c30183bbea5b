"""
Host wrapper of the dense block-matvec kernel (`xt_block_matvec`): the accelerated body of
`MatrixLinearOperator._mm/_mv/_rmm/_rmv` (reference: torch.matmul,
/root/reference/xitorch/_core/linop.py:692-702).  Batch broadcasting follows torch.matmul.
"""
from typing import Optional, Tuple

import torch

from xitorch_b200 import _lib
from xitorch_b200._utils import bcast_dims

_SUPPORTED = (torch.float32, torch.bfloat16, torch.float64)


def supports(mat: torch.Tensor, x: torch.Tensor) -> bool:
    return (mat.is_cuda and x.is_cuda and mat.dtype in _SUPPORTED and not x.is_complex()
            and mat.dim() >= 2 and x.dim() >= 2 and x.shape[-1] >= 1 and mat.shape[-1] >= 1 and mat.shape[-2] >= 1)


def flatten_batch(mat: torch.Tensor, batch: Tuple[int, ...]):
    """(mat3d, a_bstride) such that item b of the flattened broadcast batch reads mat3d at b * a_bstride."""
    p, q = mat.shape[-2:]
    nb = 1
    for s in batch:
        nb *= s
    ba = tuple(mat.shape[:-2])
    nba = 1
    for s in ba:
        nba *= s
    if nba == 1:
        m2 = mat.reshape(p, q)
        if m2.stride(-1) != 1:
            m2 = m2.contiguous()
        return m2, 0, m2.stride(0)
    if tuple([1] * (len(batch) - len(ba)) + list(ba)) != tuple(batch):
        mat = mat.expand(*batch, p, q)
    m3 = mat.reshape(nb, p, q)
    if m3.stride(-1) != 1:
        m3 = m3.contiguous()
    return m3, m3.stride(0), m3.stride(1)


def block_matvec(mat: torch.Tensor, x: torch.Tensor, adjoint: bool = False,
                 E: Optional[torch.Tensor] = None, Z: Optional[torch.Tensor] = None,
                 impl: int = 0) -> torch.Tensor:
    """``mat @ x`` (or ``mat^T @ x``), optionally ``- Z * E`` with ``E (*B, k)``, ``Z (*B, p, k)`` (default x)."""
    _lib.require_cuda(mat, "the dense block matvec")
    # mat^T @ x: the transposed-access kernel reads mat once, by column strips (fp32 / fp64, no shift); anything else
    # materialises the transpose as the reference's `.H` does for dense operators (linop.py:390-391)
    trans = bool(adjoint and E is None and mat.dtype in (torch.float32, torch.float64)
                 and (mat.shape[-1] * mat.element_size()) % 16 == 0)
    if adjoint and not trans:
        mat = mat.transpose(-2, -1).contiguous()
    p, q = mat.shape[-2:]                       # the stored matrix: p rows, q columns
    nin, nout = (p, q) if trans else (q, p)     # rows of x / of the result
    if x.shape[-2] != nin:
        raise RuntimeError("block_matvec: shape mismatch %s%s @ %s" % (tuple(mat.shape), "^T" if trans else "",
                                                                       tuple(x.shape)))
    k = x.shape[-1]
    vdt = _lib.vec_dtype(mat.dtype)
    out_dtype = x.dtype
    batch = tuple(bcast_dims(mat.shape[:-2], x.shape[:-2]))
    nb = 1
    for s in batch:
        nb *= s
    if nb == 0:
        return torch.empty((*batch, nout, k), dtype=out_dtype, device=x.device)
    m3, a_bstride, lda = flatten_batch(mat, batch)
    if m3.data_ptr() % 16 != 0:
        m3 = m3.clone()
    if trans and (lda * m3.element_size()) % 16 != 0:
        m3 = m3.contiguous()
        lda = m3.stride(-2)
        a_bstride = m3.stride(0) if (m3.dim() == 3 and a_bstride != 0) else a_bstride
    # block widths between the kernel's instantiations (1, 2, 4, 8, 16): the operand is padded with zero columns so that
    # it stays on the bulk-copy path of its slot (a k = 9..15 block ran at 3.2 TB/s against 4.3 for k = 16: the ragged
    # rows of X went through the plain-load staging warp); the copy is n x 16 elements, the pass over A is n x n
    k_user = k
    if E is None and not trans and k < 16 and (k & (k - 1)) != 0:
        kp = 1 << k.bit_length()
        xx = torch.zeros((nb, nin, kp), dtype=vdt, device=x.device)
        xx[:, :, :k] = x.to(vdt).expand(*batch, nin, k).reshape(nb, nin, k)
        k = kp
    else:
        xx = x.to(vdt).expand(*batch, nin, k).reshape(nb, nin, k).contiguous()
    y = torch.empty((nb, nout, k), dtype=vdt, device=x.device)
    a = _lib.MatvecArgs()
    a.dtype = _lib.dtype_code(mat.dtype)
    a.nbatch, a.nrows, a.ncolsA, a.k = nb, p, q, k
    a.A, a.lda, a.a_bstride = m3.data_ptr(), lda, a_bstride
    a.X, a.ldx, a.x_bstride = xx.data_ptr(), k, nin * k
    a.Y, a.ldy, a.y_bstride = y.data_ptr(), k, nout * k
    a.trans = 1 if trans else 0
    keep = [m3, xx, y]
    if E is not None:
        ee = E.to(vdt).expand(*batch, k).reshape(nb, k).contiguous()
        a.E, a.e_bstride = ee.data_ptr(), k
        keep.append(ee)
        if Z is not None:
            zz = Z.to(vdt).expand(*batch, p, k).reshape(nb, p, k).contiguous()
            a.Z, a.ldz, a.z_bstride = zz.data_ptr(), k, p * k
            keep.append(zz)
    a.impl = impl
    a.stream = _lib.stream_ptr(x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().xt_block_matvec(a), "block_matvec")
    if k != k_user:
        y = y[:, :, :k_user]
    y = y.reshape(*batch, nout, k_user)
    return y if out_dtype == vdt else y.to(out_dtype)


def hermitian_check(mat: torch.Tensor, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """``torch.allclose(mat, mat^T, rtol, atol)`` for a real CUDA matrix (or batch of matrices) in one pass over it
    (`xt_hermitian_check`, csrc/linop.cu) instead of several elementwise kernels over the matrix and its strided
    transpose plus their temporaries.  One device -> host read of the flag, like the library test."""
    _lib.require_cuda(mat, "the Hermiticity check")
    n = mat.shape[-1]
    if mat.shape[-2] != n or mat.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("hermitian_check: square float32 / float64 matrices only")
    batch = tuple(mat.shape[:-2])
    m3, a_bstride, lda = flatten_batch(mat, batch)
    nb = 1
    for s_ in batch:
        nb *= s_
    flag = torch.zeros(1, dtype=torch.int32, device=mat.device)
    a = _lib.HermCheckArgs()
    a.dtype = _lib.dtype_code(mat.dtype)
    a.n, a.nbatch = n, nb
    a.A, a.lda, a.a_bstride = m3.data_ptr(), lda, a_bstride
    a.rtol, a.atol = float(rtol), float(atol)
    a.mismatch = flag.data_ptr()
    a.stream = _lib.stream_ptr(mat.device)
    with torch.cuda.device(mat.device):
        _lib.check(_lib.lib().xt_hermitian_check(a), "hermitian_check")
    return int(flag.item()) == 0
