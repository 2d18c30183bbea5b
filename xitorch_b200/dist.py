"""
Multi-GPU layer of the hot path -- new work, the reference has no distributed code.
One process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

* `shard_batch` / `solve_batch_sharded` (BASELINE config 3): independent batched problems are split along the
  leading batch dimension, one slice per rank, and solved locally with the fused kernels -- NO data-path
  collective.  Every shard stops on its own stop test (the reference's stop test is global over the batch,
  solve.py:166,310; per-shard early exit gives the same solutions within the tolerance, only the iteration
  counts differ).  Two scalars (max iterations, all-converged) are all-reduced afterwards for reporting.

* `RowPartitionedOperator` (BASELINE config 5): one large dense Hermitian operator is split into row blocks
  `A_p = A[rows_p, :]`; every rank computes `Y_p = A_p X` with the block-matvec kernel from a replicated `X`
  and ONE all-gather per operator application assembles `Y` (N*k*s bytes: 4 MiB at N = 65536, k = 16).  The
  O(N m) subspace algebra is replicated on every rank (deterministic, no further collectives).
  `symeig_row_partitioned` runs the block-Lanczos/Davidson engine on top of it.  Two engines:
  - "sharded" (default on CUDA): every rank keeps only its rows of the basis; the kernels exchange two small partial
    sums and the new basis block per iteration by direct stores into the peers' memory (`PeerRegions`: cudaMalloc'ed
    regions mapped into every rank with CUDA IPC) -- no collective call and no host round trip inside the iteration;
  - "allgather": the whole O(N m) subspace algebra replicated, one NCCL all-gather per application issued from a
    host callback (round-1 path, kept for comparison and for process groups without peer access).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from xitorch_b200.linop import LinearOperator


def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [lo, hi) slice of `n` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, group=None) -> torch.Tensor:
    """this rank's slice of the leading (batch) dimension."""
    rank, world = _world(group)
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def solve_batch_sharded(A_local: torch.Tensor, B_local: torch.Tensor, method: str = "bicgstab", group=None,
                        is_hermitian: bool = False, **opts):
    """solve this rank's shard `(b_local, n, n)`, `(b_local, n, ncols)` of a batch of independent systems.
    Returns `(x_local, info)` with `info["niter_max"]` / `info["all_converged"]` reduced over the ranks."""
    from xitorch_b200.linalg import solve
    info = {}
    x = solve(LinearOperator.m(A_local, is_hermitian=is_hermitian), B_local, method=method, info=info, **opts)
    rank, world = _world(group)
    niter = torch.tensor([float(info.get("niter", 0)), 0.0 if info.get("converged", True) else 1.0],
                         dtype=torch.float64, device=B_local.device)
    if world > 1:
        dist.all_reduce(niter, op=dist.ReduceOp.MAX, group=group)
    info["niter_max"] = int(niter[0].item())
    info["all_converged"] = bool(niter[1].item() == 0.0)
    return x, info


class RowPartitionedOperator(LinearOperator):
    """Hermitian operator of order `n` stored as row blocks over the process group.

    `A_local` is this rank's `(n_local, n)` block of rows `[row_lo, row_hi)` (equal block sizes are required so
    that the collective is one `all_gather_into_tensor`).  `mm(X)` needs the same `X (n, k)` on every rank and
    returns the same `(n, k)` result on every rank."""

    def __init__(self, A_local: torch.Tensor, n: int, group=None):
        rank, world = _world(group)
        if n % world != 0:
            raise RuntimeError("RowPartitionedOperator: n=%d must be divisible by the world size %d" % (n, world))
        if tuple(A_local.shape) != (n // world, n):
            raise RuntimeError("RowPartitionedOperator: expected a local block of shape %s, got %s"
                               % ((n // world, n), tuple(A_local.shape)))
        super().__init__(shape=(n, n), is_hermitian=True, dtype=A_local.dtype, device=A_local.device,
                         _suppress_hermit_warning=True)
        self.A_local = A_local
        self.group = group
        self.rank, self.world = rank, world
        self.n_local = n // world
        self.napply = 0

    def _local_mm(self, x: torch.Tensor) -> torch.Tensor:
        if self.A_local.is_cuda:
            from xitorch_b200 import _dense
            return _dense.block_matvec(self.A_local, x)
        return torch.matmul(self.A_local, x)

    def _mm(self, x: torch.Tensor) -> torch.Tensor:
        self.napply += 1
        y_local = self._local_mm(x).contiguous()
        if self.world == 1:
            return y_local
        y = torch.empty((self.shape[-1], x.shape[-1]), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(y, y_local, group=self.group)       # the one collective per application
        return y

    def _mv(self, x: torch.Tensor) -> torch.Tensor:
        return self._mm(x.unsqueeze(-1)).squeeze(-1)

    def _getparamnames(self, prefix: str = ""):
        return [prefix + "A_local"]


class PeerRegions(object):
    """one exchange region per rank, each mapped into every rank's address space.

    `ptrs[r]` is the address in THIS process of rank r's region.  `create` allocates this rank's region with
    `xt_peer_alloc` (cudaMalloc, zeroed), exchanges the CUDA IPC handles over the process group and opens the others.
    `epoch` counts the solves that used the regions (the engine tags its flags with it; same on all ranks)."""

    def __init__(self, ptrs, nbytes, rank, world, owner=None):
        self.ptrs, self.nbytes, self.rank, self.world = list(ptrs), int(nbytes), rank, world
        self.epoch = 0
        self._owner = owner          # (own pointer, [opened peer pointers]) when created through xt_peer_*

    @classmethod
    def create(cls, nbytes: int, group=None):
        import ctypes as C
        from xitorch_b200 import _lib
        rank, world = _world(group)
        L = _lib.lib()
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(L.xt_peer_alloc(nbytes, C.byref(own), handle), "peer_alloc")
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        ptrs, opened = [], []
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
                continue
            pp = C.c_void_p()
            _lib.check(L.xt_peer_open(C.create_string_buffer(handles[r], 64), C.byref(pp)), "peer_open")
            ptrs.append(pp.value)
            opened.append(pp.value)
        if world > 1:
            dist.barrier(group=group)
        return cls(ptrs, nbytes, rank, world, owner=(own.value, opened))

    def close(self, group=None):
        if self._owner is None:
            return
        from xitorch_b200 import _lib
        L = _lib.lib()
        own, opened = self._owner
        self._owner = None
        torch.cuda.synchronize()
        for pp in opened:
            L.xt_peer_close(pp)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=group)           # nobody frees a region a peer still has mapped
        L.xt_peer_free(own)


_REGIONS = {}          # (group, device, layout signature) -> PeerRegions, at most _REGIONS_MAX of them


_REGIONS_MAX = 4


def _regions_for(nbytes: int, group=None, signature=()) -> PeerRegions:
    """exchange regions are allocated once per (group, device, problem signature) and reused by every solve with that
    signature.  The signature (dtype, n, neig, max_basis) fixes the LAYOUT of a region: its flags are monotone counters
    that are never reset, so a region must not be reinterpreted with another layout.  Collective: every rank calls this
    with the same arguments in the same order (creation and eviction contain barriers)."""
    key = (id(group), torch.cuda.current_device() if torch.cuda.is_available() else -1, tuple(signature))
    reg = _REGIONS.get(key)
    if reg is None:
        if len(_REGIONS) >= _REGIONS_MAX:
            oldest = next(iter(_REGIONS))
            _REGIONS.pop(oldest).close(group)
        reg = PeerRegions.create(nbytes, group)
        _REGIONS[key] = reg
    return reg


def release_regions(group=None):
    """free the cached exchange regions (collective: call on every rank, before destroy_process_group)"""
    for key in list(_REGIONS):
        _REGIONS.pop(key).close(group)


def symeig_row_partitioned(A_local: torch.Tensor, n: int, neig: int, mode: str = "lowest", method: str = "lanczos",
                           group=None, min_eps: float = 1e-6, max_niter: int = 1000,
                           max_basis: Optional[int] = None, check_every: Optional[int] = None,
                           info: Optional[dict] = None, engine: Optional[str] = None,
                           regions: Optional[PeerRegions] = None, restart_keep: Optional[int] = None,
                           gather: bool = True):
    """`neig` extreme eigenpairs of the row-partitioned operator; every rank returns the same eigenvalues.

    engine: "sharded" (row-sharded subspace algebra, in-kernel exchange over peer memory; needs method="lanczos") or
            "allgather" (replicated algebra, one NCCL all-gather per application); None = sharded whenever it applies.
    gather: with the sharded engine, all-gather the eigenvectors at the end (one collective per solve) so that every
            rank returns the full (n, neig) block; False returns this rank's rows only."""
    from xitorch_b200._impls.symeig import _krylov_row_partitioned, _krylov_row_sharded, _sharded_applies
    if engine is None:
        engine = "sharded" if (method == "lanczos" and _sharded_applies(A_local, n, neig, max_basis, group)) else "allgather"
    if engine == "sharded":
        return _krylov_row_sharded(A_local, n, neig, mode, group, min_eps, max_niter, max_basis, info, regions,
                                   restart_keep, gather)
    if engine != "allgather":
        raise RuntimeError("Unknown engine: %s" % engine)
    return _krylov_row_partitioned(A_local, n, neig, mode, 1 if method == "lanczos" else 0, group, min_eps,
                                   max_niter, max_basis, check_every, info)
