"""torchrun script (GPU box): row-partitioned block-Lanczos (BASELINE config 5 shape) on WORLD_SIZE GPUs.
    torchrun --nproc-per-node N tests/gpu_dist_c5.py [n] [neig]
Checks the eigenvalues against a single-GPU solve of the same matrix (rank 0, when it fits) and prints timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import oracle
import xitorch_b200 as xt
from xitorch_b200 import dist as xd, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
neig = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lo, hi = xd.shard_range(n, rank, world)
# every rank builds only its row block of make_herm(n, neig) from per-row-block seeds (deterministic, symmetric by construction)
def row_block(lo, hi):
    # G = U + U^T with U[i, j] (i < j) drawn from a generator seeded by (min(bi,bj), max(bi,bj)) block pair
    bs = 2048
    out = torch.empty(hi - lo, n, dtype=torch.float32, device=dev)
    for bi in range(lo // bs, (hi + bs - 1) // bs):
        r0, r1 = max(lo, bi * bs), min(hi, (bi + 1) * bs)
        for bj in range((n + bs - 1) // bs):
            c0, c1 = bj * bs, min(n, (bj + 1) * bs)
            a, b = min(bi, bj), max(bi, bj)
            g = torch.Generator(device=dev); g.manual_seed(1000003 * a + b)
            blk = torch.randn(bs, bs, generator=g, device=dev)
            if bi == bj:
                blk = blk + blk.t()
                sub = blk[r0 - bi * bs: r1 - bi * bs, : c1 - c0]
            elif bi < bj:
                sub = blk[r0 - bi * bs: r1 - bi * bs, : c1 - c0]
            else:
                sub = blk.t()[r0 - bi * bs: r1 - bi * bs, : c1 - c0]
            out[r0 - lo: r1 - lo, c0:c1] = sub * (0.05 / (2.0 * n) ** 0.5) * (1.0 if bi == bj else 2.0 ** 0.5)
    d = 20.0 + 10.0 * torch.linspace(0, 1, n, device=dev)
    d[:2 * neig] = 1.0 + torch.arange(2 * neig, device=dev)
    idx = torch.arange(lo, hi, device=dev)
    out[idx - lo, idx] += d[lo:hi]
    return out
A_loc = row_block(lo, hi)
# (LinearOperator.m(..., is_hermitian=True) validates symmetry with a full transposed pass over A, as the reference
# does -- built once, outside the timed region)
op1 = xt.LinearOperator.m(A_loc, True) if world == 1 else None
torch.cuda.synchronize()
for rep in range(3):
    info = {}
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev, vec = xd.symeig_row_partitioned(A_loc, n, neig, "lowest", method="lanczos", min_eps=1e-4, info=info) if world > 1 else \
        xt.linalg.symeig(op1, neig=neig, method="lanczos", min_eps=1e-4, info=info)
    torch.cuda.synchronize(); t1 = time.perf_counter()
if world > 1:
    evs = [torch.empty_like(ev) for _ in range(world)]
    dist.all_gather(evs, ev)
    same = max((e - evs[0]).abs().max().item() for e in evs)
else:
    same = 0.0
# residual identity on this rank's rows:  (A_loc @ X - X[lo:hi] * lambda)
R = A_loc.double() @ vec.double() - vec.double()[lo:hi] * ev.double()
rmax = torch.tensor([R.abs().max().item()], device=dev)
if world > 1: dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print("C5 n=%d neig=%d world=%d: %s  %.2f ms  %.1f it/s  max|AX-XL|=%.2e  cross-rank eig diff %.1e" %
          (n, neig, world, info, (t1 - t0) * 1e3, info["niter"] / (t1 - t0), rmax.item(), same))
    print("evals", ev[:4].cpu().numpy(), "... expected ~", [1, 2, 3, 4])
if world > 1:
    dist.destroy_process_group()
