"""torchrun script (GPU box): row-partitioned block-Lanczos (BASELINE config 5 shape) on WORLD_SIZE GPUs.
    [torchrun --nproc-per-node N] tests/gpu_dist_c5.py [n] [neig] [engine: sharded|allgather|single] [restart_keep] [reps]
Prints the time per solve, the per-application time and the residual identity; all ranks' eigenvalues are compared."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import oracle
import xitorch_b200 as xt
from xitorch_b200 import dist as xd, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
neig = int(sys.argv[2]) if len(sys.argv) > 2 else 16
engine = sys.argv[3] if len(sys.argv) > 3 else "sharded"
keep = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lo, hi = xd.shard_range(n, rank, world)
A_loc = oracle.make_herm_row_block(n, neig, lo, hi, dev)
op1 = xt.LinearOperator.m(A_loc, True) if (world == 1 and engine == "single") else None
torch.cuda.synchronize()
times = []
for rep in range(reps):
    info = {}
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if engine == "single":
        ev, vec = xt.linalg.symeig(op1, neig=neig, method="lanczos", min_eps=1e-4, info=info)
    else:
        ev, vec = xd.symeig_row_partitioned(A_loc, n, neig, "lowest", method="lanczos", min_eps=1e-4, info=info,
                                            engine=engine, restart_keep=(keep or None))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    times.append(t1 - t0)
dt = min(times[1:]) if len(times) > 1 else times[0]
if world > 1:
    evs = [torch.empty_like(ev) for _ in range(world)]
    dist.all_gather(evs, ev)
    same = max((e - evs[0]).abs().max().item() for e in evs)
    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = tt.item()
else:
    same = 0.0
# residual identity on this rank's rows:  (A_loc @ X - X[lo:hi] * lambda)
R = A_loc.double() @ vec.double() - vec.double()[lo:hi] * ev.double()
rmax = torch.tensor([R.abs().max().item()], device=dev)
orth = (vec.double().t() @ vec.double() - torch.eye(neig, device=dev, dtype=torch.float64)).abs().max().item()
if world > 1: dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print("C5 n=%d neig=%d world=%d engine=%s keep=%d: %s  %.3f ms/solve (best of %d)  %.3f ms/apply  max|AX-XL|=%.2e  "
          "|X^T X - I|=%.1e  cross-rank eig diff %.1e" %
          (n, neig, world, engine, keep, info, dt * 1e3, len(times) - 1, dt * 1e3 / max(info["napply"], 1), rmax.item(),
           orth, same))
    print("evals", ev.cpu().numpy())
if world > 1:
    xd.release_regions()
    dist.destroy_process_group()
