"""BASELINE config 3 (per-GPU shard): bicgstab on a batch of independent non-symmetric bf16 systems, fp32 vectors.
    python tests/gpu_bench_c3.py [nbatch_per_gpu=64] [n=4096]            (1 GPU)
    torchrun --nproc-per-node N tests/gpu_bench_c3.py 64 4096            (batch-sharded, N*64 systems in total)
Reports iterations, time per iteration, A-read GB/s (2.1 * s * B * N^2 bytes per iteration) and the
true residual against the bf16-rounded matrices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import xitorch_b200 as xt
from xitorch_b200 import dist as xd, _lib

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
A = torch.empty(nb, n, n, dtype=torch.bfloat16, device=dev)
for b in range(nb):
    Ab = torch.randn(n, n, generator=g, device=dev) * (0.3 / n ** 0.5)
    Ab.diagonal().add_(1.0)
    A[b] = Ab.to(torch.bfloat16)
B = torch.randn(nb, n, 1, generator=g, device=dev)
torch.cuda.synchronize()
for rep in range(3):
    if world > 1: dist.barrier()
    _lib.profile_reset(True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, info = xd.solve_batch_sharded(A, B, method="bicgstab", rtol=1e-6, posdef=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    mv_ms, n_mv, n_l = _lib.profile_read(); _lib.profile_reset(False)
res = (torch.bmm(A.float(), x) - B).norm(dim=1) / B.norm(dim=1)
tt = torch.tensor([t1 - t0], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    it = info["niter"]
    bytes_it = 2.1 * 2 * nb * n * n
    print("C3 shard: B=%d/GPU (x%d GPUs) n=%d bf16: %s" % (nb, world, n, info))
    print("  %.2f ms total, %.3f ms/iter, A-read %.0f GB/s per GPU (algorithmic 2.1*s*B*N^2 per iteration), "
          "matvec kernels: %d launches avg %.3f ms -> %.0f GB/s; max true rel. residual %.2e"
          % (tt.item() * 1e3, tt.item() * 1e3 / it, it * bytes_it / tt.item() / 1e9, n_mv, mv_ms / max(n_mv, 1),
             2.0 * nb * n * n / (mv_ms / max(n_mv, 1) * 1e-3) / 1e9, res.max().item()))
if world > 1: dist.destroy_process_group()
