"""per-iteration time of cg / bicgstab on one GPU over operator sizes (not a pytest file); run once with
XT_NO_SOLVE_PDL=1 and once without to see what the programmatic dependent launches buy."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
import xitorch_b200 as xt
dev = "cuda"
tag = "plain" if os.environ.get("XT_NO_SOLVE_PDL") == "1" else "pdl"
def run(name, A, B, method, **opts):
    op = xt.LinearOperator.m(A, is_hermitian=True)
    best = None
    for rep in range(4):
        info = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x = xt.linalg.solve(op, B, method=method, info=info, **opts)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print("%-5s %-28s %-8s niter=%4d  %.3f ms  %.2f us/iter" % (tag, name, method, info["niter"], best * 1e3,
                                                               best * 1e6 / max(info["niter"], 1)), flush=True)
g = torch.Generator(device=dev); g.manual_seed(5)
for n in (1024, 2048, 4096, 8192, 16384):
    A = oracle.make_herm(n, 8, torch.float32).to(dev)
    A = A + (abs(torch.linalg.eigvalsh(A.double().cpu())[0].item()) + 1.0) * torch.eye(n, device=dev)
    for nc in (1, 8):
        B = torch.randn(n, nc, device=dev, generator=g)
        for m in ("cg", "bicgstab"):
            run("n=%d fp32 ncols=%d" % (n, nc), A, B, m, posdef=True, rtol=1e-30, atol=0.0, max_niter=200)
Ab = torch.randn(64, 4096, 4096, device=dev, generator=g).to(torch.bfloat16) * 0.01
Ab = Ab + torch.eye(4096, device=dev, dtype=torch.bfloat16) * 2.0
Bb = torch.randn(64, 4096, 1, device=dev, generator=g)
op = xt.LinearOperator.m(Ab, is_hermitian=False)
for rep in range(3):
    info = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = xt.linalg.solve(op, Bb, method="bicgstab", info=info, rtol=1e-30, atol=0.0, max_niter=60)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("%-5s C3 shard B=64 n=4096 bf16 bicgstab niter=%d %.3f ms %.1f us/iter" % (tag, info["niter"], dt * 1e3, dt * 1e6 / info["niter"]))
