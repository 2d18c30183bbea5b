"""per-iteration time of gmres on one GPU (not a pytest file); XT_NO_SOLVE_SLICES=1 for the one-CTA Arnoldi step."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xitorch_b200 as xt
dev = "cuda"
tag = "1cta  " if os.environ.get("XT_NO_SOLVE_SLICES") == "1" else "sliced"
g = torch.Generator(device=dev); g.manual_seed(5)
for n, nc, niter in ((2048, 1, 200), (4096, 4, 128), (16384, 1, 200), (16384, 8, 64)):
    A = torch.eye(n, device=dev) * 1.0 + torch.randn(n, n, device=dev, generator=g) / n ** 0.5 * 0.99
    B = torch.randn(n, nc, device=dev, generator=g)
    op = xt.LinearOperator.m(A, is_hermitian=False)
    best = None
    for rep in range(3):
        info = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x = xt.linalg.solve(op, B, method="gmres", info=info, rtol=1e-30, atol=0.0, max_niter=niter)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print("%s gmres n=%d ncols=%d niter=%d  %.2f ms  %.1f us/iter (one pass over A: %.1f us)"
          % (tag, n, nc, info["niter"], best * 1e3, best * 1e6 / max(info["niter"], 1), n * n * 4 / 6.5e12 * 1e6), flush=True)
