"""tile-height / consumer-layout sweep of the block matvec on the GPU box (not a pytest file).
XT_MV_TILE_ROWS overrides mv_tiling(); impl 3 = row-slice (two rows per thread at k = 8), 5 = one row per thread."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense


def timeit(A, X, impl):
    for _ in range(3):
        y = _dense.block_matvec(A, X, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = _dense.block_matvec(A, X, impl=impl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = A[-200:].double() @ X.double()
    err = ((y[-200:].double() - ref).abs().max() / ref.abs().max()).item()
    return ms, err


for dtype, n in ((torch.float32, 16384), (torch.bfloat16, 16384), (torch.float64, 8192)):
    A = torch.randn(n, n, device="cuda").to(dtype)
    X = torch.randn(n, 8, device="cuda", dtype=torch.float64 if dtype == torch.float64 else torch.float32)
    for rows in (None, 104, 111, 112, 113, 120, 128):
        if rows is None:
            os.environ.pop("XT_MV_TILE_ROWS", None)
        else:
            os.environ["XT_MV_TILE_ROWS"] = str(rows)
        for impl in (3, 5):
            ms, err = timeit(A, X, impl)
            print("%-8s k=8 tile_rows=%4s impl=%d: %.1f us  %.0f GB/s  relerr %.1e"
                  % (str(dtype)[6:], rows, impl, ms * 1e3, A.element_size() * n * n / ms / 1e6, err), flush=True)
