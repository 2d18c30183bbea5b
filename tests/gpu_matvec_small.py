"""block matvec at small operator sizes: TMA pipeline (impl 1) against the plain-load kernel (impl 2) and the default
(not a pytest file)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _dense
g = torch.Generator().manual_seed(0)
for dtype in (torch.float32, torch.float64):
    for n in (256, 512, 1024, 2048, 4096, 8192):
        A = torch.randn(n, n, generator=g, dtype=dtype).cuda()
        for k in (1, 8):
            X = torch.randn(n, k, generator=g, dtype=dtype).cuda()
            out = []
            for impl in (0, 1, 2):
                for _ in range(5):
                    _dense.block_matvec(A, X, impl=impl)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(100):
                    _dense.block_matvec(A, X, impl=impl)
                e1.record(); torch.cuda.synchronize()
                out.append(e0.elapsed_time(e1) * 10.0)
            print("%s n=%5d k=%d   auto %.1f us   tma %.1f us   plain %.1f us   (A at 6.5 TB/s: %.1f us)"
                  % (str(dtype)[6:], n, k, out[0], out[1], out[2], n * n * A.element_size() / 6.5e6), flush=True)
