"""phase timing of the one-CTA eigensolver (tuning aid, not a pytest file)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _lib
L = _lib.lib(); vp = ctypes.c_void_p
for m, nev in [(8, 8), (24, 8), (48, 8), (88, 8), (104, 8), (128, 8), (128, 64)]:
    g = torch.Generator().manual_seed(m)
    T = torch.randn(m, m, generator=g, dtype=torch.float64); T = ((T + T.t()) / 2).cuda()
    w = torch.zeros(nev, dtype=torch.float64, device="cuda"); S = torch.zeros(m, nev, dtype=torch.float64, device="cuda")
    sc = torch.zeros(m * (m | 1) + 16, dtype=torch.float64, device="cuda")
    for _ in range(2):
        L.xt_small_eigh(vp(T.data_ptr()), m, nev, 0, vp(w.data_ptr()), vp(S.data_ptr()), vp(sc.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    c = sc[m * (m | 1):].view(torch.int64).cpu().tolist()
    d = [c[i + 1] - c[i] for i in range(5)]
    print("m=%3d nev=%2d  tridiag %7d  bisect %7d  invit %7d  mgs %7d  backtr %7d  total %7d clk (%.1f us @1.9GHz)" % (m, nev, *d, c[5] - c[0], (c[5] - c[0]) / 1900.0))
    print("      tridiag phases per column: A %d  B+bar %d  C1+bar %d  C2+bar %d" % tuple(x // max(m - 2, 1) for x in c[6:10]))
