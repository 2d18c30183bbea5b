import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xitorch_b200 import _lib
if os.environ.get('XT_LIB_OVERRIDE'): _lib.LIB_PATH = os.environ['XT_LIB_OVERRIDE']
L = _lib.lib(); vp = ctypes.c_void_p
def run(T, nev, mode=0):
    m = T.shape[0]
    Td = T.contiguous().cuda()
    w = torch.zeros(nev, dtype=torch.float64, device="cuda"); S = torch.zeros(m, nev, dtype=torch.float64, device="cuda")
    sc = torch.zeros(m * (m | 1) + 16, dtype=torch.float64, device="cuda")
    rc = L.xt_small_eigh(vp(Td.data_ptr()), m, nev, mode, vp(w.data_ptr()), vp(S.data_ptr()), vp(sc.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return w.cpu(), S.cpu()
for m in [int(x) for x in os.environ.get('MS', '2,3,5,8,16,24,33,48,64,65,88,96,97,104,112,113,128,160').split(',')]:
    g = torch.Generator().manual_seed(m)
    T = torch.randn(m, m, generator=g, dtype=torch.float64); T = (T + T.t()) / 2
    nev = min(8, m)
    for mode in (0, 1):
        w, S = run(T, nev, mode)
        wr = torch.linalg.eigvalsh(T); wr = wr[:nev] if mode == 0 else wr[-nev:]
        print("m=%3d mode=%d  eval err %.2e  resid %.2e  orth %.2e" % (m, mode, (w - wr).abs().max().item(),
              (T @ S - S * w).abs().max().item(), (S.t() @ S - torch.eye(nev, dtype=torch.float64)).abs().max().item()))
