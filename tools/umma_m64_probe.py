"""Decode the TMEM lane mapping of an M = 64 tcgen05.mma (diagnostic variants >= 4 of pgrf_umma_selftest)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from panogrf_b200 import _lib
lib = _lib.load()
K, N = 16, 16
A = torch.zeros(128, K, device="cuda"); A[:, 0] = torch.arange(1, 129, device="cuda").float()
W = torch.zeros(N, K, device="cuda"); W[:, 0] = torch.arange(1, N + 1, device="cuda").float()   # out[r][n] = (r+1)*(n+1)
for var in (4, 5, 6, 8):
    out = torch.full((128, N), -7.0, device="cuda")
    _lib.check(lib.pgrf_umma_selftest(_lib.ptr(A), _lib.ptr(W), _lib.ptr(out), K, N, var, _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    col0 = out[:, 0].cpu().tolist()
    col1 = out[:, 1].cpu().tolist()
    print("variant", var, "lane offset", (var - 4) * 16)
    print(" col0 per lane:", [int(x) for x in col0])
    print(" col1/col0 where nonzero:", sorted(set(round(b / a, 2) for a, b in zip(col0, col1) if a != 0)))
