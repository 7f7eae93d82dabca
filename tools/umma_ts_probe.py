"""Probe: tcgen05.mma with the A operand in TMEM (variant 2 of pgrf_umma_selftest)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from panogrf_b200 import _lib
lib = _lib.load()
import sys as _s
VAR = int(_s.argv[1]) if len(_s.argv) > 1 else 2
for K, N in [(16, 16), (32, 32), (64, 64), (80, 64), (240, 64), (32, 48)]:
    g = torch.Generator(device="cuda").manual_seed(K * 1000 + N)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    out = torch.zeros(128, N, device="cuda")
    rc = lib.pgrf_umma_selftest(_lib.ptr(A), _lib.ptr(W), _lib.ptr(out), K, N, VAR, _lib.stream_ptr())
    _lib.check(rc, "selftest")
    torch.cuda.synchronize()
    ref = A.bfloat16().float() @ W.bfloat16().float().t()
    print("variant", VAR, K, N, "max err", float((out - ref).abs().max()), "ref max", float(ref.abs().max()))
