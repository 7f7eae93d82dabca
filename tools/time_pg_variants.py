import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from panogrf_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
que, ref = bench.make_inputs(torch)
ref_d = {k: v.to(dev) for k, v in ref.items()}
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
# warm the clocks up first, then interleave the variants over several rounds (the first measurements of a cold GPU read ~8 % low)
for _ in range(20):
    bench.time_project_gather(torch, None, ref_d, flush, bench.measured_peaks())
for rnd in range(3):
  for var in (5,):
    for grid in (64, 96):
        lib.pgrf_debug_set(b"pg_variant", var); lib.pgrf_debug_set(b"pg_grid", grid)
        r = bench.time_project_gather(torch, None, ref_d, flush, bench.measured_peaks())
        print("round", rnd, "variant", var, "grid", grid, "ms %.4f frac %.4f" % (r["ms"], r["roofline"]["frac"]))
lib.pgrf_debug_set(b"pg_variant", 5); lib.pgrf_debug_set(b"pg_grid", 64)
