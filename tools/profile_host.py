"""cProfile of the Python side of net.render() on a small shard (host overhead per view)."""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import panogrf_b200 as pg
torch.manual_seed(0)
cfg = bench.cfg_dict(); cfg["mlp_dtype"] = "bf16"
net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
que, ref = bench.make_inputs(torch, (0, 8))
q = {k: v.cuda() for k, v in que.items()}; r = {k: v.cuda() for k, v in ref.items()}
for _ in range(5): net.render(q, r, False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): net.render(q, r, False)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
