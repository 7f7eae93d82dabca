import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import panogrf_b200 as pg
from panogrf_b200.renderer import to_channels_last
dev = torch.device("cuda")
torch.manual_seed(0)
cfg = bench.cfg_dict(); cfg["mlp_dtype"] = "bf16"
net = pg.NeuralRayBaseRenderer(cfg).to(dev).eval()
que, ref = bench.make_inputs(torch)
q = {k: v.to(dev) for k, v in que.items()}; r = {k: v.to(dev) for k, v in ref.items()}
for _ in range(3): net.render(q, r, False)
torch.cuda.synchronize()
def ev(fn, n=5):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
    return ts
print("cached   ", ev(lambda: net.render(q, r, False)))
for k, pad in (("imgs", 4), ("img_feats", None), ("ray_feats", None)):
    print("convert", k, ev(lambda: to_channels_last(r[k], pad)))
fr = [{k: (v.clone() if k in ("imgs", "img_feats", "ray_feats") else v) for k, v in r.items()} for _ in range(5)]
it = iter(fr)
print("fresh    ", ev(lambda: net.render(q, next(it), False)))
net.cache_maps = False
print("nocache  ", ev(lambda: net.render(q, r, False)))
