"""Error distribution of the bf16 tensor-core path against the committed reference goldens (GPU).
python tools/bf16_error_stats.py  -> one line per (case, output): mean / p99 / max of |a-e| / max|e| and of |a-e| / (|e| + 1e-3 max|e|)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import cases
from util import load_golden
import panogrf_b200 as pg


def split(g):
    que = {k[4:]: v for k, v in g.items() if k.startswith("que.")}
    ref = {k[4:]: v for k, v in g.items() if k.startswith("ref.")}
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    gold = {k[4:]: v for k, v in g.items() if k.startswith("out.")}
    return que, ref, W, gold


cu = lambda d: {k: v.cuda() for k, v in d.items()}
for name in cases.RENDER_CASES:
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, gold = split(load_golden(name))
    for dt in ("fp32", "bf16"):
        net = pg.NeuralRayBaseRenderer({**cfg, "mlp_dtype": dt}).cuda().eval()
        net.load_state_dict(W, strict=False)
        out = net.render_impl(cu(que), cu(ref), False, keep_hit_prob=True)
        torch.cuda.synchronize()
        for k in ("pixel_colors_nr", "render_depth", "hit_prob_nr", "colors_nr", "density_nr"):
            if k not in gold:
                continue
            e = gold[k].float(); a = out[k].float().cpu()
            d = (a - e).abs().flatten(); rng = float(e.abs().max())
            rel = d / (e.abs().flatten() + 1e-3 * rng)
            q = lambda x, p: float(torch.quantile(x.double(), p))
            print(f"{name:24s} {dt} {k:16s} range {rng:8.3f} | abs/range mean {float(d.mean())/rng:.2e} p99 {q(d, .99)/rng:.2e} max {float(d.max())/rng:.2e}"
                  f" | rel mean {float(rel.mean()):.2e} p99 {q(rel, .99):.2e} max {float(rel.max()):.2e}")
