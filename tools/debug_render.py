import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, cases
from test_oracle_render import split_golden
from util import load_golden
import panogrf_b200 as pg
name = sys.argv[1] if len(sys.argv) > 1 else "render_m3d_2src"
cfg, _, _ = cases.make_render_inputs(name)
cfg["mlp_dtype"] = sys.argv[2] if len(sys.argv) > 2 else "fp32"
que, ref, W, gold = split_golden(load_golden(name))
net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
net.load_state_dict(W, strict=False)
cu = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}
out = net.render_impl(cu(que), cu(ref), False, keep_hit_prob=True)
torch.cuda.synchronize()
for k, v in gold.items():
    if k.startswith("ray_mask"): continue
    d = (out[k].cpu() - v.float()).abs()
    print(k, "max abs err", float(d.max()), "max ref", float(v.abs().max()), "rel-to-max", float(d.max() / v.abs().max()))
