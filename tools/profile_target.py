"""Small deterministic workload for ncu: cost volume at configs[0] + coarse/fine render of 2 ray chunks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import panogrf_b200 as pg

torch.manual_seed(0)
dev = torch.device("cuda:0")
cfg = bench.cfg_dict()
cfg["mlp_dtype"] = os.environ.get("PGRF_MLP", "bf16")
net = pg.NeuralRayBaseRenderer(cfg).to(dev).eval()
net.rays_per_launch = int(os.environ.get("PGRF_RPL", "4096"))
que, ref = bench.make_inputs(torch, rows=(250, 250 + 2 * net.rays_per_launch // bench.W))
que_d = {k: v.to(dev) for k, v in que.items()}
ref_d = {k: v.to(dev) for k, v in ref.items()}
for _ in range(2):
    out = net.render(que_d, ref_d, False)
B, Hc, Wc, C, D = 1, 256, 512, 32, 64
images = torch.randn(B, 2, Hc, Wc, C, device=dev)
rots = torch.eye(3, device=dev).expand(B, 2, 3, 3).contiguous()
trans = torch.zeros(B, 2, 3, device=dev); trans[:, 0, 2], trans[:, 1, 2] = 0.5, -0.5
depths = torch.linspace(0.1, 10, D, device=dev)
args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
for layout in ("bdchw", "bdhwc"):
    for _ in range(2):
        pg.calculate_cost_volume_erp(args, images, depths, trans, rots, out_layout=layout)
torch.cuda.synchronize()
print("done")

# ---- stand-alone operators (one launch each after warm-up) ----
import types
from panogrf_b200 import render_ops as rops
g = torch.Generator(device=dev).manual_seed(0)
rn, dn = 16384, 64
dirs = torch.nn.functional.normalize(torch.randn(1, rn, 1, 3, device=dev, generator=g), dim=-1)
pts = (dirs * torch.linspace(0.5, 15.0, dn, device=dev).view(1, 1, dn, 1)).contiguous()
spt = types.SimpleNamespace(dataset="m3d", height=bench.H, width=bench.W)
for _ in range(2):
    rops.project_points_dict(ref_d, pts, spt)
dcfg = {"dataset_name": "m3d", "height": bench.H, "width": bench.W, "min_depth": 0.5, "max_depth": 15.0, "n_candidates": 1000,
        "n_samples": 64, "n_gaussian": 15, "backface_culling": True, "contain_uniform": False}
ref2 = dict(ref_d)
ref2["mvs_depth"] = 3.0 + torch.rand(bench.RFN, 1, bench.H, bench.W, device=dev, generator=g)
ref2["mvs_uncert"] = torch.full((bench.RFN, 1, bench.H, bench.W), 0.01, device=dev)
ref2["mvs_normal"] = torch.randn(bench.RFN, 3, bench.H, bench.W, device=dev, generator=g)
q16 = dict(que_d); q16["coords"] = que_d["coords"][:, :16384]
for _ in range(2):
    rops.depth_guided_placement(dcfg, q16, ref2)
img_g = images.clone().requires_grad_(True)
out = pg.calculate_cost_volume_erp(args, img_g, depths, trans, rots, out_layout="bdhwc")
for _ in range(2):
    torch.autograd.grad(out, img_g, torch.ones_like(out), retain_graph=True)
torch.cuda.synchronize()
print("ops done")

# ---- profiled region (ncu --profile-from-start off): ONE launch of every kernel ----
torch.cuda.profiler.start()
q1 = dict(que_d); q1["coords"] = que_d["coords"][:, :net.rays_per_launch]
net.render(q1, ref_d, False)                                           # mlp + rays kernels, coarse and fine pass
for layout in ("bdchw", "bdhwc"):
    pg.calculate_cost_volume_erp(args, images, depths, trans, rots, out_layout=layout)
rops.project_points_dict(ref_d, pts, spt)
rops.depth_guided_placement(dcfg, q16, ref2)
torch.autograd.grad(out, img_g, torch.ones_like(out), retain_graph=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled region done")
