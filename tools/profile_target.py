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
