"""Diagnostic: sample a coordinate ramp with cost_type='none' -> per-voxel source pixel coords of the kernel vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import panogrf_b200 as pg
from oracle import cost_volume as ocv
B, H, W, C, D = 1, 256, 512, 4, 64
ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
img = torch.stack([xs, ys, torch.zeros_like(xs), torch.zeros_like(xs)], -1)
images = torch.stack([img, img], 0)[None]
rots = torch.eye(3).expand(B, 2, 3, 3).contiguous()
trans = torch.tensor([[[0., 0., 0.5], [0., 0., -0.5]]])
depths = torch.linspace(0.1, 10.0, D)
args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
out = pg.calculate_cost_volume_erp(args, images.cuda(), depths.cuda(), trans.cuda(), rots.cuda(), cost_type="none", out_layout="bdhwc").cpu()
depth = depths.view(1, D, 1, 1).expand(B, D, H, W)
u, v, r = ocv.sweep_uv("m3d", depth, rots[:, 1], trans[:, 1], rots[:, 0], trans[:, 0], return_radius=True)
ix = ((u + 1) / 2) * (W - 1); iy = ((v + 1) / 2) * (H - 1)
ex = (out[..., 0] - ix).abs(); ey = (out[..., 1] - iy).abs()
# ignore seam wrap (x interpolates between W-1 and ... no wrap in zeros padding) -> just report
for name, e in (("x", ex), ("y", ey)):
    print(name, "max", float(e.max()), "mean", float(e.mean()), "frac>1e-4", float((e > 1e-4).float().mean()), "frac>1e-3", float((e > 1e-3).float().mean()))
    idx = torch.topk(e.flatten(), 8).indices
    for i in idx.tolist():
        d, y, x = (i // (H * W)) % D, (i // W) % H, i % W
        print("   d", d, "y", y, "x", x, "err", float(e.flatten()[i]), "ix", float(ix.flatten()[i]), "iy", float(iy.flatten()[i]),
              "kernel", float(out[..., 0].flatten()[i]), float(out[..., 1].flatten()[i]), "radius", float(r.flatten()[i]), "depth", float(depths[d]))
