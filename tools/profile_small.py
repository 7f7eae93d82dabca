"""Small ncu targets: python tools/profile_small.py dg|cvbwd|pg  (one or two launches of the named kernel at a reduced size)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import panogrf_b200 as pg
from panogrf_b200 import _lib
lib = _lib.load()
what = sys.argv[1]
dev = "cuda"
if what == "dg":
    from panogrf_b200.render_ops import depth_guided_placement
    H, W, rfn = 512, 1024, 2
    cfg = {"dataset_name": "m3d", "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0, "n_candidates": 1000,
           "n_samples": 64, "n_gaussian": 15, "backface_culling": True, "contain_uniform": False}
    g = torch.Generator().manual_seed(0)
    rn = 65536
    idx = torch.randperm(H * W, generator=g)[:rn].sort().values
    coords = torch.stack([idx % W, idx // W], -1).float()[None].to(dev)
    w2c = torch.eye(3, 4)[None].repeat(rfn, 1, 1); w2c[0, 2, 3], w2c[1, 2, 3] = 0.5, -0.5
    yy = torch.linspace(0, 3.14159, H)[:, None]; xx = torch.linspace(0, 6.28318, W)[None, :]
    smooth = (3.0 + 1.5 * torch.sin(xx * 2) * torch.sin(yy) + (xx > 3.0).float() * 2.0)[None, None].repeat(rfn, 1, 1, 1).to(dev)
    ref = {"imgs": torch.zeros(rfn, 3, H, W, device=dev), "w2c": w2c.to(dev), "mvs_depth": smooth,
           "mvs_uncert": torch.full((rfn, 1, H, W), 0.01, device=dev), "mvs_normal": torch.randn(rfn, 3, H, W, generator=g).to(dev)}
    que = {"coords": coords, "c2w": torch.eye(3, 4)[None].to(dev)}
    fill = torch.rand(rn, 64, device=dev); ga = torch.randn(rn, 15, device=dev)
    for _ in range(2):
        depth_guided_placement(cfg, que, ref, fill, ga)
elif what == "cvbwd":
    B, S, H, W, C, D = 1, 2, 256, 512, 32, 64
    g = torch.Generator(device=dev).manual_seed(0)
    images = torch.randn(B, S, H, W, C, device=dev, generator=g, requires_grad=True)
    rots = torch.eye(3, device=dev).expand(B, S, 3, 3).contiguous()
    trans = torch.zeros(B, S, 3, device=dev); trans[:, 0, 2] = 0.5; trans[:, 1, 2] = -0.5
    depths = torch.linspace(0.1, 10, D, device=dev)
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    out = pg.calculate_cost_volume_erp(args, images, depths, trans, rots)
    gout = torch.randn(out.shape, device=dev, generator=g)
    for v in (1, 0):
        lib.pgrf_debug_set(b"cv_bwd_variant", v)
        torch.autograd.grad(out, images, gout, retain_graph=True)
torch.cuda.synchronize()
