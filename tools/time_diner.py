"""Time the fused depth-prior sample placement kernel at the full c2 size (512x1024 rays, 1000 candidates, 2 sources)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
from panogrf_b200.render_ops import depth_guided_placement

H, W, rfn = 512, 1024, 2
cfg = {"dataset_name": "m3d", "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0, "n_candidates": 1000,
       "n_samples": 64, "n_gaussian": 15, "backface_culling": True, "contain_uniform": False}
g = torch.Generator().manual_seed(0)
dev = "cuda"
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
coords = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).float()[None].to(dev)
c2w = torch.eye(3, 4)[None].to(dev)
w2c = torch.eye(3, 4)[None].repeat(rfn, 1, 1)
w2c[0, 2, 3], w2c[1, 2, 3] = 0.5, -0.5
que = {"coords": coords, "c2w": c2w}
ref = {"imgs": torch.zeros(rfn, 3, H, W, device=dev), "w2c": w2c.to(dev),
       "mvs_depth": (3.0 + torch.rand(rfn, 1, H, W, generator=g)).to(dev), "mvs_uncert": torch.full((rfn, 1, H, W), 0.01, device=dev),
       "mvs_normal": torch.randn(rfn, 3, H, W, generator=g).to(dev)}
rn = H * W
fill = torch.rand(rn, 64, device=dev); ga = torch.randn(rn, 15, device=dev)
from panogrf_b200 import _lib
lib = _lib.load()
# a smooth prior (what an MVS network produces) next to the white-noise one above
yy = torch.linspace(0, 3.14159, H)[:, None]; xx = torch.linspace(0, 6.28318, W)[None, :]
smooth = (3.0 + 1.5 * torch.sin(xx * 2) * torch.sin(yy) + (xx > 3.0).float() * 2.0)[None, None].repeat(rfn, 1, 1, 1).to(dev)
outs = {}
for label, dmap in (("noise", ref["mvs_depth"]), ("smooth", smooth)):
    ref2 = dict(ref, mvs_depth=dmap)
    for pre in (0, 1):
        lib.pgrf_debug_set(b"dg_prefilter", pre)
        for _ in range(2):
            z = depth_guided_placement(cfg, que, ref2, fill, ga)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            z = depth_guided_placement(cfg, que, ref2, fill, ga)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        outs[(label, pre)] = z
        print(f"depth_guided_placement[{label}, prefilter={pre}] {rn} rays x 1000 candidates x {rfn} views: {ms:.2f} ms  ({rn * 1000 * rfn / ms / 1e6:.1f} G candidate-views/s)")
        print("  finite", bool(torch.isfinite(z).all()), "sorted", bool((z[..., 1:] >= z[..., :-1]).all()))
    print(f"  [{label}] prefilter on == off bit for bit:", bool(torch.equal(outs[(label, 0)], outs[(label, 1)])))

# the other three pixel conventions (same scene, smooth prior): the filters must stay effective, not only exact
for ds in ("replica_test", "residential", "CoffeeArea"):
    cfg2 = dict(cfg, dataset_name=ds)
    ref2 = dict(ref, mvs_depth=smooth)
    ts = {}
    for pre in (0, 1):
        lib.pgrf_debug_set(b"dg_prefilter", pre)
        for _ in range(2):
            z = depth_guided_placement(cfg2, que, ref2, fill, ga)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            z = depth_guided_placement(cfg2, que, ref2, fill, ga)
        e1.record(); torch.cuda.synchronize()
        ts[pre] = e0.elapsed_time(e1) / 3
        outs[(ds, pre)] = z
    print(f"[{ds}] exact {ts[0]:.2f} ms, filtered {ts[1]:.2f} ms, bit-identical: {bool(torch.equal(outs[(ds, 0)], outs[(ds, 1)]))}")
lib.pgrf_debug_set(b"dg_prefilter", 1)
