"""Run the larger BASELINE configs once for crashes / NaNs and print timings (configs[3], configs[4] shapes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import panogrf_b200 as pg

def render_case(H, W, rfn, dtype, dn=64):
    torch.manual_seed(0)
    cfg = {"dataset_name": "m3d", "batch_size": 1, "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0, "use_disp": True,
           "use_hierarchical_sampling": True, "depth_sample_num": dn, "fine_depth_sample_num": dn, "render_depth": True,
           "dist_decoder_cfg": {"use_vis": False}, "fine_dist_decoder_cfg": {"use_vis": False}, "mlp_dtype": dtype}
    net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(rfn, 3, H, W, generator=g)
    ref = {"imgs": imgs.cuda(), "img_feats": torch.randn(rfn, 32, H // 4, W // 4, generator=g).cuda(),
           "ray_feats": torch.randn(rfn, 32, H // 8, W // 8, generator=g).cuda(),
           "depth_range": torch.tensor([[0.5, 15.0]]).repeat(rfn, 1).cuda()}
    w2c = torch.zeros(rfn, 3, 4); w2c[:, :, :3] = torch.eye(3)
    offs = [(0, 0, -0.5), (0, 0, 0.5), (-0.5, 0, 0), (0.5, 0, 0)]
    for i in range(rfn): w2c[i, :, 3] = torch.tensor(offs[i])
    ref["w2c"] = w2c.cuda()
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    que = {"coords": torch.stack([xs, ys], -1).reshape(1, -1, 2).float().cuda(), "c2w": torch.eye(4)[None, :3].cuda(),
           "depth_range": torch.tensor([[0.5, 15.0]]).cuda()}
    for _ in range(2):
        out = net.render(que, ref, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = net.render(que, ref, False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = all(bool(torch.isfinite(v).all()) for v in out.values() if v.dtype.is_floating_point)
    print(f"render {H}x{W} rfn={rfn} {dtype}: {dt*1e3:.1f} ms, {H*W/dt/1e6:.2f} M rays/s, finite={ok}, "
          f"rgb mean {float(out['pixel_colors_nr_fine'].mean()):.4f}", flush=True)

def cv_case(B, S, H, W, C, D):
    images = torch.randn(B, S, H, W, C, device="cuda")
    rots = torch.eye(3, device="cuda").expand(B, S, 3, 3).contiguous()
    trans = torch.randn(B, S, 3, device="cuda") * 0.3
    dv = torch.sort(torch.rand(B, D, H, W, device="cuda") * 9 + 0.5, 1)[0]
    args = {"dataset_name": "m3d", "contain_dnet": True, "mono_uncertainty": False}
    f = (lambda: pg.calculate_cost_volume_erp_multiview(args, images, None, trans, rots, depth_volume=dv, curr_idx=0, groups=8)) if S > 2 else \
        (lambda: pg.calculate_cost_volume_erp(args, images, None, trans, rots, depth_volume=dv, out_layout="bdhwc"))
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"cost volume B{B} S{S} {H}x{W} C{C} D{D}: {dt*1e3:.2f} ms, {B*D*H*W/dt/1e9:.2f} Gvoxel/s, finite={bool(torch.isfinite(out).all())}", flush=True)
    del out

if __name__ == "__main__":
    render_case(512, 1024, 4, "bf16")          # configs[3]: 4 source panoramas
    render_case(512, 1024, 4, "fp32")
    render_case(1024, 2048, 4, "bf16")         # configs[4] render part
    render_case(512, 1024, 3, "bf16")
    render_case(512, 1024, 1, "bf16")
    cv_case(1, 2, 512, 1024, 32, 128)          # configs[2] per-GPU shape
    cv_case(1, 5, 256, 512, 32, 64)            # multi-view, group-wise epilogue
    cv_case(1, 2, 1024, 2048, 32, 192)         # configs[4] cost volume: 51.5 GB output
