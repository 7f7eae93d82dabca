"""Cost-volume backward: one-lane-per-(pixel, channel group) kernel vs the run-merged kernel (CUDA events, L2 flushed)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import panogrf_b200 as pg
from panogrf_b200 import _lib

lib = _lib.load()


def run(B, H, W, C, D, S=2, per_pixel=False, cost_type="abs_diff", iters=5, smooth=False):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    images = torch.randn(B, S, H, W, C, device=dev, generator=g)
    if smooth:
        images = torch.nn.functional.avg_pool2d(images.permute(0, 1, 4, 2, 3).reshape(B * S, C, H, W), 3, 1, 1).reshape(B, S, C, H, W).permute(0, 1, 3, 4, 2).contiguous()
    images.requires_grad_(True)
    rots = torch.eye(3, device=dev).expand(B, S, 3, 3).contiguous()
    trans = torch.zeros(B, S, 3, device=dev); trans[:, 0, 2] = 0.5; trans[:, 1, 2] = -0.5
    if S > 2:
        trans[:, 2:, 0] = 0.4
    depths = torch.linspace(0.1, 10, D, device=dev)
    dv = depths.view(1, D, 1, 1).expand(B, D, H, W).contiguous() if per_pixel else None
    args = {"dataset_name": "m3d", "contain_dnet": per_pixel, "mono_uncertainty": False}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    if S > 2:
        out = pg.calculate_cost_volume_erp_multiview(args, images, depths, trans, rots, depth_volume=dv, cost_type=cost_type)
    else:
        out = pg.calculate_cost_volume_erp(args, images, depths, trans, rots, depth_volume=dv, cost_type=cost_type)
    gout = torch.randn(out.shape, device=dev, generator=g)
    res = {}
    grads = {}
    for label, variant, L, minb in (("lane", 0, 0, 8), ("lane6", 0, 0, 6), ("run8", 1, 8, 4), ("run16", 1, 16, 4)):
        lib.pgrf_debug_set(b"cv_bwd_variant", variant); lib.pgrf_debug_set(b"cv_bwd_run", L); lib.pgrf_debug_set(b"cv_bwd_minb", minb)
        f = lambda: torch.autograd.grad(out, images, gout, retain_graph=True)[0]
        for _ in range(2): gr = f()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr = f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[label] = round(sorted(ts)[len(ts) // 2], 4)
        grads[label] = gr
    lib.pgrf_debug_set(b"cv_bwd_variant", 0); lib.pgrf_debug_set(b"cv_bwd_run", 0); lib.pgrf_debug_set(b"cv_bwd_minb", 8)
    scale = grads["lane"].abs().max().item()
    err = {k: round((grads[k] - grads["lane"]).abs().max().item() / scale, 9) for k in grads if k != "lane"}
    print(json.dumps(dict(B=B, H=H, W=W, C=C, D=D, S=S, per_pixel=per_pixel, cost=cost_type, smooth=smooth, ms=res, max_err_vs_lane=err)))


if __name__ == "__main__":
    run(1, 256, 512, 32, 64)
    run(1, 256, 512, 32, 64, smooth=True)
    run(1, 256, 512, 32, 64, per_pixel=True)
    run(1, 256, 512, 32, 64, cost_type="dot")
    run(1, 64, 128, 32, 64, per_pixel=True)
    run(1, 256, 512, 16, 64)
    run(1, 256, 512, 64, 32)
    run(1, 128, 256, 32, 64, S=4)
    run(1, 512, 1024, 32, 128)
