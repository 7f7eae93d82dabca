"""Quick device-side timing of the cost-volume kernel (CUDA events, L2 flushed between runs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import panogrf_b200 as pg

def run(B, H, W, C, D, layout, per_pixel=False, groups=0, S=2, iters=10):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    images = torch.randn(B, S, H, W, C, device=dev, generator=g)
    rots = torch.eye(3, device=dev).expand(B, S, 3, 3).contiguous()
    trans = torch.zeros(B, S, 3, device=dev); trans[:, 0, 2] = 0.5; trans[:, 1, 2] = -0.5
    depths = torch.linspace(0.1, 10, D, device=dev)
    dv = depths.view(1, D, 1, 1).expand(B, D, H, W).contiguous() if per_pixel else None
    args = {"dataset_name": "m3d", "contain_dnet": per_pixel, "mono_uncertainty": False}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    f = lambda: pg.calculate_cost_volume_erp(args, images, depths, trans, rots, depth_volume=dv, out_layout=layout, groups=groups)
    for _ in range(3): out = f()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    vox = B * D * H * W
    oc = groups if groups else C
    bytes_ = vox * oc * 4 + B * S * H * W * C * 4 + (vox * 4 if per_pixel else 0)
    print(json.dumps(dict(B=B, H=H, W=W, C=C, D=D, layout=layout, per_pixel=per_pixel, groups=groups, ms=round(ms, 4),
                          gvox_s=round(vox / ms / 1e6, 2), gbs=round(bytes_ / ms / 1e6, 1))))

if __name__ == "__main__":
    from panogrf_b200 import _lib
    lib = _lib.load()
    for jb in (2, 4):
        for mb in (0, 1):
            lib.pgrf_debug_set(b"cv_jb", jb); lib.pgrf_debug_set(b"cv_minb", mb)
            print("jb", jb, "minb", mb)
            run(1, 256, 512, 32, 64, "bdhwc"); run(1, 256, 512, 32, 64, "bdchw")
    lib.pgrf_debug_set(b"cv_jb", 0); lib.pgrf_debug_set(b"cv_minb", 0)
    for layout in ["bdhwc", "bdchw", "bcdhw"]:
        run(1, 256, 512, 32, 64, layout)
    run(1, 256, 512, 32, 64, "bdhwc", per_pixel=True)
    run(1, 256, 512, 32, 64, "bcdhw", groups=8)
    run(1, 512, 1024, 32, 128, "bdhwc")
    run(1, 512, 1024, 32, 128, "bdchw")
    run(1, 64, 128, 32, 64, "bdchw", per_pixel=True)
