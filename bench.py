#!/usr/bin/env python
"""Headline benchmark of the PanoGRF render-time hot path (BASELINE.json configs[1]).

One "step" = render ONE 512x1024 novel view (524 288 rays, 64 coarse + 64 fine samples per ray)
from 2 synthetic source panoramas 1.0 m apart with a random-init renderer, through the fused
sm_100a kernels of `panogrf_b200` (pre-encoded feature maps; the CNN encoders are outside the hot
path).  Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  For N>1 launch under torchrun; the view's rows are sharded across ranks and the
output tiles (rgb + depth) are all-gathered with NCCL ("strong" scaling: total work fixed).

`--impl reference` times the CPU restatement of the reference path (oracle/render.py, the only
place the oracle is executed here) on the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 1024
RFN = 2
DN = 64
N_RAYS = H * W
WORKLOAD = "render 512x1024 novel view, 2 source panoramas 1.0 m baseline, 64 coarse + 64 fine samples, random-init"
# algorithmic FLOPs per (view,sample) row / per sample (2 per MAC), from the layer table of DESIGN.md
MAC_ROW_R1 = 3 * (32 * 32 * 2 + 32 * 2) - 32 + (34 * 32 + 32 * 32) + (4 * 16 + 16 * 35) + (32 * 8 + 8)
MAC_ROW_R2 = 67 * 64 + 64 * 32 + (32 * 32 + 32 * 33) + (32 * 32 + 32) + (37 * 16 + 16 * 8 + 8)
MAC_SAMPLE_R2 = 140 * 64
MAC_SAMPLE_R3 = 65 * 64 + 64 * 16 + 3 * 256 + 2 * 4 * 64 * 4 + 256 + 256 + 16


def cfg_dict():
    return {
        "dataset_name": "m3d", "batch_size": 1, "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0,
        "use_disp": True, "use_hierarchical_sampling": True, "fine_depth_use_all": False,
        "depth_sample_num": DN, "fine_depth_sample_num": DN, "ray_batch_num": 2048, "render_depth": True,
        "render_uncert": False, "use_ray_mask": True, "debug": False,
        "dist_decoder_cfg": {"use_vis": False}, "fine_dist_decoder_cfg": {"use_vis": False},
        "agg_net_cfg": {}, "fine_agg_net_cfg": {},
    }


def make_inputs(torch, rows=None):
    """Seeded synthetic scene: smooth RGB panoramas, randn feature maps at H/4 and H/8."""
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(RFN, 3, H // 8, W // 8, generator=g)
    imgs = torch.nn.functional.interpolate(imgs, size=(H, W), mode="bilinear", align_corners=False).contiguous()
    img_feats = torch.randn(RFN, 32, H // 4, W // 4, generator=g)
    ray_feats = torch.randn(RFN, 32, H // 8, W // 8, generator=g)
    w2c = torch.zeros(RFN, 3, 4)
    w2c[:, :, :3] = torch.eye(3)
    w2c[0, 2, 3], w2c[1, 2, 3] = -0.5, 0.5            # source cameras at z = +0.5 / -0.5 (1.0 m baseline)
    r0, r1 = rows if rows else (0, H)
    ys, xs = torch.meshgrid(torch.arange(r0, r1), torch.arange(W), indexing="ij")
    coords = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    que = {"coords": coords, "c2w": torch.eye(4)[None, :3].contiguous(), "depth_range": torch.tensor([[0.5, 15.0]])}
    ref = {"imgs": imgs, "w2c": w2c, "depth_range": torch.tensor([[0.5, 15.0]]).repeat(RFN, 1),
           "ray_feats": ray_feats, "img_feats": img_feats}
    return que, ref


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained"),
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


# --------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# --------------------------------------------------------------------------------------------------

def oracle_rays_per_s(n_rays, repeats=1, warm=0):
    import torch
    from oracle import render as orender
    from panogrf_b200.renderer import NeuralRayBaseRenderer
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    cfg = cfg_dict()
    net = NeuralRayBaseRenderer(cfg)            # parameter container only (random init), stays on the CPU
    Wd = {k: v.detach() for k, v in net.state_dict().items()}
    que, ref = make_inputs(torch)
    g = torch.Generator().manual_seed(1)
    idx = torch.randperm(N_RAYS, generator=g)[:n_rays]
    que = dict(que)
    que["coords"] = que["coords"][:, idx]
    ocfg = dict(cfg)
    ocfg["sample_num"] = DN
    times = []
    with torch.no_grad():
        for i in range(warm + repeats):
            t0 = time.perf_counter()
            orender.render(ocfg, Wd, que, ref, ray_batch_num=2048)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return n_rays / min(times), min(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 4096
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    for i in range(args.warmup + args.steps):
        v, t = oracle_rays_per_s(sample)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"{sample}-ray sample of the view (2 ray batches of 2048)"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{sample} random rays of the 512x1024 view per step, oracle/render.py (torch CPU fp32)"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rays-per-launch", type=int, default=0)
    ap.add_argument("--mlp-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: tcgen05 tensor-core MLP (rtol 1e-2); fp32: SIMT parity path (rtol 1e-4)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import panogrf_b200 as pg
    from panogrf_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun", file=sys.stderr)
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    assert H % world == 0
    rows = (rank * H // world, (rank + 1) * H // world)
    torch.manual_seed(0)
    cfg = cfg_dict()
    cfg["mlp_dtype"] = args.mlp_dtype
    net = pg.NeuralRayBaseRenderer(cfg).to(dev).eval()           # same seed on every rank -> identical weights
    if args.rays_per_launch:
        net.rays_per_launch = args.rays_per_launch
    que, ref = make_inputs(torch, rows)
    que_d = {k: v.to(dev) for k, v in que.items()}
    ref_d = {k: v.to(dev) for k, v in ref.items()}
    rn_local = que["coords"].shape[1]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    gathered_rgb = torch.empty(world, rn_local, 3, device=dev) if world > 1 else None
    gathered_depth = torch.empty(world, rn_local, device=dev) if world > 1 else None

    def step():
        out = net.render(que_d, ref_d, False)
        if world > 1:   # the only collective of the path: output tiles (rgb + depth), 16 B/ray
            dist.all_gather_into_tensor(gathered_rgb, out["pixel_colors_nr_fine"][0])
            dist.all_gather_into_tensor(gathered_depth, out["render_depth_fine"][0])
        return out

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    times = []
    for _ in range(args.steps):
        flush.zero_()                                            # L2 flush between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        dist.barrier()
    total_ms = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = N_RAYS / (ms_per_step / 1e3)

    # ---- per-kernel device times on one resident ray chunk (CUDA events on the launching stream) ----
    kern = {}
    if rank == 0:
        kern = time_stages(torch, net, que_d, ref_d, flush)

    # ---- e2e: the C-ABI host entry point, pinned host buffers in, rgb+depth out ----
    e2e_ms, h2d, d2h = time_e2e(torch, net, que, ref, cfg, args.steps)
    e2e_t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = N_RAYS / (float(e2e_t.item()) / 1e3)

    if rank == 0:
        peaks = measured_peaks()
        line = {
            "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.mlp_dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mlp": args.mlp_dtype + (" tcgen05 (fp32 accumulate, rtol 1e-2)" if args.mlp_dtype == "bf16" else " SIMT (rtol 1e-4)"), "views_per_s": 1e3 / ms_per_step, "rays_per_launch": rays_per_launch(net),
                       "l2": "256 MiB flush buffer written between timed iterations", "sharding": f"{H // world} rows/rank",
                       "collective": "all_gather(rgb, depth)" if world > 1 else "none"},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
        }
        line.update(kern.get("roofline_objects", {}))
        line["kernels"] = kern.get("kernels")
        if not args.no_cpu_baseline and world == 1:        # the CPU baseline is an N=1 measurement (rank 0, all host cores)
            v, t = oracle_rays_per_s(8192)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"8192 random rays of the view, oracle/render.py torch-CPU fp32, {t:.1f} s"}
        line["cost_volume"] = time_cost_volume(torch, pg, flush, peaks)
        line["project_gather"] = time_project_gather(torch, que_d, ref_d, flush, peaks)
        line["depth_guided"] = time_depth_guided(torch, que_d, ref_d)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def rays_per_launch(net):
    return int(net.rays_per_launch or (131072 if net.mlp_dtype == "bf16" else 32768))


def time_stages(torch, net, que_d, ref_d, flush):
    """Device time of each of the three kernels of the coarse pass on one chunk of rays_per_launch rays."""
    from panogrf_b200 import _lib
    from panogrf_b200.renderer import coarse_depth_table
    lib = _lib.load()
    cfg = net.cfg
    ctx = net._context(que_d, ref_d)
    rn = min(rays_per_launch(net), que_d["coords"].shape[1])
    # whole ERP rows spread evenly over the latitudes (the first rows alone are all next to a pole, where the source
    # footprints scatter and the pass is ~1.5x slower than the view average)
    n_rows = max(1, rn // W)
    rn = n_rows * W
    total_rows = que_d["coords"].shape[1] // W
    row_ids = ((torch.arange(n_rows, dtype=torch.float64) + 0.5) * total_rows / n_rows).long().clamp(max=total_rows - 1)
    coords = que_d["coords"][0].reshape(total_rows, W, 2)[row_ids.to(que_d["coords"].device)].reshape(rn, 2).float().contiguous()
    dev = coords.device
    depth = coarse_depth_table(cfg, DN, cfg["use_disp"]).to(dev)
    outs = {"pixel_colors_nr": torch.empty(1, rn, 3, device=dev), "density_nr": torch.empty(1, rn, DN, device=dev),
            "colors_nr": torch.empty(1, rn, DN, 3, device=dev), "render_depth": torch.empty(1, rn, device=dev)}
    res = {}
    names = {1: "render_rows_kernel", 2: "render_samples_kernel", 4: "render_rays_kernel"}
    if net.mlp_dtype == "bf16":
        names = {3: "render_mlp_bf16_kernel", 4: "render_rays_bf16_kernel"}
    net._pass(ctx, coords, depth, 0, False, False, outs, 0, False)      # populate workspaces
    torch.cuda.synchronize()
    for mask, name in names.items():
        ts = []
        reps = 5                       # back-to-back launches inside one event pair: amortises the host launch path
        for i in range(4):
            ctx["stage_mask"] = mask
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                net._pass(ctx, coords, depth, 0, False, False, outs, 0, False)
            e1.record()
            torch.cuda.synchronize()
            if i >= 1:
                ts.append(e0.elapsed_time(e1) / reps)
        res[name] = sum(ts) / len(ts)
    ctx["stage_mask"] = 0
    samples = rn * DN
    rows = samples * RFN
    flops = {
        "render_rows_kernel": 2.0 * rows * MAC_ROW_R1,
        "render_samples_kernel": 2.0 * (rows * MAC_ROW_R2 + samples * MAC_SAMPLE_R2),
        "render_mlp_bf16_kernel": 2.0 * (rows * (MAC_ROW_R1 + MAC_ROW_R2) + samples * MAC_SAMPLE_R2),
        "render_rays_kernel": 2.0 * samples * MAC_SAMPLE_R3,
        "render_rays_bf16_kernel": 2.0 * samples * MAC_SAMPLE_R3,
    }
    peaks = measured_peaks()
    kernels = {k: {"ms": round(v, 4), "tflops": round(flops[k] / v / 1e9, 2), "rays": rn} for k, v in res.items()}
    top = max(res, key=res.get)
    peak = peaks["bf16_sustained"] or peaks["bf16_tflops"]
    # DRAM bytes per ray of the two bf16 kernels from the committed ncu capture (profiles/r1_final_ncu_summary.md: 16 384-ray launch,
    # dram__bytes_read.sum + dram__bytes_write.sum), scaled to this launch; None for the fp32 kernels (not captured this round)
    ncu_dram_per_ray = {"render_mlp_bf16_kernel": (0.006649e9 + 0.233480e9) / 16384, "render_rays_bf16_kernel": (0.285318e9 + 0.019544e9) / 16384}
    traffic = ncu_dram_per_ray[top] * rn if top in ncu_dram_per_ray else None
    roof = {"kernel": top, "bound": "tensor", "achieved": flops[top] / res[top] / 1e9, "peak": peak, "unit": "TFLOP/s",
            "frac": flops[top] / res[top] / 1e9 / peak, "traffic": traffic,
            "note": (f"algorithmic (unpadded) MLP FLOPs of the kernel / its CUDA-event time, against the {peaks['src']} bf16 "
                     "tensor peak (sustained); " + ("tcgen05 bf16 path" if net.mlp_dtype == "bf16" else
                                                   "fp32 SIMT parity path, fp32 FMA peak of B200 is ~74 TFLOP/s"))}
    return {"kernels": kernels, "roofline_objects": {"roofline": roof}}


def time_e2e(torch, net, que, ref, cfg, steps):
    """pgrf_render_view_host: HOST buffers in (NCHW maps, poses, weights), rgb+depth of the fine pass out."""
    from panogrf_b200 import _lib
    from panogrf_b200.renderer import coarse_depth_table, fine_u_table
    lib = _lib.load()
    pin = lambda t: t.float().contiguous().pin_memory()
    coords = pin(que["coords"][0])
    rn = coords.shape[0]
    imgs, imf, rf = pin(ref["imgs"]), pin(ref["img_feats"]), pin(ref["ray_feats"])
    w2c, rng, c2w = pin(ref["w2c"]), pin(ref["depth_range"]), pin(que["c2w"].reshape(3, 4))
    depth, fine_u = pin(coarse_depth_table(cfg, DN, True)), pin(fine_u_table(DN))
    dev = next(net.parameters()).device
    wc, wf = pin(net._blob(False, dev).cpu()), pin(net._blob(True, dev).cpu())
    w16c = w16f = None
    if net.mlp_dtype == "bf16":
        w16c, w16f = net._blob16(False, dev).cpu().pin_memory(), net._blob16(True, dev).cpu().pin_memory()
    rgb_c = torch.empty(rn, 3).pin_memory()
    rgb = torch.empty(rn, 3).pin_memory()
    dep = torch.empty(rn).pin_memory()
    va = _lib.RenderViewArgs()
    a = va.pass_
    a.dataset, a.H, a.W, a.rfn, a.rn, a.dn, a.use_vis, a.bias_val = 0, H, W, RFN, rn, DN, 0, 0.05
    a.coords, a.depth, a.depth_ray_stride = _lib.ptr(coords), _lib.ptr(depth), 0
    a.que_c2w, a.que_near, a.que_far = _lib.ptr(c2w), 0.5, 15.0
    a.ref_w2c, a.ref_depth_range = _lib.ptr(w2c), _lib.ptr(rng)
    a.imgs_cl, a.img_h, a.img_w = _lib.ptr(imgs), H, W
    a.img_feats_cl, a.if_h, a.if_w = _lib.ptr(imf), imf.shape[2], imf.shape[3]
    a.ray_feats_cl, a.rf_h, a.rf_w = _lib.ptr(rf), rf.shape[2], rf.shape[3]
    a.weights = _lib.ptr(wc)
    if w16c is not None:
        a.mlp_bf16, a.weights16, va.weights16_fine = 1, _lib.ptr(w16c), _lib.ptr(w16f)
    a.pixel_colors = _lib.ptr(rgb_c)
    a.fine_dn, a.fine_u, a.fine_use_all, a.use_disp = DN, _lib.ptr(fine_u), 0, 1
    va.hierarchical, va.weights_fine, va.bias_val_fine = 1, _lib.ptr(wf), 0.05
    va.rays_per_launch = rays_per_launch(net)
    va.pixel_colors_fine, va.render_depth_fine = _lib.ptr(rgb), _lib.ptr(dep)
    h2d = sum(t.numel() * 4 for t in (coords, imgs, imf, rf, w2c, rng, c2w, depth, fine_u, wc, wf))
    if w16c is not None:
        h2d += w16c.numel() + w16f.numel()
    d2h = sum(t.numel() * 4 for t in (rgb_c, rgb, dep))
    times = []
    for i in range(2 + steps):
        t0 = time.perf_counter()
        rc = lib.pgrf_render_view_host(ctypes.byref(va))
        dt = time.perf_counter() - t0
        _lib.check(rc, "pgrf_render_view_host")
        if i >= 2:
            times.append(dt)
    assert torch.isfinite(rgb).all() and float(rgb.abs().sum()) > 0
    return 1e3 * sum(times) / len(times), h2d, d2h


def time_project_gather(torch, que_d, ref_d, flush, peaks):
    """Stand-alone K2 (project_points_dict + get_img_feats): 65536 rays x 64 samples x 2 views, HBM roofline on its outputs."""
    import types
    from panogrf_b200 import render_ops as rops
    dev = flush.device
    rn, dn = 65536, DN
    g = torch.Generator(device=dev).manual_seed(0)
    dirs = torch.nn.functional.normalize(torch.randn(1, rn, 1, 3, device=dev, generator=g), dim=-1)
    depth = torch.linspace(0.5, 15.0, dn, device=dev).view(1, 1, dn, 1)
    pts = (dirs * depth).contiguous()                                   # (1,rn,dn,3) world points around the origin
    spt = types.SimpleNamespace(dataset="m3d", height=H, width=W)
    f = lambda: rops.project_points_dict(ref_d, pts, spt)
    for _ in range(3):
        out = f()
    ts = []
    reps = 5          # back-to-back calls (each writes 2.4 GB of outputs, >> L2), see time_cost_volume
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ms = sorted(ts)[len(ts) // 2]
    rows = rn * dn * RFN
    alg = rows * (2 + 1 + 3 + 32 + 3 + 32) * 4 + rn * dn * 12
    return {"workload": f"{rn} rays x {dn} samples x {RFN} views -> pts,depth,dir,ray_feats,rgb,img_feats (292 B/row)",
            "rows_per_s": rows / ms * 1e3, "ms": ms,
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg / ms / 1e6 / peaks["hbm_gbs"], "traffic": (0.056696e9 + 0.554805e9) * rn / 16384,   # ncu, 16 384-ray capture scaled
                         "bytes_per_row": 292}}


def time_cost_volume(torch, pg, flush, peaks):
    """Second metric of BASELINE.json: cost-volume voxels/s at configs[0] (256x512, C32, D64, 2 views)."""
    dev = flush.device
    B, Hc, Wc, C, D = 1, 256, 512, 32, 64
    g = torch.Generator(device=dev).manual_seed(0)
    images = torch.randn(B, 2, Hc, Wc, C, device=dev, generator=g)
    rots = torch.eye(3, device=dev).expand(B, 2, 3, 3).contiguous()
    trans = torch.zeros(B, 2, 3, device=dev)
    trans[:, 0, 2], trans[:, 1, 2] = 0.5, -0.5
    depths = torch.linspace(0.1, 10, D, device=dev)
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    f = lambda: pg.calculate_cost_volume_erp(args, images, depths, trans, rots)
    for _ in range(3):
        f()
    REPS = 5   # back-to-back calls inside one event pair: keeps the queue fed, so the host-side wrapper (output allocation, the
               # err-flag fill, ctypes) is not timed as GPU idle; every call streams 1.1 GB through the 126 MB L2, so no call
               # finds its inputs cached by the previous one

    def median_ms(fn, n=7, reps=REPS):
        for _ in range(2):
            fn()
        t = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) / reps)
        return sorted(t)[len(t) // 2]

    from panogrf_b200 import spherical_cost_volume as scv
    f()                              # one checked call: raises if any uv left [-1,1] (reads the device flag = a host sync)
    scv._CHECK_UV = False            # timed calls: device work only, no per-call flag read-back
    ms = median_ms(f, n=9)
    vox = B * D * Hc * Wc
    alg = vox * C * 4 + B * 2 * Hc * Wc * C * 4 + D * 4

    ms_cl = median_ms(lambda: pg.calculate_cost_volume_erp(args, images, depths, trans, rots, out_layout="bdhwc"))
    # backward: d/d(images) of the same volume (reads the 1.07 GB upstream gradient once, vector atomics into 2 maps)
    img_g = images.clone().requires_grad_(True)
    out = pg.calculate_cost_volume_erp(args, img_g, depths, trans, rots, out_layout="bdhwc")
    gout = torch.ones_like(out)
    ms_bwd = median_ms(lambda: torch.autograd.grad(out, img_g, gout, retain_graph=True), n=5)
    del out, gout
    scv._CHECK_UV = True
    return {"workload": "configs[0]: 2 views 256x512 C32 D64 abs_diff, reference layout (B,D,C,H,W)",
            "timing": f"median of groups of {REPS} back-to-back calls, 256 MiB L2 flush before each group",
            "voxels_per_s": vox / ms * 1e3, "ms": ms,
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg / ms / 1e6 / peaks["hbm_gbs"], "traffic": 0.066685e9 + 1.016099e9,   # ncu, profiles/r1_final_ncu_summary.md
                         "bytes_per_voxel": alg / vox, "peak_src": peaks["src"]},
            "channels_last": {"ms": ms_cl, "voxels_per_s": vox / ms_cl * 1e3, "frac": alg / ms_cl / 1e6 / peaks["hbm_gbs"]},
            "backward": {"ms": ms_bwd, "voxels_per_s": vox / ms_bwd * 1e3, "frac": alg / ms_bwd / 1e6 / peaks["hbm_gbs"],
                         "note": "grad w.r.t. feature maps through torch.autograd.Function (includes the zero-fill of grad_images)"}}


def time_depth_guided(torch, que_d, ref_d):
    """Depth-prior sample placement (diner branch): every ray x 1000 linear candidates x RFN views -> 64 samples, one kernel."""
    from panogrf_b200.render_ops import depth_guided_placement
    dev = que_d["coords"].device
    rn = que_d["coords"].shape[1]
    g = torch.Generator(device=dev).manual_seed(0)
    cfg = {"dataset_name": "m3d", "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0, "n_candidates": 1000,
           "n_samples": 64, "n_gaussian": 15, "backface_culling": True, "contain_uniform": False}
    ref = dict(ref_d)
    ref["mvs_depth"] = 3.0 + torch.rand(RFN, 1, H, W, device=dev, generator=g)
    ref["mvs_uncert"] = torch.full((RFN, 1, H, W), 0.01, device=dev)
    ref["mvs_normal"] = torch.randn(RFN, 3, H, W, device=dev, generator=g)
    fill, ga = torch.rand(rn, 64, device=dev, generator=g), torch.randn(rn, 15, device=dev, generator=g)
    f = lambda: depth_guided_placement(cfg, que_d, ref, fill, ga)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {"workload": f"{rn} rays x 1000 candidates x {RFN} views -> 64 samples/ray (49 likelihood + 15 gaussian, fill-up, sort)",
            "ms": ms, "candidate_views_per_s": rn * 1000.0 * RFN / ms * 1e3,
            "note": "compute bound (atan2/acos/erf per candidate-view); writes only (rn,64) floats"}


if __name__ == "__main__":
    main()
