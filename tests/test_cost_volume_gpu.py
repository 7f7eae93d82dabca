"""GPU parity: fused sm_100a cost-volume kernel (through the C ABI) vs the CPU oracle and the
reference's golden outputs (HOT 1)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import assert_close, load_golden  # noqa: E402

pytestmark = pytest.mark.gpu

from oracle import cost_volume as ocv  # noqa: E402


def _dev(x):
    return None if x is None else x.cuda()


def _run_kernel(c, g, **kw):
    import panogrf_b200 as pg
    args = {"dataset_name": c["dataset"], "contain_dnet": c["per_pixel"], "mono_uncertainty": False}
    dv = _dev(g["depth_volume"]) if c["per_pixel"] else None
    if c.get("mv"):
        return pg.calculate_cost_volume_erp_multiview(args, _dev(g["images"]), _dev(g["depths"]), _dev(g["trans"]),
                                                      _dev(g["rots"]), depth_volume=dv, cost_type=c["cost_type"],
                                                      curr_idx=c["curr_idx"], **kw)
    return pg.calculate_cost_volume_erp(args, _dev(g["images"]), _dev(g["depths"]), _dev(g["trans"]),
                                        _dev(g["rots"]), depth_volume=dv, cost_type=c["cost_type"], **kw)


@pytest.mark.parametrize("name", list(cases.CV_CASES))
@pytest.mark.parametrize("layout", ["bdchw", "bdhwc", "bcdhw"])
def test_kernel_matches_reference_golden(name, layout):
    c = cases.CV_CASES[name]
    g = load_golden(name)
    out = _run_kernel(c, g, out_layout=layout)
    assert tuple(out.shape) == tuple(g["out"].shape)
    if layout == "bdchw" and not c.get("mv"):
        B, D, H, W, C = out.shape
        assert out.stride() == (D * C * H * W, C * H * W, W, 1, H * W)   # reference's permuted view (:340)
    assert_close(out, g["out"], atol=2e-4, max_bad_frac=2e-4, what=f"{name}/{layout}")


@pytest.mark.parametrize("name", ["cv_m3d_volume", "cv_mv5_volume"])
def test_group_wise_epilogue(name):
    c = cases.CV_CASES[name]
    g = load_golden(name)
    out = _run_kernel(c, g, groups=8)
    expect = ocv.group_mean(g["out"], 8)
    assert_close(out, expect, atol=1e-4, max_bad_frac=2e-4, what=name + "/groups")


def test_smooth_features_tight_tolerance():
    """Low-pass features: bilinear error no longer dominated by white noise -> pure rtol 1e-4 (+1e-5)."""
    name = "cv_m3d_volume"
    c = cases.CV_CASES[name]
    inp = cases.make_cv_inputs(name, seed=3, smooth_feats=True)
    expect = ocv.calculate_cost_volume_erp(inp["args"], inp["images"], inp["depths"], inp["trans"], inp["rots"],
                                           depth_volume=inp["depth_volume"])
    g = dict(images=inp["images"], depths=inp["depths"], trans=inp["trans"], rots=inp["rots"],
             depth_volume=inp["depth_volume"])
    out = _run_kernel(c, g)
    assert_close(out, expect, atol=1e-5, max_bad_frac=2e-4, what="smooth")


def test_config1_size_vs_oracle():
    """BASELINE config 1 shape (256x512, C32, D64, 2 views): full comparison against the CPU oracle."""
    import panogrf_b200 as pg
    gen = torch.Generator().manual_seed(0)
    B, H, W, C, D = 1, 256, 512, 32, 64
    # CNN-like (band-limited) features: with white noise the comparison measures nothing but the fp32
    # rounding of pixel coordinates of magnitude ~500 (1 ulp = 3e-5 px) times the feature gradient
    images = cases.smooth(torch.randn(B, 2, H, W, C, generator=gen), passes=4) * 4.0   # std 0.68, grad std 0.28
    rots = torch.eye(3).expand(B, 2, 3, 3).contiguous()
    trans = torch.tensor([[[0., 0., 0.5], [0., 0., -0.5]]])
    depths = torch.linspace(0.1, 10.0, D)
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    expect = ocv.calculate_cost_volume_erp(args, images, depths, trans, rots)
    out = pg.calculate_cost_volume_erp(args, images.cuda(), depths.cuda(), trans.cuda(), rots.cuda())
    # Spherical coordinates are ill-conditioned in two places, where two correct fp32 implementations
    # (the reference on CPU vs on CUDA included) disagree by more than 1e-4 px in the sampled position:
    #  (a) voxels that land within a few cm of the source camera centre (next to the epipole at depth ~
    #      baseline): the direction of a ~0-length vector, angle error ~ 1e-7 * depth / radius;
    #  (b) voxels that project next to a pole of the source panorama: longitude = atan2 of two ~0 numbers
    #      (measured on B200: <= 7e-4 px in x for |iy - pole| < 0.05 px, tools/debug_cv_coords.py).
    # Everything else must meet the stated tolerance; the ill-conditioned rest is bounded in count and size.
    depth = depths.view(1, D, 1, 1).expand(B, D, H, W)
    _, v, radius = ocv.sweep_uv("m3d", depth, rots[:, 1], trans[:, 1], rots[:, 0], trans[:, 0], return_radius=True)
    well = ((radius > 0.1 * depth) & (v.abs() < 1 - 8.0 / H))[..., None].expand_as(expect)
    assert float(well.float().mean()) > 0.95
    o = out.cpu()
    assert_close(o[well], expect[well], atol=1e-4, max_bad_frac=1e-5, what="config1/well-conditioned")
    assert_close(o, expect, atol=1e-4, max_bad_frac=1e-3, what="config1/all")
    assert float((o - expect).abs().max()) < 0.05
    out_cl = pg.calculate_cost_volume_erp(args, images.cuda(), depths.cuda(), trans.cuda(), rots.cuda(),
                                          out_layout="bdhwc")
    assert torch.equal(out_cl, out.contiguous())          # layouts are bit-identical


def test_sampled_coordinates_config1():
    """Warp a coordinate ramp with cost_type='none': the output IS the sampled source position, so the
    geometry (ray, rigid transform, atan2/acos, uv mapping) is compared with the oracle in pixel units."""
    import panogrf_b200 as pg
    B, H, W, D = 1, 256, 512, 64
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.stack([xs, ys, torch.zeros_like(xs), torch.zeros_like(xs)], -1)
    images = torch.stack([img, img], 0)[None]
    rots = torch.eye(3).expand(B, 2, 3, 3).contiguous()
    trans = torch.tensor([[[0., 0., 0.5], [0., 0., -0.5]]])
    depths = torch.linspace(0.1, 10.0, D)
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    out = pg.calculate_cost_volume_erp(args, images.cuda(), depths.cuda(), trans.cuda(), rots.cuda(),
                                       cost_type="none", out_layout="bdhwc").cpu()
    depth = depths.view(1, D, 1, 1).expand(B, D, H, W)
    u, v, radius = ocv.sweep_uv("m3d", depth, rots[:, 1], trans[:, 1], rots[:, 0], trans[:, 0], return_radius=True)
    ix, iy = ((u + 1) / 2) * (W - 1), ((v + 1) / 2) * (H - 1)
    ex, ey = (out[..., 0] - ix).abs(), (out[..., 1] - iy).abs()
    # longitude error scales with 1/sin(phi_src) and both with depth/radius: strict bound where both are benign
    well = (radius > 0.25 * depth) & (v.abs() < 0.8)
    assert float(well.float().mean()) > 0.7
    assert float(ex[well].max()) < 5e-4 and float(ey[well].max()) < 5e-4      # 1e-6 of the 512-px extent (1 ulp = 3e-5 px)
    assert float(ex.mean()) < 3e-5 and float(ey.mean()) < 3e-5
    assert float(ex.max()) < 5e-3 and float(ey.max()) < 5e-3                  # poles / epipole, see config1 test


def test_full_size_properties():
    """512x1024, D128 (config 3 per-GPU shape): size-independent properties instead of an oracle run."""
    import panogrf_b200 as pg
    gen = torch.Generator(device="cuda").manual_seed(1)
    B, H, W, C, D = 1, 512, 1024, 32, 128
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    rots = torch.eye(3, device="cuda").expand(B, 2, 3, 3).contiguous()
    depths = torch.linspace(0.1, 10.0, D, device="cuda")
    # (1) constant image, any pose -> abs_diff cost is exactly zero, 'none' returns the constant
    const = torch.full((B, 2, H, W, C), 0.625, device="cuda")
    trans = torch.tensor([[[0., 0., 0.5], [0., 0., -0.5]]], device="cuda")
    out = pg.calculate_cost_volume_erp(args, const, depths, trans, rots, out_layout="bdhwc")
    assert float(out.abs().max()) <= 1e-6
    del out
    # (2) linearity of the warp in the source features: cv_none(a*x + y) == a*cv_none(x) + cv_none(y)
    D2 = 16
    x = torch.randn(B, 2, H, W, C, device="cuda", generator=gen)
    y = torch.randn(B, 2, H, W, C, device="cuda", generator=gen)
    f = lambda im: pg.calculate_cost_volume_erp(args, im, depths[:D2], trans, rots, cost_type="none",
                                                out_layout="bdhwc")
    lhs = f(2.0 * x + y)
    rhs = 2.0 * f(x) + f(y)
    assert float((lhs - rhs).abs().max()) < 1e-4
    # (3) multi-view with identical source views equals the 2-view volume
    img3 = torch.stack([x[:, 0], x[:, 1], x[:, 0], x[:, 0]], 1).contiguous()
    rots4 = torch.eye(3, device="cuda").expand(B, 4, 3, 3).contiguous()
    trans4 = torch.stack([trans[:, 0], trans[:, 1], trans[:, 0], trans[:, 0]], 1).contiguous()
    mv = pg.calculate_cost_volume_erp_multiview(args, img3, depths[:D2], trans4, rots4, curr_idx=1)
    two = pg.calculate_cost_volume_erp(args, x, depths[:D2], trans, rots, out_layout="bdhwc")
    assert float((mv - two).abs().max()) < 1e-5


def test_uv_range_assert_and_errors():
    import panogrf_b200 as pg
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    images = torch.randn(1, 2, 8, 32, 8, device="cuda")
    rots = torch.eye(3, device="cuda").expand(1, 2, 3, 3).contiguous()
    trans = torch.zeros(1, 2, 3, device="cuda")
    from panogrf_b200 import spherical_cost_volume as scv
    # depth 0 with identical poses -> radius 0 -> NaN uv -> the reference's assert (:191) fires.
    # default ("deferred"): the call itself never blocks; the flag is looked at by check_pending() or by the next call
    scv.check_pending()
    pg.calculate_cost_volume_erp(args, images, torch.zeros(2, device="cuda"), trans, rots)
    with pytest.raises(AssertionError, match="Wrong UV mapping"):
        scv.check_pending()
    scv.check_pending()                                   # reported once
    pg.calculate_cost_volume_erp(args, images, torch.zeros(2, device="cuda"), trans, rots)
    torch.cuda.synchronize()
    with pytest.raises(AssertionError, match="Wrong UV mapping"):
        pg.calculate_cost_volume_erp(args, images, torch.ones(2, device="cuda"), trans, rots)   # the NEXT call raises
    scv.check_pending()
    try:                                                  # "sync": the reference's behaviour, one host sync per call
        scv.UV_CHECK = "sync"
        with pytest.raises(AssertionError, match="Wrong UV mapping"):
            pg.calculate_cost_volume_erp(args, images, torch.zeros(2, device="cuda"), trans, rots)
    finally:
        scv.UV_CHECK = "deferred"
    with pytest.raises(ValueError):
        pg.calculate_cost_volume_erp(args, images, torch.ones(2, device="cuda"), trans, rots, cost_type="ssd")
    with pytest.raises(Exception):
        pg.calculate_cost_volume_erp({"dataset_name": "nope", "contain_dnet": False}, images,
                                     torch.ones(2, device="cuda"), trans, rots)


def test_host_entry_point_matches_device_path():
    """The *_host C-ABI call (H2D + kernel + D2H) returns the same bits as the device-pointer call."""
    import ctypes
    import numpy as np
    import panogrf_b200 as pg
    from panogrf_b200 import _lib
    name = "cv_m3d_scalar"
    c = cases.CV_CASES[name]
    g = load_golden(name)
    dev_out = _run_kernel(c, g, out_layout="bdhwc").cpu().numpy()
    lib = _lib.load()
    B, S, H, W, C = g["images"].shape
    D = g["depths"].numel()
    out = np.empty((B, D, H, W, C), np.float32)
    arr = lambda t: np.ascontiguousarray(t.numpy(), np.float32)
    im, dp, ro, tr = arr(g["images"]), arr(g["depths"]), arr(g["rots"]), arr(g["trans"])
    views = (ctypes.c_int * 1)(0)
    rc = lib.pgrf_cost_volume_host(im.ctypes.data, B, S, H, W, C, dp.ctypes.data, None, D, ro.ctypes.data,
                                   tr.ctypes.data, 1, views, 1, 0.0, 0, 0, 1, 0, out.ctypes.data)
    assert rc == 0, _lib.last_error()
    assert np.array_equal(out, dev_out)


# ---- backward (torch.autograd.Function around pgrf_cost_volume_bwd) -------------------------------------------

@pytest.mark.parametrize("name", cases.CV_BWD_CASES)
@pytest.mark.parametrize("layout", ["bdchw", "bdhwc"])
def test_backward_matches_reference_autograd(name, layout):
    """d(sum(out*w))/d(images) against the gradient the reference's own autograd graph produced (tests/golden/*_bwd)."""
    from panogrf_b200 import spherical_cost_volume as scv
    inp = cases.make_cv_inputs(name)
    images = inp["images"].cuda().requires_grad_(True)
    kw = dict(depth_volume=None if inp["depth_volume"] is None else inp["depth_volume"].cuda(), cost_type=inp["cost_type"],
              out_layout=layout)
    if inp["mv"]:
        out = scv.calculate_cost_volume_erp_multiview(inp["args"], images, inp["depths"], inp["trans"].cuda(), inp["rots"].cuda(),
                                                      curr_idx=inp["curr_idx"], **kw)
    else:
        out = scv.calculate_cost_volume_erp(inp["args"], images, inp["depths"], inp["trans"].cuda(), inp["rots"].cuda(), **kw)
    w = cases.cv_bwd_weight(name, out.shape).cuda()
    (out * w).sum().backward()
    gold = load_golden(name + "_bwd")["grad_images"]
    assert images.grad.shape == gold.shape
    # white-noise features: a 1-ulp footprint difference flips sign(warp - ref) for a few taps near ties
    assert_close(images.grad, gold, rtol=1e-4, atol=2e-4, max_bad_frac=3e-3, what=f"{name}/{layout}")


def test_backward_group_mean_and_linearity():
    """groups>0 (fused group-wise mean) backward == backward of the explicit mean; gradient is linear in grad_out."""
    from panogrf_b200 import spherical_cost_volume as scv
    name = "cv_m3d_volume"
    inp = cases.make_cv_inputs(name)
    dv = inp["depth_volume"].cuda()
    t, r = inp["trans"].cuda(), inp["rots"].cuda()

    def grad_of(groups, scale):
        images = inp["images"].cuda().requires_grad_(True)
        out = scv.calculate_cost_volume_erp(inp["args"], images, inp["depths"], t, r, depth_volume=dv, groups=groups)
        if groups == 0:
            B, D, H, W, C = out.shape
            out = out.reshape(B, D, H, W, 8, C // 8).mean(-1).permute(0, 4, 1, 2, 3)     # (B,G,D,H,W)
        gen = torch.Generator().manual_seed(7)
        w = torch.randn(out.shape, generator=gen).cuda() * scale
        (out * w).sum().backward()
        return images.grad

    g_fused, g_explicit = grad_of(8, 1.0), grad_of(0, 1.0)
    assert_close(g_fused, g_explicit, rtol=1e-4, atol=1e-5, what="group-mean backward")
    assert_close(grad_of(8, 3.0), 3.0 * g_fused, rtol=1e-4, atol=1e-5, what="linearity")


def test_backward_full_size_properties():
    """configs[0] size: 'none' cost -> every source texel receives exactly the sum of bilinear weights that the forward
    reads it with, so sum(grad_src) == sum(grad_out) (partition of unity), and the reference view gets no gradient."""
    from panogrf_b200 import spherical_cost_volume as scv
    B, H, W, C, D = 1, 256, 512, 32, 64
    gen = torch.Generator().manual_seed(0)
    images = torch.randn(B, 2, H, W, C, generator=gen).cuda().requires_grad_(True)
    rots = torch.eye(3).repeat(B, 2, 1, 1).cuda()
    trans = torch.tensor([[[0.0, 0.0, -0.5], [0.0, 0.0, 0.5]]]).cuda()
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    out = scv.calculate_cost_volume_erp(args, images, torch.linspace(0.5, 10, D), trans, rots, cost_type="none", out_layout="bdhwc")
    w = torch.rand(out.shape, device="cuda")
    (out * w).sum().backward()
    g = images.grad
    assert float(g[:, 1].abs().max()) == 0.0
    total_in, total_out = float(w.double().sum()), float(g[:, 0].double().sum())
    assert abs(total_in - total_out) <= 1e-4 * total_in
    # abs_diff: ref gradient = -(sum over d of the warped-side gradient signs): antisymmetry of the two sides
    images2 = images.detach().clone().requires_grad_(True)
    out2 = scv.calculate_cost_volume_erp(args, images2, torch.linspace(0.5, 10, D), trans, rots, cost_type="abs_diff",
                                         out_layout="bdhwc")
    out2.sum().backward()
    g2 = images2.grad
    assert abs(float(g2[:, 0].double().sum()) + float(g2[:, 1].double().sum())) <= 1e-3 * float(g2[:, 1].double().abs().sum())


@pytest.mark.parametrize("layout", ["bdchw", "bcdhw"])
@pytest.mark.parametrize("mv", [False, True])
def test_planar_tma_store_path_is_bit_identical(layout, mv):
    """W % 128 == 0, C = 32: the planar layouts can leave through the TMA engine (register transpose + swizzled warp tile +
    cp.async.bulk.tensor.5d, csrc/cost_volume.cu: cost_volume_planar_tma_kernel; opt-in, debug knob cv_tma).  Same arithmetic as
    the LSU path -> identical bits,
    incl. per-pixel depth volumes, a ragged last depth chunk and the multi-view mean."""
    import panogrf_b200 as pg
    from panogrf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    B, S, H, W, C, D = 2, (4 if mv else 2), 16, 256, 32, 11
    images = torch.randn(B, S, H, W, C, device="cuda", generator=g)
    rots = cases.small_rotations(torch.Generator().manual_seed(1), B, S, 4.0).cuda()
    trans = (torch.randn(B, S, 3, generator=torch.Generator().manual_seed(2)) * 0.3).cuda()
    dvol = (0.5 + 8.0 * torch.rand(B, D, H, W, device="cuda", generator=g)).sort(dim=1).values
    args = {"dataset_name": "m3d", "contain_dnet": True, "mono_uncertainty": False}
    fn = pg.calculate_cost_volume_erp_multiview if mv else pg.calculate_cost_volume_erp
    outs = {}
    for tma in (1, 0):
        _lib.check(lib.pgrf_debug_set(b"cv_tma", tma), "pgrf_debug_set")
        try:
            outs[tma] = fn(args, images, None, trans, rots, depth_volume=dvol, out_layout=layout).contiguous()
            torch.cuda.synchronize()
        finally:
            lib.pgrf_debug_set(b"cv_tma", 0)
    assert torch.equal(outs[1], outs[0])
    assert float(outs[1].abs().sum()) > 0


def test_bf16_channels_last_layout_and_regulariser_pipeline():
    """out_layout="bdhwc_bf16": the sweep stores exactly bf16(round-to-nearest) of its fp32 channels-last result, and the tensor-core
    regulariser consumes that storage in place with the same result as the fp32 volume + its own conversion pass."""
    import panogrf_b200 as pg
    from panogrf_b200 import regulariser as reg
    torch.manual_seed(3)
    B, H, W, C, D = 1, 16, 32, 16, 8
    images = torch.randn(B, 2, H, W, C, device="cuda")
    rots = torch.eye(3, device="cuda").expand(B, 2, 3, 3).contiguous()
    trans = torch.zeros(B, 2, 3, device="cuda")
    trans[:, 0, 0] = 0.3
    depths = torch.linspace(0.6, 9.0, D, device="cuda")
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    cv32 = pg.calculate_cost_volume_erp(args, images, depths, trans, rots, out_layout="bdhwc")
    cv16 = pg.calculate_cost_volume_erp(args, images, depths, trans, rots, out_layout="bdhwc_bf16")
    assert cv16.dtype == torch.bfloat16 and cv16.shape == cv32.shape and cv16.is_contiguous()
    assert torch.equal(cv16, cv32.to(torch.bfloat16))
    net = reg.CostRegulariser3D(3).cuda()                      # 2^(3+1) = 16 input channels
    y32 = net(cv32.permute(0, 4, 1, 2, 3))                     # fp32 view -> conversion kernel -> convolutions
    y16 = net(cv16.permute(0, 4, 1, 2, 3))                     # bf16 storage consumed in place
    assert torch.equal(y16, y32)
    with pytest.raises(Exception):
        pg.calculate_cost_volume_erp(args, images[..., :8].contiguous(), depths, trans, rots, out_layout="bdhwc_bf16")   # C % 16


@pytest.mark.parametrize("C", [32, 64])
def test_backward_run_merged_variant_agrees(C):
    """The opt-in run-merged backward kernel (debug knob cv_bwd_variant = 1) against the default kernel: same gradient up to the
    summation order of the float atomics."""
    from panogrf_b200 import _lib, spherical_cost_volume as scv
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    B, S, H, W, D = 1, 2, 32, 96, 12                      # W is not a multiple of the CTA tile: ragged runs
    images = torch.randn(B, S, H, W, C, generator=g).cuda().requires_grad_(True)
    rots = torch.eye(3).expand(B, S, 3, 3).contiguous().cuda()
    trans = torch.zeros(B, S, 3)
    trans[:, 0, 2], trans[:, 1, 0] = 0.4, -0.3
    depths = torch.linspace(0.3, 8, D).cuda()
    args = {"dataset_name": "m3d", "contain_dnet": False, "mono_uncertainty": False}
    out = scv.calculate_cost_volume_erp(args, images, depths, trans.cuda(), rots)
    gout = torch.randn(out.shape, generator=g).cuda()
    grads = []
    try:
        for variant in (0, 1):
            lib.pgrf_debug_set(b"cv_bwd_variant", variant)
            grads.append(torch.autograd.grad(out, images, gout, retain_graph=True)[0])
    finally:
        lib.pgrf_debug_set(b"cv_bwd_variant", 0)
    assert_close(grads[1], grads[0], rtol=1e-4, atol=1e-4 * float(grads[0].abs().max()), what=f"run-merged backward C={C}")
