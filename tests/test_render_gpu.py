"""GPU parity: fused render-path kernels (through the C ABI) vs the reference's golden outputs and
the CPU oracle (HOT 2-4)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from test_oracle_render import split_golden  # noqa: E402
from util import assert_close, load_golden  # noqa: E402

from oracle import render as orender  # noqa: E402

pytestmark = pytest.mark.gpu


def cuda_dict(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


def build_renderer(cfg, W):
    import panogrf_b200 as pg
    net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
    missing, unexpected = net.load_state_dict(W, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    return net


@pytest.mark.parametrize("name", list(cases.RENDER_CASES))
def test_render_matches_reference_golden(name):
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, gold = split_golden(load_golden(name))
    net = build_renderer(cfg, W)
    out = net.render_impl(cuda_dict(que), cuda_dict(ref), False, keep_hit_prob=True)
    torch.cuda.synchronize()
    for k, v in gold.items():
        if k.startswith("ray_mask"):
            assert out[k].dtype == torch.bool and torch.equal(out[k].cpu(), v.bool())
            continue
        # fp32 rtol 1e-4 (north_star); atol covers values that are sums of O(1) terms cancelling to ~0.
        # render_c2f_all: merged coarse/fine depths that are EQUAL (or equal to ~1 ulp) sort either way under the reference's
        # unstable sort (CPU and CUDA sorts already disagree): the per-sample outputs of such a pair swap places, the
        # composited pixel / depth / variance (checked strictly) do not change
        frac = 0.2 if (cfg.get("render_c2f_all") and k in ("hit_prob_nr_fine", "colors_nr_fine", "density_nr_fine")) else 0.0
        assert_close(out[k], v.float(), rtol=1e-4, atol=5e-5, max_bad_frac=frac, what=f"{name}/{k}")


@pytest.mark.parametrize("name", ["render_m3d_2src", "render_replica", "render_m3d_vis_nodisp"])
def test_prj_dict_intermediates(name):
    """project_points_dict / predict_proj_ray_prob / get_img_feats intermediates of the coarse pass."""
    cfg, _, _ = cases.make_render_inputs(name)
    cfg = {**cfg, "use_hierarchical_sampling": False}     # coarse pass only: the dumps are per pass
    que, ref, W, _ = split_golden(load_golden(name))
    W = {k: v for k, v in W.items() if not k.startswith("fine_")}
    net = build_renderer(cfg, W)
    rn = que["coords"].shape[1]
    dn = cfg["depth_sample_num"]
    depth = orender.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, dn, cfg["use_disp"])
    o = orender.render_by_depth(cfg, W, que, ref, depth, False, return_prj=True)
    prj = o["prj"]
    rfn = ref["imgs"].shape[0]
    q, r = cuda_dict(que), cuda_dict(ref)
    ctx = net._context(q, r)
    ctx["prj_dbg"] = torch.zeros(rfn, rn * dn, 6, device="cuda")
    ctx["feat_dbg"] = torch.zeros(rfn, rn * dn, 67, device="cuda")
    ctx["prob_dbg"] = torch.zeros(rfn, rn * dn, 3, device="cuda")
    net.render_impl(q, r, False, _ctx=ctx)
    torch.cuda.synchronize()
    pd, fd, pr = ctx["prj_dbg"].cpu(), ctx["feat_dbg"].cpu(), ctx["prob_dbg"].cpu()
    sh = lambda t: t.reshape(rfn, rn * dn, -1)
    assert_close(pd[..., 0:2], sh(prj["pts"]), rtol=1e-4, atol=2e-3, what="pts")       # pixels: |x|<=W, 1e-4 rel of W
    assert_close(pd[..., 2:3], sh(prj["depth"]), rtol=1e-4, atol=1e-5, what="depth")
    assert_close(pd[..., 3:6], sh(prj["dir"]), rtol=1e-4, atol=1e-5, what="dir")
    assert_close(fd[..., 0:32], sh(prj["ray_feats"]), rtol=1e-4, atol=2e-4, what="ray_feats")
    assert_close(fd[..., 32:35], sh(prj["rgb"]), rtol=1e-4, atol=2e-5, what="rgb")
    assert_close(fd[..., 35:67], sh(prj["img_feats"]), rtol=1e-4, atol=2e-4, what="img_feats")
    assert_close(pr[..., 0:1], sh(prj["alpha"]), rtol=1e-4, atol=2e-4, what="alpha")
    assert_close(pr[..., 1:2], sh(prj["vis"]), rtol=1e-4, atol=1e-5, what="vis")
    assert_close(pr[..., 2:3], sh(prj["hit_prob"]), rtol=1e-4, atol=1e-5, what="hit_prob")


def test_fine_sampling_bins_bit_exact():
    """Depth-bin indices and sample ordering must match the oracle bit-exactly when both consume the
    SAME coarse hit_prob: run the coarse pass on the GPU, feed its hit_prob to the oracle sampler."""
    name = "render_m3d_2src"
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, _ = split_golden(load_golden(name))
    net = build_renderer(cfg, W)
    rn = que["coords"].shape[1]
    dn, fdn = cfg["depth_sample_num"], cfg["fine_depth_sample_num"]
    q, r = cuda_dict(que), cuda_dict(ref)
    ctx = net._context(q, r)
    ctx["fine_inds"] = torch.zeros(rn, fdn, dtype=torch.int32, device="cuda")
    out = net.render_impl(q, r, False, _ctx=ctx, keep_hit_prob=True)
    torch.cuda.synchronize()
    hit = out["hit_prob_nr"].cpu()
    depth = orender.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, dn, cfg["use_disp"])
    fine, inds = orender.sample_fine_depth(depth, hit, que["depth_range"], fdn, cfg["use_disp"], return_indices=True)
    assert torch.equal(ctx["fine_inds"].cpu().long(), inds[0]), "searchsorted bin indices differ"
    expect = torch.sort(fine, -1)[0]
    got = out["que_depth_fine"].cpu()
    assert bool((got[..., 1:] >= got[..., :-1]).all()), "fine samples not sorted"
    # same bins and the same fp32 formula; the affine map back from normalised inverse depth cancels
    # (x*1.93 - 2.0 ~ -0.07), so a last-place difference upstream is worth ~30 ulp in the depth
    assert_close(got, expect, rtol=1e-5, atol=0.0, what="fine depths")


def test_render_full_view_chunking_and_properties():
    """Whole 64x128 view through render(): chunked launches == single launch; outputs are convex
    combinations (colours within the source range), hit_prob in [0,1], sum <= 1; depth within range."""
    name = "render_m3d_2src"
    _, _, _ = cases.make_render_inputs(name)
    cfg = cases.render_cfg(height=64, width=128, sample_num=16)
    _, ref, W, _ = split_golden(load_golden(name))
    # reuse the golden weights; maps are re-made at the larger size
    gen = torch.Generator().manual_seed(11)
    h, w, rfn = 64, 128, 2
    imgs = cases.smooth(torch.rand(rfn, h, w, 3, generator=gen), 1).permute(0, 3, 1, 2).contiguous()
    ref2 = {"imgs": imgs, "w2c": ref["w2c"], "depth_range": ref["depth_range"],
            "ray_feats": torch.randn(rfn, 32, h // 4, w // 4, generator=gen),
            "img_feats": torch.randn(rfn, 32, h // 2, w // 2, generator=gen)}
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    que = {"coords": coords, "c2w": torch.eye(4)[None, :3], "depth_range": torch.tensor([[0.5, 15.0]])}
    net = build_renderer(cfg, W)
    net.rays_per_launch = 1000          # ragged chunks
    a = net.render(cuda_dict(que), cuda_dict(ref2), False, keep_hit_prob=True)
    net.rays_per_launch = None
    b = net.render(cuda_dict(que), cuda_dict(ref2), False, keep_hit_prob=True)
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    hit = a["hit_prob_nr_fine"]
    assert float(hit.min()) >= 0 and float(hit.sum(-1).max()) <= 1 + 1e-5
    col = a["pixel_colors_nr_fine"]
    assert float(col.min()) >= -1e-5 and float(col.max()) <= float(imgs.max()) + 1e-5
    d = a["render_depth_fine"]
    assert float(d.min()) >= 0 and float(d.max()) <= 15.0 + 1e-3
    # spot-check 300 rays of the full view against the oracle
    idx = torch.randperm(h * w, generator=gen)[:300]
    q2 = dict(que)
    q2["coords"] = coords[:, idx]
    o = orender.render_rays(cfg, W, q2, ref2)
    assert_close(a["pixel_colors_nr_fine"][:, idx.cuda()], o["pixel_colors_nr_fine"], rtol=1e-4, atol=5e-5, what="view/rgb")
    assert_close(a["render_depth_fine"][:, idx.cuda()], o["render_depth_fine"], rtol=1e-4, atol=5e-5, what="view/depth")


def test_perpoint_loss_outputs():
    """cfg['perpoint_loss'] (renderer.py:314-316): per-sample weights and depths of both passes next to the composited outputs"""
    name = "render_m3d_2src"
    cfg, _, _ = cases.make_render_inputs(name)
    g = load_golden(name)
    que, ref, W, gold = split_golden(g)
    net = build_renderer({**cfg, "perpoint_loss": True}, W)
    out = net.render(cuda_dict(que), cuda_dict(ref), False)
    expect = orender.render_rays(cfg, W, que, ref, keep_hit_prob=True)
    assert_close(out["render_weights"], expect["hit_prob_nr"], rtol=1e-4, atol=2e-5, what="render_weights")
    assert_close(out["render_weights_fine"], expect["hit_prob_nr_fine"], rtol=1e-4, atol=2e-4, max_bad_frac=2e-3, what="render_weights_fine")
    assert_close(out["render_dvals_fine"], expect["que_depth_fine"], rtol=1e-4, atol=1e-4, max_bad_frac=2e-3, what="render_dvals_fine")
    assert out["render_dvals"].shape == out["render_weights"].shape
    assert "hit_prob_nr" not in out


def test_errors():
    import panogrf_b200 as pg
    cfg = cases.render_cfg()
    net = pg.NeuralRayBaseRenderer(cfg).cuda()
    _, que, ref = cases.make_render_inputs("render_m3d_2src")
    with pytest.raises(NotImplementedError):
        net.render_impl(cuda_dict(que), cuda_dict(ref), True)
    with pytest.raises(pg._lib.PanoGRFError):
        net.render_impl(que, ref, False)                       # CPU tensors: no fallback
    with pytest.raises(Exception):
        pg.NeuralRayBaseRenderer({**cfg, "dataset_name": "nope"})
    bad = cases.render_cfg(sample_num=16)
    bad["depth_sample_num"] = 24                               # posenc length mismatch (ibrnet.py:358)
    with pytest.raises(RuntimeError):
        pg.NeuralRayBaseRenderer(bad).cuda().render_impl(cuda_dict(que), cuda_dict(ref), False)


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core MLP path (tcgen05): north-star tolerance rtol 1e-2
# ------------------------------------------------------------------------------------------------

# Natural scale of every output: colours, probabilities and alphas live in [0, 1]; depths in the 0.5 .. 15 m sampling range.
DEPTH_SCALE = 14.5
BF16_ATOL = 5e-3      # of the natural scale: 0.5 % of full-scale colour (1.3 / 255), 7 cm of depth


def _close_bf16(actual, expected, what, scale=1.0, atol=BF16_ATOL, max_bad_frac=0.0):
    """north-star bf16 tolerance: |a-e| <= 1e-2*|e| + atol*scale.

    `scale` is the NATURAL scale of the quantity (1 for colours / hit probabilities, the depth range for depths), not the largest
    value that happens to occur in the test case.  The operands of every Linear layer are rounded to bf16 (unit roundoff 2^-9 =
    2e-3) and the random-init network chains ~12 such layers; measured on the goldens (tools/bf16_error_stats.py, B200):
    composited colour error <= 3.4e-3, depth error <= 4.1 cm, hit_prob error <= 6.6e-3 in the worst element, ~1e-3 / 1 cm / 5e-4 on
    average.  `density` has no natural scale (pre-activation of alpha): its tolerance is relative to max|e| (callers pass it)."""
    e = torch.as_tensor(expected).float().cpu()
    assert_close(actual, e, rtol=1e-2, atol=atol * scale, max_bad_frac=max_bad_frac, what=what)


def _check_bf16_pass(out, gold, name, suffix=""):
    _close_bf16(out["pixel_colors_nr"], gold["pixel_colors_nr" + suffix], f"{name}/pixel_colors_nr{suffix}")
    _close_bf16(out["render_depth"], gold["render_depth" + suffix], f"{name}/render_depth{suffix}", scale=DEPTH_SCALE)
    # per-sample quantities (no averaging over the ray): 1 % of the probability / colour scale in the worst element
    _close_bf16(out["hit_prob_nr"], gold["hit_prob_nr" + suffix], f"{name}/hit_prob_nr{suffix}", atol=1e-2)
    _close_bf16(out["colors_nr"], gold["colors_nr" + suffix], f"{name}/colors_nr{suffix}", atol=1e-2)
    d = gold["density_nr" + suffix]
    _close_bf16(out["density_nr"], d, f"{name}/density_nr{suffix}", scale=float(d.abs().max()), atol=2.5e-2)


@pytest.mark.parametrize("name", list(cases.RENDER_CASES))
def test_bf16_coarse_pass_matches_reference_golden(name):
    cfg, _, _ = cases.make_render_inputs(name)
    cfg = {**cfg, "mlp_dtype": "bf16"}
    que, ref, W, gold = split_golden(load_golden(name))
    net = build_renderer(cfg, W)
    out = net.render_impl(cuda_dict(que), cuda_dict(ref), False, keep_hit_prob=True)
    torch.cuda.synchronize()
    _check_bf16_pass(out, gold, name)


@pytest.mark.parametrize("name", ["render_m3d_2src", "render_m3d_vis_nodisp", "render_replica", "render_m3d_4src_all"])
def test_bf16_fine_pass_on_reference_sample_positions(name):
    """Fine networks on the SAME sample depths as the oracle (the end-to-end fine pass resamples from the coarse
    hit_prob, so with white-noise feature maps a 1 % change of hit_prob moves samples onto different features)."""
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, _ = split_golden(load_golden(name))
    o = orender.render_rays(cfg, W, que, ref, keep_hit_prob=True)
    fdepth = o["que_depth_fine"]
    net = build_renderer({**cfg, "mlp_dtype": "bf16"}, W)
    out = net.render_by_depth(fdepth.cuda(), cuda_dict(que), cuda_dict(ref), False, True)
    torch.cuda.synchronize()
    _check_bf16_pass(out, o, name, "_fine")


def test_bf16_full_view_end_to_end():
    """Whole 64x128 view, coarse + resampled fine pass, smooth (CNN-like) feature maps: final colours within 1e-2."""
    cfg = cases.render_cfg(height=64, width=128, sample_num=16)
    _, ref, W, _ = split_golden(load_golden("render_m3d_2src"))
    gen = torch.Generator().manual_seed(12)
    h, w, rfn = 64, 128, 2
    sm = lambda x: cases.smooth(x.permute(0, 2, 3, 1), 2).permute(0, 3, 1, 2).contiguous()
    ref2 = {"imgs": sm(torch.rand(rfn, 3, h, w, generator=gen)), "w2c": ref["w2c"], "depth_range": ref["depth_range"],
            "ray_feats": sm(torch.randn(rfn, 32, h // 4, w // 4, generator=gen)) * 3,
            "img_feats": sm(torch.randn(rfn, 32, h // 2, w // 2, generator=gen)) * 3}
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    que = {"coords": coords, "c2w": torch.eye(4)[None, :3], "depth_range": torch.tensor([[0.5, 15.0]])}
    a = build_renderer({**cfg, "mlp_dtype": "bf16"}, W).render(cuda_dict(que), cuda_dict(ref2), False)
    b = build_renderer(cfg, W).render(cuda_dict(que), cuda_dict(ref2), False)
    torch.cuda.synchronize()
    # coarse pass: same sample positions in both paths -> the plain bf16 tolerance
    _close_bf16(a["pixel_colors_nr"], b["pixel_colors_nr"], "view/pixel_colors_nr")
    _close_bf16(a["render_depth"], b["render_depth"], "view/render_depth", scale=DEPTH_SCALE)
    # fine pass: the inverse-CDF resampling is discontinuous in the coarse hit_prob (a 1e-3 change moves a sample across an
    # occlusion boundary of the scene), so single pixels may land on a different surface: bound the distribution —
    # mean within 2e-3 of the natural scale, 99 % of the pixels within the bf16 tolerance, none off by more than 10 % of scale
    for k, scale in (("pixel_colors_nr_fine", 1.0), ("render_depth_fine", DEPTH_SCALE)):
        e = b[k].float().cpu()
        err = (a[k].float().cpu() - e).abs()
        assert float(err.mean()) < 2e-3 * scale, (k, float(err.mean()))
        _close_bf16(a[k], e, "view/" + k, scale=scale, max_bad_frac=1e-2)
        assert float(err.max()) < 1e-1 * scale, (k, float(err.max()))


@pytest.mark.parametrize("rfn", [2, 4])
def test_production_shapes_max_samples(rfn):
    """The bench's sample counts at the kernel limits: 64 coarse + 64 fine samples merged (fine pass dn = 128 = the
    maximum), 2 and 4 source views (64- and 32-sample tiles), ragged ray count, against the oracle; fp32 1e-4, bf16 1e-2."""
    import panogrf_b200 as pg
    cfg = cases.render_cfg(height=64, width=128, hierarchical=True, fine_use_all=True, use_vis=True)
    cfg.pop("sample_num")
    cfg.update(depth_sample_num=64, fine_depth_sample_num=64, agg_net_cfg={"sample_num": 64}, fine_agg_net_cfg={"sample_num": 128})
    gen = torch.Generator().manual_seed(100 + rfn)
    torch.manual_seed(100 + rfn)
    net = pg.NeuralRayBaseRenderer(cfg)
    with torch.no_grad():
        for n_, p_ in net.named_parameters():
            if n_.endswith("bias"):
                p_.copy_(torch.randn(p_.shape, generator=gen) * 0.1)
    W = {k: v.detach().clone() for k, v in net.state_dict().items()}
    h, w = 64, 128
    imgs = cases.smooth(torch.rand(rfn, h, w, 3, generator=gen), 1).permute(0, 3, 1, 2).contiguous()
    rots = cases.small_rotations(gen, 1, rfn, 6.0)[0]
    trans = torch.randn(rfn, 3, generator=gen) * 0.3
    ref = {"imgs": imgs, "w2c": torch.cat([rots, trans[:, :, None]], -1).contiguous(),
           "depth_range": torch.tensor([[0.5, 15.0]]).repeat(rfn, 1),
           "ray_feats": cases.smooth(torch.randn(rfn, h // 4, w // 4, 32, generator=gen), 1).permute(0, 3, 1, 2).contiguous(),
           "img_feats": cases.smooth(torch.randn(rfn, h // 2, w // 2, 32, generator=gen), 1).permute(0, 3, 1, 2).contiguous()}
    perm = torch.randperm(h * w, generator=gen)[:45]
    coords = torch.stack([(perm % w).float(), (perm // w).float()], -1)[None]
    que = {"coords": coords, "c2w": torch.eye(4)[None, :3], "depth_range": torch.tensor([[0.5, 15.0]])}
    o = orender.render_rays(cfg, W, que, ref, keep_hit_prob=True)
    assert o["hit_prob_nr_fine"].shape == (1, 45, 128)
    net = net.cuda().eval()
    out = net.render(cuda_dict(que), cuda_dict(ref), False, keep_hit_prob=True)
    for k in ("pixel_colors_nr", "render_depth", "hit_prob_nr", "density_nr", "pixel_colors_nr_fine", "render_depth_fine",
              "que_depth_fine", "hit_prob_nr_fine", "density_nr_fine", "colors_nr_fine"):
        # inverse-CDF resampling divides by the cdf increment of a bin: in a nearly empty bin a 1e-7 difference of the coarse
        # hit_prob moves a fine sample by a few 1e-4 (and its per-sample outputs with it); one sample in a thousand may do so
        frac = 2e-3 if k in ("que_depth_fine", "hit_prob_nr_fine", "density_nr_fine", "colors_nr_fine") else 0.0
        assert_close(out[k], o[k], rtol=1e-4, atol=1e-4, max_bad_frac=frac, what=f"max-samples rfn={rfn} fp32 {k}")
    net16 = build_renderer({**cfg, "mlp_dtype": "bf16"}, W)
    out16 = net16.render(cuda_dict(que), cuda_dict(ref), False, keep_hit_prob=True)
    _close_bf16(out16["pixel_colors_nr"], o["pixel_colors_nr"], f"rfn={rfn} bf16 pixel_colors_nr")
    _close_bf16(out16["render_depth"], o["render_depth"], f"rfn={rfn} bf16 render_depth", scale=DEPTH_SCALE)
    _close_bf16(out16["hit_prob_nr"], o["hit_prob_nr"], f"rfn={rfn} bf16 hit_prob_nr", atol=1e-2)
    # the fine pass resamples from the (bf16) coarse hit_prob: smooth maps keep the composited colour within 1 % of full scale
    _close_bf16(out16["pixel_colors_nr_fine"], o["pixel_colors_nr_fine"], f"rfn={rfn} bf16 pixel_colors_nr_fine", atol=1e-2)
    assert out16["hit_prob_nr_fine"].shape == (1, 45, 128) and bool(torch.isfinite(out16["colors_nr_fine"]).all())


@pytest.mark.parametrize("variant", ["plain", "huge_logits", "one_view", "fine128"])
def test_bf16_tensor_core_attention_rays_kernel(variant):
    """dn = 64 / 128 take the rays kernel whose attention runs on tcgen05 (csrc/render_rays_tc.cu): a ragged ray count (odd:
    the last 2-ray tile is half empty), the exact-maximum pre-pass (query / key projections scaled so that |q||k| leaves the
    safe range of the Cauchy-Schwarz shift), the < 2 views mask rule (ibrnet.py:359-360: uniform attention) and one ray of 128
    samples per tile (two key blocks), each against the oracle and against the SIMT-attention kernel (debug knob rays_tc=0)."""
    import panogrf_b200 as pg
    from panogrf_b200 import _lib
    rfn = 1 if variant == "one_view" else 2
    dn = 128 if variant == "fine128" else 64
    cfg = cases.render_cfg(height=64, width=128, hierarchical=False)
    cfg.pop("sample_num")
    cfg.update(depth_sample_num=dn, agg_net_cfg={"sample_num": dn}, fine_agg_net_cfg={"sample_num": dn})
    gen = torch.Generator().manual_seed(300 + len(variant))
    torch.manual_seed(300 + len(variant))
    net = pg.NeuralRayBaseRenderer(cfg)
    with torch.no_grad():
        if variant == "huge_logits":
            net.agg_net.agg_impl.ray_attention.w_qs.weight.mul_(40.0)
            net.agg_net.agg_impl.ray_attention.w_ks.weight.mul_(40.0)
    W = {k: v.detach().clone() for k, v in net.state_dict().items()}
    h, w = 64, 128
    imgs = cases.smooth(torch.rand(rfn, h, w, 3, generator=gen), 1).permute(0, 3, 1, 2).contiguous()
    rots = cases.small_rotations(gen, 1, rfn, 6.0)[0]
    trans = torch.randn(rfn, 3, generator=gen) * 0.3
    ref = {"imgs": imgs, "w2c": torch.cat([rots, trans[:, :, None]], -1).contiguous(),
           "depth_range": torch.tensor([[0.5, 15.0]]).repeat(rfn, 1),
           "ray_feats": cases.smooth(torch.randn(rfn, h // 4, w // 4, 32, generator=gen), 1).permute(0, 3, 1, 2).contiguous(),
           "img_feats": cases.smooth(torch.randn(rfn, h // 2, w // 2, 32, generator=gen), 1).permute(0, 3, 1, 2).contiguous()}
    perm = torch.randperm(h * w, generator=gen)[:37]
    coords = torch.stack([(perm % w).float(), (perm // w).float()], -1)[None]
    que = {"coords": coords, "c2w": torch.eye(4)[None, :3], "depth_range": torch.tensor([[0.5, 15.0]])}
    o = orender.render_rays(cfg, W, que, ref, keep_hit_prob=True)
    net16 = build_renderer({**cfg, "mlp_dtype": "bf16"}, W)
    lib = _lib.load()
    outs = {}
    for tc in (1, 0):
        _lib.check(lib.pgrf_debug_set(b"rays_tc", tc), "pgrf_debug_set")
        try:
            outs[tc] = net16.render_impl(cuda_dict(que), cuda_dict(ref), False, keep_hit_prob=True)
            torch.cuda.synchronize()
        finally:
            lib.pgrf_debug_set(b"rays_tc", 1)
    # with huge logits the softmax is nearly one-hot: a bf16 rounding of q / k flips the winner for a few samples, so the
    # per-sample density is compared on the composited quantities only
    for tc in (1, 0):
        out = outs[tc]
        _close_bf16(out["pixel_colors_nr"], o["pixel_colors_nr"], f"{variant} tc={tc} pixel_colors_nr", atol=1e-2 if variant == "huge_logits" else BF16_ATOL)
        _close_bf16(out["render_depth"], o["render_depth"], f"{variant} tc={tc} render_depth", scale=DEPTH_SCALE, atol=1e-2 if variant == "huge_logits" else BF16_ATOL)
        if variant != "huge_logits":
            _close_bf16(out["hit_prob_nr"], o["hit_prob_nr"], f"{variant} tc={tc} hit_prob_nr", atol=1e-2)
            d = o["density_nr"]
            _close_bf16(out["density_nr"], d, f"{variant} tc={tc} density_nr", scale=float(d.abs().max()), atol=2.5e-2)
    assert bool(torch.isfinite(outs[1]["density_nr"]).all())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_perspective_query_rays_match_reference_golden(dtype):
    """is_perspec=True (network/render_ops.py:37-74): pinhole query rays (K, pose) instead of the ERP table — the kernels take the
    world-space directions from the caller (pgrf_render_args.ray_dirs); golden from the reference's render_impl(is_perspec=True).
    Both entry points: render_impl (per-pass calls) and render (whole-view C call)."""
    name = "render_m3d_perspec"
    cfg, _, _ = cases.make_perspec_inputs(name)
    que, ref, W, gold = split_golden(load_golden(name))
    net = build_renderer({**cfg, "mlp_dtype": dtype}, W)
    for call in ("render_impl", "render"):
        out = getattr(net, call)(cuda_dict(que), cuda_dict(ref), False, is_perspec=True)
        torch.cuda.synchronize()
        for k, scale in (("pixel_colors_nr", 1.0), ("render_depth", DEPTH_SCALE), ("pixel_colors_nr_fine", 1.0), ("render_depth_fine", DEPTH_SCALE)):
            if dtype == "fp32":
                assert_close(out[k], gold[k], rtol=1e-4, atol=5e-5, what=f"{call}/{k}")
            elif not k.endswith("_fine"):
                _close_bf16(out[k], gold[k], f"{call}/{k}", scale=scale)


@pytest.mark.parametrize("name", ["render_m3d_ft_range", "render_m3d_ft_range_all"])
def test_ft_depth_range_full_view_path(name):
    """que_imgs_info['ft_depth_range'] (fine_render_impl, renderer.py:438-456): render() runs the ray-batch loop with the prior-guided
    samples written between the passes == render_impl on all rays; without any valid prior it equals the plain render."""
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, gold = split_golden(load_golden(name))
    net = build_renderer({**cfg, "fused_ray_batch": 16}, W)          # several ray batches
    q, r = cuda_dict(que), cuda_dict(ref)
    full = net.render(dict(q), dict(r), False)
    for k in ("pixel_colors_nr_fine", "render_depth_fine", "density_nr_fine", "pixel_colors_nr"):
        assert_close(full[k], gold[k].float(), rtol=1e-4, atol=5e-5, what=f"{name}/render()/{k}")
    assert not any(k.startswith("hit_prob") for k in full)
    q0 = dict(q)
    q0["ft_depth_range"] = torch.zeros_like(q["ft_depth_range"])      # marker 0 < min_depth: no ray has a prior
    a = net.render(q0, dict(r), False)
    q1 = {k: v for k, v in q.items() if k != "ft_depth_range"}
    b = net.render(q1, dict(r), False)
    for k in ("pixel_colors_nr_fine", "render_depth_fine"):
        assert_close(a[k], b[k], rtol=1e-5, atol=1e-6, what=f"{name}/no prior/{k}")
