"""Seeded synthetic inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Everything is derived from (name, seed) so a fixture file only has to store the *reference output*
plus the inputs (kept too, so the fixtures stay valid if torch's RNG stream ever changes).
"""
import numpy as np
import torch


def small_rotations(gen, B, S, deg):
    ang = torch.randn(B, S, 3, generator=gen) * np.deg2rad(deg)
    K = torch.zeros(B, S, 3, 3)
    K[..., 0, 1], K[..., 0, 2] = -ang[..., 2], ang[..., 1]
    K[..., 1, 0], K[..., 1, 2] = ang[..., 2], -ang[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -ang[..., 1], ang[..., 0]
    return torch.matrix_exp(K)


def smooth(x, passes=2):
    """3x3 box blur over (H,W) of a channels-last (...,H,W,C) tensor with ERP wrap in W."""
    for _ in range(passes):
        x = (torch.roll(x, 1, -2) + x + torch.roll(x, -1, -2)) / 3
        up = torch.cat([x[..., :1, :, :], x[..., :-1, :, :]], -3)
        dn = torch.cat([x[..., 1:, :, :], x[..., -1:, :, :]], -3)
        x = (up + x + dn) / 3
    return x


# name -> dict(dataset, B,S,H,W,C,D, per_pixel, cost_type, mv, curr_idx)
CV_CASES = {
    "cv_m3d_scalar":      dict(dataset="m3d", B=2, S=2, H=16, W=32, C=8, D=6, per_pixel=False, cost_type="abs_diff"),
    "cv_m3d_volume":      dict(dataset="m3d", B=1, S=2, H=16, W=64, C=32, D=5, per_pixel=True, cost_type="abs_diff"),
    "cv_m3d_dot":         dict(dataset="m3d", B=1, S=2, H=12, W=40, C=16, D=4, per_pixel=False, cost_type="dot"),
    "cv_m3d_none":        dict(dataset="m3d", B=1, S=2, H=12, W=40, C=4, D=4, per_pixel=True, cost_type="none"),
    "cv_replica_scalar":  dict(dataset="replica_test", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=False, cost_type="abs_diff"),
    "cv_residential_vol": dict(dataset="residential", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=True, cost_type="abs_diff"),
    "cv_coffee_scalar":   dict(dataset="CoffeeArea", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=False, cost_type="abs_diff"),
    "cv_mv4_scalar":      dict(dataset="m3d", B=1, S=4, H=16, W=32, C=8, D=5, per_pixel=False, cost_type="abs_diff", mv=True, curr_idx=0),
    "cv_mv5_volume":      dict(dataset="m3d", B=2, S=5, H=8, W=32, C=32, D=3, per_pixel=True, cost_type="abs_diff", mv=True, curr_idx=1),
}


def make_cv_inputs(name, seed=0, smooth_feats=False):
    c = CV_CASES[name]
    gen = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    B, S, H, W, C, D = c["B"], c["S"], c["H"], c["W"], c["C"], c["D"]
    images = torch.randn(B, S, H, W, C, generator=gen)
    if smooth_feats:
        images = smooth(images)
    rots = small_rotations(gen, B, S, 5.0)
    trans = torch.randn(B, S, 3, generator=gen) * 0.3
    depths = torch.linspace(0.5, 10.0, D)
    depth_volume = None
    if c["per_pixel"]:
        depth_volume = torch.sort(torch.rand(B, D, H, W, generator=gen) * 9.5 + 0.5, dim=1)[0]
    args = {"dataset_name": c["dataset"], "contain_dnet": c["per_pixel"], "mono_uncertainty": False}
    return dict(args=args, images=images, depths=depths, trans=trans, rots=rots, depth_volume=depth_volume,
                cost_type=c["cost_type"], mv=c.get("mv", False), curr_idx=c.get("curr_idx", 0))


CV_BWD_CASES = ["cv_m3d_scalar", "cv_m3d_volume", "cv_m3d_dot", "cv_m3d_none", "cv_mv4_scalar", "cv_residential_vol"]


def cv_bwd_weight(name, shape):
    """Upstream gradient of the backward parity cases: d(loss)/d(out) with loss = sum(out * weight)."""
    gen = torch.Generator().manual_seed(1000 + sum(map(ord, name)))
    return torch.randn(*shape, generator=gen)


# ------------------------------------------------------------------------------------------------
# render path
# ------------------------------------------------------------------------------------------------

def render_cfg(height=32, width=64, dataset="m3d", sample_num=16, use_vis=False, use_disp=True,
               hierarchical=True, fine_use_all=False, ray_batch_num=128, **extra):
    """Config dict with every key the reference renderer reads (render.py:102-124 defaults + yaml)."""
    cfg = {
        "dataset_name": dataset, "batch_size": 1, "height": height, "width": width,
        "min_depth": 0.5, "max_depth": 15.0, "use_disp": use_disp,
        "use_wrap_padding": True, "autoencoder": False, "debug": False,
        "use_hierarchical_sampling": hierarchical, "fine_depth_use_all": fine_use_all,
        "depth_sample_num": sample_num, "fine_depth_sample_num": sample_num, "sample_num": sample_num,
        "ray_batch_num": ray_batch_num, "render_depth": True, "render_uncert": False,
        "use_polar_weighted_loss": False, "use_ray_mask": True,
        "dist_decoder_cfg": {"use_vis": use_vis}, "fine_dist_decoder_cfg": {"use_vis": use_vis},
        "agg_net_cfg": {}, "fine_agg_net_cfg": {},
        "local_feature_type": "ERP",
        # render.py:102-124 defaults read by the (out-of-scope) encoders' constructors
        "handle_distort": False, "handle_distort_all": False, "handle_distort_input_all": False,
        "with_sin": False, "wo_mono_feat": False, "uncert_tune": False, "use_depth": False,
    }
    cfg.update(extra)
    return cfg


RENDER_CASES = {
    # name: (cfg kwargs, rfn, n_rays, feature-map scales)
    "render_m3d_2src": dict(cfg=dict(), rfn=2, n_rays=96),
    "render_m3d_vis_nodisp": dict(cfg=dict(use_vis=True, use_disp=False), rfn=2, n_rays=64),
    "render_m3d_4src_all": dict(cfg=dict(fine_use_all=True, sample_num=32, hierarchical=True), rfn=4, n_rays=48,
                                fine_sample_num=16),
    "render_residential": dict(cfg=dict(dataset="residential"), rfn=2, n_rays=64),
    "render_replica": dict(cfg=dict(dataset="replica_test"), rfn=3, n_rays=64),
    "render_coffee": dict(cfg=dict(dataset="CoffeeArea", hierarchical=False), rfn=2, n_rays=64),
    # optional outputs: depth variance (renderer.py:299-301) and the coarse+fine re-compositing of render_c2f_all (:484-521)
    "render_m3d_c2f_all": dict(cfg=dict(render_uncert=True, render_c2f_all=True), rfn=2, n_rays=48),
    # one source view: ibrnet.py:359-360 masks every attention key (uniform weights); ragged sizes (37 rays, 24 samples)
    "render_m3d_1src": dict(cfg=dict(sample_num=24, use_vis=True), rfn=1, n_rays=37),
    # ablation switches of DefaultAggregationNet (aggregate_net.py:60-62, 79-81)
    "render_m3d_wo_geometry": dict(cfg=dict(wo_geometry=True), rfn=2, n_rays=40),
    "render_m3d_wo_appearance": dict(cfg=dict(wo_appearance=True), rfn=3, n_rays=40),
    # fine samples of rays with a valid depth prior from sample_3sigma (fine_render_impl, renderer.py:438-456; sample_utils.py:6-60)
    "render_m3d_ft_range": dict(cfg=dict(), rfn=2, n_rays=56, ft=True),
    "render_m3d_ft_range_all": dict(cfg=dict(fine_use_all=True, sample_num=32, hierarchical=True), rfn=2, n_rays=40,
                                    fine_sample_num=16, ft=True),
}


def make_render_inputs(name, seed=0):
    c = RENDER_CASES[name]
    kw = dict(c["cfg"])
    if "fine_sample_num" in c:
        # fine_depth_use_all: coarse dn + fine fdn must equal the fine agg net's sample_num
        kw["depth_sample_num"] = kw["sample_num"] - c["fine_sample_num"]
        kw["fine_depth_sample_num"] = c["fine_sample_num"]
    cfg = render_cfg(**{k: v for k, v in kw.items() if k not in ("depth_sample_num", "fine_depth_sample_num")})
    for k in ("depth_sample_num", "fine_depth_sample_num"):
        if k in kw:
            cfg[k] = kw[k]
    if "fine_sample_num" in c:
        # coarse net sees depth_sample_num samples, the fine net depth_sample_num + fine_depth_sample_num.
        # A top-level "sample_num" would override both (renderer.py:67-69), so it is dropped here.
        total = cfg.pop("sample_num")
        cfg["agg_net_cfg"] = {"sample_num": cfg["depth_sample_num"]}
        cfg["fine_agg_net_cfg"] = {"sample_num": total}
    gen = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    h, w, rfn = cfg["height"], cfg["width"], c["rfn"]
    imgs = smooth(torch.rand(rfn, h, w, 3, generator=gen), 1).permute(0, 3, 1, 2).contiguous()
    img_feats = torch.randn(rfn, 32, h // 2, w // 2, generator=gen)
    ray_feats = torch.randn(rfn, 32, h // 4, w // 4, generator=gen)
    rots = small_rotations(gen, 1, rfn + 1, 8.0)[0]
    trans = torch.randn(rfn + 1, 3, generator=gen) * 0.4
    w2c = torch.cat([rots, trans[:, :, None]], -1)                       # (rfn+1,3,4)
    que_w2c = w2c[-1]
    R, t = que_w2c[:, :3], que_w2c[:, 3]
    c2w = torch.cat([R.t(), (-R.t() @ t)[:, None]], -1)[None]
    perm = torch.randperm(h * w, generator=gen)[:c["n_rays"]]
    coords = torch.stack([(perm % w).float(), (perm // w).float()], -1)[None]
    que = {"coords": coords, "c2w": c2w, "w2c": que_w2c[None], "depth_range": torch.tensor([[0.5, 15.0]])}
    if c.get("ft"):
        # (1,rn,3): [validity marker (>= min_depth: valid), low, high]; a third of the rays have no prior, some intervals stick out
        # of [min_depth, max_depth] (the clamp of sample_utils.py:9 makes zero-width bins there)
        n = c["n_rays"]
        mu = 0.3 + 14.0 * torch.rand(n, generator=gen)
        half = 0.05 + 1.5 * torch.rand(n, generator=gen)
        mark = torch.where(torch.rand(n, generator=gen) < 0.33, torch.zeros(n), mu)
        que["ft_depth_range"] = torch.stack([mark, mu - half, mu + half], -1)[None]
    ref = {"imgs": imgs, "w2c": w2c[:rfn].contiguous(), "depth_range": torch.tensor([[0.5, 15.0]]).repeat(rfn, 1),
           "ray_feats": ray_feats, "img_feats": img_feats}
    return cfg, que, ref


# perspective (pinhole) query rays, renderer.render_impl(..., is_perspec=True): render_ops.py:37-74
PERSPEC_CASES = {"render_m3d_perspec": dict(base="render_m3d_2src", img_hw=(24, 32), focal=20.0, n_rays=50)}


def make_perspec_inputs(name):
    c = PERSPEC_CASES[name]
    cfg, que, ref = make_render_inputs(c["base"])
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    ph, pw = c["img_hw"]
    perm = torch.randperm(ph * pw, generator=gen)[:c["n_rays"]]
    coords = torch.stack([(perm % pw).float() + 0.25, (perm // pw).float() + 0.5], -1)[None]      # non-integer pixel centres
    K = torch.tensor([[c["focal"], 0.0, pw / 2.0], [0.0, c["focal"] * 1.1, ph / 2.0], [0.0, 0.0, 1.0]])[None]
    que = {"coords": coords, "poses": que["w2c"].clone(), "Ks": K, "depth_range": que["depth_range"]}
    return cfg, que, ref


RENDER_WEIGHT_PREFIXES = ("dist_decoder.", "fine_dist_decoder.", "agg_net.", "fine_agg_net.")


# ------------------------------------------------------------------------------------------------
# depth-prior sample placement ("diner" branch of render_impl, renderer.py:570-600)
# ------------------------------------------------------------------------------------------------

DINER_CASES = {
    # few candidates survive |mu - depth| < 0.05: most slots are filled up uniformly
    "diner_sparse": dict(cfg=dict(n_candidates=400, n_samples=16, n_gaussian=4, contain_uniform=True, n_uniform=8,
                                  inv_uniform=True, sample_num=24), rfn=2, n_rays=64, map_scale=1, radius=3.0),
    # small depth range: more surviving candidates than slots -> the top-k cut is exercised; with a real fine pass
    "diner_dense_c2f": dict(cfg=dict(n_candidates=600, n_samples=16, n_gaussian=0, contain_uniform=False, c2f=True,
                                     max_depth=4.0, sample_num=16, diner_sigma=0.03, hierarchical=True), rfn=3, n_rays=48, map_scale=2,
                            radius=2.0),
    # N_uniform + one_mlp: an extra uniform pass is merged with the depth-guided samples and re-composited (renderer.py:526-565)
    "diner_merge_uniform": dict(cfg=dict(n_candidates=500, n_samples=16, n_gaussian=2, contain_uniform=False, N_uniform=16, one_mlp=True,
                                         max_depth=6.0, sample_num=16, render_uncert=True), rfn=2, n_rays=40, map_scale=1, radius=2.5),
}


def make_diner_inputs(name, seed=0):
    """Inputs of the depth-guided branch: the usual render inputs + per-source MVS depth / variance / normal maps of
    a synthetic sphere around the origin (so candidates near the surface agree with the depth prior)."""
    from oracle import render as R
    c = DINER_CASES[name]
    kw = dict(c["cfg"])
    kw.setdefault("hierarchical", False)
    cfg = render_cfg(diner_depth_guided_sampling=True, backface_culling=True, **kw)
    gen = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    h, w, rfn = cfg["height"], cfg["width"], c["rfn"]
    imgs = smooth(torch.rand(rfn, h, w, 3, generator=gen), 1).permute(0, 3, 1, 2).contiguous()
    img_feats = torch.randn(rfn, 32, h // 2, w // 2, generator=gen)
    ray_feats = torch.randn(rfn, 32, h // 4, w // 4, generator=gen)
    rots = small_rotations(gen, 1, rfn + 1, 8.0)[0]
    trans = torch.randn(rfn + 1, 3, generator=gen) * 0.25
    w2c = torch.cat([rots, trans[:, :, None]], -1)
    que_w2c = w2c[-1]
    Rq, tq = que_w2c[:, :3], que_w2c[:, 3]
    c2w = torch.cat([Rq.t(), (-Rq.t() @ tq)[:, None]], -1)[None]
    perm = torch.randperm(h * w, generator=gen)[:c["n_rays"]]
    coords = torch.stack([(perm % w).float(), (perm // w).float()], -1)[None]
    # sphere |p| = radius seen from every source camera
    mh, mw = h // c["map_scale"], w // c["map_scale"]
    dirs = R.equi_to_unit_dirs(cfg["dataset_name"], mh, mw).reshape(-1, 3)          # camera-frame unit rays
    depth_maps, normal_maps = [], []
    for v in range(rfn):
        Rv, tv = w2c[v, :, :3], w2c[v, :, 3]
        cam = -Rv.t() @ tv
        dw = dirs @ Rv                                                              # world rays (R^T d)
        b = dw @ cam
        t = -b + torch.sqrt(b * b - cam.dot(cam) + c["radius"] ** 2)
        p = cam[None] + t[:, None] * dw
        n_cam = (p / p.norm(dim=1, keepdim=True)) @ Rv.t()                          # outward normal in the camera frame
        depth_maps.append(t.reshape(1, mh, mw))
        normal_maps.append(n_cam.reshape(mh, mw, 3).permute(2, 0, 1))
    mvs_depth = torch.stack(depth_maps) + 0.01 * smooth(torch.randn(rfn, mh, mw, 1, generator=gen), 2).permute(0, 3, 1, 2)
    flip = torch.sign(smooth(torch.randn(rfn, mh, mw, 1, generator=gen), 3).permute(0, 3, 1, 2) + 0.05)
    mvs_normal = torch.stack(normal_maps) * flip + 0.05 * torch.randn(rfn, 3, mh, mw, generator=gen)
    mvs_uncert = 0.0004 + 0.01 * torch.rand(rfn, 1, mh, mw, generator=gen)
    que = {"coords": coords, "c2w": c2w, "w2c": que_w2c[None], "depth_range": torch.tensor([[cfg["min_depth"], cfg["max_depth"]]])}
    ref = {"imgs": imgs, "w2c": w2c[:rfn].contiguous(),
           "depth_range": torch.tensor([[cfg["min_depth"], cfg["max_depth"]]]).repeat(rfn, 1),
           "ray_feats": ray_feats, "img_feats": img_feats,
           "mvs_depth": mvs_depth.contiguous(), "mvs_uncert": mvs_uncert, "mvs_normal": mvs_normal.contiguous()}
    rn = c["n_rays"]
    fill_rand = torch.rand(rn, cfg["n_samples"], generator=gen)
    gauss = torch.randn(rn, max(cfg["n_gaussian"], 1), generator=gen)[:, :cfg["n_gaussian"]]
    return cfg, que, ref, fill_rand, gauss


# ------------------------------------------------------------------------------------------------
# depth-hypothesis builder (pipeline3_model.py:537-545, 717-733, 774-821)
# ------------------------------------------------------------------------------------------------
_HYP_ARGS = {"min_depth": 0.1, "max_depth": 10.0, "mono_uncertainty": False, "mono_uncert_tune": False, "fixed_sigma": 0.5,
             "basic_sigma": 0.05, "relaxation_factor": 1.0, "wo_hdh": False, "use_depth_sampling": True, "revise_range": False,
             "fixed_dist": 1.5}
HYP_CASES = {
    # name: (args overrides, n_samples, cost_volume_channels, contain_dnet)
    "hyp_fixed_sigma_linear": ({}, 5, 64, True),
    "hyp_mono_uncert_basic_sigma": ({"mono_uncertainty": True}, 5, 64, True),
    "hyp_inverse_linear": ({"use_depth_sampling": False}, 5, 32, True),
    "hyp_revise_range": ({"revise_range": True}, 5, 24, True),
    "hyp_wo_hdh": ({"wo_hdh": True}, 5, 64, True),
    "hyp_no_mono_samples": ({}, 0, 16, True),
    "hyp_scalar_linear": ({}, 0, 48, False),
    "hyp_scalar_inverse": ({"use_depth_sampling": False}, 0, 48, False),
}


def make_hyp_inputs(name):
    over, n_samples, channels, contain_dnet = HYP_CASES[name]
    g = torch.Generator().manual_seed(900 + sorted(HYP_CASES).index(name))
    B, h, w = 2, 6, 11
    mu = 0.3 + 9.5 * torch.rand(B, 1, h, w, generator=g)          # a few pixels clamp at both ends
    sigma = 0.6 * torch.rand(B, 1, h, w, generator=g)             # some below basic_sigma
    return {"args": {**_HYP_ARGS, **over}, "n_samples": n_samples, "sampling_range": 3, "cost_volume_channels": channels,
            "contain_dnet": contain_dnet, "ref_gmms": torch.cat([mu, sigma], 1)}


# ------------------------------------------------------------------------------------------------
# depth2normal (network/orig_diner_depth2normal.py) — prior normals of the depth-guided placement (backface_culling)
# ------------------------------------------------------------------------------------------------
NORMAL_CASES = {
    # name: (dataset, h, w, fraction of zero-depth holes)
    "normal_m3d": ("m3d", 16, 32, 0.0),
    "normal_m3d_holes": ("m3d", 12, 24, 0.08),
    "normal_residential": ("residential", 10, 20, 0.03),
}


def make_normal_inputs(name):
    ds, h, w, holes = NORMAL_CASES[name]
    g = torch.Generator().manual_seed(700 + sorted(NORMAL_CASES).index(name))
    d = smooth(1.0 + 4.0 * torch.rand(2, h, w, 1, generator=g), 1).permute(0, 3, 1, 2).contiguous()
    if holes > 0:
        d = d * (torch.rand(2, 1, h, w, generator=g) > holes)
    cfg = {"dataset_name": ds, "batch_size": 1, "height": h, "width": w}
    return cfg, d


# ------------------------------------------------------------------------------------------------
# equirect -> cubemap (UniFuse util.py Equirec2Cube, pipeline3_model.py:262-283 e2c_process)
# ------------------------------------------------------------------------------------------------
E2C_CASES = {"e2c_32x64_f16": (32, 64, 16), "e2c_24x48_f9": (24, 48, 9), "e2c_64x128_f32": (64, 128, 32)}


def make_e2c_input(name):
    import numpy as np
    h, w, _ = E2C_CASES[name]
    g = torch.Generator().manual_seed(500 + sorted(E2C_CASES).index(name))
    return torch.rand(h, w, 3, generator=g).numpy().astype(np.float32)


# ------------------------------------------------------------------------------------------------
# 3-D cost regulariser `unet3d` (models/test_models.py:81-146, common_blocks.py:187-242, 366-503)
# ------------------------------------------------------------------------------------------------
UNET3D_CASES = {
    # name: (size, input shape (B, 2^(size+1), D, H, W)); D, H, W divisible by 8 (three poolings)
    "unet3d_s1": (1, (1, 4, 8, 8, 16)),
    "unet3d_s2": (2, (2, 8, 8, 16, 16)),
    "unet3d_s3": (1, (1, 4, 8, 16, 128)),      # W = 128: the row variant of the convolution kernel on the full-resolution levels
}


UNET3D_STORE_WEIGHTS = {"unet3d_s1": True, "unet3d_s2": False, "unet3d_s3": False}


def unet3d_seed(name):
    return sum(map(ord, name))


def make_unet3d_input(name):
    size, shape = UNET3D_CASES[name]
    g = torch.Generator().manual_seed(300 + sorted(UNET3D_CASES).index(name))
    return torch.rand(*shape, generator=g) * 2.0      # abs-diff costs are non-negative


# ------------------------------------------------------------------------------------------------
# 2-D heads after the regulariser: decoders1 / decoders2 (models/test_models.py:147-205, pipeline3_model.py:866-905)
# ------------------------------------------------------------------------------------------------
DEC2D_CASES = {
    # name: (size, cost_volume_channels = D, (B, H, W), out_type)
    "dec2d_s1": (1, 8, (2, 8, 16), "depth"),
    "dec2d_s2": (2, 16, (1, 8, 32), "disparity"),        # 4x upscaled width = 128: the row variant of the convolution kernel
}


def make_dec2d_inputs(name):
    """regularised cost (B, D, H, W) as unet3d returns it ((B,1,D,H,W)[:, 0]) and the mono feature map (B, 2^(size+1), H, W)"""
    size, D, (B, H, W), _ = DEC2D_CASES[name]
    g = torch.Generator().manual_seed(400 + sorted(DEC2D_CASES).index(name))
    cost_reg = torch.randn(B, 1, D, H, W, generator=g)[:, 0]
    mono = torch.randn(B, 2 ** (size + 1), H, W, generator=g)
    return cost_reg, mono


# ------------------------------------------------------------------------------------------------
# DefaultVisEncoder (network/vis_encoder.py:6-33)
# ------------------------------------------------------------------------------------------------
VISENC_CASES = {
    # name: (use_wrap_padding, n views, ray_feats (h, w), img_feats (hi, wi))
    "visenc_wrap": (True, 2, (16, 32), (16, 32)),
    "visenc_wrap_resize": (True, 1, (8, 128), (16, 256)),      # W = 128: row variant of the convolution; img_feats resized
    "visenc_zero": (False, 2, (12, 24), (24, 48)),
}


def make_visenc_inputs(name):
    _, n, (h, w), (hi, wi) = VISENC_CASES[name]
    g = torch.Generator().manual_seed(500 + sorted(VISENC_CASES).index(name))
    return torch.randn(n, 32, h, w, generator=g), torch.randn(n, 32, hi, wi, generator=g)


# CostVolumeInitNet's convolution stacks (network/init_net.py:540-574, 606-636)
INITCONV_CASES = {
    # name: (use_wrap_padding, n views, (h, w))
    "initconv_wrap": (True, 2, (16, 32)),
    "initconv_zero": (False, 1, (12, 128)),
}


def make_initconv_inputs(name):
    _, n, (h, w) = INITCONV_CASES[name]
    g = torch.Generator().manual_seed(600 + sorted(INITCONV_CASES).index(name))
    return torch.randn(n, 32, h, w, generator=g), 0.5 + 9.0 * torch.rand(n, 1, h, w, generator=g)


# image encoder ResUNetLight (network/ops.py:235-455, built as in network/renderer.py:106)
RESUNET_CASES = {
    # name: (use_wrap_padding, n views, (H, W)); H, W multiples of 16
    "resunet_wrap": (True, 2, (32, 64)),
    "resunet_zero": (False, 1, (48, 32)),
}


def make_resunet_input(name):
    _, n, (h, w) = RESUNET_CASES[name]
    g = torch.Generator().manual_seed(700 + sorted(RESUNET_CASES).index(name))
    return torch.rand(n, 3, h, w, generator=g)
