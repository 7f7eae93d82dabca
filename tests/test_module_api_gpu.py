"""GPU: the reference's MODULE-level API (dist decoder, aggregation net, network_rendering, predict_proj_ray_prob, get_img_feats,
interpolate_feature_map, depth2points_spherical) evaluated by the CUDA kernels on a caller-provided prj_dict, vs the oracle."""
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import assert_close, load_golden  # noqa: E402
from test_oracle_render import split_golden  # noqa: E402
from test_render_gpu import build_renderer, cuda_dict  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = ["render_m3d_vis_nodisp", "render_replica", "render_m3d_2src"]


def _oracle_pass(name):
    from oracle import render as R
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, _ = split_golden(load_golden(name))
    rn = que["coords"].shape[1]
    depth = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg["depth_sample_num"], cfg["use_disp"])
    out = R.render_by_depth(cfg, W, que, ref, depth, False, return_prj=True)
    return cfg, que, ref, W, depth, out


@pytest.mark.parametrize("name", CASES)
def test_functional_geometry_ops(name):
    from oracle import render as R
    from panogrf_b200 import render_ops as rops
    cfg, que, ref, W, depth, out = _oracle_pass(name)
    spt = types.SimpleNamespace(dataset=cfg["dataset_name"], height=cfg["height"], width=cfg["width"])
    pts_o, dir_o = R.depth2points_spherical(cfg["dataset_name"], cfg["height"], cfg["width"], que["c2w"], que["coords"], depth)
    pts, dirs = rops.depth2points_spherical(cuda_dict(que), depth.cuda(), spt)
    assert_close(pts, pts_o, rtol=1e-5, atol=1e-5, what="que_pts")
    assert_close(dirs, dir_o, rtol=1e-5, atol=1e-6, what="que_dir")
    rfn, _, h, w = ref["imgs"].shape
    pix = out["prj"]["pts"].reshape(rfn, -1, 2)
    for key in ("ray_feats", "img_feats", "imgs"):
        got = rops.interpolate_feature_map(ref[key].cuda(), pix.cuda(), h, w)
        assert_close(got, R.bilinear_border(ref[key], pix, h, w), rtol=1e-4, atol=1e-4, what=key)


@pytest.mark.parametrize("name", CASES)
def test_dist_decoder_module(name):
    from oracle import render as R
    cfg, que, ref, W, depth, out = _oracle_pass(name)
    net = build_renderer(cfg, W)
    prj = out["prj"]
    use_vis = cfg["dist_decoder_cfg"].get("use_vis", True)
    mean_o, var_o, vis_o, aw_o = R.dist_decoder_forward(W, "dist_decoder", prj["ray_feats"], use_vis)
    mean, var, vis, aw = net.dist_decoder(prj["ray_feats"].cuda())
    assert_close(mean, mean_o, rtol=1e-4, atol=1e-5, what="mean")
    assert_close(var, var_o, rtol=1e-4, atol=1e-5, what="var")
    assert_close(aw, aw_o, rtol=1e-4, atol=1e-5, what="aw")
    assert (vis is None) == (vis_o is None)
    if vis is not None:
        assert_close(vis, vis_o, rtol=1e-4, atol=1e-5, what="vis")
    # compute_prob on the ORACLE's decoder outputs: same inputs on both sides
    c = lambda t: None if t is None else t.cuda()
    a, v, h = net.dist_decoder.compute_prob(prj["depth"].squeeze(-1).cuda(), out["dists"].unsqueeze(0).cuda(), c(mean_o), c(var_o),
                                            c(vis_o), c(aw_o), True, ref["depth_range"].cuda())
    a_o, v_o, h_o = R.compute_prob_ref(prj["depth"].squeeze(-1), out["dists"].unsqueeze(0), mean_o, var_o, vis_o, aw_o,
                                       ref["depth_range"], use_vis)
    assert_close(v, v_o, rtol=1e-4, atol=1e-5, what="visibility")
    assert_close(h, h_o, rtol=1e-4, atol=1e-5, what="hit_prob")
    assert_close(a, a_o, rtol=1e-4, atol=2e-3, what="alpha")       # log of a ratio of differences: ill-conditioned near 0


@pytest.mark.parametrize("tag", ["a", "b"])
def test_compute_prob_is_ref_false_matches_reference_golden(tag):
    """the query rays' own hit probability (is_ref=False, training losses): golden produced by the reference class"""
    import panogrf_b200 as pg
    g = load_golden("prob_que")
    c = {k[2:]: v for k, v in g.items() if k.startswith(tag + ".")}
    cfg, _, _ = cases.make_render_inputs("render_m3d_2src")
    uv = bool(c["use_vis"])
    net = pg.NeuralRayBaseRenderer({**cfg, "dist_decoder_cfg": {"use_vis": uv}, "fine_dist_decoder_cfg": {"use_vis": uv}}).cuda()
    d = lambda k: c[k].cuda()
    a, v, h = net.dist_decoder.compute_prob(d("depth"), d("interval"), d("mean"), d("var"), d("vis"), d("aw"), False, d("depth_range"))
    assert_close(v, c["visibility"], rtol=1e-4, atol=1e-5, what="visibility")
    assert_close(h, c["hit_prob"], rtol=1e-4, atol=1e-5, what="hit_prob")
    assert_close(a, c["alpha"], rtol=1e-4, atol=2e-3, what="alpha")


@pytest.mark.parametrize("name", CASES)
def test_predict_prob_agg_net_and_network_rendering(name):
    cfg, que, ref, W, depth, out = _oracle_pass(name)
    net = build_renderer(cfg, W)
    prj_o = out["prj"]
    # 1. predict_proj_ray_prob + get_img_feats from the geometric part of the dict
    prj = {k: prj_o[k].cuda() for k in ("pts", "depth", "dir", "ray_feats", "rgb")}
    prj = net.predict_proj_ray_prob(prj, cuda_dict(ref), out["dists"].cuda(), False)
    prj = net.get_img_feats(cuda_dict(ref), prj)
    assert_close(prj["vis"], prj_o["vis"], rtol=1e-4, atol=1e-5, what="vis")
    assert_close(prj["hit_prob"], prj_o["hit_prob"], rtol=1e-4, atol=1e-5, what="hit_prob")
    assert_close(prj["alpha"], prj_o["alpha"], rtol=1e-4, atol=2e-3, what="alpha")
    assert_close(prj["img_feats"], prj_o["img_feats"], rtol=1e-4, atol=1e-4, what="img_feats")
    # 2. the aggregation network on the ORACLE's dict (identical inputs)
    prj_in = {k: v.cuda() for k, v in prj_o.items()}
    density, colors = net.agg_net(prj_in, out["que_dir"].cuda())
    assert density.shape == out["density_nr"].shape and colors.shape == out["colors_nr"].shape
    assert_close(density, out["density_nr"], rtol=1e-4, atol=1e-4, what="density")
    assert_close(colors, out["colors_nr"], rtol=1e-4, atol=1e-4, what="colors")
    # 3. network_rendering
    hit, col, pix, den = net.network_rendering(prj_in, out["que_dir"].cuda(), False)
    assert_close(hit, out["hit_prob_nr"], rtol=1e-4, atol=1e-5, what="hit_prob_nr")
    assert_close(pix, out["pixel_colors_nr"], rtol=1e-4, atol=1e-4, what="pixel_colors_nr")
    assert_close(den, out["density_nr"], rtol=1e-4, atol=1e-4, what="density_nr")
    # the positional table is tied to sample_num like in the reference (ibrnet.py:358)
    bad = {k: v[:, :, :, :-1] for k, v in prj_in.items()}
    with pytest.raises(RuntimeError):
        net.agg_net(bad, out["que_dir"][:, :, :-1].cuda())


def test_gen_renderer_forward_and_registry():
    """NeuralRayGenRenderer.forward(data) (renderer.py:777-786): init_net hook, render, depth-mean outputs in eval."""
    import panogrf_b200 as pg
    from oracle import render as R
    name = "render_m3d_2src"
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, gold = split_golden(load_golden(name))
    assert pg.name2network["neuray_gen"] is pg.NeuralRayGenRenderer
    net = pg.NeuralRayGenRenderer({**cfg, "depth_loss_coords_num": 256}).cuda().eval()
    net.load_state_dict(W, strict=True)
    ref_c = cuda_dict(ref)
    ray_feats = ref_c.pop("ray_feats")
    calls = []

    def init_net(ref_imgs_info, src_imgs_info, is_train):          # stands in for the (out-of-scope) CNN init net
        calls.append((src_imgs_info is not None, is_train))
        return {"ray_feats": ray_feats, "mvs_depth": torch.ones(2, 1, 8, 16, device="cuda")}

    with pytest.raises(pg._lib.PanoGRFError):
        net({"que_imgs_info": cuda_dict(que), "ref_imgs_info": dict(ref_c), "eval": True})
    net.init_net = init_net
    out = net({"que_imgs_info": cuda_dict(que), "ref_imgs_info": dict(ref_c), "src_imgs_info": {}, "eval": True})
    assert calls == [(True, False)]
    assert_close(out["pixel_colors_nr_fine"], gold["pixel_colors_nr_fine"].float(), rtol=1e-4, atol=5e-5, what="gen/pixel_colors_nr_fine")
    assert out["depth_mean"].shape == (2, 256) and out["depth_coords"].shape == (2, 256, 2)
    # depth_mean == first mixture mean of the coarse decoder at those source pixels
    h, w = ref["imgs"].shape[-2:]
    feats = R.bilinear_border(ref["ray_feats"], out["depth_coords"].float().cpu(), h, w)
    mean_o = R.dist_decoder_forward(W, "dist_decoder", feats, cfg["dist_decoder_cfg"].get("use_vis", True))[0]
    assert_close(out["depth_mean"], mean_o[..., 0], rtol=1e-4, atol=1e-5, what="gen/depth_mean")
    assert_close(out["depth_mean_fine_2"],
                 R.dist_decoder_forward(W, "fine_dist_decoder", feats, cfg["fine_dist_decoder_cfg"].get("use_vis", True))[0][..., 1],
                 rtol=1e-4, atol=1e-5, what="gen/depth_mean_fine_2")


def test_render_emits_pixel_colors_gt_like_the_reference():
    """renderer.py:278-286 / 398-405: whenever the query carries its image, the output dict has `pixel_colors_gt[_fine]`
    (interpolate_feats(imgs, coords, align_corners=True)); network/metrics.py reads it.  Also: maps created under
    torch.inference_mode (no version counter) must not crash the map cache."""
    import torch.nn.functional as F
    import panogrf_b200 as pg
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import cases
    cfg, que, ref = cases.make_render_inputs("render_m3d_2src")
    net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
    gen = torch.Generator().manual_seed(3)
    h, w = int(cfg["height"]), int(cfg["width"])
    qimg = torch.rand(1, 3, h, w, generator=gen)
    que = {**{k: v.cuda() for k, v in que.items()}, "imgs": qimg.cuda()}
    with torch.inference_mode():
        ref_c = {k: (v.cuda() * 1.0) for k, v in ref.items()}        # inference tensors: `_version` raises on them
    out = net.render(que, ref_c, False)
    c = que["coords"].cpu()
    grid = torch.stack([c[..., 0] / (w - 1) * 2 - 1, c[..., 1] / (h - 1) * 2 - 1], -1)[:, None]
    want = F.grid_sample(qimg, grid, mode="bilinear", padding_mode="border", align_corners=True)[:, :, 0].permute(0, 2, 1)
    for k in ("pixel_colors_gt", "pixel_colors_gt_fine"):
        assert k in out and float((out[k].cpu() - want).abs().max()) < 1e-5
    assert bool(torch.isfinite(out["pixel_colors_nr_fine"]).all())


def test_pose_loop_driver_renders_views():
    """f4: the reference's render script loop (render.py:249-291) around the fused renderer: build_render_imgs_info per pose,
    `renderer(data)`, uint8 image + normalised depth back; the first view equals a direct render() of the same pose."""
    import numpy as np
    import panogrf_b200 as pg
    from panogrf_b200 import driver
    cfg, que, ref = cases.make_render_inputs("render_m3d_2src")
    h, w = int(cfg["height"]), int(cfg["width"])
    net = pg.NeuralRayBaseRenderer(cfg).cuda().eval()
    poses = [que["w2c"][0].numpy().astype(np.float64), np.concatenate([np.eye(3), [[0.1], [0.0], [0.2]]], 1)]
    imgs, depths = driver.render_poses(net, ref, poses, [(h, w)] * 2, [(0.5, 15.0)] * 2)
    assert len(imgs) == 2 and imgs[0].shape == (h, w, 3) and imgs[0].dtype == np.uint8 and depths[1].shape == (h, w)
    q0 = driver.to_cuda(driver.imgs_info_to_torch(driver.build_render_imgs_info(poses[0], (h, w), (0.5, 15.0))))
    direct = net.render(q0, {k: v.cuda() for k, v in ref.items()}, False)["pixel_colors_nr_fine"].cpu().numpy().reshape(h, w, 3)
    assert np.array_equal(imgs[0], driver.color_map_backward(direct))
    m = driver.WSPSNR()
    a = torch.from_numpy(imgs[0].astype(np.float32) / 255)[None]
    assert float(m.ws_psnr(a, a + 0.01)[0]) > 39.0
