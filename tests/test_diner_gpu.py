"""GPU: depth-prior sample placement (diner branch) — fused kernel and functional drop-ins vs the oracle and the
reference's goldens."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import assert_close, load_golden  # noqa: E402
from test_oracle_render import split_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


class _Spt:
    def __init__(self, cfg):
        self.dataset, self.height, self.width = cfg["dataset_name"], cfg["height"], cfg["width"]


def _rows_equal(a, e, rtol=1e-5, atol=1e-6):
    """per-ray agreement mask for (1,rn,n) depth tables"""
    return ((a - e).abs() <= rtol * e.abs() + atol).all(-1)[0]


@pytest.mark.parametrize("name", list(cases.DINER_CASES))
def test_placement_matches_reference_golden(name):
    from panogrf_b200.render_ops import depth_guided_placement
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
    g = load_golden(name)
    gold = g["out.que_depth" if cfg.get("c2f") else "out.que_depth_fine"].float()
    z, lik = depth_guided_placement(cfg, _cuda(que), _cuda(ref), fill_rand.cuda(), gauss.cuda(), return_likelihood=True)
    assert z.shape == gold.shape
    assert bool((z[..., 1:] >= z[..., :-1]).all())
    ok = _rows_equal(z.cpu(), gold)
    # a candidate within 1 ulp of the |mu - depth| < 0.05 cut may fall on the other side: allow 1 ray in 32
    assert float(ok.float().mean()) >= 1 - 1 / 32, f"{name}: {int((~ok).sum())}/{ok.numel()} rays differ"


@pytest.mark.parametrize("name", list(cases.DINER_CASES))
def test_likelihood_and_placement_match_oracle(name):
    from oracle import depth_guided as odg, render as R
    from panogrf_b200.render_ops import depth_guided_placement
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
    rn = que["coords"].shape[1]
    cand = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg["n_candidates"], use_disp=False)
    pts, qd = R.depth2points_spherical(cfg["dataset_name"], cfg["height"], cfg["width"], que["c2w"], que["coords"], cand)
    prj = odg.project_points_dict_diner(cfg["dataset_name"], cfg["height"], cfg["width"], ref, pts, True)
    lik_o = odg.point_likelihood(cfg, ref["w2c"], prj, qd, cfg["n_candidates"])
    z, lik = depth_guided_placement(cfg, _cuda(que), _cuda(ref), fill_rand.cuda(), gauss.cuda(), return_likelihood=True)
    lik = lik.cpu()
    same_support = (lik > 0) == (lik_o > 0)
    assert float(same_support.float().mean()) > 0.9995
    both = (lik > 0) & (lik_o > 0)
    assert int(both.sum()) > 20
    assert_close(lik[both], lik_o[both], rtol=2e-3, atol=1e-6, what=f"{name}/likelihood")
    z_o = odg.diner_sample_placement(cfg, que, ref, fill_rand, gauss)
    ok = _rows_equal(z.cpu(), z_o)
    assert float(ok.float().mean()) >= 1 - 1 / 32


def test_functional_dict_variant_matches_oracle():
    """project_points_dict_diner + sample_depthguided (the reference's two-call form) against the oracle."""
    from oracle import depth_guided as odg, render as R
    from panogrf_b200.render_ops import project_points_dict_diner, sample_depthguided
    name = "diner_sparse"
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
    rn, nc = que["coords"].shape[1], cfg["n_candidates"]
    cand = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, nc, use_disp=False)
    pts, qd = R.depth2points_spherical(cfg["dataset_name"], cfg["height"], cfg["width"], que["c2w"], que["coords"], cand)
    prj_o = odg.project_points_dict_diner(cfg["dataset_name"], cfg["height"], cfg["width"], ref, pts, True)
    prj = project_points_dict_diner(_cuda(ref), pts.cuda(), _Spt(cfg), include_norm=True)
    assert set(prj) == set(prj_o)
    for k in prj_o:
        assert prj[k].shape == prj_o[k].shape, k
    assert_close(prj["depth"], prj_o["depth"], rtol=1e-5, atol=1e-6, what="depth")
    assert_close(prj["ref_mvs_depths"], prj_o["ref_mvs_depths"], rtol=1e-4, atol=1e-4, max_bad_frac=1e-3, what="mu")
    assert_close(prj["ref_mvs_normal"], prj_o["ref_mvs_normal"], rtol=1e-4, atol=2e-3, max_bad_frac=2e-3, what="normal")
    # feed the ORACLE's dict to the CUDA selection: identical inputs -> identical decisions
    z = sample_depthguided(cfg, _cuda(ref), _cuda(prj_o), cand.cuda(), qd.cuda(), cfg["n_samples"], nc, cfg["n_gaussian"],
                           include_norm=True, fill_rand=fill_rand.cuda(), gauss=gauss.cuda())
    z_o = odg.sample_depthguided(cfg, ref["w2c"], prj_o, cand, qd, cfg["n_samples"], nc, cfg["n_gaussian"], fill_rand, gauss)
    ok = _rows_equal(z.cpu(), z_o, rtol=2e-5)
    assert bool(ok.all()), f"{int((~ok).sum())} rays differ"


@pytest.mark.parametrize("name", list(cases.DINER_CASES))
@pytest.mark.parametrize("mlp_dtype", ["fp32", "bf16"])
def test_renderer_diner_branch_matches_reference_golden(name, mlp_dtype):
    from panogrf_b200.renderer import NeuralRayBaseRenderer
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
    g = load_golden(name)
    _, _, W, gold = split_golden(g)
    net = NeuralRayBaseRenderer({**cfg, "mlp_dtype": mlp_dtype}).cuda()
    net.load_state_dict(W, strict=True)
    q = _cuda(que)
    q["diner_fill_rand"], q["diner_gauss"] = fill_rand.cuda(), gauss.cuda()
    out = net.render_impl(q, _cuda(ref), False)
    assert set(gold) <= set(out), set(gold) - set(out)
    dkey = "que_depth" if cfg.get("c2f") else "que_depth_fine"
    ok = _rows_equal(out[dkey].cpu(), gold[dkey].float())
    assert float(ok.float().mean()) >= 1 - 1 / 32
    tol = dict(rtol=1e-4, atol=1e-4) if mlp_dtype == "fp32" else dict(rtol=1e-2, atol=3e-2)
    for k in ("pixel_colors_nr", "render_depth", "pixel_colors_nr_fine", "render_depth_fine"):
        if k in gold:
            a, e = out[k].cpu()[0][ok], gold[k].float()[0][ok]
            scale = float(e.abs().max())
            assert_close(a, e, rtol=tol["rtol"], atol=tol["atol"] * max(scale, 1.0), max_bad_frac=0.02 if mlp_dtype == "bf16" else 0.0,
                         what=f"{name}/{mlp_dtype}/{k}")


def test_diner_random_tables_default_and_validation():
    from panogrf_b200 import _lib
    from panogrf_b200.render_ops import depth_guided_placement
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs("diner_sparse")
    gen = torch.Generator(device="cuda").manual_seed(3)
    z1 = depth_guided_placement(cfg, _cuda(que), _cuda(ref), generator=gen)
    gen = torch.Generator(device="cuda").manual_seed(3)
    z2 = depth_guided_placement(cfg, _cuda(que), _cuda(ref), generator=gen)
    assert torch.equal(z1, z2) and bool(torch.isfinite(z1).all())
    bad = dict(cfg, n_gaussian=cfg["n_samples"] + 1)
    with pytest.raises(_lib.PanoGRFError):
        depth_guided_placement(bad, _cuda(que), _cuda(ref))
    with pytest.raises(_lib.PanoGRFError):
        depth_guided_placement(cfg, que, ref)          # CPU tensors: no fallback


@pytest.mark.parametrize("name", list(cases.NORMAL_CASES))
def test_depth2normal_matches_reference(name):
    """f3: depth2normal (network/orig_diner_depth2normal.py) on the GPU == the reference's output: same NaN pattern (degenerate
    cross products the reference leaves as NaN), values within 1e-5 (unit vectors; the pixel -> ray trig differs by 1 ulp)."""
    import types
    from panogrf_b200 import render_ops as rops
    ds, h, w, _ = cases.NORMAL_CASES[name]
    g = load_golden(name)
    got = rops.depth2normal({"mvs_depth": g["mvs_depth"].cuda()}, types.SimpleNamespace(dataset=ds, height=h, width=w)).cpu()
    want = g["normal"]
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert float((torch.nan_to_num(got) - torch.nan_to_num(want)).abs().max()) < 1e-5


@pytest.mark.parametrize("kind,dataset", [("smooth", "m3d"), ("edges", "m3d"), ("noise", "m3d"), ("edges", "replica_test"),
                                          ("edges", "residential"), ("noise", "CoffeeArea"), ("smooth", "CoffeeArea")])
def test_prefilter_is_exact(kind, dataset):
    """The conservative pre-filter of phase 1 (approximate projection, csrc/depth_guided.cu:dg_certainly_far) must not change a
    single bit: likelihoods and placed samples with the filter == without it, on priors with smooth surfaces, depth edges and noise."""
    from panogrf_b200 import _lib
    from panogrf_b200.render_ops import depth_guided_placement
    lib = _lib.load()
    H, W, rfn, rn = 64, 128, 3, 4096
    cfg = {"dataset_name": dataset, "height": H, "width": W, "min_depth": 0.5, "max_depth": 15.0, "n_candidates": 1000,
           "n_samples": 64, "n_gaussian": 15, "backface_culling": True, "contain_uniform": False}
    g = torch.Generator().manual_seed(3)
    idx = torch.randperm(H * W, generator=g)[:rn]
    coords = torch.stack([idx % W, idx // W], -1).float()[None].cuda()
    w2c = torch.eye(3, 4)[None].repeat(rfn, 1, 1)
    w2c[0, 2, 3], w2c[1, 2, 3], w2c[2, 0, 3] = 0.5, -0.5, 0.3
    w2c[2, 1, 3] = 0.2                                   # every axis carries an offset in some view (the conventions permute the axes)
    yy = torch.linspace(0, 3.14159, H)[:, None]
    xx = torch.linspace(0, 6.28318, W)[None, :]
    depth = 3.0 + 1.5 * torch.sin(xx * 2) * torch.sin(yy)
    if kind == "edges":
        depth = depth + (xx > 3.0).float() * 2.0 + (yy > 1.2).float() * 0.7
    if kind == "noise":
        depth = depth + torch.rand(H, W, generator=g)
    for map_hw in ((H, W), (H // 2, W // 2)):          # align_corners and half-resolution priors
        d = torch.nn.functional.interpolate(depth[None, None], size=map_hw, mode="bilinear")[0, 0]
        ref = {"imgs": torch.zeros(rfn, 3, H, W).cuda(), "w2c": w2c.cuda(), "mvs_depth": d[None, None].repeat(rfn, 1, 1, 1).cuda(),
               "mvs_uncert": torch.full((rfn, 1, *map_hw), 0.02).cuda(), "mvs_normal": torch.randn(rfn, 3, *map_hw, generator=g).cuda()}
        que = {"coords": coords, "c2w": torch.eye(3, 4)[None].cuda()}
        fill, ga = torch.rand(rn, 64, generator=g).cuda(), torch.randn(rn, 15, generator=g).cuda()
        res = []
        try:
            for pre in (0, 1):
                lib.pgrf_debug_set(b"dg_prefilter", pre)
                res.append(depth_guided_placement(cfg, que, ref, fill, ga, return_likelihood=True))
        finally:
            lib.pgrf_debug_set(b"dg_prefilter", 1)
        assert int((res[0][1] > 0).sum()) > rn          # the scene does hit the prior surfaces
        assert torch.equal(res[0][1], res[1][1]), f"{kind} {map_hw}: likelihood changed"
        assert torch.equal(res[0][0], res[1][0]), f"{kind} {map_hw}: placement changed"


def test_diner_c2f_with_ft_depth_range_matches_oracle():
    """Depth-prior branch with a real fine pass (c2f) AND per-ray priors for the fine samples (que_imgs_info['ft_depth_range'],
    fine_render_impl renderer.py:438-456): against the oracle (its parts are pinned by the reference goldens `diner_dense_c2f` and
    `render_m3d_ft_range`)."""
    from oracle import depth_guided as odg
    from panogrf_b200.renderer import NeuralRayBaseRenderer
    name = "diner_dense_c2f"
    cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
    _, _, W, _ = split_golden(load_golden(name))
    rn = que["coords"].shape[1]
    gen = torch.Generator().manual_seed(5)
    mu = 0.6 + 3.0 * torch.rand(rn, generator=gen)
    half = 0.05 + 0.5 * torch.rand(rn, generator=gen)
    mark = torch.where(torch.rand(rn, generator=gen) < 0.4, torch.zeros(rn), mu)
    que = dict(que, ft_depth_range=torch.stack([mark, mu - half, mu + half], -1)[None])
    exp = odg.render_rays_diner(cfg, W, que, ref, fill_rand, gauss)
    net = NeuralRayBaseRenderer({**cfg, "mlp_dtype": "fp32"}).cuda()
    net.load_state_dict(W, strict=True)
    q = _cuda(que)
    q["diner_fill_rand"], q["diner_gauss"] = fill_rand.cuda(), gauss.cuda()
    out = net.render_impl(q, _cuda(ref), False)
    ok = _rows_equal(out["que_depth"].cpu(), exp["que_depth"]) & _rows_equal(out["que_depth_fine"].cpu(), exp["que_depth_fine"], atol=1e-4)
    assert float(ok.float().mean()) >= 1 - 2 / 32
    valid = (que["ft_depth_range"][0, :, 0] >= cfg["min_depth"])
    assert int(valid.sum()) > 5 and int((~valid).sum()) > 5
    for k in ("pixel_colors_nr_fine", "render_depth_fine"):
        a, e = out[k].cpu()[0][ok], exp[k][0][ok]
        assert_close(a, e, rtol=1e-4, atol=1e-4 * max(float(e.abs().max()), 1.0), what=f"diner c2f + ft/{k}")
