"""GPU parity of the stand-alone operators (reference functional API, network/render_ops.py) against the oracle."""
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from test_oracle_render import split_golden  # noqa: E402
from util import assert_close, load_golden  # noqa: E402

from oracle import cost_volume as ocv  # noqa: E402
from oracle import render as orender  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["render_m3d_2src", "render_replica", "render_residential", "render_coffee", "render_m3d_4src_all"])
def test_project_points_dict(name):
    from panogrf_b200 import render_ops as rops
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, _ = split_golden(load_golden(name))
    rn, dn = que["coords"].shape[1], cfg["depth_sample_num"]
    depth = orender.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, dn, cfg["use_disp"])
    pts, _ = orender.depth2points_spherical(cfg["dataset_name"], cfg["height"], cfg["width"], que["c2w"], que["coords"], depth)
    o = orender.render_by_depth(cfg, W, que, ref, depth, False, return_prj=True)["prj"]
    spt = types.SimpleNamespace(dataset=cfg["dataset_name"], height=cfg["height"], width=cfg["width"])
    refc = {k: v.cuda() for k, v in ref.items()}
    got = rops.project_points_dict(refc, pts.cuda(), spt)
    torch.cuda.synchronize()
    tol = {"pts": 2e-3, "depth": 1e-5, "dir": 1e-5, "ray_feats": 2e-4, "rgb": 2e-5, "img_feats": 2e-4}
    for k, atol in tol.items():
        assert got[k].shape == o[k].shape, k
        assert_close(got[k], o[k], rtol=1e-4, atol=atol, what=f"{name}/{k}")


def test_alpha_compositing_and_fine_sampling_bit_exact_bins():
    from panogrf_b200 import render_ops as rops
    gen = torch.Generator().manual_seed(7)
    rn, dn = 333, 64
    alpha = torch.rand(1, rn, dn, generator=gen) ** 3
    hit = rops.alpha_values2hit_prob(alpha.cuda()).cpu()
    assert torch.equal(hit, orender.alpha_values2hit_prob(alpha))          # same sequential fp32 order -> identical bits
    density = torch.randn(1, rn, dn, generator=gen)
    colors = torch.rand(1, rn, dn, 3, generator=gen)
    depth = orender.sample_depth(0.5, 15.0, rn, dn, True)
    h, pix, rd = rops.composite(density.cuda(), colors.cuda(), depth.cuda())
    eh, ep, ed = orender.composite(density, colors, depth)
    assert_close(h, eh, rtol=1e-5, atol=1e-7, what="hit")
    assert_close(pix, ep, rtol=1e-4, atol=1e-6, what="pixel")
    assert_close(rd, ed, rtol=1e-4, atol=1e-5, what="depth")
    args = {"use_disp": True}
    rng = torch.tensor([[0.5, 15.0]])
    hit2 = eh.clone()
    hit2[:, :10] = 0                                                        # uniform pdf rows
    hit2[:, 10:20, 5] = 40.0                                                # dominant bin (denom < 1e-5 branch)
    fine, inds = rops.sample_fine_depth(args, depth.cuda(), hit2.cuda(), rng, 64, False, return_indices=True)
    efine, einds = orender.sample_fine_depth(depth, hit2, rng, 64, True, return_indices=True)
    assert torch.equal(inds.cpu(), einds), "searchsorted bin indices differ"
    assert_close(fine, efine, rtol=1e-5, atol=0.0, what="fine depths")
    fine_lin = rops.sample_fine_depth({"use_disp": False}, depth.cuda(), hit2.cuda(), rng, 32, False)
    assert_close(fine_lin, orender.sample_fine_depth(depth, hit2, rng, 32, False), rtol=1e-5, atol=0.0, what="fine/no-disp")


def test_sample_depth_and_inv_dists():
    from panogrf_b200 import render_ops as rops
    args = {"min_depth": 0.5, "max_depth": 15.0}
    coords = torch.zeros(1, 7, 2, device="cuda")
    for use_disp in (True, False):
        d, dist = rops.sample_depth(args, coords, 64, False, use_disp)
        assert torch.equal(d.cpu(), orender.sample_depth(0.5, 15.0, 7, 64, use_disp))
        assert float(dist[0, 0, -1]) > 9e5
    rng = torch.tensor([[0.5, 15.0]])
    inv = rops.depth2inv_dists(d, rng.cuda())
    assert_close(inv, orender.depth2inv_dists(d.cpu(), rng), rtol=1e-6, atol=1e-7, what="inv dists")


def test_depth_hypotheses_bit_exact():
    from panogrf_b200 import render_ops as rops
    gen = torch.Generator().manual_seed(3)
    mu = torch.rand(2, 1, 24, 40, generator=gen) * 11 - 0.5        # also outside [min,max] -> clamps
    ks = ocv.magnet_k_list(5, 3)
    got = rops.mono_guided_hypotheses(mu.cuda(), ks, 0.5, 0.1, 10.0, 59).cpu()
    expect = ocv.mono_guided_hypotheses(mu, ks, 0.5, 0.1, 10.0, 59)
    assert got.shape == (2, 64, 24, 40)
    assert torch.equal(got, expect), "hypothesis values / order differ"


# ---- backward passes of the cheap differentiable stages -----------------------------------------------------

def test_composite_backward_matches_torch_autograd():
    """d/d(density), d/d(colors) of (hit_prob, pixel_colors, render_depth) vs torch autograd on the reference formulas in fp64."""
    from panogrf_b200 import render_ops as rops
    gen = torch.Generator().manual_seed(3)
    qn, rn, dn = 1, 70, 48
    density = (torch.randn(qn, rn, dn, generator=gen) * 2).requires_grad_(True)        # ~half negative: relu branch
    colors = torch.rand(qn, rn, dn, 3, generator=gen).requires_grad_(True)
    depth = torch.sort(torch.rand(qn, rn, dn, generator=gen) * 10 + 0.5, -1)[0]
    wh, wp, wd = torch.randn(qn, rn, dn, generator=gen), torch.randn(qn, rn, 3, generator=gen), torch.randn(qn, rn, generator=gen)

    def ref(d, c, z):
        alpha = 1.0 - torch.exp(-torch.relu(d))
        T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., :1]), 1 - alpha + 1e-10], -1), -1)[..., :-1]
        hit = alpha * T
        return hit, (hit[..., None] * c).sum(2), (hit * z).sum(-1)

    d64, c64 = density.detach().double().requires_grad_(True), colors.detach().double().requires_grad_(True)
    h, p, r = ref(d64, c64, depth.double())
    ((h * wh).sum() + (p * wp).sum() + (r * wd).sum()).backward()
    dc, cc = density.detach().cuda().requires_grad_(True), colors.detach().cuda().requires_grad_(True)
    h2, p2, r2 = rops.composite(dc, cc, depth.cuda())
    ((h2 * wh.cuda()).sum() + (p2 * wp.cuda()).sum() + (r2 * wd.cuda()).sum()).backward()
    assert_close(h2, h.float(), rtol=1e-5, atol=1e-6, what="hit")
    assert_close(dc.grad, d64.grad.float(), rtol=1e-4, atol=1e-5, what="grad density")
    assert_close(cc.grad, c64.grad.float(), rtol=1e-4, atol=1e-6, what="grad colors")
    # alpha-input form (alpha_values2hit_prob)
    a64 = torch.rand(5, 33, generator=gen).double().requires_grad_(True)
    T = torch.cumprod(torch.cat([torch.ones_like(a64[..., :1]), 1 - a64 + 1e-10], -1), -1)[..., :-1]
    w = torch.randn(5, 33, generator=gen)
    ((a64 * T) * w).sum().backward()
    ac = a64.detach().float().cuda().requires_grad_(True)
    (rops.alpha_values2hit_prob(ac) * w.cuda()).sum().backward()
    assert_close(ac.grad, a64.grad.float(), rtol=1e-4, atol=1e-5, what="grad alpha")


@pytest.mark.parametrize("scale", [1, 4])
def test_interpolate_feature_map_backward_matches_torch_autograd(scale):
    """d/d(feats) of the bilinear border gather vs torch autograd through F.grid_sample (the reference's own op)."""
    import torch.nn.functional as F
    from panogrf_b200 import render_ops as rops
    gen = torch.Generator().manual_seed(5 + scale)
    rfn, f, h, w, pn = 2, 7, 32, 64, 500
    feats = torch.randn(rfn, f, h // scale, w // scale, generator=gen)
    pix = torch.stack([torch.rand(rfn, pn, generator=gen) * (w + 3) - 2, torch.rand(rfn, pn, generator=gen) * (h + 3) - 2], -1)
    wgt = torch.randn(rfn, pn, f, generator=gen)
    f64 = feats.double().requires_grad_(True)
    grid = torch.stack([pix[..., 0] / (w - 1) * 2 - 1, pix[..., 1] / (h - 1) * 2 - 1], -1).unsqueeze(1).double()
    ref = F.grid_sample(f64, grid, mode="bilinear", padding_mode="border", align_corners=(scale == 1)).squeeze(2).permute(0, 2, 1)
    (ref * wgt.double()).sum().backward()
    fc = feats.cuda().requires_grad_(True)
    out = rops.interpolate_feature_map(fc, pix.cuda(), h, w)
    (out * wgt.cuda()).sum().backward()
    assert_close(out, ref.float(), rtol=1e-4, atol=1e-5, what="fwd")
    assert_close(fc.grad, f64.grad.float(), rtol=1e-4, atol=1e-4, what="grad feats")


@pytest.mark.parametrize("name", list(cases.HYP_CASES))
def test_depth_hypotheses_every_variant_bit_exact(name):
    """a5 with every switch of the reference (fixed_sigma / mono_uncertainty + basic_sigma, linear / inverse-linear / revise_range
    centres, wo_hdh, n_samples = 0, scalar hypotheses): the CUDA builder equals the outputs of the reference's own source lines
    (tests/golden/make_golden_hypotheses.py) bit for bit — bin order of the sweep depends on it."""
    from panogrf_b200 import render_ops as rops
    case = cases.make_hyp_inputs(name)
    g = load_golden(name)
    k_list = rops.magnet_k_list(case["n_samples"], case["sampling_range"]) if case["n_samples"] > 0 else []
    assert all(abs(a - float(b)) < 1e-13 for a, b in zip(k_list, g["k_list"].double()))
    vol, cen = rops.depth_hypotheses(case["args"], g["ref_gmms"].cuda(), k_list, case["cost_volume_channels"], case["contain_dnet"])
    if "depth_volume" in g:
        assert torch.equal(vol.cpu(), g["depth_volume"]), float((vol.cpu() - g["depth_volume"]).abs().max())
    else:
        assert vol is None and torch.equal(cen.cpu(), g["d_centers"])


@pytest.mark.parametrize("name", list(cases.E2C_CASES))
def test_e2c_matches_reference_class(name):
    """f2 (first slice): equirect -> cubemap on the GPU == the reference's Equirec2Cube (scipy map_coordinates on the CPU), for a
    batch of (B, S) images in one launch; coordinate tables identical to the reference's."""
    import numpy as np
    from panogrf_b200 import e2c
    h, w, f = cases.E2C_CASES[name]
    g = load_golden(name)
    inst = e2c.Equirec2Cube(h, w, f)
    equ = g["equ"]
    batch = torch.stack([equ, equ.flip(1), equ * 0.5, equ + 1.0]).reshape(2, 2, h, w, 3).cuda()
    got = e2c.e2c_process(batch, inst).cpu()
    assert got.shape == (2, 2, f, 6 * f, 3)
    assert float((got[0, 0] - g["cube"]).abs().max()) <= 1.2e-7
    assert float((got[1, 0] - 0.5 * g["cube"]).abs().max()) <= 1.2e-7            # linear in the image
    from scipy.ndimage import map_coordinates                                     # the second image against scipy directly
    e = equ.flip(1).numpy()
    for c in range(3):
        ch = e[..., c]
        pad = np.concatenate([ch, np.roll(ch[[-1]], w // 2, 1), np.roll(ch[[0]], w // 2, 1)], 0)
        want = map_coordinates(pad, [inst.coor_y, inst.coor_x], order=1, mode="wrap")[..., 0]
        assert np.abs(got[0, 1, :, :, c].numpy() - want).max() <= 1.2e-7


def test_sample_3sigma_matches_reference_golden():
    """render_ops.sample_3sigma (pgrf_sample_3sigma_fwd) against the reference function's outputs and the oracle."""
    from oracle import render as R
    from panogrf_b200.render_ops import sample_3sigma
    g = load_golden("sample_3sigma")
    for tag in "abc":
        n, near, far = int(g[tag + ".n"]), float(g[tag + ".near"]), float(g[tag + ".far"])
        low, high = torch.as_tensor(g[tag + ".low"]), torch.as_tensor(g[tag + ".high"])
        z = sample_3sigma(low.cuda(), high.cuda(), n, True, near, far).cpu()
        assert_close(z, torch.as_tensor(g[tag + ".z"]), rtol=1e-5, atol=1e-4, what=f"sample_3sigma/{tag} vs reference")
        assert_close(z, R.sample_3sigma(low, high, n, near, far), rtol=1e-5, atol=2e-5, what=f"sample_3sigma/{tag} vs oracle")
    with pytest.raises(NotImplementedError):
        sample_3sigma(low.cuda(), high.cuda(), 8, False, near, far)
