"""CPU restatement of the depth-prior sample placement of the render path.  TEST INFRASTRUCTURE ONLY:
imported by tests/, never by panogrf_b200/.

Follows (reference file:line):
  project_points_dict_diner   network/render_ops.py:260-290
  sample_depthguided          network/original_depth_guided_sample.py:45-297
  fill_up_uniform_samples     network/original_depth_guided_sample.py:333-366
  diner_render_by_depth       network/renderer.py:318-436   (sample placement part, :318-355)
  render_impl (diner branch)  network/renderer.py:570-600

The reference draws random numbers in two places (uniform jitter of filled-up slots, Gaussian samples
around the occlusion-aware mean).  Here — and in the CUDA kernel — they are explicit inputs:
  fill_rand (rn, n_samples)   U[0,1) value used if slot (ray, slot) is empty after the first sort
  gauss     (rn, n_gaussian)  N(0,1) value for (ray, j)
The golden generator patches torch.rand_like / torch.randn_like so that the real reference consumes
exactly these tables (tests/golden/make_golden_diner.py); parity pinned by tests/test_oracle_diner.py.
"""
import math

import torch

from . import render as R


def project_points_dict_diner(dataset, height, width, ref, que_pts, include_norm=True):
    """render_ops.py:260-290. que_pts (qn,rn,dn,3) -> dict of (rfn,qn,rn,dn,·)."""
    qn, rn, dn, _ = que_pts.shape
    pix, pdepth, _ = R.project_points(dataset, height, width, ref["w2c"], que_pts.reshape(qn * rn * dn, 3))
    rfn, _, h, w = ref["imgs"].shape
    out = {
        "ref_mvs_depths": R.bilinear_border(ref["mvs_depth"], pix, h, w),
        "ref_mvs_uncert": R.bilinear_border(ref["mvs_uncert"], pix, h, w),
        "pts": pix, "depth": pdepth[..., None],
    }
    if include_norm:
        out["ref_mvs_normal"] = R.bilinear_border(ref["mvs_normal"], pix, h, w)
    return {k: v.reshape(rfn, qn, rn, dn, -1) for k, v in out.items()}


def point_likelihood(cfg, w2c, prj, que_dir, n_candidates, depth_diff_max=0.05, include_norm=True, var=True):
    """original_depth_guided_sample.py:80-196: per-candidate surface likelihood, max over source views.
    Returns (rn, n_candidates)."""
    mu = prj["ref_mvs_depths"].squeeze(-1)                                # (rfn,1,rn,nc)
    uncert = prj["ref_mvs_uncert"].squeeze(-1)
    pd = prj["depth"].squeeze(-1)
    if cfg.get("diner_sigma", 0) > 0:
        sigma = torch.ones_like(mu) * cfg["diner_sigma"]
    else:
        sigma = torch.sqrt(uncert) if var else uncert
    step = torch.ones_like(mu) * (cfg["max_depth"] - cfg["min_depth"]) / n_candidates
    if include_norm:
        d = -que_dir                                                      # (1,rn,nc,3)
        d_cam = torch.einsum("vij,qrcj->vqrci", w2c[:, :3, :3], d)        # (rfn,1,rn,nc,3)
        cos = (d_cam * prj["ref_mvs_normal"]).sum(-1)
        ok = cos <= 0
    else:
        ok = torch.ones_like(mu, dtype=torch.bool)
    mask = ((mu - pd).abs() < depth_diff_max) & ok
    s2 = sigma * math.sqrt(2)
    lik = 0.5 * (torch.erf((pd + step / 2 - mu) / s2) - torch.erf((pd - step / 2 - mu) / s2)).abs()
    lik = torch.where(mask, lik, torch.zeros_like(lik))
    return lik.max(dim=0).values.reshape(mu.shape[2], mu.shape[3])


def fill_up_uniform_samples(cfg, z, fill_rand):
    """original_depth_guided_sample.py:333-366. z (rn,n) with 0 = empty; fill_rand (rn,n)."""
    z = z.sort(dim=-1).values
    missing = z == 0
    n_missing = missing.sum(-1, keepdim=True).to(torch.float32)          # (rn,1)
    near = torch.full_like(n_missing, cfg["min_depth"])
    far = torch.full_like(n_missing, cfg["max_depth"])
    step = (far - near) / n_missing                                      # inf where nothing is missing (unused)
    slot = torch.arange(z.shape[-1], dtype=torch.float32)[None, :]
    filled = near + slot * step
    filled = filled + fill_rand * step
    z = torch.where(missing, filled, z)
    return z.sort(dim=-1).values


def sample_depthguided(cfg, w2c, prj, que_depth, que_dir, n_samples, n_candidates, n_gaussian, fill_rand, gauss=None,
                       depth_diff_max=0.05, include_norm=True, var=True, return_aux=False):
    """original_depth_guided_sample.py:45-297. que_depth (1,rn,nc) -> (1,rn,n_samples)."""
    assert n_samples >= n_gaussian
    lik = point_likelihood(cfg, w2c, prj, que_dir, n_candidates, depth_diff_max, include_norm, var)
    zc = que_depth[0]
    # stable descending order: ties keep the lower candidate index first
    idx = torch.sort(lik, dim=-1, descending=True, stable=True).indices[:, :n_samples]
    sel_lik = torch.gather(lik, -1, idx)
    z = torch.gather(zc, -1, idx)
    z = torch.where(sel_lik == 0, torch.zeros_like(z), z)
    if n_gaussian > 0:
        opaque = lik.clone()
        opaque[:, 1:] = opaque[:, 1:] * R.seq_cumprod(1.0 - lik)[:, :-1]
        ray_mask = (opaque != 0).any(-1)
        wn = opaque / opaque.sum(-1, keepdim=True)
        mean = (zc * wn).sum(-1, keepdim=True)
        std = ((zc - mean).pow(2) * wn).sum(-1, keepdim=True).sqrt()
        g = gauss * std + mean
        g = torch.where(ray_mask[:, None], g, torch.zeros_like(g))
        z = torch.cat([z[:, :n_samples - n_gaussian], g], -1)
    out = fill_up_uniform_samples(cfg, z, fill_rand)[None]
    if return_aux:
        return out, {"likelihood": lik, "selected": idx}
    return out


def diner_sample_placement(cfg, que, ref, fill_rand, gauss=None):
    """renderer.py:570-572 + :318-355: candidates -> depth-guided samples (+ optional uniform samples, sorted)."""
    ds, h, w = cfg["dataset_name"], cfg["height"], cfg["width"]
    rn = que["coords"].shape[1]
    cand = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg["n_candidates"], use_disp=False)
    pts, que_dir = R.depth2points_spherical(ds, h, w, que["c2w"], que["coords"], cand)
    include_norm = bool(cfg.get("backface_culling", False))
    prj = project_points_dict_diner(ds, h, w, ref, pts, include_norm)
    z = sample_depthguided(cfg, ref["w2c"], prj, cand, que_dir, cfg["n_samples"], cfg["n_candidates"],
                           cfg["n_gaussian"], fill_rand, gauss, 0.05, include_norm)
    if cfg.get("contain_uniform", False):
        uni = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg["n_uniform"],
                             use_disp=bool(cfg.get("inv_uniform", False)))
        z = torch.sort(torch.cat([z, uni], -1), -1)[0]
    return z


def render_rays_diner(cfg, W, que, ref, fill_rand, gauss=None):
    """render_impl's diner branch (renderer.py:570-600) without N_uniform merging: the coarse networks evaluated on
    the depth-guided samples; outputs carry the `_fine` suffix unless c2f adds a real fine pass."""
    depth = diner_sample_placement(cfg, que, ref, fill_rand, gauss)
    d_out = R.render_by_depth(cfg, W, que, ref, depth, False)
    d_out["que_depth"] = depth
    if cfg.get("N_uniform", 0) > 0 and cfg.get("one_mlp", False):       # merge_uniform_diner, renderer.py:526-565
        rn = que["coords"].shape[1]
        udepth = R.sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg["depth_sample_num"], use_disp=True)
        u_out = R.render_by_depth(cfg, W, que, ref, udepth, False)
        z, idx = torch.cat([depth, udepth], 2).sort()
        col = torch.gather(torch.cat([d_out["colors_nr"], u_out["colors_nr"]], 2), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3))
        den = torch.gather(torch.cat([d_out["density_nr"], u_out["density_nr"]], 2), 2, idx)
        hit, pix, rdepth = R.composite(den, col, z)
        d_out.update({"pixel_colors_nr": pix, "hit_prob_nr": hit, "colors_nr": col, "density_nr": den, "render_depth": rdepth})
        if cfg.get("render_uncert", False):
            d_out["render_uncert"] = ((z - rdepth.unsqueeze(-1)).pow(2) * hit).sum(-1) + 1e-5
    if cfg.get("c2f", False):
        fine = R.sample_fine_depth(depth, d_out["hit_prob_nr"], que["depth_range"], cfg.get("fine_depth_sample_num", 64),
                                   cfg["use_disp"])
        if que.get("ft_depth_range") is not None:          # fine_render_impl's prior-guided samples also apply here (renderer.py:438-456)
            fdepth = R.fine_depth_with_ft_range(fine, depth, que["ft_depth_range"], cfg["min_depth"], cfg["max_depth"],
                                                cfg.get("fine_depth_use_all", False))
        elif cfg.get("fine_depth_use_all", False):
            fdepth = torch.sort(torch.cat([depth, fine], -1), -1)[0]
        else:
            fdepth = torch.sort(fine, -1)[0]
        f_out = R.render_by_depth(cfg, W, que, ref, fdepth, not cfg.get("one_mlp", False))
        f_out["que_depth"] = fdepth
        out = dict(d_out)
        for k, v in f_out.items():
            out[k + "_fine"] = v
        return out
    return {k + "_fine": v for k, v in d_out.items()}


def depth2normal(dataset, dmap):
    """network/orig_diner_depth2normal.py:7-110 restated: normals of the back-projected depth map by central differences over the
    panorama (zero rows above / below, longitude wrap left / right), then the reference's "cleaning": where the x coordinate of a
    neighbouring point is exactly 0 (a hole or the zero padding) the normal is replaced by the RAW normal of the pixel one step to
    the opposite side, and pixels without depth get a zero normal.  dmap (N,1,H,W) -> (N,3,H,W).  Pinned by
    tests/golden/normal_*.npz (produced by the reference)."""
    from .render import equi_to_unit_dirs
    N, _, H, W = dmap.shape
    pts = equi_to_unit_dirs(dataset, H, W).unsqueeze(0) * dmap.view(N, H, W, 1)          # (N,H,W,3)
    zero = torch.zeros(N, 1, W, 3)
    down = torch.cat([pts[:, 1:], zero], 1)
    up = torch.cat([zero, pts[:, :-1]], 1)
    right = torch.roll(pts, -1, 2)
    left = torch.roll(pts, 1, 2)
    vdiff, hdiff = down - up, right - left
    normal = torch.linalg.cross(vdiff, hdiff, dim=-1)
    normal = normal / torch.norm(normal, p=2, dim=-1, keepdim=True)
    off_y = (up[..., 0] == 0).long() - (down[..., 0] == 0).long()
    off_x = (left[..., 0] == 0).long() - (right[..., 0] == 0).long()
    ys = (torch.arange(H).view(1, H, 1) + off_y).clamp(0, H - 1)
    xs = (torch.arange(W).view(1, 1, W) + off_x).clamp(0, W - 1)
    moved = (off_y != 0) | (off_x != 0)
    n_idx = torch.arange(N).view(N, 1, 1).expand(N, H, W)
    cleaned = torch.where(moved.unsqueeze(-1), normal[n_idx, ys, xs], normal)
    cleaned = torch.where((dmap[:, 0] == 0).unsqueeze(-1), torch.zeros_like(cleaned), cleaned)
    return cleaned.permute(0, 3, 1, 2)
