"""CPU oracle for the spherical-sweep cost volume (HOT 1).

TEST INFRASTRUCTURE ONLY — imported by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py`; never by the product path (`panogrf_b200/`).

Restates, in explicit fp32 tensor arithmetic on the CPU (no `grid_sample`, no python loop over
depths), what the reference computes in
  models/spherical_cost_volume.py:135-230   get_cv_per_depth
  models/spherical_cost_volume.py:231-341   calculate_cost_volume_erp
  models/spherical_cost_volume_mv.py:219-347 calculate_cost_volume_erp_multiview
  helpers/my_torch_helpers.py:12-59 / 62-120 spherical_to_cartesian / cartesian_to_spherical
  network/omni_mvsnet/pipeline3_model.py:849-853  group-wise mean
Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` and committed under `tests/golden/`.
"""
import math

import numpy as np
import torch

DATASETS = ("m3d", "replica_test", "residential", "CoffeeArea")
COST_TYPES = ("abs_diff", "dot", "none")


def pixel_angles(dataset_name, height, width, dtype=torch.float32):
    """ERP pixel -> (theta[W], phi[H]).  spherical_cost_volume.py:272-295."""
    phi = torch.arange(0, height, dtype=dtype)
    theta = torch.arange(0, width, dtype=dtype)
    if dataset_name == "m3d":
        phi = (phi + 0.5) * (np.pi / height)
        theta = (theta + 0.5) * (2 * np.pi / width) - np.pi / 2
    elif dataset_name == "replica_test":
        theta = (2 * np.pi / width) * (theta + 0.5) - np.pi
        phi = -(phi + 0.5) * np.pi / height + np.pi * 0.5
    elif dataset_name == "residential":
        theta = np.pi * (2 * theta / (width - 1) - 1.5)
        phi = np.pi * (0.5 - phi / (height - 1))
    elif dataset_name == "CoffeeArea":
        theta = (-2 * np.pi / (width - 1)) * theta + 2 * np.pi
        phi = (np.pi / (height - 1)) * phi
    else:
        raise Exception(f"unknown dataset_name {dataset_name}")
    return theta, phi


def unit_rays(dataset_name, height, width):
    """(H,W,3) unit ray per ERP pixel.  my_torch_helpers.py:12-59 with r=1."""
    theta, phi = pixel_angles(dataset_name, height, width)
    phi, theta = phi[:, None].expand(height, width), theta[None, :].expand(height, width)
    if dataset_name == "m3d":
        tmp = 1 * torch.sin(phi)
        x, y, z = tmp * torch.cos(theta), 1 * torch.cos(phi), tmp * torch.sin(theta)
    elif dataset_name == "replica_test":
        x = 1 * torch.sin(theta) * torch.cos(phi)
        y = -1 * torch.sin(phi)
        z = 1 * torch.cos(theta) * torch.cos(phi)
    elif dataset_name == "residential":
        x = 1 * torch.cos(theta) * torch.cos(phi)
        z = 1 * torch.sin(theta) * torch.cos(phi)
        y = 1 * torch.sin(phi)
    elif dataset_name == "CoffeeArea":
        x = 1 * torch.sin(phi) * torch.cos(theta)
        y = 1 * torch.sin(phi) * torch.sin(theta)
        z = 1 * torch.cos(phi)
    else:
        raise Exception
    return torch.stack((x, y, z), -1)


def cartesian_to_uv(dataset_name, c):
    """camera-frame point (...,3) -> normalised grid coords (u,v) in [-1,1].

    my_torch_helpers.py:62-120 followed by spherical_cost_volume.py:153-190.
    """
    lin = float(np.deg2rad(10))
    cos_deg = float(np.cos(lin))
    x, y, z = c.unbind(-1)
    radius = torch.sqrt(x * x + y * y + z * z)
    if dataset_name == "m3d":
        theta = torch.atan2(z, x)
        yr = y / radius
        valid = yr.abs() < cos_deg
        phi = torch.where(
            valid,
            torch.acos(torch.where(valid, yr, torch.zeros_like(yr))),
            torch.where(y >= 0, lin * (1 - yr) / (1 - cos_deg), np.pi - lin * (yr + 1) / (-cos_deg + 1)))
        u = torch.fmod(theta + np.pi / 2 + 2 * np.pi, 2 * np.pi)
        v = phi
    elif dataset_name == "replica_test":
        theta = torch.atan2(x, z)
        phi = -torch.asin(z / radius)          # sic: the reference uses z here (:99)
        u = torch.fmod(theta + np.pi + 2 * np.pi, 2 * np.pi)
        v = -phi + 0.5 * np.pi
    elif dataset_name == "residential":
        theta = -torch.atan2(-z, x)
        phi = torch.asin(y / radius)
        theta = torch.where((theta > np.pi * 0.5) & (theta <= 2 * np.pi), theta - 2 * np.pi, theta)
        u = torch.fmod(theta + 3 / 4.0 * 2 * np.pi, 2 * np.pi)
        v = 0.5 * np.pi - phi
    elif dataset_name == "CoffeeArea":
        theta = torch.atan2(y, x)
        phi = torch.acos(z / radius)
        theta = torch.where(theta < 0, theta + 2 * np.pi, theta)
        u = 2 * np.pi - theta
        v = phi
    else:
        raise Exception
    u = u / np.pi - 1
    v = 2 * v / np.pi - 1
    return u, v


def bilinear_zeros_align(src_cl, u, v):
    """grid_sample(bilinear, zeros padding, align_corners=True) on a channels-last map.

    src_cl: (B,H,W,C); u,v: (B,...) normalised.  Returns (B,...,C).
    Weight/tap order follows ATen's grid_sampler_2d (nw,ne,sw,se).
    """
    B, H, W, C = src_cl.shape
    ix = ((u + 1) / 2) * (W - 1)
    iy = ((v + 1) / 2) * (H - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = src_cl.reshape(B, H * W, C)
    shape = u.shape
    out = torch.zeros(*shape, C, dtype=src_cl.dtype)

    def tap(xi, yi, w):
        ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().reshape(B, -1)
        val = torch.gather(flat, 1, idx[..., None].expand(-1, -1, C)).reshape(*shape, C)
        return val * (w * ok.to(w.dtype))[..., None]

    out = tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)
    return out


def _depth_tensor(depths, depth_volume, B, H, W):
    if depth_volume is not None:
        return depth_volume.to(torch.float32)                       # (B,D,H,W)
    d = torch.as_tensor(depths, dtype=torch.float32).reshape(1, -1, 1, 1)
    return d.expand(B, d.shape[1], H, W)


def sweep_uv(dataset_name, depth, rot_ref, tran_ref, rot_src, tran_src, return_radius=False):
    """(u,v) of every (b,d,y,x) voxel in the source panorama; get_cv_per_depth :137-159."""
    B, D, H, W = depth.shape
    xyz = unit_rays(dataset_name, H, W)                             # (H,W,3)
    m = depth[..., None] * xyz[None, None]                          # (B,D,H,W,3)
    inv_ref = torch.inverse(rot_ref)                                # (B,3,3)
    w = torch.einsum("bij,bdhwj->bdhwi", inv_ref, m - tran_ref[:, None, None, None, :])
    c = torch.einsum("bij,bdhwj->bdhwi", rot_src, w) + tran_src[:, None, None, None, :]
    if return_radius:
        return cartesian_to_uv(dataset_name, c) + (torch.linalg.norm(c, dim=-1),)
    return cartesian_to_uv(dataset_name, c)


def cost_volume_erp(dataset_name, images, depths, trans, rots, depth_volume=None, cost_type="abs_diff",
                    ref_idx=1, src_views=(0,), divisor=None, return_uv=False):
    """Generalised sweep: sum over `src_views` of (per-view cost / divisor).

    images (B,S,H,W,C) channels-last, rots (B,S,3,3), trans (B,S,3) world->camera.
    Returns (B,D,H,W,C) contiguous.
    """
    if cost_type not in COST_TYPES:
        raise ValueError("Unknown cost type")
    images = images.to(torch.float32)
    B, S, H, W, C = images.shape
    depth = _depth_tensor(depths, depth_volume, B, H, W)
    ref = images[:, ref_idx]                                        # (B,H,W,C)
    total = None
    uvs = []
    for s in src_views:
        u, v = sweep_uv(dataset_name, depth, rots[:, ref_idx], trans[:, ref_idx], rots[:, s], trans[:, s])
        uvs.append((u, v))
        assert bool(((u >= -1) & (u <= 1) & (v >= -1) & (v <= 1)).all()), \
            "Wrong UV mapping, UV must be in [-1, 1]!"
        warped = bilinear_zeros_align(images[:, s], u, v)           # (B,D,H,W,C)
        if cost_type == "abs_diff":
            cost = (warped - ref[:, None]).abs()
        elif cost_type == "dot":
            cost = warped * ref[:, None]
        else:
            cost = warped
        if divisor is not None:
            cost = cost / divisor                                    # _mv.py:331 `/(seq_len-2)`
        total = cost if total is None else total + cost
    if return_uv:
        return total, uvs
    return total


def calculate_cost_volume_erp(args, images, depths, trans, rots, depth_volume=None, cost_type="abs_diff", **_):
    """Oracle twin of models/spherical_cost_volume.py:231 (images[:,0]=source, [:,1]=reference)."""
    dv = depth_volume if args.get("contain_dnet") else None
    return cost_volume_erp(args["dataset_name"], images, depths, trans, rots, dv, cost_type,
                           ref_idx=1, src_views=(0,))


def calculate_cost_volume_erp_multiview(args, images, depths, trans, rots, depth_volume=None,
                                        cost_type="abs_diff", curr_idx=0, **_):
    """Oracle twin of models/spherical_cost_volume_mv.py:219: views 0..S-2 except curr_idx, each /(S-2)."""
    S = images.shape[1]
    assert S > 2
    views = tuple(v for v in range(S - 1) if v != curr_idx)
    dv = depth_volume if args.get("contain_dnet") else None
    return cost_volume_erp(args["dataset_name"], images, depths, trans, rots, dv, cost_type,
                           ref_idx=curr_idx, src_views=views, divisor=S - 2)


def group_mean(cost_volume_bdhwc, groups):
    """pipeline3_model.py:847-853: (B,D,H,W,C) -> (B,G,D,H,W) mean over C/G consecutive channels."""
    B, D, H, W, C = cost_volume_bdhwc.shape
    cv = cost_volume_bdhwc.permute(0, 4, 1, 2, 3).reshape(B, groups, C // groups, D, H, W)
    return cv.mean(2)


def mono_guided_hypotheses(ref_mu, k_list, fixed_sigma, min_depth, max_depth, n_linear):
    """pipeline3_model.py:733,800-815: clamp(mu+k*sigma) ++ linspace, sorted along D. (B,1,h,w)->(B,D,h,w)."""
    B, _, H, W = ref_mu.shape
    mono = [torch.clamp(ref_mu + k * fixed_sigma, min=min_depth, max=max_depth) for k in k_list]
    lin = torch.linspace(min_depth, max_depth, n_linear).reshape(1, n_linear, 1, 1).repeat(B, 1, H, W)
    vol = torch.cat(mono + [lin], 1)
    return torch.sort(vol, dim=1)[0]


def depth_hypotheses(args, ref_gmms, k_list, cost_volume_channels, contain_dnet=True):
    """pipeline3_model.py:717-733, 774-821 restated (every switch).  Pinned by tests/golden/hyp_*.npz, which are produced by
    executing the reference's own source lines (tests/golden/make_golden_hypotheses.py).  Returns (depth_volume, d_centers)."""
    lo, hi = args["min_depth"], args["max_depth"]
    n_samples = len(k_list)
    if not contain_dnet:
        n = cost_volume_channels
        return None, (torch.linspace(lo, hi, n) if args["use_depth_sampling"] else 1.0 / torch.linspace(1 / lo, 1 / hi, n))
    ref_mu = ref_gmms[:, :1]
    vol = None
    if n_samples > 0:
        if args["mono_uncertainty"] or args.get("mono_uncert_tune"):
            ref_sigma = ref_gmms[:, 1:]
            if args.get("relaxation_factor") in args:          # sic (:726): the value looked up as a key
                mono = [torch.clamp(ref_mu + ref_sigma * float(k) * args["relaxation_factor"], min=lo, max=hi) for k in k_list]
            else:
                mono = [torch.clamp(ref_mu + torch.clamp(ref_sigma, min=args["basic_sigma"]) * float(k), min=lo, max=hi) for k in k_list]
        else:
            mono = [torch.clamp(ref_mu + float(k) * args["fixed_sigma"], min=lo, max=hi) for k in k_list]
        vol = torch.cat(mono, 1)
    if args["wo_hdh"]:
        return vol, None
    n = cost_volume_channels - n_samples
    B, _, H, W = ref_mu.shape
    if args["use_depth_sampling"]:
        if args["revise_range"]:
            d_min = torch.clamp(ref_mu - args["fixed_dist"], min=lo)
            d_max = torch.clamp(ref_mu + args["fixed_dist"], max=hi)
            interval = (d_max - d_min) / (n - 1)
            cen = torch.cat([d_min, d_min + interval * torch.arange(0, n - 1).reshape(1, n - 1, 1, 1)], 1)
        else:
            cen = torch.linspace(lo, hi, n).reshape(1, n, 1, 1)
    else:
        cen = 1.0 / torch.linspace(1 / lo, 1 / hi, n).reshape(1, n, 1, 1)
    if not args["revise_range"]:
        cen = cen.repeat(B, 1, H, W)
    vol = cen if vol is None else torch.cat([vol, cen], 1)
    return torch.sort(vol, dim=1)[0], cen


def magnet_k_list(n_samples=5, sampling_range=3):
    """pipeline3_model.py:537-545 without scipy: midpoints of equal-probability normal quantiles."""
    from statistics import NormalDist
    p_total = math.erf(sampling_range / math.sqrt(2))
    ps = [(1 - p_total) / 2 + (i / n_samples) * p_total for i in range(n_samples + 1)]
    ks = [NormalDist().inv_cdf(p) for p in ps]
    return [(a + b) / 2 for a, b in zip(ks[1:], ks[:-1])]
