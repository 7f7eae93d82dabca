"""CPU oracle for the per-ray render path (HOT 2, 3, 4) of PanoGRF.

TEST INFRASTRUCTURE ONLY — imported by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py`; never by the product path (`panogrf_b200/`).

A functional fp32 restatement (explicit index arithmetic and matmuls on CPU tensors; no
`grid_sample`, no `nn.Module`) of what the reference computes per ray batch in
  network/render_ops.py:76-122,126-153,158-257,292-339,413-473
  network/ray_utils.py:4-22,53-71      network/spt_utils.py:37-199 (all four dataset conventions)
  network/ops.py:32-52 (interpolate_feats)
  network/dist_decoder.py:6-51,99-140  network/aggregate_net.py:8-14,41-89
  network/ibrnet.py:15-27,52-102,112-116,305-373
  network/renderer.py:120-136,180-188,210-219,223-317,435-524,567-633
Weights are addressed by the reference's own `state_dict()` names (e.g.
`agg_net.agg_impl.base_fc.0.weight`), so the same checkpoint / random init feeds the reference,
this oracle and the CUDA path.

Parity pinning: the reference has no tests or golden vectors for this path; the
oracle is pinned against outputs of the reference itself (`tests/golden/make_golden_render.py`,
fixtures `tests/golden/render_*.npz`).  Accumulation order of `cumsum`/`cumprod` is sequential
left-to-right fp32 (what torch does on CPU), which is the order the CUDA scan reproduces.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# ERP conventions of the render path (network/spt_utils.py) — (W-1),(H-1) pixel convention
# ------------------------------------------------------------------------------------------------


def equi_to_unit_dirs(dataset, height, width):
    """get_sphere_ray_directions (ray_utils.py:4-16): (H,W,3) unit directions in the camera frame."""
    x = torch.linspace(0, width - 1, width).view(1, width).expand(height, width)
    y = torch.linspace(0, height - 1, height).view(height, 1).expand(height, width)
    if dataset == "m3d":
        x = x.clamp(0, width - 1)
        y = y.clamp(0, height - 1)
        theta = x / (width - 1) * 2 * math.pi
        theta = theta - 0.5 * math.pi
        phi = y / (height - 1) * math.pi
    elif dataset == "replica_test":
        theta = x * 2 * math.pi / (width - 1) - math.pi
        phi = -y * math.pi / (height - 1) + math.pi * 0.5
    elif dataset == "residential":
        theta = math.pi * (2 * x / (width - 1) - 1.5)
        phi = math.pi * (0.5 - y / (height - 1))
    elif dataset == "CoffeeArea":
        theta = (-2 * math.pi / (width - 1)) * x + 2 * math.pi
        phi = (math.pi / (height - 1)) * y
    else:
        raise Exception
    rad = torch.ones_like(theta)
    if dataset == "residential":
        xs, zs, ys = rad * torch.cos(theta) * torch.cos(phi), rad * torch.sin(theta) * torch.cos(phi), rad * torch.sin(phi)
    elif dataset == "m3d":
        tmp = rad * torch.sin(phi)
        xs, ys, zs = tmp * torch.cos(theta), rad * torch.cos(phi), tmp * torch.sin(theta)
    elif dataset == "replica_test":
        xs, ys, zs = rad * torch.sin(theta) * torch.cos(phi), -rad * torch.sin(phi), rad * torch.cos(theta) * torch.cos(phi)
    else:
        xs, ys, zs = rad * torch.sin(phi) * torch.cos(theta), rad * torch.sin(phi) * torch.sin(theta), rad * torch.cos(phi)
    d = torch.stack([xs, ys, zs], -1)
    return d / torch.norm(d, p=2, dim=-1).unsqueeze(-1)


def cartesian_to_equi(dataset, pts_cam, height, width):
    """cartesian_2_equi (ray_utils.py:18-22): (...,3) -> depth (...), pixel (...,2) with (x,y) order."""
    xc, yc, zc = pts_cam.unbind(-1)
    radius = torch.linalg.norm(pts_cam, dim=-1)
    if dataset == "m3d":
        theta = torch.atan2(zc, xc)
        phi = torch.acos(yc / (radius + 1e-5))
        theta = torch.remainder(theta + 0.5 * math.pi, 2 * math.pi)
        px = theta / (2 * math.pi) * (width - 1)
        py = phi / math.pi * (height - 1)
    elif dataset == "replica_test":
        theta = torch.atan2(xc, zc)
        phi = -torch.asin(yc / radius)
        px = ((width - 1) / (2 * math.pi)) * (theta + math.pi)
        py = (height - 1) / math.pi * (-phi + 0.5 * math.pi)
    elif dataset == "residential":
        theta = -torch.atan2(-zc, xc)
        phi = torch.asin(yc / radius)
        theta = torch.where((theta > math.pi * 0.5) & (theta <= 2 * math.pi), theta - 2 * math.pi, theta)
        px = ((1 / (2.0 * math.pi)) * theta + (3 / 4.0)) * (width - 1)
        py = (0.5 - phi / math.pi) * (height - 1)
    elif dataset == "CoffeeArea":
        theta = torch.atan2(yc, xc)
        phi = torch.acos(zc / radius)
        theta = torch.where(theta < 0, theta + 2 * math.pi, theta)
        px = (width - 1) * (1 - theta / (2.0 * math.pi))
        py = phi * (height - 1) / math.pi
    else:
        raise Exception
    return radius, torch.stack([px, py], -1)


# ------------------------------------------------------------------------------------------------
# sampling along the ray
# ------------------------------------------------------------------------------------------------


def sample_depth(min_depth, max_depth, rn, sample_num, use_disp=True):
    """render_ops.py:292-339 with random_sample=False. Returns (1,rn,dn) depths."""
    near = torch.ones(1) * min_depth
    far = torch.ones(1) * max_depth
    dn = sample_num
    assert dn > 2
    val = torch.arange(1, dn - 1, dtype=torch.float32)[None, None, :] + torch.zeros(1, rn, dn - 2)
    if not use_disp:
        interval = (far - near) / (dn - 1)
        ticks = interval[:, None, None] * val
        diff = far - near
        ticks = torch.cat([torch.zeros(1, rn, 1), ticks, diff[:, None, None].repeat(1, rn, 1)], -1)
        return near[:, None, None] + ticks
    interval = (1 / far - 1 / near) / (dn - 1)
    ticks = interval[:, None, None] * val
    diff = 1 / far - 1 / near
    ticks = torch.cat([torch.zeros(1, rn, 1), ticks, diff[:, None, None].repeat(1, rn, 1)], -1)
    return 1 / (1 / near[:, None, None] + ticks)


def depth2inv_dists(depth, depth_range):
    """render_ops.py:110-122: normalised inverse-depth intervals, last = 1e6."""
    near, far = -1 / depth_range[:, 0], -1 / depth_range[:, 1]
    near, far = near[:, None, None], far[:, None, None]
    depth_inv = (-1 / depth - near) / (far - near)
    dists = depth_inv[..., 1:] - depth_inv[..., :-1]
    return torch.cat([dists, torch.full([*depth.shape[:-1], 1], 1e6, dtype=torch.float32)], -1)


#: bench.py's eager-GPU baseline sets this: use torch.cumsum / torch.cumprod (one launch each, as the reference does,
#: render_ops.py:152,438-439) instead of the stated sequential order (64 python iterations).  Never set by the parity checks.
TORCH_SCANS = False


def seq_cumsum(x):
    """Sequential left-to-right fp32 cumulative sum over the last dim — the STATED accumulation order
    of this build.  torch.cumsum accumulates in fp64 on the CPU and with a parallel scan
    on CUDA, so neither is a fixed fp32 order; the CUDA kernels reproduce exactly this loop."""
    if TORCH_SCANS:
        return torch.cumsum(x, -1)
    out = torch.empty_like(x)
    acc = torch.zeros_like(x[..., 0])
    for i in range(x.shape[-1]):
        acc = acc + x[..., i]
        out[..., i] = acc
    return out


def seq_cumprod(x):
    """Sequential left-to-right fp32 cumulative product over the last dim (see seq_cumsum)."""
    if TORCH_SCANS:
        return torch.cumprod(x, -1)
    out = torch.empty_like(x)
    acc = torch.ones_like(x[..., 0])
    for i in range(x.shape[-1]):
        acc = acc * x[..., i]
        out[..., i] = acc
    return out


def fine_sample_u(fdn):
    """The deterministic u-table of sample_fine_depth (render_ops.py:442-445)."""
    interval = 1 / fdn
    return 0.5 * interval + torch.arange(fdn) * interval


def sample_fine_depth(depth, hit_prob, depth_range, sample_num, use_disp=True, return_indices=False):
    """render_ops.py:413-473 with random_sample=False (inv_mode == use_disp)."""
    inv_mode = bool(use_disp)
    if inv_mode:
        near, far = depth_range[0, 0], depth_range[0, 1]
        near, far = -1 / near, -1 / far
        depth = (-1 / depth - near) / (far - near)
    center = (depth[..., 1:] + depth[..., :-1]) / 2
    center = torch.cat([depth[..., 0:1], center, depth[..., -1:]], -1)
    hp = hit_prob + 1e-5
    # stated accumulation order: sequential left-to-right fp32 for the normaliser as well as the cdf
    pdf = hp / seq_cumsum(hp)[..., -1:]
    cdf = seq_cumsum(pdf)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = fine_sample_u(sample_num).expand(list(cdf.shape[:-1]) + [sample_num]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(center, -1, below), torch.gather(center, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    fine = bin_b + t * (bin_a - bin_b)
    if inv_mode:
        fine = fine * (far - near) + near
        fine = -1 / fine
    if return_indices:
        return fine, inds
    return fine


def sample_3sigma_tables(n):
    """Per-call constants of sample_3sigma (sample_utils.py:7,11-12), built with the reference's own torch ops:
    t = linspace(0,1,n) (bin positions AND the deterministic u of sample_pdf), g = N(0,1) density at linspace(-3,3,n-1)."""
    t = torch.linspace(0., 1., steps=n)
    x = torch.linspace(-3., 3., steps=n - 1)
    g = 1. / math.sqrt(2 * np.pi) * torch.exp(-0.5 * x.pow(2))
    return t, g


def sample_3sigma(low, high, n, near, far, u=None):
    """sample_utils.py:6-15 + sample_pdf :18-60 with det=True (u = linspace(0,1,n); pass `u` (R,n) for the random branch).
    low, high (R,) -> (R,n) depths inside [low, high] clamped to [near, far], Gaussian-weighted bins, unsorted like the reference.
    Accumulation order stated like sample_fine_depth: sequential left-to-right fp32 for the normaliser and the cdf."""
    t, g = sample_3sigma_tables(n)
    step = (high - low) / (n - 1)
    edges = (low.unsqueeze(-1) * (1. - t) + high.unsqueeze(-1) * t).clamp(near, far)
    factor = (edges[..., 1:] - edges[..., :-1]) / step.unsqueeze(-1)
    w = factor * g.unsqueeze(0) + 1e-5
    pdf = w / seq_cumsum(w)[..., -1:]
    cdf = seq_cumsum(pdf)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    uu = t.expand(list(cdf.shape[:-1]) + [n]).contiguous() if u is None else u.contiguous()
    inds = torch.searchsorted(cdf, uu, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(edges, -1, below), torch.gather(edges, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    tt = (uu - cdf_b) / denom
    return bin_b + tt * (bin_a - bin_b)


def fine_depth_with_ft_range(fine_depth, coarse_depth, ft_depth_range, min_depth, max_depth, use_all):
    """fine_render_impl, renderer.py:438-470 (eval): rays whose ft_depth_range[...,0] >= min_depth take sample_3sigma between
    ft_depth_range[...,1] and [...,2] instead of the inverse-CDF samples; then the sort (with the coarse depths when use_all).
    fine_depth, coarse_depth (1,rn,dn); ft_depth_range (1,rn,3)."""
    valid = ft_depth_range[..., 0] >= min_depth
    z2 = fine_depth.clone()
    n = coarse_depth.shape[-1]
    if bool(valid.any()):
        z2[valid] = sample_3sigma(ft_depth_range[valid][:, 1], ft_depth_range[valid][:, 2], n, min_depth, max_depth)
    if use_all:
        return torch.sort(torch.cat([coarse_depth, z2], -1), -1)[0]
    return torch.sort(z2, -1)[0]


# ------------------------------------------------------------------------------------------------
# geometry: rays, projection, gathers
# ------------------------------------------------------------------------------------------------


def depth2points_spherical(dataset, height, width, c2w, coords, depth):
    """render_ops.py:76-106. c2w (1,3,4), coords (1,rn,2) (x,y), depth (1,rn,dn) -> pts, dir (1,rn,dn,3)."""
    dirs = equi_to_unit_dirs(dataset, height, width)                     # (H,W,3)
    c2w = c2w.reshape(3, 4)
    rays_d = torch.matmul(c2w[:3, :3].view(1, 1, 3, 3), dirs.view(height, width, 3, 1)).view(height, width, 3)
    iy, ix = coords[:, :, 1].long(), coords[:, :, 0].long()
    centers = c2w[:, 3].view(1, 1, 3).expand(coords.shape[0], coords.shape[1], 3)
    directions = rays_d[iy, ix, :]                                       # (1,rn,3)
    pts = centers.unsqueeze(2) + directions.unsqueeze(2) * depth.unsqueeze(3)
    que_dir = -directions / torch.norm(directions, dim=2, keepdim=True)
    return pts, que_dir.unsqueeze(2).repeat(1, 1, depth.shape[2], 1)


def depth2points_perspec(coords, poses, Ks, depth):
    """coords2rays + depth2points_perspec (render_ops.py:37-74): pinhole query rays.  coords (qn,rn,2), poses (qn,3,4) world->camera,
    Ks (qn,3,3), depth (qn,rn,dn) -> pts, dir (qn,rn,dn,3); directions are not normalised before the points are formed."""
    rot = poses[:, :, :3].unsqueeze(1).permute(0, 1, 3, 2)
    trans = -rot @ poses[:, :, 3:].unsqueeze(1)
    qn, rn, _ = coords.shape
    centers = trans.repeat(1, rn, 1, 1).squeeze(-1)
    hom = torch.cat([coords, torch.ones([qn, rn, 1], dtype=torch.float32)], 2)
    cam_xyz = rot @ (torch.inverse(Ks).unsqueeze(1) @ hom.unsqueeze(3)) + trans
    directions = cam_xyz.squeeze(3) - centers
    pts = centers.unsqueeze(2) + directions.unsqueeze(2) * depth.unsqueeze(3)
    que_dir = -directions / torch.norm(directions, dim=2, keepdim=True)
    return pts, que_dir.unsqueeze(2).repeat(1, 1, depth.shape[2], 1)


def project_points(dataset, height, width, w2c, pts):
    """project_points_ref_views (render_ops.py:158-230). w2c (rfn,3,4), pts (pn,3) ->
    pixel (rfn,pn,2), depth (rfn,pn), dir (rfn,pn,3)."""
    pn = pts.shape[0]
    hpts = torch.cat([pts, torch.ones(pn, 1)], 1)
    rfn = w2c.shape[0]
    last = torch.zeros(rfn, 1, 4)
    last[:, :, 3] = 1.0
    Hm = torch.cat([w2c, last], 1)
    pts_cam = (Hm[:, None] @ hpts[None, :, :, None])[:, :, :3, 0]
    depth, pix = cartesian_to_equi(dataset, pts_cam, height, width)
    cam = -w2c[:, :, :3].permute(0, 2, 1) @ w2c[:, :, 3:]               # (rfn,3,1) camera centres
    d = pts.unsqueeze(0) - cam.permute(0, 2, 1)
    d = -d / torch.clamp_min(torch.norm(d, dim=2, keepdim=True), min=1e-5)
    return pix, depth, d


def bilinear_border(feats, pix, h, w):
    """interpolate_feature_map -> interpolate_feats -> grid_sample (render_ops.py:126-143, ops.py:32-52).

    feats (rfn,f,fh,fw), pix (rfn,pn,2) in FULL-RES pixel units -> (rfn,pn,f).
    bilinear, padding_mode='border', align_corners = (fh==h and fw==w).
    """
    rfn, f, fh, fw = feats.shape
    align = (fh == h and fw == w)
    xn = pix[:, :, 0] / (w - 1) * 2 - 1
    yn = pix[:, :, 1] / (h - 1) * 2 - 1
    if align:
        ix = ((xn + 1) / 2) * (fw - 1)
        iy = ((yn + 1) / 2) * (fh - 1)
    else:
        ix = ((xn + 1) * fw - 1) / 2
        iy = ((yn + 1) * fh - 1) / 2
    ix = ix.clamp(0, fw - 1)                                             # border padding: clip_coordinates
    iy = iy.clamp(0, fh - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    w_nw, w_ne = (x1 - ix) * (y1 - iy), (ix - x0) * (y1 - iy)
    w_sw, w_se = (x1 - ix) * (iy - y0), (ix - x0) * (iy - y0)
    flat = feats.reshape(rfn, f, fh * fw)

    def tap(xi, yi, wgt):
        ok = (xi >= 0) & (xi <= fw - 1) & (yi >= 0) & (yi <= fh - 1)
        idx = (yi.clamp(0, fh - 1) * fw + xi.clamp(0, fw - 1)).long()
        val = torch.gather(flat, 2, idx[:, None, :].expand(-1, f, -1))  # (rfn,f,pn)
        return val * (wgt * ok.to(wgt.dtype))[:, None, :]

    out = tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)
    return out.permute(0, 2, 1)


# ------------------------------------------------------------------------------------------------
# small MLP helpers addressed by state_dict names
# ------------------------------------------------------------------------------------------------


def _lin(W, prefix, x):
    y = x @ W[prefix + ".weight"].t()
    b = W.get(prefix + ".bias")
    return y if b is None else y + b


def _mlp3(W, prefix, x, last):
    """Linear-ELU-Linear-ELU-Linear-<last> (dist_decoder.py:64-97)."""
    x = F.elu(_lin(W, prefix + ".0", x))
    x = F.elu(_lin(W, prefix + ".2", x))
    return last(_lin(W, prefix + ".4", x))


def dist_decoder_forward(W, prefix, feats, use_vis, bias_val=0.05):
    mean = _mlp3(W, prefix + ".mean_decoder", feats, F.softplus)
    var = _mlp3(W, prefix + ".var_decoder", feats, F.softplus) + bias_val
    aw = _mlp3(W, prefix + ".aw_decoder", feats, torch.sigmoid)
    vis = _mlp3(W, prefix + ".vis_decoder", feats, torch.sigmoid) if use_vis else None
    return mean, var, vis, aw


def compute_prob_ref(depth, interval, mean, var, vis, aw, depth_range, use_vis):
    """dist_decoder.py:109-140 with is_ref=True. depth (rfn,qn,rn,dn), interval (1,qn,rn,dn)."""
    near_r = -1 / depth_range[:, 0][:, None, None, None]
    far_r = -1 / depth_range[:, 1][:, None, None, None]
    d = -1 / torch.clamp(depth, min=1e-5)
    d = (d - near_r) / (far_r - near_r)
    half = interval / 2
    ext = torch.cat([half[..., 0:1], half], -1)
    near = (d - ext[..., :-1])[..., None]
    far = (d + ext[..., 1:])[..., None]
    mix = torch.cat([aw, 1 - aw], -1)
    cdf0 = 0.5 + 0.5 * torch.tanh((near - mean) * var)
    cdf1 = 0.5 + 0.5 * torch.tanh((far - mean) * var)
    if use_vis:
        cdf0, cdf1 = cdf0 * vis, cdf1 * vis
    visibility = torch.sum((1 - cdf0) * mix, -1)
    hit_prob = torch.sum((cdf1 - cdf0) * mix, -1)
    eps = 1e-5
    alpha = torch.log(hit_prob / (visibility - hit_prob + eps) + eps)
    return alpha, visibility, hit_prob


def compute_prob_que(depth, interval, mean, var, vis, aw, depth_range, use_vis):
    """dist_decoder.py:109-140 with is_ref=False (the query rays' own depth distribution, training only): depth, interval (qn,rn,dn),
    mean / var (qn,rn,1|dn,2), vis / aw (qn,rn,1|dn,1), depth_range (qn,2).  Near / far are the midpoints between consecutive
    normalised inverse depths (dist_decoder.py:37-45)."""
    near_q = -1 / depth_range[:, 0][:, None, None]
    far_q = -1 / depth_range[:, 1][:, None, None]
    d = -1 / torch.clamp(depth, min=1e-5)
    d = (d - near_q) / (far_q - near_q)
    half = interval / 2
    first = d[..., 0] - half[..., 0]
    last = d[..., -1] + half[..., -1]
    ext = torch.cat([first[..., None], (d[..., :-1] + d[..., 1:]) / 2, last[..., None]], -1)
    near, far = ext[..., :-1, None], ext[..., 1:, None]
    mix = torch.cat([aw, 1 - aw], -1)
    cdf0 = 0.5 + 0.5 * torch.tanh((near - mean) * var)
    cdf1 = 0.5 + 0.5 * torch.tanh((far - mean) * var)
    if use_vis:
        cdf0, cdf1 = cdf0 * vis, cdf1 * vis
    visibility = torch.sum((1 - cdf0) * mix, -1)
    hit_prob = torch.sum((cdf1 - cdf0) * mix, -1)
    eps = 1e-5
    alpha = torch.log(hit_prob / (visibility - hit_prob + eps) + eps)
    return alpha, visibility, hit_prob


def posenc_table(d_hid, n_samples):
    """ibrnet.py:305-313 (fp64 numpy table cast to fp32)."""
    pos = np.arange(n_samples)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table).float().unsqueeze(0).to(torch.empty(0).device)     # follows `with torch.device(...)`


def _mean_var(x, weight):
    mean = torch.sum(x * weight, dim=2, keepdim=True)
    var = torch.sum(weight * (x - mean) ** 2, dim=2, keepdim=True)
    return mean, var


def ibrnet_forward(W, p, rgb_feat, neuray_feat, ray_diff, mask, n_samples):
    """IBRNetWithNeuRay.forward (ibrnet.py:315-373). Inputs (rn,dn,rfn,*). Returns (rn,dn,4)."""
    num_views = rgb_feat.shape[2]
    direction_feat = F.elu(_lin(W, p + ".ray_dir_fc.2", F.elu(_lin(W, p + ".ray_dir_fc.0", ray_diff))))
    rgb_in = rgb_feat[..., :3]
    rgb_feat = rgb_feat + direction_feat
    weight = mask / (torch.sum(mask, dim=2, keepdim=True) + 1e-8)
    nf = _lin(W, p + ".neuray_fc.2", F.elu(_lin(W, p + ".neuray_fc.0", neuray_feat)))
    weight0 = torch.sigmoid(nf) * weight
    mean0, var0 = _mean_var(rgb_feat, weight0)
    mean1, var1 = _mean_var(rgb_feat, weight)
    globalfeat = torch.cat([mean0, var0, mean1, var1], dim=-1)
    x = torch.cat([globalfeat.expand(-1, -1, num_views, -1), rgb_feat, neuray_feat], dim=-1)
    x = F.elu(_lin(W, p + ".base_fc.2", F.elu(_lin(W, p + ".base_fc.0", x))))
    x_vis = F.elu(_lin(W, p + ".vis_fc.2", F.elu(_lin(W, p + ".vis_fc.0", x * weight))))
    x_res, vis = torch.split(x_vis, [x_vis.shape[-1] - 1, 1], dim=-1)
    vis = torch.sigmoid(vis) * mask
    x = x + x_res
    vis = torch.sigmoid(_lin(W, p + ".vis_fc2.2", F.elu(_lin(W, p + ".vis_fc2.0", x * vis)))) * mask
    weight = vis / (torch.sum(vis, dim=2, keepdim=True) + 1e-8)
    mean, var = _mean_var(x, weight)
    globalfeat = torch.cat([mean.squeeze(2), var.squeeze(2), weight.mean(dim=2)], dim=-1)
    globalfeat = F.elu(_lin(W, p + ".geometry_fc.2", F.elu(_lin(W, p + ".geometry_fc.0", globalfeat))))
    num_valid_obs = torch.sum(mask, dim=2)
    globalfeat = globalfeat + posenc_table(16, n_samples)
    # ---- MultiHeadAttention(4, 16, 4, 4) over the dn samples of a ray (ibrnet.py:52-102)
    a = p + ".ray_attention"
    rn, dn, _ = globalfeat.shape
    q = (globalfeat @ W[a + ".w_qs.weight"].t()).view(rn, dn, 4, 4).transpose(1, 2)
    k = (globalfeat @ W[a + ".w_ks.weight"].t()).view(rn, dn, 4, 4).transpose(1, 2)
    v = (globalfeat @ W[a + ".w_vs.weight"].t()).view(rn, dn, 4, 4).transpose(1, 2)
    attn = torch.matmul(q / (4 ** 0.5), k.transpose(2, 3))
    amask = (num_valid_obs > 1).float().unsqueeze(1)                     # (rn,1,dn,1): broadcast over keys
    attn = attn.masked_fill(amask == 0, -1e9)
    attn = F.softmax(attn, dim=-1)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(rn, dn, -1)
    o = o @ W[a + ".fc.weight"].t()
    o = o + globalfeat
    o = F.layer_norm(o, (16,), W[a + ".layer_norm.weight"], W[a + ".layer_norm.bias"], eps=1e-6)
    sigma = F.relu(_lin(W, p + ".out_geometry_fc.2", F.elu(_lin(W, p + ".out_geometry_fc.0", o))))
    sigma_out = sigma.masked_fill(num_valid_obs < 1, 0.)
    x = torch.cat([x, vis, ray_diff], dim=-1)
    x = _lin(W, p + ".rgb_fc.4", F.elu(_lin(W, p + ".rgb_fc.2", F.elu(_lin(W, p + ".rgb_fc.0", x)))))
    x = x.masked_fill(mask == 0, -1e9)
    blend = F.softmax(x, dim=2)
    rgb_out = torch.sum(rgb_in * blend, dim=2)
    return torch.cat([rgb_out, sigma_out], dim=-1)


def agg_net_forward(W, p, prj, que_dir, n_samples, wo_geometry=False, wo_appearance=False):
    """DefaultAggregationNet.forward (aggregate_net.py:41-89). Returns density (qn,rn,dn), colors (qn,rn,dn,3)."""
    hit = (prj["hit_prob"] - 0.5) * 2
    vis = (prj["vis"] - 0.5) * 2
    rfn, qn, rn, dn, _ = hit.shape
    emb = torch.cat([prj["ray_feats"], hit, vis], -1)
    emb = _lin(W, p + ".prob_embed.2", F.relu(_lin(W, p + ".prob_embed.0", emb)))
    if wo_geometry:                                                       # aggregate_net.py:60-62
        emb = torch.zeros_like(emb)
    dir_diff = prj["dir"] - que_dir.unsqueeze(0)
    dir_dot = torch.sum(prj["dir"] * que_dir.unsqueeze(0), -1, keepdim=True)
    dir_diff = torch.cat([dir_diff, dir_dot], -1).reshape(rfn, qn * rn, dn, -1).permute(1, 2, 0, 3)
    mask = torch.ones(qn * rn, dn, rfn, 1)
    img = torch.cat([prj["rgb"], prj["img_feats"]], -1).reshape(rfn, qn * rn, dn, -1).permute(1, 2, 0, 3)
    if wo_appearance:                                                     # aggregate_net.py:79-81
        img = torch.zeros_like(img)
    emb = emb.reshape(rfn, qn * rn, dn, -1).permute(1, 2, 0, 3)
    outs = ibrnet_forward(W, p + ".agg_impl", img, emb, dir_diff, mask, n_samples)
    return outs[..., 3].reshape(qn, rn, dn), outs[..., :3].reshape(qn, rn, dn, 3)


def alpha_values2hit_prob(alpha):
    """render_ops.py:145-153."""
    no_hit = torch.cat([torch.ones((*alpha.shape[:-1], 1)), 1. - alpha + 1e-10], -1)
    return alpha * seq_cumprod(no_hit)[..., :-1]


def composite(density, colors, depth):
    """renderer.py:210-219 + :302-304. Returns hit_prob, pixel_colors (qn,rn,3), render_depth (qn,rn)."""
    alpha = 1.0 - torch.exp(-torch.relu(density))
    hit = alpha_values2hit_prob(alpha)
    return hit, torch.sum(hit.unsqueeze(-1) * colors, 2), torch.sum(hit * depth, -1)


# ------------------------------------------------------------------------------------------------
# one render_by_depth pass and the coarse+fine driver
# ------------------------------------------------------------------------------------------------


def agg_sample_num(cfg, is_fine):
    """Length of the positional table of the (fine) aggregation net: a top-level cfg["sample_num"]
    overrides both nets (renderer.py:67-69), else `[fine_]agg_net_cfg.sample_num`, default 64."""
    if "sample_num" in cfg:
        return cfg["sample_num"]
    sub = cfg.get("fine_agg_net_cfg" if is_fine else "agg_net_cfg", {}) or {}
    return sub.get("sample_num", 64)


def render_by_depth(cfg, W, que, ref, depth, is_fine, return_prj=False, is_perspec=False):
    """renderer.py:223-317 (eval, non-debug). `ref` holds imgs, w2c, depth_range, ray_feats, img_feats."""
    ds, h, w = cfg["dataset_name"], cfg["height"], cfg["width"]
    dists = depth2inv_dists(depth, que["depth_range"])
    if is_perspec:                                                        # renderer.py:231-232
        pts, que_dir = depth2points_perspec(que["coords"], que["poses"], que["Ks"], depth)
    else:
        pts, que_dir = depth2points_spherical(ds, h, w, que["c2w"], que["coords"], depth)
    qn, rn, dn, _ = pts.shape
    rfn, _, ih, iw = ref["imgs"].shape
    pix, pdepth, pdir = project_points(ds, h, w, ref["w2c"], pts.reshape(qn * rn * dn, 3))
    prj = {
        "dir": pdir, "pts": pix, "depth": pdepth[..., None],
        "ray_feats": bilinear_border(ref["ray_feats"], pix, ih, iw),
        "rgb": bilinear_border(ref["imgs"], pix, ih, iw),
    }
    prj = {k: v.reshape(rfn, qn, rn, dn, -1) for k, v in prj.items()}
    dd = "fine_dist_decoder" if is_fine else "dist_decoder"
    use_vis = cfg["dist_decoder_cfg"].get("use_vis", True) if not is_fine else \
        cfg.get("fine_dist_decoder_cfg", {}).get("use_vis", True)
    mean, var, vis, aw = dist_decoder_forward(W, dd, prj["ray_feats"], use_vis)
    # the reference always calls self.dist_decoder.compute_prob (renderer.py:129), i.e. the COARSE cfg's use_vis
    alpha, visibility, hit_prob = compute_prob_ref(prj["depth"].squeeze(-1), dists.unsqueeze(0), mean, var, vis, aw,
                                                   ref["depth_range"], cfg["dist_decoder_cfg"].get("use_vis", True))
    prj["alpha"], prj["vis"], prj["hit_prob"] = alpha[..., None], visibility[..., None], hit_prob[..., None]
    prj["img_feats"] = bilinear_border(ref["img_feats"], pix, ih, iw).reshape(rfn, qn, rn, dn, -1)
    agg = "fine_agg_net" if is_fine else "agg_net"
    n_samples = agg_sample_num(cfg, is_fine)
    sub = cfg.get("fine_agg_net_cfg" if is_fine else "agg_net_cfg") or {}     # renderer.py:67-78 copies the top-level switches
    density, colors = agg_net_forward(W, agg, prj, que_dir, n_samples,
                                      bool(cfg.get("wo_geometry", sub.get("wo_geometry", False))),
                                      bool(cfg.get("wo_appearance", sub.get("wo_appearance", False))))
    hit, pixel_colors, render_depth = composite(density, colors, depth)
    out = {"pixel_colors_nr": pixel_colors, "hit_prob_nr": hit, "colors_nr": colors, "density_nr": density,
           "render_depth": render_depth}
    if cfg.get("render_uncert", False):                                   # renderer.py:299-301
        out["render_uncert"] = ((depth - render_depth.unsqueeze(-1)).pow(2) * hit).sum(-1) + 1e-5
    if return_prj:
        out["prj"] = prj
        out["que_dir"] = que_dir
        out["dists"] = dists
    return out


def render_rays(cfg, W, que, ref, keep_hit_prob=False, is_perspec=False):
    """render_impl (renderer.py:567-633), default (non-diner) eval branch, one ray batch."""
    rn = que["coords"].shape[1]
    depth = sample_depth(cfg["min_depth"], cfg["max_depth"], rn, cfg.get("depth_sample_num", 64), cfg["use_disp"])
    out = render_by_depth(cfg, W, que, ref, depth, False, is_perspec=is_perspec)
    out["que_depth"] = depth
    if cfg.get("use_hierarchical_sampling", False):
        fine = sample_fine_depth(depth, out["hit_prob_nr"], que["depth_range"], cfg.get("fine_depth_sample_num", 64),
                                 cfg["use_disp"])
        if que.get("ft_depth_range") is not None:                        # renderer.py:438-456: prior-guided samples for valid rays
            fdepth = fine_depth_with_ft_range(fine, depth, que["ft_depth_range"], cfg["min_depth"], cfg["max_depth"],
                                              cfg.get("fine_depth_use_all", False))
        elif cfg.get("fine_depth_use_all", False):
            fdepth = torch.sort(torch.cat([depth, fine], -1), -1)[0]
        else:
            fdepth = torch.sort(fine, -1)[0]
        fout = render_by_depth(cfg, W, que, ref, fdepth, not cfg.get("one_mlp", False), is_perspec=is_perspec)
        if cfg.get("render_c2f_all", False):                              # renderer.py:484-521: coarse + fine samples together
            z, idx = torch.cat([depth, fdepth], 2).sort()
            col = torch.gather(torch.cat([out["colors_nr"], fout["colors_nr"]], 2), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3))
            den = torch.gather(torch.cat([out["density_nr"], fout["density_nr"]], 2), 2, idx)
            hit, pix, rdepth = composite(den, col, z)
            fout.update({"pixel_colors_nr": pix, "hit_prob_nr": hit, "colors_nr": col, "density_nr": den, "render_depth": rdepth})
            if cfg.get("render_uncert", False):
                fout["render_uncert"] = ((z - rdepth.unsqueeze(-1)).pow(2) * hit).sum(-1) + 1e-5
        fout["que_depth"] = fdepth
        for k, v in fout.items():
            out[k + "_fine"] = v
    if not keep_hit_prob:
        out = {k: v for k, v in out.items() if not k.startswith("hit_prob")}
    return out


def render(cfg, W, que, ref, ray_batch_num=None):
    """NeuralRayBaseRenderer.render's ray-batch loop (renderer.py:647-683) on pre-encoded feature maps."""
    ray_batch_num = ray_batch_num or cfg.get("ray_batch_num", 2048)
    coords = que["coords"]
    outs = {}
    for r0 in range(0, coords.shape[1], ray_batch_num):
        q = dict(que)
        q["coords"] = coords[:, r0:r0 + ray_batch_num]
        if que.get("ft_depth_range") is not None:
            q["ft_depth_range"] = que["ft_depth_range"][:, r0:r0 + ray_batch_num]
        o = render_rays(cfg, W, q, ref)
        for k, v in o.items():
            outs.setdefault(k, []).append(v)
    return {k: torch.cat(v, 1) for k, v in outs.items()}
