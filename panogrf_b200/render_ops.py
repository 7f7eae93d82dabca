"""Drop-ins for the functional operators of the reference's `network/render_ops.py` (same names, argument meaning and
result shapes), each one CUDA kernel through the C ABI (csrc/render_ops.cu).  The fused renderer
(`panogrf_b200.renderer`) does not call these; they exist for callers of the reference's functional API and for
per-kernel measurements."""
import math

import torch

from . import _lib
from .renderer import coarse_depth_table, fine_u_table, tensor_version, to_channels_last


def _f32(t):
    return t.contiguous().float()


_CL_CACHE = {}


def _channels_last_cached(name, t, pad_to=None):
    """NCHW -> channels-last copy, reused while the caller keeps passing the same unmodified tensor."""
    ver = tensor_version(t)
    if ver is None:
        return to_channels_last(t, pad_to)                  # untracked (inference-mode) tensor: never cached
    key = (t.data_ptr(), ver, tuple(t.shape), str(t.device))
    hit = _CL_CACHE.get(name)
    if hit is None or hit[0] != key:
        _CL_CACHE[name] = (key, to_channels_last(t, pad_to), t)
    return _CL_CACHE[name][1]


def sample_depth(args, coords, sample_num, random_sample, use_disp=True):
    """render_ops.py:292-339 (deterministic branch): (qn,rn,dn) depths and the (qn,rn,dn) forward differences."""
    if random_sample:
        raise NotImplementedError("random_sample=True (training) is not part of the render-time path")
    qn, rn, _ = coords.shape
    table = coarse_depth_table(args, sample_num, use_disp).to(coords.device)
    depth = table.view(1, 1, -1).expand(qn, rn, -1).contiguous()
    dists = torch.cat([depth[..., 1:], torch.full_like(depth[..., :1], 1e6)], -1) - depth
    return depth, dists


def depth2inv_dists(depth, depth_range):
    """render_ops.py:115-122 (tiny element-wise op kept in torch: it is not on the fused path)."""
    near, far = -1 / depth_range[:, 0], -1 / depth_range[:, 1]
    near, far = near[:, None, None], far[:, None, None]
    inv = (-1 / depth - near) / (far - near)
    return torch.cat([inv[..., 1:] - inv[..., :-1], torch.full_like(inv[..., :1], 1e6)], -1)


class _CompositeFn(torch.autograd.Function):
    """pgrf_composite_fwd / pgrf_composite_bwd.  `x` is the density (from_density) or the alpha values; colors / depth may be
    None.  Gradients flow to x and colors (the sample depths carry none on this path)."""

    @staticmethod
    def forward(ctx, x, colors, depth, from_density):
        lib = _lib.load()
        rn, dn = x.shape
        hit = torch.empty_like(x)
        pix = torch.empty(rn, 3, device=x.device) if colors is not None else None
        rd = torch.empty(rn, device=x.device) if depth is not None else None
        with torch.cuda.device(x.device):
            rc = lib.pgrf_composite_fwd(_lib.ptr(x) if from_density else None, None if from_density else _lib.ptr(x), _lib.ptr(colors),
                                        _lib.ptr(depth), dn if depth is not None else 0, rn, dn, _lib.ptr(hit), _lib.ptr(pix),
                                        _lib.ptr(rd), _lib.stream_ptr())
        _lib.check(rc, "pgrf_composite_fwd")
        ctx.from_density = from_density
        ctx.has = (colors is not None, depth is not None)
        ctx.save_for_backward(x, colors if colors is not None else x.new_empty(0), depth if depth is not None else x.new_empty(0))
        outs = [hit]
        outs.append(pix if pix is not None else x.new_empty(0))
        outs.append(rd if rd is not None else x.new_empty(0))
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_hit, g_pix, g_rd):
        lib = _lib.load()
        x, colors, depth = ctx.saved_tensors
        has_c, has_d = ctx.has
        rn, dn = x.shape
        gx = torch.empty_like(x)
        gc = torch.empty_like(colors) if has_c else None
        f = lambda t: None if t is None else t.contiguous().float()
        g_hit, g_pix, g_rd = f(g_hit), f(g_pix) if has_c else None, f(g_rd) if has_d else None
        with torch.cuda.device(x.device):
            rc = lib.pgrf_composite_bwd(_lib.ptr(x) if ctx.from_density else None, None if ctx.from_density else _lib.ptr(x),
                                        _lib.ptr(colors) if has_c else None, _lib.ptr(depth) if has_d else None, dn if has_d else 0,
                                        rn, dn, _lib.ptr(g_hit), _lib.ptr(g_pix), _lib.ptr(g_rd), _lib.ptr(gx), _lib.ptr(gc),
                                        _lib.stream_ptr())
        _lib.check(rc, "pgrf_composite_bwd")
        return gx, gc, None, None


def alpha_values2hit_prob(alpha_values):
    """render_ops.py:145-153: (...,dn) -> (...,dn); sequential fp32 cumprod (the stated accumulation order); differentiable."""
    _lib.require_cuda(alpha_values)
    shape = alpha_values.shape
    a = _f32(alpha_values).reshape(-1, shape[-1])
    hit, _, _ = _CompositeFn.apply(a, None, None, False)
    return hit.reshape(shape)


def composite(density, colors, depth):
    """network_rendering's tail (renderer.py:214-218) + render_depth (:302-304): density (qn,rn,dn), colors (qn,rn,dn,3),
    depth (qn,rn,dn) -> hit_prob (qn,rn,dn), pixel_colors (qn,rn,3), render_depth (qn,rn).  Differentiable w.r.t. density and
    colors (pgrf_composite_bwd)."""
    _lib.require_cuda(density, colors, depth)
    qn, rn, dn = density.shape
    d, c, z = _f32(density).reshape(-1, dn), _f32(colors).reshape(-1, dn, 3), _f32(depth).detach().reshape(-1, dn)
    hit, pix, rd = _CompositeFn.apply(d, c, z, True)
    return hit.reshape(qn, rn, dn), pix.reshape(qn, rn, 3), rd.reshape(qn, rn)


def sample_fine_depth(args, depth, hit_prob, depth_range, sample_num, random_sample, inv_mode=True, return_indices=False):
    """render_ops.py:413-473 (deterministic branch): (qn,rn,dn) x2 -> (qn,rn,sample_num), unsorted like the reference."""
    if random_sample:
        raise NotImplementedError("random_sample=True (training) is not part of the render-time path")
    _lib.require_cuda(depth, hit_prob)
    lib = _lib.load()
    if not args["use_disp"]:
        inv_mode = False
    qn, rn, dn = depth.shape
    d, h = _f32(depth).reshape(-1, dn), _f32(hit_prob).reshape(-1, dn)
    u = fine_u_table(sample_num).to(d.device)
    dr = depth_range.float().cpu()
    out = torch.empty(d.shape[0], sample_num, device=d.device)
    inds = torch.empty(d.shape[0], sample_num, device=d.device, dtype=torch.int32) if return_indices else None
    with torch.cuda.device(d.device):
        rc = lib.pgrf_fine_sample_fwd(_lib.ptr(d), dn, _lib.ptr(h), _lib.ptr(u), float(dr[0, 0]), float(dr[0, 1]), int(bool(inv_mode)),
                                      d.shape[0], dn, sample_num, 0, 0, _lib.ptr(out), _lib.ptr(inds), _lib.stream_ptr())
    _lib.check(rc, "pgrf_fine_sample_fwd")
    out = out.reshape(qn, rn, sample_num)
    return (out, inds.reshape(qn, rn, sample_num).long()) if return_indices else out


def sample_3sigma_tables(n, device):
    """Constants of sample_3sigma (sample_utils.py:7,11-12) with the reference's own torch ops: t = linspace(0,1,n) (bin positions and
    the deterministic u of sample_pdf), g = unit-Gaussian density at linspace(-3,3,n-1)."""
    t = torch.linspace(0., 1., steps=n)
    x = torch.linspace(-3., 3., steps=n - 1)
    g = 1. / math.sqrt(2 * math.pi) * torch.exp(-0.5 * x.pow(2))
    return t.to(device), g.to(device)


def sample_3sigma(low_3sigma, high_3sigma, N, det, near, far):
    """network/sample_utils.py:6-15 (+ sample_pdf :18-60), deterministic branch: (R,) x2 -> (R,N) depths, unsorted like the reference."""
    if not det:
        raise NotImplementedError("sample_3sigma(det=False) draws torch.rand: not part of the render-time path")
    _lib.require_cuda(low_3sigma, high_3sigma)
    lib = _lib.load()
    lh = torch.stack([_f32(low_3sigma).reshape(-1), _f32(high_3sigma).reshape(-1)], -1).contiguous()
    t, g = sample_3sigma_tables(int(N), lh.device)
    out = torch.empty(lh.shape[0], int(N), device=lh.device, dtype=torch.float32)
    with torch.cuda.device(lh.device):
        rc = lib.pgrf_sample_3sigma_fwd(_lib.ptr(lh), 2, 0.0, _lib.ptr(t), _lib.ptr(g), int(N), float(near), float(far), None, 0, 0, 0,
                                        lh.shape[0], _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "pgrf_sample_3sigma_fwd")
    return out.reshape(*low_3sigma.shape, int(N))


def project_points_dict(ref_imgs_info, que_pts, spt_utils, with_img_feats=True):
    """render_ops.py:234-257 (+ get_img_feats, renderer.py:180-188, when `img_feats` is present): que_pts (qn,rn,dn,3)
    -> dict of (rfn,qn,rn,dn,*) tensors: dir, pts, depth, ray_feats, rgb[, img_feats].  `spt_utils` only has to expose
    `.dataset`, `.height`, `.width` like network/spt_utils.Utils."""
    _lib.require_cuda(que_pts, ref_imgs_info["imgs"], ref_imgs_info["ray_feats"])
    lib = _lib.load()
    qn, rn, dn, _ = que_pts.shape
    pts = _f32(que_pts).reshape(-1, 3)
    pn = pts.shape[0]
    imgs = ref_imgs_info["imgs"]
    rfn, _, ih, iw = imgs.shape
    dev = pts.device
    imgs_cl = _channels_last_cached("imgs", imgs, 4)
    rf_cl = _channels_last_cached("ray_feats", ref_imgs_info["ray_feats"])
    has_if = with_img_feats and "img_feats" in ref_imgs_info
    if_cl = _channels_last_cached("img_feats", ref_imgs_info["img_feats"]) if has_if else None
    w2c = _f32(ref_imgs_info["w2c"]).to(dev)
    e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    pix, dep, dr, rf, rgb = e(rfn, pn, 2), e(rfn, pn, 1), e(rfn, pn, 3), e(rfn, pn, 32), e(rfn, pn, 3)
    imf = e(rfn, pn, 32) if has_if else None
    name = spt_utils.dataset
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset {name!r}")
    with torch.cuda.device(dev):
        rc = lib.pgrf_project_gather_fwd(
            _lib.ptr(pts), pn, _lib.ptr(w2c), rfn, _lib.DATASET_IDS[name], int(spt_utils.height), int(spt_utils.width),
            _lib.ptr(imgs_cl), ih, iw, _lib.ptr(if_cl), if_cl.shape[1] if has_if else 0, if_cl.shape[2] if has_if else 0,
            _lib.ptr(rf_cl), rf_cl.shape[1], rf_cl.shape[2], _lib.ptr(pix), _lib.ptr(dep), _lib.ptr(dr), _lib.ptr(rf), _lib.ptr(rgb),
            _lib.ptr(imf), _lib.stream_ptr())
    _lib.check(rc, "pgrf_project_gather_fwd")
    out = {"dir": dr, "pts": pix, "depth": dep, "ray_feats": rf, "rgb": rgb}
    if has_if:
        out["img_feats"] = imf
    return {k: v.reshape(rfn, qn, rn, dn, -1) for k, v in out.items()}


class _InterpolateFn(torch.autograd.Function):
    """pgrf_interpolate_feature_map_fwd / _bwd (gradient w.r.t. the map; pixel coordinates carry none on this path)."""

    @staticmethod
    def forward(ctx, feats, pix, h, w):
        lib = _lib.load()
        rfn, f, fh, fw = feats.shape
        pn = pix.shape[1]
        out = torch.empty(rfn, pn, f, device=feats.device, dtype=torch.float32)
        with torch.cuda.device(feats.device):
            rc = lib.pgrf_interpolate_feature_map_fwd(_lib.ptr(feats), rfn, f, fh, fw, _lib.ptr(pix), pn, h, w, _lib.ptr(out),
                                                      _lib.stream_ptr())
        _lib.check(rc, "pgrf_interpolate_feature_map_fwd")
        ctx.save_for_backward(pix)
        ctx.meta = (rfn, f, fh, fw, pn, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (pix,) = ctx.saved_tensors
        rfn, f, fh, fw, pn, h, w = ctx.meta
        g = g.contiguous().float()
        gf = torch.zeros(rfn, f, fh, fw, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            rc = lib.pgrf_interpolate_feature_map_bwd(_lib.ptr(g), rfn, f, fh, fw, _lib.ptr(pix), pn, h, w, _lib.ptr(gf), _lib.stream_ptr())
        _lib.check(rc, "pgrf_interpolate_feature_map_bwd")
        return gf, None, None, None


def interpolate_feature_map(ray_feats, coords, h, w, border_type="border"):
    """render_ops.py:126-143: ray_feats (rfn,f,fh,fw), coords (rfn,pn,2) in (h,w) pixel units -> (rfn,pn,f); bilinear,
    border padding, align_corners only when the map is at full resolution.  Differentiable w.r.t. ray_feats."""
    if border_type != "border":
        raise NotImplementedError("only border padding is used on the render path")
    _lib.require_cuda(ray_feats, coords)
    rfn = ray_feats.shape[0]
    pn = coords.shape[1]
    return _InterpolateFn.apply(_f32(ray_feats), _f32(coords).detach().reshape(rfn, pn, 2), int(h), int(w))


def depth2points_spherical(que_imgs_info, que_depth, spt_utils):
    """render_ops.py:76-106: que_depth (qn=1,rn,dn) -> que_pts (1,rn,dn,3), que_dir (1,rn,dn,3)."""
    c2w = que_imgs_info["c2w"]
    assert c2w.shape[0] == 1, "que_imgs_info c2w.shape[0]=1"
    _lib.require_cuda(que_depth, que_imgs_info["coords"])
    lib = _lib.load()
    qn, rn, dn = que_depth.shape
    assert qn == 1
    dev = que_depth.device
    name = spt_utils.dataset
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset {name!r}")
    depth = _f32(que_depth).reshape(rn, dn)
    coords = _f32(que_imgs_info["coords"]).reshape(rn, 2)
    c = _f32(c2w).reshape(3, 4).to(dev)
    pts, dirs = torch.empty(1, rn, dn, 3, device=dev), torch.empty(1, rn, dn, 3, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgrf_depth2points_fwd(_lib.ptr(coords), _lib.ptr(depth), dn, _lib.ptr(c), _lib.DATASET_IDS[name],
                                       int(spt_utils.height), int(spt_utils.width), rn, dn, _lib.ptr(pts), _lib.ptr(dirs),
                                       _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth2points_fwd")
    return pts, dirs


def depth2normal(ref_imgs_info, spt_utils):
    """network/orig_diner_depth2normal.py:7-110: ref_imgs_info['mvs_depth'] (N,1,H,W) -> normals (N,3,H,W) by central
    differences of the back-projected panorama (the renderer fills ref_imgs_info['mvs_normal'] with it, renderer.py:714)."""
    dmap = ref_imgs_info["mvs_depth"]
    _lib.require_cuda(dmap)
    name = spt_utils.dataset
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset {name!r}")
    N, _, H, W = dmap.shape
    d = _f32(dmap)
    dev = d.device
    raw = torch.empty(N, H, W, 3, device=dev)
    offs = torch.empty(N, H, W, 2, device=dev, dtype=torch.int8)
    out = torch.empty(N, 3, H, W, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().pgrf_depth2normal_fwd(_lib.ptr(d), N, H, W, _lib.DATASET_IDS[name], _lib.ptr(raw), _lib.ptr(offs),
                                               _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth2normal_fwd")
    return out


def mono_guided_hypotheses(ref_mu, k_list, fixed_sigma, min_depth, max_depth, n_linear):
    """Depth hypotheses of the MVS net (pipeline3_model.py:723-733, 774-815): clamp(ref_mu + k*sigma) for k in k_list,
    concatenated with linspace(min,max,n_linear) and sorted per pixel.  ref_mu (B,1,h,w) -> (B,len(k_list)+n_linear,h,w)."""
    _lib.require_cuda(ref_mu)
    lib = _lib.load()
    B, _, h, w = ref_mu.shape
    mu = _f32(ref_mu)
    ks = torch.tensor(sorted(float(k) * float(fixed_sigma) for k in k_list), dtype=torch.float32, device=mu.device)
    lin = torch.linspace(min_depth, max_depth, n_linear).to(mu.device)
    out = torch.empty(B, ks.numel() + n_linear, h, w, device=mu.device)
    with torch.cuda.device(mu.device):
        rc = lib.pgrf_depth_hypotheses_fwd(_lib.ptr(mu), B, h, w, _lib.ptr(ks), ks.numel(), _lib.ptr(lin), n_linear,
                                           float(min_depth), float(max_depth), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth_hypotheses_fwd")
    return out


def magnet_k_list(n_samples=5, sampling_range=3):
    """pipeline3_model.py:537-545 (`depth_sampling`): midpoints of equal-probability normal quantiles covering
    +-sampling_range sigma.  fp64, without scipy: statistics.NormalDist().inv_cdf agrees with scipy's norm.ppf to 1e-15."""
    import math
    from statistics import NormalDist
    p_total = math.erf(sampling_range / math.sqrt(2))
    ps = [(1 - p_total) / 2 + (i / n_samples) * p_total for i in range(n_samples + 1)]
    ks = [NormalDist().inv_cdf(p) for p in ps]
    return [(a + b) / 2 for a, b in zip(ks[1:], ks[:-1])]


def depth_hypotheses(args, ref_gmms, k_list, cost_volume_channels, contain_dnet=True):
    """The MVS net's hypothesis builder with every switch of the reference (pipeline3_model.py:717-733, 774-821).

    args keys read: min_depth, max_depth, mono_uncertainty, mono_uncert_tune, fixed_sigma, basic_sigma, relaxation_factor, wo_hdh,
    use_depth_sampling, revise_range, fixed_dist.  `k_list` = magnet_k_list(MAGNET_num_samples, MAGNET_sampling_range) ([] when
    n_samples == 0).  Returns (depth_volume (B,D,h,w) or None, d_centers): with `contain_dnet` the per-pixel sorted volume (and the
    centres the reference keeps alongside), else (None, 1-D centres) exactly as the reference passes them to the cost volume."""
    lib = _lib.load()
    lo, hi = args["min_depth"], args["max_depth"]
    n_samples = len(k_list)
    if not contain_dnet:
        n = int(cost_volume_channels)
        cen = torch.linspace(lo, hi, n) if args["use_depth_sampling"] else 1.0 / torch.linspace(1 / lo, 1 / hi, n)
        return None, cen.to(ref_gmms.device if ref_gmms is not None else "cpu")
    _lib.require_cuda(ref_gmms)
    dev = ref_gmms.device
    B, _, h, w = ref_gmms.shape
    mu = _f32(ref_gmms[:, :1])
    use_sigma = bool(args["mono_uncertainty"] or args.get("mono_uncert_tune"))
    sigma = _f32(ref_gmms[:, 1:2]) if (use_sigma and n_samples > 0) else None
    mono_mode, ks = 0, None
    if n_samples > 0:
        if use_sigma:
            # the reference tests `args["relaxation_factor"] in self.args` — the VALUE as a key (:726) — kept as written
            mono_mode = 2 if args.get("relaxation_factor") in args else 1
            ks = torch.tensor([float(k) for k in k_list], dtype=torch.float32, device=dev)
        else:
            ks = torch.tensor([float(k) * float(args["fixed_sigma"]) for k in k_list], dtype=torch.float32, device=dev)
    wo_hdh = bool(args["wo_hdh"])
    n_cen = int(cost_volume_channels) - n_samples
    cen_mode, cen, d_centers = 2, None, None
    if not wo_hdh:
        if args["use_depth_sampling"] and args["revise_range"]:
            cen_mode = 1
        else:
            cen_mode = 0
            cen = (torch.linspace(lo, hi, n_cen) if args["use_depth_sampling"] else 1.0 / torch.linspace(1 / lo, 1 / hi, n_cen)).to(dev)
            d_centers = cen.reshape(1, n_cen, 1, 1)
    D = n_samples + (0 if wo_hdh else n_cen)
    out = torch.empty(B, D, h, w, device=dev)
    with torch.cuda.device(dev):
        rc = lib.pgrf_depth_hypotheses2_fwd(_lib.ptr(mu), _lib.ptr(sigma), B, h, w, _lib.ptr(ks), n_samples, mono_mode,
                                            float(args.get("basic_sigma", 0.0)), float(args.get("relaxation_factor", 1.0)),
                                            _lib.ptr(cen), n_cen, cen_mode, float(args.get("fixed_dist", 0.0)), float(lo), float(hi),
                                            0 if wo_hdh else 1, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth_hypotheses2_fwd")
    return out, d_centers


# ------------------------------------------------------------------------------------------------
# depth-prior sample placement (the "diner" branch of render_impl)
# ------------------------------------------------------------------------------------------------

def _cand_step(cfg, n_candidates):
    """(max_depth - min_depth) / n_candidates exactly as the reference's fp32 tensor arithmetic evaluates it
    (original_depth_guided_sample.py:131)."""
    return float((torch.ones(1) * (cfg["max_depth"] - cfg["min_depth"]) / n_candidates)[0])


def _random_tables(rn, n_samples, n_gaussian, dev, fill_rand, gauss, generator=None):
    """The reference draws these with torch.rand_like / torch.randn_like inside the op; here they are arguments (drawn
    on the device when not given) so that a run is reproducible and checkable."""
    if fill_rand is None:
        fill_rand = torch.rand(rn, n_samples, device=dev, generator=generator)
    if n_gaussian > 0 and gauss is None:
        gauss = torch.randn(rn, n_gaussian, device=dev, generator=generator)
    fill_rand = _f32(fill_rand).to(dev).reshape(rn, n_samples)
    gauss = _f32(gauss).to(dev).reshape(rn, n_gaussian) if n_gaussian > 0 else None
    return fill_rand, gauss


def project_points_dict_diner(ref_imgs_info, diner_que_pts, spt_utils, include_norm=False):
    """render_ops.py:260-290: candidates (qn,rn,dn,3) -> {'ref_mvs_depths','ref_mvs_uncert','pts','depth'[,'ref_mvs_normal']}
    of (rfn,qn,rn,dn,*)."""
    _lib.require_cuda(diner_que_pts, ref_imgs_info["mvs_depth"], ref_imgs_info["mvs_uncert"])
    lib = _lib.load()
    qn, rn, dn, _ = diner_que_pts.shape
    pts = _f32(diner_que_pts).reshape(-1, 3)
    pn = pts.shape[0]
    rfn, _, ih, iw = ref_imgs_info["imgs"].shape
    dev = pts.device
    md, mu_ = _f32(ref_imgs_info["mvs_depth"]), _f32(ref_imgs_info["mvs_uncert"])
    mn = _f32(ref_imgs_info["mvs_normal"]) if include_norm else None
    mh, mw = md.shape[-2:]
    w2c = _f32(ref_imgs_info["w2c"]).to(dev)
    e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    pix, dep, mu, unc = e(rfn, pn, 2), e(rfn, pn, 1), e(rfn, pn, 1), e(rfn, pn, 1)
    nrm = e(rfn, pn, 3) if include_norm else None
    name = spt_utils.dataset
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset {name!r}")
    with torch.cuda.device(dev):
        rc = lib.pgrf_project_gather_diner_fwd(
            _lib.ptr(pts), pn, _lib.ptr(w2c), rfn, _lib.DATASET_IDS[name], int(spt_utils.height), int(spt_utils.width),
            _lib.ptr(md), _lib.ptr(mu_), _lib.ptr(mn), mh, mw, ih, iw, _lib.ptr(pix), _lib.ptr(dep), _lib.ptr(mu), _lib.ptr(unc),
            _lib.ptr(nrm), _lib.stream_ptr())
    _lib.check(rc, "pgrf_project_gather_diner_fwd")
    out = {"ref_mvs_depths": mu, "ref_mvs_uncert": unc, "pts": pix, "depth": dep}
    if include_norm:
        out["ref_mvs_normal"] = nrm
    return {k: v.reshape(rfn, qn, rn, dn, -1) for k, v in out.items()}


def sample_depthguided(cfg, ref_imgs_info, prj_depth_info_dict, que_depth, que_dir, n_samples, n_candidates, n_gaussian,
                       depth_diff_max=0.05, include_norm=False, var=True, fill_rand=None, gauss=None, generator=None):
    """original_depth_guided_sample.py:45-297 on a precomputed project_points_dict_diner result: (1,rn,n_candidates)
    candidate depths -> (1,rn,n_samples) sorted sample depths.  `fill_rand` (rn,n_samples) / `gauss` (rn,n_gaussian)
    replace the reference's in-op random draws."""
    assert n_samples >= n_gaussian
    _lib.require_cuda(que_depth, que_dir, prj_depth_info_dict["depth"])
    lib = _lib.load()
    dev = que_depth.device
    rn = que_depth.shape[1]
    rfn = prj_depth_info_dict["depth"].shape[0]
    fill_rand, gauss = _random_tables(rn, n_samples, n_gaussian, dev, fill_rand, gauss, generator)
    a = _lib.DinerArgs()
    a.rfn, a.rn = rfn, rn
    a.n_candidates, a.n_samples, a.n_gaussian, a.n_uniform = n_candidates, n_samples, n_gaussian, 0
    a.include_norm, a.sigma_is_var = int(bool(include_norm)), int(bool(var))
    a.diner_sigma = float(cfg["diner_sigma"]) if cfg.get("diner_sigma", 0) > 0 else 0.0
    a.cand_step = _cand_step(cfg, n_candidates)
    a.min_depth, a.max_depth, a.depth_diff_max = float(cfg["min_depth"]), float(cfg["max_depth"]), float(depth_diff_max)
    cand = _f32(que_depth).reshape(rn, n_candidates)
    w2c = _f32(ref_imgs_info["w2c"]).to(dev)
    keep = [cand, w2c, fill_rand, gauss]
    mu = _f32(prj_depth_info_dict["ref_mvs_depths"]).reshape(rfn, rn, n_candidates)
    unc = _f32(prj_depth_info_dict["ref_mvs_uncert"]).reshape(rfn, rn, n_candidates)
    pd = _f32(prj_depth_info_dict["depth"]).reshape(rfn, rn, n_candidates)
    nrm = _f32(prj_depth_info_dict["ref_mvs_normal"]).reshape(rfn, rn, n_candidates, 3) if include_norm else None
    qd = _f32(que_dir).reshape(rn, n_candidates, 3)
    keep += [mu, unc, pd, nrm, qd]
    a.cand_depth, a.cand_ray_stride, a.ref_w2c = _lib.ptr(cand), n_candidates, _lib.ptr(w2c)
    a.fill_rand, a.gauss = _lib.ptr(fill_rand), _lib.ptr(gauss)
    a.prj_mu, a.prj_uncert, a.prj_depth, a.prj_normal, a.que_dir = (_lib.ptr(mu), _lib.ptr(unc), _lib.ptr(pd), _lib.ptr(nrm),
                                                                    _lib.ptr(qd))
    out = torch.empty(1, rn, n_samples, device=dev, dtype=torch.float32)
    a.out_depth = _lib.ptr(out)
    with torch.cuda.device(dev):
        rc = lib.pgrf_depth_guided_sample_fwd(a, _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth_guided_sample_fwd")
    del keep
    return out


def depth_guided_placement(cfg, que_imgs_info, ref_imgs_info, fill_rand=None, gauss=None, generator=None,
                           return_likelihood=False):
    """The sample placement of `diner_render_by_depth` (renderer.py:318-349) fused into ONE kernel: linear candidates
    (`n_candidates`), projection into the source panoramas, MVS-prior gathers, likelihood, selection of `n_samples`,
    Gaussian samples, fill-up, optional `n_uniform` uniform samples (cfg contain_uniform / inv_uniform), sort.
    Returns (1,rn,n_samples[+n_uniform]) depths (and the (rn,n_candidates) likelihood when asked for)."""
    coords = que_imgs_info["coords"]
    _lib.require_cuda(coords, ref_imgs_info["mvs_depth"], ref_imgs_info["mvs_uncert"])
    lib = _lib.load()
    dev = coords.device
    rn = coords.shape[1]
    nc, ns, ng = int(cfg["n_candidates"]), int(cfg["n_samples"]), int(cfg["n_gaussian"])
    include_norm = bool(cfg.get("backface_culling", False))
    name = cfg["dataset_name"]
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset {name!r}")
    fill_rand, gauss = _random_tables(rn, ns, ng, dev, fill_rand, gauss, generator)
    nu = int(cfg["n_uniform"]) if cfg.get("contain_uniform", False) else 0
    uni = coarse_depth_table(cfg, nu, bool(cfg.get("inv_uniform", False))).to(dev) if nu > 0 else None
    cand = coarse_depth_table(cfg, nc, False).to(dev)
    rfn, _, ih, iw = ref_imgs_info["imgs"].shape
    md, mu_ = _f32(ref_imgs_info["mvs_depth"]), _f32(ref_imgs_info["mvs_uncert"])
    mn = _f32(ref_imgs_info["mvs_normal"]) if include_norm else None
    c2w = _f32(que_imgs_info["c2w"]).to(dev).reshape(3, 4)
    w2c = _f32(ref_imgs_info["w2c"]).to(dev)
    xy = _f32(coords).reshape(rn, 2)
    a = _lib.DinerArgs()
    a.dataset, a.H, a.W, a.rfn, a.rn = _lib.DATASET_IDS[name], int(cfg["height"]), int(cfg["width"]), rfn, rn
    a.n_candidates, a.n_samples, a.n_gaussian, a.n_uniform = nc, ns, ng, nu
    a.include_norm, a.sigma_is_var = int(include_norm), 1
    a.diner_sigma = float(cfg["diner_sigma"]) if cfg.get("diner_sigma", 0) > 0 else 0.0
    a.cand_step = _cand_step(cfg, nc)
    a.min_depth, a.max_depth, a.depth_diff_max = float(cfg["min_depth"]), float(cfg["max_depth"]), 0.05
    a.coords, a.cand_depth, a.cand_ray_stride = _lib.ptr(xy), _lib.ptr(cand), 0
    a.que_c2w, a.ref_w2c = _lib.ptr(c2w), _lib.ptr(w2c)
    a.mvs_depth, a.mvs_uncert, a.mvs_normal = _lib.ptr(md), _lib.ptr(mu_), _lib.ptr(mn)
    a.map_h, a.map_w, a.img_h, a.img_w = md.shape[-2], md.shape[-1], ih, iw
    a.fill_rand, a.gauss, a.uniform_depth = _lib.ptr(fill_rand), _lib.ptr(gauss), _lib.ptr(uni)
    out = torch.empty(1, rn, ns + nu, device=dev, dtype=torch.float32)
    lik = torch.empty(rn, nc, device=dev, dtype=torch.float32) if return_likelihood else None
    a.out_depth, a.likelihood = _lib.ptr(out), _lib.ptr(lik)
    with torch.cuda.device(dev):
        rc = lib.pgrf_depth_guided_sample_fwd(a, _lib.stream_ptr())
    _lib.check(rc, "pgrf_depth_guided_sample_fwd")
    return (out, lik) if return_likelihood else out
