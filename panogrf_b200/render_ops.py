"""Drop-ins for the functional operators of the reference's `network/render_ops.py` (same names, argument meaning and
result shapes), each one CUDA kernel through the C ABI (csrc/render_ops.cu).  The fused renderer
(`panogrf_b200.renderer`) does not call these; they exist for callers of the reference's functional API and for
per-kernel measurements."""
import math

import torch

from . import _lib
from .renderer import coarse_depth_table, fine_u_table, to_channels_last


def _f32(t):
    return t.contiguous().float()


_CL_CACHE = {}


def _channels_last_cached(name, t, pad_to=None):
    """NCHW -> channels-last copy, reused while the caller keeps passing the same unmodified tensor."""
    key = (t.data_ptr(), t._version, tuple(t.shape), str(t.device))
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


def alpha_values2hit_prob(alpha_values):
    """render_ops.py:145-153: (...,dn) -> (...,dn); sequential fp32 cumprod (the stated accumulation order)."""
    _lib.require_cuda(alpha_values)
    lib = _lib.load()
    shape = alpha_values.shape
    a = _f32(alpha_values).reshape(-1, shape[-1])
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        rc = lib.pgrf_composite_fwd(None, _lib.ptr(a), None, None, 0, a.shape[0], a.shape[1], _lib.ptr(out), None, None,
                                    _lib.stream_ptr())
    _lib.check(rc, "pgrf_composite_fwd")
    return out.reshape(shape)


def composite(density, colors, depth):
    """network_rendering's tail (renderer.py:214-218) + render_depth (:302-304): density (qn,rn,dn), colors (qn,rn,dn,3),
    depth (qn,rn,dn) -> hit_prob (qn,rn,dn), pixel_colors (qn,rn,3), render_depth (qn,rn)."""
    _lib.require_cuda(density, colors, depth)
    lib = _lib.load()
    qn, rn, dn = density.shape
    d, c, z = _f32(density).reshape(-1, dn), _f32(colors).reshape(-1, dn, 3), _f32(depth).reshape(-1, dn)
    hit = torch.empty_like(d)
    pix = torch.empty(d.shape[0], 3, device=d.device)
    rd = torch.empty(d.shape[0], device=d.device)
    with torch.cuda.device(d.device):
        rc = lib.pgrf_composite_fwd(_lib.ptr(d), None, _lib.ptr(c), _lib.ptr(z), dn, d.shape[0], dn, _lib.ptr(hit), _lib.ptr(pix),
                                    _lib.ptr(rd), _lib.stream_ptr())
    _lib.check(rc, "pgrf_composite_fwd")
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
