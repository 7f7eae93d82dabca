"""Drop-in for the render-time hot path of the reference's `network/renderer.py`.

`NeuralRayBaseRenderer(cfg)` keeps the reference's constructor signature, config keys, method names
(`render`, `render_impl`, `render_by_depth`) and the `state_dict()` names of the hot-path
parameters (`[fine_]dist_decoder.*`, `[fine_]agg_net.*`), so a reference checkpoint loads with
`load_state_dict(..., strict=False)`.  The per-ray-batch python/ATen pipeline
(network/renderer.py:223-317, 435-524, 567-633, 635-686) is replaced by three persistent sm_100a
kernels per pass (csrc/render_kernels.cu) reached through the C ABI.

The per-call CNN encoders (SURVEY.md §8f "next"): `render()` takes `ref_imgs_info['img_feats']` and the already
vis-encoded `ref_imgs_info['ray_feats']`, or — when `ref_imgs_info` has no `img_feats` — runs the attached encoders like
network/renderer.py:639-642: `net.vis_encoder = panogrf_b200.vis_encoder.DefaultVisEncoder(cfg)` is the tensor-core drop-in
(its parameters then appear as `vis_encoder.*`, the reference's names), `net.image_encoder =
panogrf_b200.image_encoder.ResUNetLight(cfg, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=...)` the image encoder;
`init_net` (mono / MVS depth networks) stays a user-supplied callable.
Only the eval path (`is_train=False`, deterministic sampling, no autograd) is implemented.
"""
import ctypes

import math

import torch
import torch.nn as nn

from . import _lib
from .weights import pack_blob, pack_blob16


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's names / shapes / default initialisation
# ------------------------------------------------------------------------------------------------

def _seq(dims, act_slots=True):
    """Linear layers at even indices (0,2,4,...) like the reference's nn.Sequential stacks."""
    mods = []
    for i in range(len(dims) - 1):
        mods.append(nn.Linear(dims[i], dims[i + 1]))
        mods.append(nn.Identity())
    return nn.Sequential(*mods)


def _kaiming(module):
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.kaiming_normal_(m.weight.data)
            if m.bias is not None:
                nn.init.zeros_(m.bias.data)


class MixtureLogisticsDistDecoder(nn.Module):
    """Parameters of network/dist_decoder.py:53-97."""
    default_cfg = {"feats_dim": 32, "bias_val": 0.05, "use_vis": True}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        d = self.cfg["feats_dim"]
        if d != 32:
            raise _lib.PanoGRFError("panogrf_b200 kernels are built for feats_dim == 32")
        self.mean_decoder = _seq([d, d, d, 2])
        self.var_decoder = _seq([d, d, d, 2])
        self.aw_decoder = _seq([d, d, d, 1])
        if self.cfg["use_vis"]:
            self.vis_decoder = _seq([d, d, d, 1])

    def forward(self, feats):
        """dist_decoder.py:99-107: feats (...,32) -> prj_mean (...,2), prj_var (...,2), prj_vis (...,1) | None, prj_aw (...,1)."""
        lead = feats.shape[:-1]
        x = feats.detach().float().reshape(-1, 32)
        n0 = x.shape[0]
        n = (n0 + 3) // 4 * 4                                     # rows are independent here: any (rn, dn=4) factorisation
        feat = torch.zeros(1, n, 67, device=x.device)
        feat[0, :n0, :32] = x
        prj = torch.zeros(1, n, 6, device=x.device)
        prj[..., 2] = 1.0
        out = _module_pass(_module_blob(self, "dist_decoder.", self), 1, n // 4, 4, self.cfg["use_vis"], self.cfg["bias_val"],
                           prj=prj, feat=feat, stage_mask=1, want_dec=True)["dec"][0, :n0]
        mean, var = out[:, 0:2].reshape(*lead, 2), out[:, 2:4].reshape(*lead, 2)
        vis = out[:, 4:5].reshape(*lead, 1) if self.cfg["use_vis"] else None
        return mean, var, vis, out[:, 5:6].reshape(*lead, 1)

    def predict_mean(self, prj_ray_feats):
        """dist_decoder.py:146-148."""
        return self.forward(prj_ray_feats)[0]

    def predict_aw(self, prj_ray_feats):
        """dist_decoder.py:150-151."""
        return self.forward(prj_ray_feats)[3]

    def decode_alpha_value(self, alpha_value):
        """dist_decoder.py:142-144."""
        return torch.sigmoid(alpha_value)

    def compute_prob(self, depth, interval, mean, var, vis, aw, is_ref, depth_range):
        """dist_decoder.py:109-140.  is_ref=True (the render path): depth (rfn,qn,rn,dn), interval (1|rfn,qn,rn,dn),
        mean/var (rfn,qn,rn,dn,2), vis/aw (rfn,qn,rn,dn,1), depth_range (rfn,2) -> alpha, visibility, hit_prob (rfn,qn,rn,dn).
        is_ref=False (the query rays' own distribution, used by the training losses): depth, interval (qn,rn,dn), mean/var
        (qn,rn,1|dn,2), vis/aw (qn,rn,1|dn,1), depth_range (qn,2) -> (qn,rn,dn)."""
        _lib.require_cuda(depth, interval, mean, var, aw)
        lib = _lib.load()
        if not is_ref:
            # the query rays' own distribution: depth, interval (qn,rn,dn); mean / var (qn,rn,1|dn,2); vis / aw (qn,rn,1|dn,1)
            shape = depth.shape
            qn, dn = shape[0], shape[-1]
            n = depth[0].numel()
            fq = lambda t, c: t.detach().float().expand(*shape, c).reshape(qn, n, c).contiguous()
            d = depth.detach().float().reshape(qn, n).contiguous()
            iv = interval.detach().float().expand(*shape).reshape(qn, n).contiguous()
            m2, v2, a1 = fq(mean, 2), fq(var, 2), fq(aw, 1)
            vs = fq(vis, 1) if (self.cfg["use_vis"] and vis is not None) else None
            rng = depth_range.detach().float().contiguous().to(d.device)
            outs = [torch.empty(qn, n, device=d.device) for _ in range(3)]
            with torch.cuda.device(d.device):
                rc = lib.pgrf_compute_prob_que_fwd(_lib.ptr(d), _lib.ptr(iv), _lib.ptr(m2), _lib.ptr(v2), _lib.ptr(vs), _lib.ptr(a1),
                                                   _lib.ptr(rng), qn, n, dn, *[_lib.ptr(o) for o in outs], _lib.stream_ptr())
            _lib.check(rc, "pgrf_compute_prob_que_fwd")
            return tuple(o.reshape(shape) for o in outs)
        shape = depth.shape
        rfn, dn = shape[0], shape[-1]
        n = depth[0].numel()
        f = lambda t, c: t.detach().float().expand(*shape, c).reshape(rfn, n, c).contiguous()
        d = depth.detach().float().reshape(rfn, n).contiguous()
        per_view = interval.shape[0] == rfn and rfn > 1
        iv = interval.detach().float().expand(rfn if per_view else 1, *shape[1:]).reshape(-1).contiguous()
        m2, v2, a1 = f(mean, 2), f(var, 2), f(aw, 1)
        vs = f(vis, 1) if (self.cfg["use_vis"] and vis is not None) else None
        rng = depth_range.detach().float().contiguous().to(d.device)
        outs = [torch.empty(rfn, n, device=d.device) for _ in range(3)]
        with torch.cuda.device(d.device):
            rc = lib.pgrf_compute_prob_fwd(_lib.ptr(d), _lib.ptr(iv), int(per_view), _lib.ptr(m2), _lib.ptr(v2), _lib.ptr(vs),
                                           _lib.ptr(a1), _lib.ptr(rng), rfn, n, dn, *[_lib.ptr(o) for o in outs],
                                           _lib.stream_ptr())
        _lib.check(rc, "pgrf_compute_prob_fwd")
        return tuple(o.reshape(shape) for o in outs)


class _RayAttention(nn.Module):
    def __init__(self):
        super().__init__()
        self.w_qs = nn.Linear(16, 16, bias=False)
        self.w_ks = nn.Linear(16, 16, bias=False)
        self.w_vs = nn.Linear(16, 16, bias=False)
        self.fc = nn.Linear(16, 16, bias=False)
        self.layer_norm = nn.LayerNorm(16, eps=1e-6)


class IBRNetWithNeuRay(nn.Module):
    """Parameters of network/ibrnet.py:239-300."""

    def __init__(self, neuray_in_dim=32, in_feat_ch=32, n_samples=64):
        super().__init__()
        if neuray_in_dim != 32 or in_feat_ch != 32:
            raise _lib.PanoGRFError("panogrf_b200 kernels are built for neuray_dim == in_feat_ch == 32")
        self.n_samples = n_samples
        self.ray_dir_fc = _seq([4, 16, in_feat_ch + 3])
        self.base_fc = _seq([(in_feat_ch + 3) * 5 + neuray_in_dim, 64, 32])
        self.vis_fc = _seq([32, 32, 33])
        self.vis_fc2 = _seq([32, 32, 1])
        self.geometry_fc = _seq([32 * 2 + 1, 64, 16])
        self.ray_attention = _RayAttention()
        self.out_geometry_fc = _seq([16, 16, 1])
        self.rgb_fc = _seq([32 + 1 + 4, 16, 8, 1])
        self.neuray_fc = _seq([neuray_in_dim, 8, 1])
        for m in (self.base_fc, self.vis_fc2, self.vis_fc, self.geometry_fc, self.rgb_fc, self.neuray_fc):
            _kaiming(m)


class DefaultAggregationNet(nn.Module):
    """Parameters of network/aggregate_net.py:16-39."""
    default_cfg = {"sample_num": 64, "neuray_dim": 32, "use_img_feats": False}

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.default_cfg, **cfg}
        if self.cfg.get("level") in [-1]:
            raise _lib.PanoGRFError("level=-1 (16-channel image features) is not supported")
        dim = self.cfg["neuray_dim"]
        self.agg_impl = IBRNetWithNeuRay(dim, in_feat_ch=32, n_samples=self.cfg["sample_num"])
        self.prob_embed = nn.Sequential(nn.Linear(2 + 32, dim), nn.Identity(), nn.Linear(dim, dim))

    def _run(self, prj_dict, que_dir):
        rfn, qn, rn, dn, _ = prj_dict["hit_prob"].shape
        if dn != self.cfg["sample_num"]:
            raise RuntimeError(f"The size of tensor a ({dn}) must match the size of tensor b "
                               f"({self.cfg['sample_num']}) at non-singleton dimension 1")   # ibrnet.py:358
        n = qn * rn * dn
        dev = prj_dict["hit_prob"].device
        prj = torch.zeros(rfn, n, 6, device=dev)
        prj[..., 2] = 1.0
        prj[..., 3:6] = _rows(prj_dict["dir"], rfn, n, 3)
        feat = torch.cat([_rows(prj_dict["ray_feats"], rfn, n, 32), _rows(prj_dict["rgb"], rfn, n, 3),
                          _rows(prj_dict["img_feats"], rfn, n, 32)], -1).contiguous()
        alpha = prj_dict["alpha"] if "alpha" in prj_dict else torch.zeros_like(prj_dict["vis"])
        prob = torch.cat([_rows(alpha, rfn, n, 1), _rows(prj_dict["vis"], rfn, n, 1), _rows(prj_dict["hit_prob"], rfn, n, 1)],
                         -1).contiguous()
        qd = que_dir.detach().float().reshape(n, 3).contiguous()
        blob = _module_blob(self, "agg_net.", self, self.cfg["sample_num"])
        out = _module_pass(blob, rfn, qn * rn, dn, False, 0.05, prj=prj, feat=feat, prob=prob, que_dir=qd)
        return out, (qn, rn, dn)

    def forward(self, prj_dict, que_dir):
        """aggregate_net.py:41-89: prj_dict of (rfn,qn,rn,dn,*) tensors (ray_feats, hit_prob, vis, rgb, dir, img_feats),
        que_dir (qn,rn,dn,3) -> density (qn,rn,dn), colors (qn,rn,dn,3)."""
        out, (qn, rn, dn) = self._run(prj_dict, que_dir)
        return out["density"].reshape(qn, rn, dn), out["colors"].reshape(qn, rn, dn, 3)


# ------------------------------------------------------------------------------------------------
# module-level API of the reference (dist decoder / aggregation net evaluated on a caller-provided prj_dict):
# the fp32 kernels of the fused path with their producers switched off (pgrf_render_args.prj_in / feat_in / prob_in)
# ------------------------------------------------------------------------------------------------

def _module_blob(module, prefix, cache_owner, n_samples=64):
    """fp32 blob holding only `module`'s parameters under the reference prefix (`dist_decoder.` / `agg_net.`)."""
    params = list(module.parameters())
    dev = params[0].device
    key = (str(dev), tuple((p.data_ptr(), p._version) for p in params))
    hit = getattr(cache_owner, "_mod_blob", None)
    if hit is None or hit[0] != key:
        state = {prefix + k: v for k, v in module.state_dict().items()}
        cache_owner._mod_blob = (key, pack_blob(state, False, n_samples, dev, allow_missing=True))
    return cache_owner._mod_blob[1]


def _rows(t, rfn, n, c):
    return t.detach().float().reshape(rfn, n, c).contiguous()


def _module_pass(blob, rfn, rn, dn, use_vis, bias_val, *, prj, feat, prob=None, que_dir=None, interval=None, ref_range=None,
                 stage_mask=7, want_dec=False, want_prob=False):
    """One pgrf_render_pass_fwd / pgrf_agg_mlp_fwd call on per-row inputs. prj (rfn,n,6), feat (rfn,n,67), prob (rfn,n,3),
    que_dir (n,3), interval (n); n = rn*dn. Returns a dict of outputs."""
    lib = _lib.load()
    dev = feat.device
    _lib.require_cuda(prj, feat, prob, que_dir, interval)
    n = rn * dn
    if not (1 <= rfn <= 4) or not (3 <= dn <= 128):
        raise _lib.PanoGRFError(f"module-level kernels need 1..4 views and 3..128 samples per ray (got rfn={rfn}, dn={dn})")
    a = _lib.RenderArgs()
    a.dataset, a.H, a.W = 0, 2, 2
    a.rfn, a.rn, a.dn = rfn, rn, dn
    a.use_vis, a.bias_val = int(bool(use_vis)), float(bias_val)
    a.img_h = a.img_w = a.if_h = a.if_w = a.rf_h = a.rf_w = 2
    a.que_near, a.que_far = 1.0, 2.0
    e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    if que_dir is None:
        que_dir = torch.zeros(n, 3, device=dev)
    if interval is None and prob is None:
        interval = torch.zeros(n, device=dev)
    if ref_range is None:
        ref_range = torch.tensor([[1.0, 2.0]], device=dev).repeat(rfn, 1)
    ref_range = ref_range.detach().float().contiguous().to(dev)
    a.prj_in, a.feat_in, a.que_dir_in = _lib.ptr(prj), _lib.ptr(feat), _lib.ptr(que_dir)
    a.prob_in, a.interval_in, a.ref_depth_range = _lib.ptr(prob), _lib.ptr(interval), _lib.ptr(ref_range)
    a.weights = _lib.ptr(blob)
    f1n, f2n = ctypes.c_longlong(), ctypes.c_longlong()
    _lib.check(lib.pgrf_render_workspace(rfn, n, ctypes.byref(f1n), ctypes.byref(f2n)), "pgrf_render_workspace")
    f1, f2 = e(f1n.value), e(f2n.value)
    a.f1, a.f2 = _lib.ptr(f1), _lib.ptr(f2)
    out = {"pixel_colors": e(rn, 3), "hit_prob": e(rn, dn), "density": e(rn, dn), "colors": e(rn, dn, 3)}
    a.pixel_colors, a.hit_prob = _lib.ptr(out["pixel_colors"]), _lib.ptr(out["hit_prob"])
    a.density, a.colors = _lib.ptr(out["density"]), _lib.ptr(out["colors"])
    if want_dec:
        out["dec"] = e(rfn, n, 6)
        a.dec_dbg = _lib.ptr(out["dec"])
    if want_prob:
        out["prob"] = e(rfn, n, 3)
        a.prob_dbg = _lib.ptr(out["prob"])
    a.stage_mask = stage_mask
    with torch.cuda.device(dev):
        if stage_mask == 7 and prob is not None:
            rc = lib.pgrf_agg_mlp_fwd(ctypes.byref(a), _lib.stream_ptr())
        else:
            rc = lib.pgrf_render_pass_fwd(ctypes.byref(a), _lib.stream_ptr())
    _lib.check(rc, "pgrf_render_pass_fwd")
    return out


name2dist_decoder = {"mixture_logistics": MixtureLogisticsDistDecoder}
name2agg_net = {"default": DefaultAggregationNet}


# ------------------------------------------------------------------------------------------------
# host-side tables that must be bit-identical to the reference's (tiny, computed once on the CPU)
# ------------------------------------------------------------------------------------------------

def coarse_depth_table(cfg, sample_num, use_disp):
    """sample_depth with random_sample=False (network/render_ops.py:292-339): one (dn,) table —
    the deterministic samples are identical for every ray."""
    near = torch.ones(1) * cfg["min_depth"]
    far = torch.ones(1) * cfg["max_depth"]
    dn = sample_num
    assert dn > 2
    val = torch.arange(1, dn - 1, dtype=torch.float32)[None, None, :] + torch.zeros(1, 1, dn - 2)
    if not use_disp:
        interval = (far - near) / (dn - 1)
        ticks = interval[:, None, None] * val
        diff = far - near
        ticks = torch.cat([torch.zeros(1, 1, 1), ticks, diff[:, None, None]], -1)
        return (near[:, None, None] + ticks).reshape(-1)
    interval = (1 / far - 1 / near) / (dn - 1)
    ticks = interval[:, None, None] * val
    diff = 1 / far - 1 / near
    ticks = torch.cat([torch.zeros(1, 1, 1), ticks, diff[:, None, None]], -1)
    return (1 / (1 / near[:, None, None] + ticks)).reshape(-1)


def fine_u_table(fdn):
    """render_ops.py:442-445."""
    interval = 1 / fdn
    return (0.5 * interval + torch.arange(fdn) * interval).float()


def to_channels_last(x, pad_to=None):
    """(N,C,H,W) -> contiguous (N,H,W,C[+pad]) with the library's tiled transpose kernel (pgrf_nchw_to_nhwc: one launch per map,
    coalesced on both sides; the torch permute + cat + contiguous chain it replaces cost 19 ms for the three maps of a
    512x1024 view)."""
    x = x.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    _lib.require_cuda(x)
    n, c, h, w = x.shape
    cpad = max(c, int(pad_to or 0))
    out = torch.empty((n, h, w, cpad), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().pgrf_nchw_to_nhwc(_lib.ptr(x), _lib.ptr(out), n, c, h, w, cpad, _lib.stream_ptr())
    _lib.check(rc, "pgrf_nchw_to_nhwc")
    return out


def perspective_ray_dirs(que_imgs_info):
    """coords2rays (network/render_ops.py:37-59) for ONE query view: pixel coords (1,rn,2) + world->camera `poses` (1,3,4) + `Ks`
    (1,3,3) -> (c2w (3,4), directions (rn,3)) with the reference's op order (K^-1 @ [x,y,1], rotate + translate, subtract the
    centre).  The directions are NOT normalised, exactly like the reference's."""
    coords, poses, Ks = que_imgs_info["coords"].float(), que_imgs_info["poses"].float(), que_imgs_info["Ks"].float()
    assert poses.shape[0] == 1, "que_imgs_info poses.shape[0]=1"
    rot = poses[:, :, :3].unsqueeze(1).permute(0, 1, 3, 2)
    trans = -rot @ poses[:, :, 3:].unsqueeze(1)
    rfn, rn, _ = coords.shape
    centers = trans.repeat(1, rn, 1, 1).squeeze(-1)
    hom = torch.cat([coords, torch.ones([rfn, rn, 1], dtype=torch.float32, device=coords.device)], 2)
    cam_xyz = torch.inverse(Ks).unsqueeze(1) @ hom.unsqueeze(3)
    cam_xyz = rot @ cam_xyz + trans
    directions = cam_xyz.squeeze(3) - centers
    c2w = torch.cat([rot[0, 0], trans[0, 0]], 1)                      # (3,4): [R^T | -R^T t]
    return c2w.contiguous(), directions[0].contiguous()


def tensor_version(t):
    """In-place version counter of a tensor, or None when it is not tracked (tensors created under torch.inference_mode):
    such tensors are never cached."""
    try:
        return t._version
    except RuntimeError:
        return None


class NeuralRayBaseRenderer(nn.Module):
    base_cfg = {
        "dist_decoder_type": "mixture_logistics", "dist_decoder_cfg": {},
        "agg_net_type": "default", "agg_net_cfg": {},
        "use_hierarchical_sampling": False, "fine_agg_net_cfg": {}, "fine_dist_decoder_cfg": {},
        "fine_depth_sample_num": 64, "fine_depth_use_all": False,
        "ray_batch_num": 2048, "depth_sample_num": 64,
        "use_ray_mask": True, "ray_mask_view_num": 1, "ray_mask_point_num": 8,
        "render_depth": False, "render_uncert": False, "debug": False, "use_disp": True,
    }
    #: rays per kernel launch (the reference's ray_batch_num only bounds ITS activation memory; here it bounds the
    #: inter-kernel workspaces).  None = 524288 for the bf16 path (only the 160 B/sample F2 tiles exist: 5.4 GB at 64 samples; larger
    #: launches amortise the per-CTA weight load and the tail: 46.63 / 46.41 / 46.27 ms per view at 131072 / 262144 / 524288), 32768 for the fp32 path (F1 tiles: 38.9 KB per 64 samples, 1.3 GB of workspace).
    rays_per_launch = None
    #: reuse the channels-last copies of the source maps while the caller passes the same, unmodified tensors (keyed on storage
    #: pointer + in-place version).  Writes through `.data` do not bump the version: set False (or call invalidate_map_cache())
    #: if the maps are updated that way.
    cache_maps = True

    def __init__(self, cfg):
        super().__init__()
        self.cfg = {**self.base_cfg, **cfg}
        for k in ("agg_net_cfg", "fine_agg_net_cfg", "dist_decoder_cfg", "fine_dist_decoder_cfg"):
            self.cfg[k] = dict(self.cfg.get(k) or {})
        for key in ("sample_num", "level", "wo_geometry", "wo_appearance"):      # renderer.py:67-78
            if key in cfg:
                self.cfg["agg_net_cfg"][key] = self.cfg[key]
                self.cfg["fine_agg_net_cfg"][key] = self.cfg[key]
        if self.cfg["dataset_name"] not in _lib.DATASET_IDS:
            raise Exception(f"unknown dataset_name {self.cfg['dataset_name']!r}")
        if self.cfg.get("debug"):
            raise _lib.PanoGRFError("cfg['debug'] (network bypass) is not part of the hot path")
        self.dist_decoder = name2dist_decoder[self.cfg["dist_decoder_type"]](self.cfg["dist_decoder_cfg"])
        self.agg_net = name2agg_net[self.cfg["agg_net_type"]](self.cfg["agg_net_cfg"])
        if self.cfg["use_hierarchical_sampling"] and not self.cfg.get("one_mlp"):
            self.fine_dist_decoder = name2dist_decoder[self.cfg["dist_decoder_type"]](self.cfg["fine_dist_decoder_cfg"])
            self.fine_agg_net = name2agg_net[self.cfg["agg_net_type"]](self.cfg["fine_agg_net_cfg"])
        self.image_encoder = None    # optional user-supplied encoders (out of scope, see module docstring)
        self.vis_encoder = None
        self._blob_cache = {}
        self._param_lists = {}
        self._key_memo = None          # per-call memo of _param_key (set by _render_view)
        self._ws = {}
        self._cl_cache = {}
        self._tables = {}
        #: "fp32": SIMT parity path (rtol 1e-4); "bf16": tcgen05 tensor-core MLP (bf16 operands, fp32 accumulate, rtol 1e-2)
        self.mlp_dtype = str(self.cfg.get("mlp_dtype", "fp32"))
        if self.mlp_dtype not in ("fp32", "bf16"):
            raise _lib.PanoGRFError(f"mlp_dtype must be 'fp32' or 'bf16', got {self.mlp_dtype!r}")

    # ---- weights --------------------------------------------------------------------------------
    def _param_key(self, fine, device):
        """Cheap identity of the hot-path parameters of one net pair: (storage pointer, in-place version) of every tensor.
        The parameter LIST is collected once (module traversal costs ~0.25 ms per call, as much as 3 % of a 64-row shard);
        in-place updates, load_state_dict and .to()/.cuda() are seen through version / pointer; after REPLACING a Parameter
        object call invalidate_weight_cache()."""
        memo = self._key_memo
        if memo is not None and (fine, str(device)) in memo:
            return memo[(fine, str(device))]
        lst = self._param_lists.get(fine)
        if lst is None:
            pre = ("fine_dist_decoder.", "fine_agg_net.") if fine else ("dist_decoder.", "agg_net.")
            lst = self._param_lists[fine] = [p for n, p in self.named_parameters() if n.startswith(pre)]
        key = (fine, str(device), tuple((p.data_ptr(), p._version) for p in lst))
        if memo is not None:
            memo[(fine, str(device))] = key
        return key

    def invalidate_weight_cache(self):
        """Drop the packed weight blobs.  REQUIRED after updating parameters through `.data` (p.data.copy_, EMA updates,
        nn.init.*(p.data)): such writes do not bump the version counter the cache key uses.  Optimizer steps, load_state_dict
        and .to()/.cuda() are detected automatically."""
        self._param_lists.clear()
        self._blob_cache.clear()

    def invalidate_map_cache(self):
        self._cl_cache.clear()

    def _blob(self, fine, device):
        key = self._param_key(fine, device)
        hit = self._blob_cache.get(fine)
        if hit is None or hit[0] != key:
            agg = self.fine_agg_net if fine else self.agg_net
            blob = pack_blob(self.state_dict(), fine, agg.cfg["sample_num"], device)
            self._blob_cache[fine] = (key, blob)
        return self._blob_cache[fine][1]

    def _blob16(self, fine, device):
        key = self._param_key(fine, device)
        hit = self._blob_cache.get(("w16", fine))
        if hit is None or hit[0] != key:
            self._blob_cache[("w16", fine)] = (key, pack_blob16(self.state_dict(), fine, device))
        return self._blob_cache[("w16", fine)][1]

    # ---- one pass -------------------------------------------------------------------------------
    def _pass(self, ctx, coords, depth, depth_stride, fine_net, want_fine, outs, r0, keep_hit):
        """Launch rows/samples/rays kernels for `coords` (rn,2) with sample depths `depth`."""
        cfg, lib = self.cfg, _lib.load()
        rn = coords.shape[0]
        dn = depth.shape[-1]
        dev = coords.device
        agg = self.fine_agg_net if fine_net else self.agg_net
        if agg.cfg["sample_num"] != dn:
            raise RuntimeError(f"The size of tensor a ({dn}) must match the size of tensor b "
                               f"({agg.cfg['sample_num']}) at non-singleton dimension 1")   # ibrnet.py:358
        a = _lib.RenderArgs()
        a.dataset = _lib.DATASET_IDS[cfg["dataset_name"]]
        a.H, a.W = int(cfg["height"]), int(cfg["width"])
        a.rfn, a.rn, a.dn = ctx["rfn"], rn, dn
        a.use_vis = int(bool(self.dist_decoder.cfg["use_vis"]))
        a.bias_val = float((self.fine_dist_decoder if fine_net else self.dist_decoder).cfg["bias_val"])
        a.coords, a.depth, a.depth_ray_stride = _lib.ptr(coords), _lib.ptr(depth), depth_stride
        if ctx.get("ray_dirs") is not None:
            a.ray_dirs = _lib.ptr(ctx["ray_dirs"][r0:r0 + rn])
        a.que_c2w, a.que_near, a.que_far = _lib.ptr(ctx["c2w"]), ctx["que_near"], ctx["que_far"]
        a.ref_w2c, a.ref_depth_range = _lib.ptr(ctx["w2c"]), _lib.ptr(ctx["ref_range"])
        a.imgs_cl, a.img_h, a.img_w = _lib.ptr(ctx["imgs"]), ctx["imgs"].shape[1], ctx["imgs"].shape[2]
        a.img_feats_cl, a.if_h, a.if_w = _lib.ptr(ctx["img_feats"]), ctx["img_feats"].shape[1], ctx["img_feats"].shape[2]
        a.ray_feats_cl, a.rf_h, a.rf_w = _lib.ptr(ctx["ray_feats"]), ctx["ray_feats"].shape[1], ctx["ray_feats"].shape[2]
        a.weights = _lib.ptr(self._blob(fine_net, dev))
        if self.mlp_dtype == "bf16":
            a.mlp_bf16, a.weights16 = 1, _lib.ptr(self._blob16(fine_net, dev))
            a.sched = _lib.ptr(self._sched(dev))
        f1n, f2n = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(lib.pgrf_render_workspace(a.rfn, rn * dn, ctypes.byref(f1n), ctypes.byref(f2n)), "pgrf_render_workspace")
        ws = ctx["ws"]
        if ws.get("f1") is None or ws["f1"].numel() < f1n.value:
            ws["f1"] = torch.empty(f1n.value, device=dev, dtype=torch.float32)
        if ws.get("f2") is None or ws["f2"].numel() < f2n.value:
            ws["f2"] = torch.empty(f2n.value, device=dev, dtype=torch.float32)
        a.f1, a.f2 = _lib.ptr(ws["f1"]), _lib.ptr(ws["f2"])
        sl = slice(r0, r0 + rn)
        a.pixel_colors = _lib.ptr(outs["pixel_colors_nr"][0, sl])
        a.render_depth = _lib.ptr(outs["render_depth"][0, sl]) if "render_depth" in outs else None
        a.density = _lib.ptr(outs["density_nr"][0, sl])
        a.colors = _lib.ptr(outs["colors_nr"][0, sl])
        hit = None
        if keep_hit:
            a.hit_prob = _lib.ptr(outs["hit_prob_nr"][0, sl])
        fine_depth = None
        if want_fine:
            fdn = int(cfg["fine_depth_sample_num"])
            use_all = bool(cfg["fine_depth_use_all"])
            fine_depth = torch.empty(rn, fdn + (dn if use_all else 0), device=dev, dtype=torch.float32)
            a.fine_depth, a.fine_dn, a.fine_u = _lib.ptr(fine_depth), fdn, _lib.ptr(ctx["fine_u"])
            a.fine_use_all, a.use_disp = int(use_all), int(bool(cfg["use_disp"]))
        for k in ("prob_dbg", "prj_dbg", "feat_dbg", "fine_inds"):
            if ctx.get(k) is not None:
                setattr(a, k, _lib.ptr(ctx[k]))
        a.stage_mask = int(ctx.get("stage_mask", 0))
        a.wo_geometry, a.wo_appearance = int(bool(agg.cfg.get("wo_geometry"))), int(bool(agg.cfg.get("wo_appearance")))
        with torch.cuda.device(dev):
            rc = lib.pgrf_render_pass_fwd(ctypes.byref(a), _lib.stream_ptr())
        _lib.check(rc, "pgrf_render_pass_fwd")
        return fine_depth

    def _context(self, que_imgs_info, ref_imgs_info, is_perspec=False):
        ctx = self._context_erp(que_imgs_info, ref_imgs_info, is_perspec)
        if is_perspec:      # perspective / cube query rays (render_ops.py:61-74): explicit world-space directions
            c2w, dirs = perspective_ray_dirs(que_imgs_info)
            ctx["c2w"], ctx["ray_dirs"] = c2w.to(ctx["w2c"].device), dirs.to(ctx["w2c"].device)
        return ctx

    def _context_erp(self, que_imgs_info, ref_imgs_info, is_perspec=False):
        imgs = ref_imgs_info["imgs"]
        _lib.require_cuda(imgs, ref_imgs_info["ray_feats"], ref_imgs_info["img_feats"], que_imgs_info["coords"])
        dev = imgs.device
        c2w = que_imgs_info["c2w"] if not is_perspec else que_imgs_info["poses"]
        assert c2w.shape[0] == 1, "que_imgs_info c2w.shape[0]=1"                 # render_ops.py:89
        if ref_imgs_info["ray_feats"].shape[1] != 32 or ref_imgs_info["img_feats"].shape[1] != 32:
            raise _lib.PanoGRFError("ray_feats / img_feats must have 32 channels")
        dr = self._cached_scalars("que_range", que_imgs_info["depth_range"])
        return {
            "rfn": imgs.shape[0],
            "imgs": self._cached_cl("imgs", imgs, 4),
            "img_feats": self._cached_cl("img_feats", ref_imgs_info["img_feats"], None),
            "ray_feats": self._cached_cl("ray_feats", ref_imgs_info["ray_feats"], None),
            "w2c": ref_imgs_info["w2c"].float().contiguous().to(dev),
            "ref_range": ref_imgs_info["depth_range"].float().contiguous().to(dev),
            "c2w": c2w.float().reshape(3, 4).contiguous().to(dev),
            "que_near": dr[0], "que_far": dr[1],
            "fine_u": self._cached_table(("fine_u", int(self.cfg["fine_depth_sample_num"])), dev,
                                         lambda: fine_u_table(int(self.cfg["fine_depth_sample_num"]))),
            "ws": {},
        }

    def _cached_scalars(self, name, t):
        """(near, far) of a (1,2) depth_range tensor as python floats without a device sync on every call."""
        ver = tensor_version(t)
        key = (t.data_ptr(), ver, str(t.device))
        hit = self._cl_cache.get(name)
        if ver is None or hit is None or hit[0] != key:
            v = t.detach().float().cpu()
            self._cl_cache[name] = (key, (float(v[0, 0]), float(v[0, 1])), t)
        return self._cl_cache[name][1]

    def _const_mask(self, value, rn, dev):
        """(1,rn) bool mask of one value as an expanded view of a cached scalar (the ERP path's ray_mask is a constant,
        renderer.py:289-293): no fill kernel per call."""
        k = ("mask", value, str(dev))
        if k not in self._tables:
            self._tables[k] = torch.full((1, 1), value, device=dev)
        return self._tables[k].expand(1, rn)

    def _cached_table(self, key, dev, make):
        k = (key, str(dev))
        if k not in self._tables:
            self._tables[k] = make().to(dev)
        return self._tables[k]

    @staticmethod
    def _ws_key(dev):
        """Workspaces (inter-kernel tiles, tile counters) are per device AND per stream: two renders enqueued on different
        streams of one device must not share them."""
        return (str(dev), torch.cuda.current_stream(dev).cuda_stream)

    def _sched(self, dev):
        ws = self._ws.setdefault(self._ws_key(dev), {})
        if ws.get("sched") is None:
            ws["sched"] = torch.zeros(4, device=dev, dtype=torch.int32)
        return ws["sched"]

    def _cached_cl(self, name, t, pad_to):
        """Channels-last copy of a source map, reused while the caller keeps passing the same (unmodified) tensor —
        the source panoramas do not change between the query poses of a video render (render.py:249-291)."""
        ver = tensor_version(t)
        if ver is None or not self.cache_maps:
            return to_channels_last(t, pad_to)              # untracked (inference-mode) tensor or caching switched off
        key = (t.data_ptr(), ver, tuple(t.shape), str(t.device))
        hit = self._cl_cache.get(name)
        if hit is None or hit[0] != key:
            self._cl_cache[name] = (key, to_channels_last(t, pad_to), t)      # keep `t` alive so data_ptr stays unique
        return self._cl_cache[name][1]

    # ---- reference API --------------------------------------------------------------------------
    def render_impl(self, que_imgs_info, ref_imgs_info, is_train, is_perspec=False, _ctx=None, _outs=None, _r0=0,
                    keep_hit_prob=False):
        """network/renderer.py:567-633 (default, non-diner branch) for the rays in que_imgs_info['coords']."""
        if is_train:
            raise NotImplementedError("panogrf_b200 renderer: only the eval path (is_train=False) is implemented")
        cfg = self.cfg
        if cfg.get("diner_depth_guided_sampling", False):
            if is_perspec:
                raise NotImplementedError("depth-guided placement is implemented for ERP query rays only")
            return self._render_diner(que_imgs_info, ref_imgs_info, keep_hit_prob=True)
        ctx = _ctx or self._context(que_imgs_info, ref_imgs_info, is_perspec)
        coords = que_imgs_info["coords"]
        assert coords.shape[0] == 1
        coords2 = coords[0].float().contiguous()
        rn = coords2.shape[0]
        dev = coords2.device
        dn = int(cfg["depth_sample_num"])
        hier = bool(cfg["use_hierarchical_sampling"])
        own = _outs is None
        if _outs is None:
            _outs = self._alloc_outputs(rn, dev, keep_hit_prob or hier or self._needs_post(), ctx['rfn'])
            _r0 = 0
        depth_table = coarse_depth_table(cfg, dn, cfg["use_disp"]).to(dev)
        coarse = {k: v for k, v in _outs.items() if not k.endswith("_fine")}
        fine_depth = self._pass(ctx, coords2, depth_table, 0, False, hier, coarse, _r0, "hit_prob_nr" in coarse)
        if hier and que_imgs_info.get("ft_depth_range") is not None:
            self._ft_range_samples(que_imgs_info["ft_depth_range"], depth_table, fine_depth)
        if hier:
            fine = {k[:-5]: v for k, v in _outs.items() if k.endswith("_fine")}
            self._pass(ctx, coords2, fine_depth, fine_depth.shape[1], not cfg.get("one_mlp", False), False, fine, _r0,
                       "hit_prob_nr" in fine)
            if "que_depth_fine" in _outs:
                _outs["que_depth_fine"][0, _r0:_r0 + rn] = fine_depth
        if own:
            self._post(_outs, depth_table, True)
            self._add_gt(_outs, que_imgs_info, self._gt_suffixes())
        return _outs

    def _ft_range_samples(self, ft_depth_range, depth_table, fine_depth, coarse_stride=0, dn=None):
        """fine_render_impl (network/renderer.py:438-456, eval): rays with a valid depth prior (ft_depth_range[...,0] >= min_depth) take
        sample_3sigma between ft_depth_range[...,1] and [...,2] (network/sample_utils.py:6-60) instead of the inverse-CDF samples; rows
        of `fine_depth` (rn, fdn [+ dn]) are rewritten in place, sorted (with the coarse depths when fine_depth_use_all).  The coarse
        depths are one shared table (`coarse_stride` 0) or one row per ray (the depth-prior branch)."""
        cfg, lib = self.cfg, _lib.load()
        rn, dev = fine_depth.shape[0], fine_depth.device
        dn, fdn = int(cfg["depth_sample_num"]) if dn is None else int(dn), int(cfg["fine_depth_sample_num"])
        if fdn != dn:      # the reference writes both kinds of rows into empty_like(coarse depth)
            raise RuntimeError(f"shape mismatch: ft_depth_range needs fine_depth_sample_num ({fdn}) == depth_sample_num ({dn})")
        ft = ft_depth_range.reshape(-1, ft_depth_range.shape[-1]).float().contiguous()
        assert ft.shape[0] == rn and ft.shape[1] >= 3, "ft_depth_range must be (1, rn, 3)"
        t, g = self._cached_table(("3sigma_t", dn), dev, lambda: torch.linspace(0., 1., steps=dn)), \
            self._cached_table(("3sigma_g", dn), dev, lambda: 1. / math.sqrt(2 * math.pi) * torch.exp(-0.5 * torch.linspace(-3., 3., steps=dn - 1).pow(2)))
        use_all = bool(cfg["fine_depth_use_all"])
        with torch.cuda.device(dev):
            rc = lib.pgrf_sample_3sigma_fwd(_lib.ptr(ft), ft.shape[1], float(cfg["min_depth"]), _lib.ptr(t), _lib.ptr(g), dn,
                                            float(cfg["min_depth"]), float(cfg["max_depth"]), _lib.ptr(depth_table) if use_all else None,
                                            int(coarse_stride), dn if use_all else 0, 1, rn, _lib.ptr(fine_depth), _lib.stream_ptr())
        _lib.check(rc, "pgrf_sample_3sigma_fwd")

    def _add_gt(self, outs, que_imgs_info, suffixes=("",)):
        """`pixel_colors_gt[_fine]` (+ `polar_weights`) of the reference's output dict (renderer.py:278-286, 398-405): the
        query image sampled at the ray coordinates, interpolate_feats(align_corners=True).  network/metrics.py reads them."""
        if "imgs" not in que_imgs_info and "cube_imgs" not in que_imgs_info:
            return outs
        from .render_ops import interpolate_feature_map
        src = que_imgs_info["imgs"] if "imgs" in que_imgs_info else que_imgs_info["cube_imgs"]
        coords = que_imgs_info["coords"]
        gt = interpolate_feature_map(src, coords, src.shape[2], src.shape[3])          # full resolution -> align_corners=True
        pw = None
        if "imgs" in que_imgs_info and self.cfg.get("use_polar_weighted_loss"):
            m = que_imgs_info["polar_weights"]
            pw = interpolate_feature_map(m, coords, m.shape[2], m.shape[3])
        for sfx in suffixes:
            outs["pixel_colors_gt" + sfx] = gt
            if pw is not None:
                outs["polar_weights" + sfx] = pw
        return outs

    def _gt_suffixes(self):
        return ("", "_fine") if self.cfg["use_hierarchical_sampling"] else ("",)

    def _needs_post(self):
        return (bool(self.cfg.get("render_uncert")) or bool(self.cfg.get("perpoint_loss"))
                or bool(self.cfg.get("render_c2f_all") and self.cfg["use_hierarchical_sampling"]))

    def _post(self, outs, depth_table, keep_hit_prob):
        """Optional outputs assembled from the per-sample results: `render_c2f_all` re-composites the coarse and the
        fine samples together (renderer.py:484-521), `render_uncert` is the depth variance under hit_prob (:299-301)."""
        cfg = self.cfg
        if not self._needs_post():
            return outs
        from .render_ops import composite
        z_c = depth_table.view(1, 1, -1).expand_as(outs["hit_prob_nr"])
        unc = lambda z, d, h: ((z - d.unsqueeze(-1)).pow(2) * h).sum(-1) + 1e-5
        if cfg.get("render_uncert"):
            if "render_depth" not in outs:
                raise KeyError("render_depth")                       # the reference reads outputs['render_depth'] (:300)
            outs["render_uncert"] = unc(z_c, outs["render_depth"], outs["hit_prob_nr"])
        if cfg["use_hierarchical_sampling"]:
            z_f = outs["que_depth_fine"]
            if cfg.get("render_c2f_all"):
                z_f, idx = torch.cat([z_c, z_f], 2).sort()
                col = torch.gather(torch.cat([outs["colors_nr"], outs["colors_nr_fine"]], 2), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3))
                den = torch.gather(torch.cat([outs["density_nr"], outs["density_nr_fine"]], 2), 2, idx)
                hit, pix, rdepth = composite(den, col, z_f)
                outs.update({"pixel_colors_nr_fine": pix, "hit_prob_nr_fine": hit, "colors_nr_fine": col, "density_nr_fine": den})
                if "render_depth_fine" in outs:
                    outs["render_depth_fine"] = rdepth
            if cfg.get("render_uncert"):
                outs["render_uncert_fine"] = unc(z_f, outs["render_depth_fine"], outs["hit_prob_nr_fine"])
        if cfg.get("perpoint_loss"):                                  # per-sample weights and depths for the losses (renderer.py:314-316)
            outs["render_weights"], outs["render_dvals"] = outs["hit_prob_nr"], z_c
            if cfg["use_hierarchical_sampling"]:
                outs["render_weights_fine"], outs["render_dvals_fine"] = outs["hit_prob_nr_fine"], z_f
        if not keep_hit_prob:
            for k in ("hit_prob_nr", "hit_prob_nr_fine", "que_depth_fine"):
                outs.pop(k, None)
        return outs

    def _alloc_outputs(self, rn, dev, keep_hit_prob, rfn):
        cfg = self.cfg
        dn = int(cfg["depth_sample_num"])
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        outs = {"pixel_colors_nr": e(1, rn, 3), "colors_nr": e(1, rn, dn, 3), "density_nr": e(1, rn, dn)}
        if keep_hit_prob:
            outs["hit_prob_nr"] = e(1, rn, dn)
        if cfg["use_ray_mask"]:
            # every projection is "valid" in the ERP path: mask of ones (renderer.py:289-293)
            views_ok = rfn >= cfg["ray_mask_view_num"]
            outs["ray_mask"] = self._const_mask(bool(views_ok and dn > cfg["ray_mask_point_num"]), rn, dev)
        if cfg["render_depth"]:
            outs["render_depth"] = e(1, rn)
        if cfg["use_hierarchical_sampling"]:
            fdn = int(cfg["fine_depth_sample_num"]) + (dn if cfg["fine_depth_use_all"] else 0)
            outs.update({"pixel_colors_nr_fine": e(1, rn, 3), "colors_nr_fine": e(1, rn, fdn, 3),
                         "density_nr_fine": e(1, rn, fdn)})
            if keep_hit_prob:
                outs["hit_prob_nr_fine"] = e(1, rn, fdn)
                outs["que_depth_fine"] = e(1, rn, fdn)
            if cfg["use_ray_mask"]:
                outs["ray_mask_fine"] = self._const_mask(bool(views_ok and fdn > cfg["ray_mask_point_num"]), rn, dev)
            if cfg["render_depth"]:
                outs["render_depth_fine"] = e(1, rn)
        return outs

    def render(self, que_imgs_info, ref_imgs_info, is_train, is_perspec=False, keep_hit_prob=False):
        """network/renderer.py:635-686: all rays of the query view; returns the reference's output dict."""
        if is_train:
            raise NotImplementedError("panogrf_b200 renderer: only the eval path (is_train=False) is implemented")
        ref_imgs_info = dict(ref_imgs_info)
        if "img_feats" not in ref_imgs_info:
            if self.image_encoder is None or self.vis_encoder is None:
                raise _lib.PanoGRFError(
                    "ref_imgs_info has no 'img_feats': attach image_encoder/vis_encoder callables or pass "
                    "pre-encoded maps (the CNN encoders are outside the hot path)")
            feats = self.image_encoder(ref_imgs_info["imgs"])
            ref_imgs_info["img_feats"] = feats
            ref_imgs_info["ray_feats"] = self.vis_encoder(ref_imgs_info["ray_feats"], feats)
        if self.cfg.get("diner_depth_guided_sampling", False):
            if is_perspec:
                raise NotImplementedError("depth-guided placement is implemented for ERP query rays only")
            return self._render_diner(que_imgs_info, ref_imgs_info, keep_hit_prob)
        ctx = self._context(que_imgs_info, ref_imgs_info, is_perspec)
        coords = que_imgs_info["coords"]
        assert coords.shape[0] == 1
        rn = coords.shape[1]
        ft = que_imgs_info.get("ft_depth_range")
        two_pass = ft is not None and bool(self.cfg["use_hierarchical_sampling"])
        outs = self._alloc_outputs(rn, coords.device, keep_hit_prob or self._needs_post(), ctx['rfn'])
        if two_pass:
            # per-ray depth priors change the fine samples between the passes (renderer.py:438-456): the ray-batch loop runs here,
            # two kernel passes per batch with the prior-guided samples written in between
            step = int(self.cfg.get("fused_ray_batch", 524288))
            for r0 in range(0, rn, step):
                q = dict(que_imgs_info)
                q["coords"], q["ft_depth_range"] = coords[:, r0:r0 + step], ft[:, r0:r0 + step]
                self.render_impl(q, ref_imgs_info, False, is_perspec, _ctx=ctx, _outs=outs, _r0=r0)
        else:
            self._render_view(ctx, coords[0].float().contiguous(), outs)
        if self._needs_post():
            dn = int(self.cfg["depth_sample_num"])
            table = self._cached_table(("coarse", dn, bool(self.cfg["use_disp"]), float(self.cfg["min_depth"]),
                                        float(self.cfg["max_depth"])), coords.device,
                                       lambda: coarse_depth_table(self.cfg, dn, self.cfg["use_disp"]))
            self._post(outs, table, keep_hit_prob)
        self._add_gt(outs, que_imgs_info, self._gt_suffixes())
        return outs

    def _render_diner(self, que_imgs_info, ref_imgs_info, keep_hit_prob=False):
        """Depth-prior sample placement branch of render_impl (network/renderer.py:570-600) with
        diner_render_by_depth (:318-436): one fused placement kernel (candidates -> likelihood -> selection -> fill-up
        -> sort), then the usual pass on the per-ray depths with the COARSE networks; `c2f` adds the hierarchical fine
        pass.  Without `c2f` every output carries the `_fine` suffix, as in the reference.  The reference's in-op
        random draws are taken from que_imgs_info['diner_fill_rand'] (rn,n_samples) / ['diner_gauss'] (rn,n_gaussian)
        when present, else drawn on the device."""
        from .render_ops import depth_guided_placement
        cfg = self.cfg
        merge_uniform = cfg.get("N_uniform", 0) > 0 and bool(cfg.get("one_mlp", False))   # else the uniform pass is dead code (:526-528)
        if merge_uniform and cfg.get("c2f", False):
            raise NotImplementedError("N_uniform + one_mlp + c2f: the reference hands mismatched depth / hit_prob shapes to "
                                      "sample_fine_depth (renderer.py:586-589)")
        for k in ("mvs_depth", "mvs_uncert"):
            if k not in ref_imgs_info:
                raise _lib.PanoGRFError(f"diner_depth_guided_sampling needs ref_imgs_info[{k!r}]")
        ctx = self._context(que_imgs_info, ref_imgs_info)
        coords = que_imgs_info["coords"]
        assert coords.shape[0] == 1
        coords2 = coords[0].float().contiguous()
        rn, dev = coords2.shape[0], coords2.device
        depth = depth_guided_placement(cfg, que_imgs_info, ref_imgs_info, que_imgs_info.get("diner_fill_rand"),
                                       que_imgs_info.get("diner_gauss"))                     # (1,rn,N)
        N = depth.shape[-1]
        c2f = bool(cfg.get("c2f", False))
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        rfn = ctx["rfn"]

        def alloc(n):
            o = {"pixel_colors_nr": e(1, rn, 3), "colors_nr": e(1, rn, n, 3), "density_nr": e(1, rn, n)}
            if keep_hit_prob or c2f or cfg.get("render_uncert"):
                o["hit_prob_nr"] = e(1, rn, n)
            if cfg["use_ray_mask"]:
                ok = rfn >= cfg["ray_mask_view_num"] and n > cfg["ray_mask_point_num"]
                o["ray_mask"] = torch.full((1, rn), bool(ok), device=dev)
            if cfg["render_depth"]:
                o["render_depth"] = e(1, rn)
            return o

        coarse = alloc(N)
        fdn = int(cfg["fine_depth_sample_num"])
        fine_total = fdn + (N if cfg["fine_depth_use_all"] else 0)
        fine = alloc(fine_total) if c2f else None
        rpl = self.rays_per_launch or (524288 if self.mlp_dtype == "bf16" else 32768)
        d2 = depth[0]
        for r0 in range(0, rn, int(rpl)):
            n = min(int(rpl), rn - r0)
            fd = self._pass(ctx, coords2[r0:r0 + n], d2[r0:r0 + n], N, False, c2f, coarse, r0, "hit_prob_nr" in coarse)
            if c2f and que_imgs_info.get("ft_depth_range") is not None:
                self._ft_range_samples(que_imgs_info["ft_depth_range"][:, r0:r0 + n], d2[r0:r0 + n], fd, coarse_stride=N, dn=N)
            if c2f:
                self._pass(ctx, coords2[r0:r0 + n], fd, fine_total, not cfg.get("one_mlp", False), False, fine, r0,
                           "hit_prob_nr" in fine)
                if keep_hit_prob or cfg.get("render_uncert"):
                    fine.setdefault("que_depth", e(1, rn, fine_total))[0, r0:r0 + n] = fd
        coarse["que_depth"] = depth
        if merge_uniform:
            # merge_uniform_diner (renderer.py:526-565): a uniform-in-disparity pass of the same (coarse) networks, merged with the
            # depth-guided samples by depth and composited again
            from .render_ops import composite
            dn_u = int(cfg["depth_sample_num"])
            table = self._cached_table(("coarse", dn_u, True, float(cfg["min_depth"]), float(cfg["max_depth"])), dev,
                                       lambda: coarse_depth_table(cfg, dn_u, True))
            uni = {"pixel_colors_nr": e(1, rn, 3), "colors_nr": e(1, rn, dn_u, 3), "density_nr": e(1, rn, dn_u)}
            if cfg["render_depth"]:
                uni["render_depth"] = e(1, rn)
            for r0 in range(0, rn, int(rpl)):
                n = min(int(rpl), rn - r0)
                self._pass(ctx, coords2[r0:r0 + n], table, 0, False, False, uni, r0, False)
            z, idx = torch.cat([depth, table.view(1, 1, -1).expand(1, rn, -1)], 2).sort()
            col = torch.gather(torch.cat([coarse["colors_nr"], uni["colors_nr"]], 2), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 3))
            den = torch.gather(torch.cat([coarse["density_nr"], uni["density_nr"]], 2), 2, idx)
            hit, pix, rdepth = composite(den, col, z)
            coarse.update({"pixel_colors_nr": pix, "hit_prob_nr": hit, "colors_nr": col, "density_nr": den})
            if cfg["render_depth"]:
                coarse["render_depth"] = rdepth
            if cfg.get("render_uncert"):
                coarse["render_uncert"] = ((z - rdepth.unsqueeze(-1)).pow(2) * hit).sum(-1) + 1e-5
        elif cfg.get("render_uncert"):
            if "render_depth" not in coarse:
                raise KeyError("render_depth")
            coarse["render_uncert"] = ((depth - coarse["render_depth"].unsqueeze(-1)).pow(2) * coarse["hit_prob_nr"]).sum(-1) + 1e-5
        if c2f and cfg.get("render_uncert"):
            fine["render_uncert"] = ((fine["que_depth"] - fine["render_depth"].unsqueeze(-1)).pow(2) * fine["hit_prob_nr"]).sum(-1) + 1e-5
        if not (keep_hit_prob or not c2f):
            coarse.pop("hit_prob_nr", None)
        if c2f:
            outs = dict(coarse)
            outs.update({k + "_fine": v for k, v in fine.items()})
            return outs
        return {k + "_fine": v for k, v in coarse.items()}

    def _render_view(self, ctx, coords2, outs):
        """One C-ABI call for the whole view: pgrf_render_view_fwd runs the ray-batch loop and the
        coarse -> fine hand-off (network/renderer.py:647-683, 600-631) on the current stream."""
        cfg, lib = self.cfg, _lib.load()
        dev = coords2.device
        self._key_memo = {}
        try:
            return self._render_view_impl(ctx, coords2, outs)
        finally:
            self._key_memo = None

    def _render_view_impl(self, ctx, coords2, outs):
        cfg, lib = self.cfg, _lib.load()
        dev = coords2.device
        rn = coords2.shape[0]
        dn = int(cfg["depth_sample_num"])
        hier = bool(cfg["use_hierarchical_sampling"])
        fdn = int(cfg["fine_depth_sample_num"])
        fine_total = fdn + (dn if cfg["fine_depth_use_all"] else 0) if hier else 0
        one_mlp = bool(cfg.get("one_mlp", False))
        for is_fine, n in ((False, dn), (True, fine_total)):
            if is_fine and not hier:
                continue
            agg = self.fine_agg_net if (is_fine and not one_mlp) else self.agg_net
            if agg.cfg["sample_num"] != n:
                raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b "
                                   f"({agg.cfg['sample_num']}) at non-singleton dimension 1")   # ibrnet.py:358
        rpl = self.rays_per_launch or (524288 if self.mlp_dtype == "bf16" else 32768)
        chunk = max(1, min(int(rpl), rn))
        va = _lib.RenderViewArgs()
        a = va.pass_
        a.dataset = _lib.DATASET_IDS[cfg["dataset_name"]]
        a.H, a.W = int(cfg["height"]), int(cfg["width"])
        a.rfn, a.rn, a.dn = ctx["rfn"], rn, dn
        a.use_vis = int(bool(self.dist_decoder.cfg["use_vis"]))
        a.bias_val = float(self.dist_decoder.cfg["bias_val"])
        depth_table = self._cached_table(("coarse", dn, bool(cfg["use_disp"]), float(cfg["min_depth"]), float(cfg["max_depth"])),
                                         dev, lambda: coarse_depth_table(cfg, dn, cfg["use_disp"]))
        a.coords, a.depth, a.depth_ray_stride = _lib.ptr(coords2), _lib.ptr(depth_table), 0
        if ctx.get("ray_dirs") is not None:
            a.ray_dirs = _lib.ptr(ctx["ray_dirs"])
        a.wo_geometry = int(bool(self.agg_net.cfg.get("wo_geometry")))
        a.wo_appearance = int(bool(self.agg_net.cfg.get("wo_appearance")))
        a.que_c2w, a.que_near, a.que_far = _lib.ptr(ctx["c2w"]), ctx["que_near"], ctx["que_far"]
        a.ref_w2c, a.ref_depth_range = _lib.ptr(ctx["w2c"]), _lib.ptr(ctx["ref_range"])
        a.imgs_cl, a.img_h, a.img_w = _lib.ptr(ctx["imgs"]), ctx["imgs"].shape[1], ctx["imgs"].shape[2]
        a.img_feats_cl, a.if_h, a.if_w = _lib.ptr(ctx["img_feats"]), ctx["img_feats"].shape[1], ctx["img_feats"].shape[2]
        a.ray_feats_cl, a.rf_h, a.rf_w = _lib.ptr(ctx["ray_feats"]), ctx["ray_feats"].shape[1], ctx["ray_feats"].shape[2]
        a.weights = _lib.ptr(self._blob(False, dev))
        if self.mlp_dtype == "bf16":
            a.mlp_bf16, a.weights16 = 1, _lib.ptr(self._blob16(False, dev))
            a.sched = _lib.ptr(self._sched(dev))
        f1n, f2n = ctypes.c_longlong(), ctypes.c_longlong()
        _lib.check(lib.pgrf_render_workspace(a.rfn, chunk * max(dn, fine_total), ctypes.byref(f1n), ctypes.byref(f2n)),
                   "pgrf_render_workspace")
        ws = self._ws.setdefault(self._ws_key(dev), {})
        if self.mlp_dtype == "bf16":
            f1n.value = 4                      # the fused tensor-core kernel has no F1 inter-kernel tiles
        for key, n in (("f1", f1n.value), ("f2", f2n.value), ("fine", chunk * max(fine_total, 1))):
            if ws.get(key) is None or ws[key].numel() < n:
                ws[key] = torch.empty(n, device=dev, dtype=torch.float32)
        a.f1, a.f2 = _lib.ptr(ws["f1"]), _lib.ptr(ws["f2"])
        opt = lambda k: _lib.ptr(outs[k][0]) if k in outs else None
        a.pixel_colors = _lib.ptr(outs["pixel_colors_nr"][0])
        a.render_depth, a.hit_prob = opt("render_depth"), opt("hit_prob_nr")
        a.density, a.colors = opt("density_nr"), opt("colors_nr")
        va.hierarchical = int(hier)
        va.rays_per_launch = chunk
        keep = [depth_table]
        if hier:
            a.fine_dn, a.fine_u = fdn, _lib.ptr(ctx["fine_u"])
            a.fine_use_all, a.use_disp = int(bool(cfg["fine_depth_use_all"])), int(bool(cfg["use_disp"]))
            va.weights_fine = _lib.ptr(self._blob(not one_mlp, dev))
            if self.mlp_dtype == "bf16":
                va.weights16_fine = _lib.ptr(self._blob16(not one_mlp, dev))
            va.bias_val_fine = float((self.dist_decoder if one_mlp else self.fine_dist_decoder).cfg["bias_val"])
            va.fine_depth_ws = _lib.ptr(ws["fine"])
            va.pixel_colors_fine = _lib.ptr(outs["pixel_colors_nr_fine"][0])
            va.render_depth_fine, va.hit_prob_fine = opt("render_depth_fine"), opt("hit_prob_nr_fine")
            va.density_fine, va.colors_fine = opt("density_nr_fine"), opt("colors_nr_fine")
            va.que_depth_fine = opt("que_depth_fine")
        with torch.cuda.device(dev):
            rc = lib.pgrf_render_view_fwd(ctypes.byref(va), _lib.stream_ptr())
        _lib.check(rc, "pgrf_render_view_fwd")
        del keep

    # ---- module-level entry points of the reference renderer ----------------------------------------
    def predict_proj_ray_prob(self, prj_dict, ref_imgs_info, que_dists, is_fine):
        """renderer.py:120-136: dist decoder on prj_dict['ray_feats'] + compute_prob -> prj_dict['alpha'|'vis'|'hit_prob']."""
        rfn, qn, rn, dn, _ = prj_dict["pts"].shape
        n = qn * rn * dn
        dec = self.fine_dist_decoder if is_fine else self.dist_decoder
        dev = prj_dict["ray_feats"].device
        prj = torch.cat([_rows(prj_dict["pts"], rfn, n, 2), _rows(prj_dict["depth"], rfn, n, 1),
                         torch.zeros(rfn, n, 3, device=dev)], -1).contiguous()
        feat = torch.zeros(rfn, n, 67, device=dev)
        feat[..., :32] = _rows(prj_dict["ray_feats"], rfn, n, 32)
        interval = que_dists.detach().float().reshape(n).contiguous()
        # the reference always evaluates self.dist_decoder.compute_prob, i.e. the COARSE cfg's use_vis (renderer.py:129)
        out = _module_pass(_module_blob(dec, "dist_decoder.", dec), rfn, qn * rn, dn, self.dist_decoder.cfg["use_vis"],
                           dec.cfg["bias_val"], prj=prj, feat=feat, interval=interval, ref_range=ref_imgs_info["depth_range"],
                           stage_mask=1, want_prob=True)["prob"]
        prj_dict["alpha"] = out[..., 0].reshape(rfn, qn, rn, dn, 1)
        prj_dict["vis"] = out[..., 1].reshape(rfn, qn, rn, dn, 1)
        prj_dict["hit_prob"] = out[..., 2].reshape(rfn, qn, rn, dn, 1)
        return prj_dict

    def get_img_feats(self, ref_imgs_info, prj_dict):
        """renderer.py:180-188."""
        from .render_ops import interpolate_feature_map
        rfn, _, h, w = ref_imgs_info["imgs"].shape
        rfn, qn, rn, dn, _ = prj_dict["pts"].shape
        feats = interpolate_feature_map(ref_imgs_info["img_feats"], prj_dict["pts"].reshape(rfn, qn * rn * dn, 2), h, w)
        prj_dict["img_feats"] = feats.reshape(rfn, qn, rn, dn, -1)
        return prj_dict

    def network_rendering(self, prj_dict, que_dir, is_fine):
        """renderer.py:210-219: aggregation net + alpha compositing -> hit_prob (qn,rn,dn), colors (qn,rn,dn,3),
        pixel_colors (qn,rn,3), density (qn,rn,dn)."""
        net = self.fine_agg_net if is_fine else self.agg_net
        out, (qn, rn, dn) = net._run(prj_dict, que_dir)
        return (out["hit_prob"].reshape(qn, rn, dn), out["colors"].reshape(qn, rn, dn, 3),
                out["pixel_colors"].reshape(qn, rn, 3), out["density"].reshape(qn, rn, dn))

    def render_by_depth(self, que_depth, que_imgs_info, ref_imgs_info, is_train, is_fine, is_perspec=False):
        """network/renderer.py:223-317 for explicit per-ray sample depths (qn=1,rn,dn)."""
        if is_train:
            raise NotImplementedError("only the eval path is implemented")
        ctx = self._context(que_imgs_info, ref_imgs_info, is_perspec)
        coords2 = que_imgs_info["coords"][0].float().contiguous()
        rn, dn = coords2.shape[0], que_depth.shape[-1]
        dev = coords2.device
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        outs = {"pixel_colors_nr": e(1, rn, 3), "hit_prob_nr": e(1, rn, dn), "colors_nr": e(1, rn, dn, 3),
                "density_nr": e(1, rn, dn), "render_depth": e(1, rn)}
        depth = que_depth.reshape(rn, dn).float().contiguous()
        self._pass(ctx, coords2, depth, dn, bool(is_fine), False, outs, 0, True)
        if self.cfg["use_ray_mask"]:
            ok = ctx["rfn"] >= self.cfg["ray_mask_view_num"] and dn > self.cfg["ray_mask_point_num"]
            outs["ray_mask"] = torch.full((1, rn), bool(ok), device=dev)
        if not self.cfg["render_depth"]:
            outs.pop("render_depth")
        return self._add_gt(outs, que_imgs_info)

    def forward(self, data, is_perspec=False):
        que = dict(data["que_imgs_info"])
        ref = dict(data["ref_imgs_info"])
        return self.render(que, ref, "eval" not in data, is_perspec)


class NeuralRayGenRenderer(NeuralRayBaseRenderer):
    """network/renderer.py:688-786.  The initialisation network (cost-volume MVS + CNNs, `name2init_net`) is outside the
    hot path: attach it as `self.init_net` (any callable with the reference's signature
    `init_net(ref_imgs_info, src_imgs_info, is_train) -> {'ray_feats', 'mvs_depth'[, 'mvs_uncert']}`), or pass
    `ref_imgs_info` that already holds 'ray_feats'."""
    gen_default_cfg = {"init_net_type": "depth", "init_net_cfg": {}, "use_depth_loss": False, "depth_loss_coords_num": 8192}

    def __init__(self, cfg):
        super().__init__({**self.gen_default_cfg, **cfg})
        self.init_net = None

    def render_call(self, que_imgs_info, ref_imgs_info, is_train, src_imgs_info=None, is_perspec=False):
        if self.init_net is not None:
            ret = self.init_net(ref_imgs_info, src_imgs_info, is_train)
            ref_imgs_info["ray_feats"] = ret["ray_feats"]
            ref_imgs_info["mvs_depth"] = ret["mvs_depth"]
            if self.cfg.get("uncert_tune"):
                ref_imgs_info["mvs_uncert"] = ret["mvs_uncert"]
        elif "ray_feats" not in ref_imgs_info:
            raise _lib.PanoGRFError("NeuralRayGenRenderer: attach init_net or pass pre-computed ref_imgs_info['ray_feats']")
        if self.cfg.get("backface_culling"):          # renderer.py:713-714
            import types
            from .render_ops import depth2normal
            if "mvs_depth" not in ref_imgs_info:
                raise _lib.PanoGRFError("backface_culling needs ref_imgs_info['mvs_depth'] (depth2normal input)")
            ref_imgs_info["mvs_normal"] = depth2normal(ref_imgs_info, types.SimpleNamespace(
                dataset=self.cfg["dataset_name"], height=int(self.cfg["height"]), width=int(self.cfg["width"])))
        return self.render(que_imgs_info, ref_imgs_info, is_train, is_perspec)

    def gen_depth_loss_coords(self, h, w, device):
        coords = torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij"), -1).reshape(-1, 2).to(device)
        return coords[torch.randperm(coords.shape[0])[:self.cfg["depth_loss_coords_num"]]]

    def predict_mean_for_depth_loss(self, ref_imgs_info):
        """renderer.py:730-775: mixture means of the (fine) dist decoder at random source pixels."""
        from .render_ops import interpolate_feature_map
        ray_feats, imgs = ref_imgs_info["ray_feats"], ref_imgs_info["imgs"]
        rfn, _, h, w = imgs.shape
        coords = self.gen_depth_loss_coords(h, w, imgs.device).unsqueeze(0).repeat(rfn, 1, 1)
        feats = interpolate_feature_map(ray_feats, coords.float(), h, w)
        mean = self.dist_decoder.predict_mean(feats)
        out = {"depth_mean": mean[..., 0], "depth_coords": coords, "depth_mean_2": mean[..., 1]}
        if self.cfg["use_hierarchical_sampling"] and not self.cfg.get("one_mlp"):
            fmean = self.fine_dist_decoder.predict_mean(feats)
            out["depth_mean_fine"], out["depth_mean_fine_2"] = fmean[..., 0], fmean[..., 1]
        return out

    def forward(self, data, is_perspec=False):
        ref = dict(data["ref_imgs_info"])
        que = dict(data["que_imgs_info"])
        is_train = "eval" not in data
        src = dict(data["src_imgs_info"]) if "src_imgs_info" in data else None
        outs = self.render_call(que, ref, is_train, src, is_perspec=is_perspec)
        if (self.cfg["use_depth_loss"] and "true_depth" in ref) or (not is_train):
            outs.update(self.predict_mean_for_depth_loss(ref))
        return outs


name2network = {"neuray_base": NeuralRayBaseRenderer, "neuray_gen": NeuralRayGenRenderer}
