"""DefaultVisEncoder on the B200 tensor cores — SURVEY.md 8 f2 (per-call CNN on the render path).

Drop-in for `network/vis_encoder.py:6-33` (`ref_imgs_info['ray_feats'] = self.vis_encoder(ref_imgs_info['ray_feats'], ref_img_feats)`,
network/renderer.py:642): same constructor (`cfg["use_wrap_padding"]`), same call signature `(ray_feats, imgs_feats)`, same parameter
names (the module tree mirrors the reference's nn.Sequential indices), so a reference state_dict loads with `load_state_dict`.
Forward: one fused resize + concatenation + bf16 channels-last conversion, the six 3x3 / 1x1 convolutions as tcgen05 implicit GEMMs
(csrc/conv3d.cu with D = 1; the skip connections are added in the convolution epilogue), InstanceNorm2d + ReLU as a statistics pass
(fp64 accumulation) + an apply pass (csrc/vis_encoder.cu).  bf16 activations between layers: the accuracy class of the bf16 render
mode (rtol 1e-2); inference only.  `level=-1` (16-channel image features) is not built.
"""
import torch
from torch import nn

from . import _lib
from .regulariser import conv3d, pack_conv


class _Pad(nn.Module):
    """placeholder with WrapPadding's position in the reference's nn.Sequential (keeps the parameter indices identical)"""


def _conv3x3(cin, cout, wrap):
    conv = nn.Conv2d(cin, cout, kernel_size=3, padding=0, bias=False)
    return nn.Sequential(_Pad(), conv) if wrap else conv


class _ResidualBlock(nn.Module):
    """Parameter holder with ResidualBlock's names (network/ops.py:61-115, use_norm=True, bias=False)."""

    def __init__(self, dim, wrap):
        super().__init__()
        norm = lambda: nn.InstanceNorm2d(dim, track_running_stats=False, affine=True)
        layers = [norm(), nn.ReLU(True)] + ([_Pad()] if wrap else []) + [nn.Conv2d(dim, dim, 3, 1, padding=0, bias=False)]
        layers += [norm(), nn.ReLU(True)] + ([_Pad()] if wrap else []) + [nn.Conv2d(dim, dim, 3, 1, padding=0, bias=False)]
        self.conv = nn.Sequential(*layers)
        self.idx = (0, 3, 4, 7) if wrap else (0, 2, 3, 5)          # norm, conv, norm, conv


class _ConvStackBase(nn.Module):
    """Shared machinery of the reference's `conv3x3 -> n x ResidualBlock -> conv1x1` stacks (network/vis_encoder.py:17-22,
    network/init_net.py:540-574): packed weights, InstanceNorm + ReLU launches, the stack itself on bf16 channels-last maps."""

    def __init__(self):
        super().__init__()
        self._packed = {}
        self._ws_cache = {}

    def invalidate_weight_cache(self):
        """REQUIRED after writing parameters through `.data` (not seen by the version counters the cache key uses)."""
        self._packed.clear()

    def _pack(self, conv, ci_pad):
        ver = (conv.weight._version, conv.weight.data_ptr(), str(conv.weight.device), ci_pad)
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != ver:
            w = conv.weight.detach().float()
            co, ci, k = w.shape[0], w.shape[1], w.shape[2]
            w3 = torch.zeros((co, ci, 3, 3, 3), device=w.device, dtype=torch.float32)
            if k == 3:
                w3[:, :, 1] = w                                 # a 2-D kernel is the kd == 1 slice of a 3-D one
            else:
                w3[:, :, 1, 1, 1] = w[:, :, 0, 0]               # 1x1: the centre tap
            hit = (ver,) + pack_conv(w3, torch.zeros(co, device=w.device), ci, 0, ci_pad, 0)
            self._packed[id(conv)] = hit
        return hit[1], hit[2]

    @staticmethod
    def _norm_relu(lib, st, x, norm, n, hw):
        y = torch.empty_like(x)
        stats = torch.empty(2 * n * x.shape[-1], device=x.device, dtype=torch.float64)
        _lib.check(lib.pgrf_instnorm_relu_fwd(_lib.ptr(x), n, hw, x.shape[-1], _lib.ptr(norm.weight.detach().float().contiguous()),
                                              _lib.ptr(norm.bias.detach().float().contiguous()), float(norm.eps), _lib.ptr(stats), _lib.ptr(y),
                                              st), "pgrf_instnorm_relu_fwd")
        return y

    def _stack(self, lib, st, seq, a, wrap):
        """seq = Sequential(conv3x3, ResidualBlock..., conv1x1); a = bf16 channels-last (n,1,h,w,Cpad) -> fp32 (n,32,h,w)"""
        n, _, h, w, _ = a.shape
        dims = (n, 1, h, w)
        c0 = seq[0][1] if wrap else seq[0]
        wk, bk = self._pack(c0, a.shape[-1])
        x = conv3d(a, None, wk, bk, 32, dims, 0, self._ws_cache, lib, st, act=False, wrap=wrap)
        for blk in list(seq)[1:-1]:
            i = blk.idx
            t = self._norm_relu(lib, st, x, blk.conv[i[0]], n, h * w)
            wk, bk = self._pack(blk.conv[i[1]], 32)
            t = conv3d(t, None, wk, bk, 32, dims, 0, self._ws_cache, lib, st, act=False, wrap=wrap)
            t = self._norm_relu(lib, st, t, blk.conv[i[2]], n, h * w)
            wk, bk = self._pack(blk.conv[i[3]], 32)
            x = conv3d(t, None, wk, bk, 32, dims, 0, self._ws_cache, lib, st, act=False, res=x, wrap=wrap)     # + skip
        wk, bk = self._pack(seq[-1], 32)
        out = torch.empty((n, 32, 1, h, w), device=a.device, dtype=torch.float32)
        _lib.check(lib.pgrf_conv3d_pointwise_fwd(_lib.ptr(x), 32, None, 0, _lib.ptr(wk), _lib.ptr(bk), None, _lib.ptr(out), 32, 32,
                                                 n, 1, h, w, 0, st), "pgrf_conv3d_pointwise_fwd")
        return out.view(n, 32, h, w)


def _stack_modules(cin, n_blocks, wrap):
    return nn.Sequential(_conv3x3(cin, 32, wrap), *[_ResidualBlock(32, wrap) for _ in range(n_blocks)],
                         nn.Conv2d(32, 32, kernel_size=1, bias=False))


class DefaultVisEncoder(_ConvStackBase):
    default_cfg = {"use_wrap_padding": True}

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = {**self.default_cfg, **(cfg or {})}
        if self.cfg.get("level") in (-1,):
            raise _lib.PanoGRFError("DefaultVisEncoder: level=-1 (16-channel image features) is not built")
        self.wrap = bool(self.cfg["use_wrap_padding"])
        self.out_conv = _stack_modules(64, 2, self.wrap)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, ray_feats, imgs_feats):
        """ray_feats (N,32,h,w), imgs_feats (N,32,hi,wi) fp32 CUDA tensors -> (N,32,h,w) fp32"""
        _lib.require_cuda(ray_feats, imgs_feats)
        lib = _lib.load()
        ray = ray_feats.detach().float().contiguous()
        img = imgs_feats.detach().float().contiguous()
        n, cr, h, w = ray.shape
        ci, hi, wi = img.shape[1:]
        if img.shape[0] != n or ci + cr != 64:
            raise RuntimeError(f"DefaultVisEncoder expects 32 + 32 channels for the same views, got {tuple(img.shape)} and {tuple(ray.shape)}")
        with torch.cuda.device(ray.device):
            st = _lib.stream_ptr()
            a = torch.empty((n, 1, h, w, 64), device=ray.device, dtype=torch.bfloat16)
            _lib.check(lib.pgrf_feats_to_bf16_cl(_lib.ptr(img), ci, hi, wi, _lib.ptr(ray), cr, n, h, w, _lib.ptr(a), st), "pgrf_feats_to_bf16_cl")
            return self._stack(lib, st, self.out_conv, a, self.wrap)


class CostVolumeInitConvs(_ConvStackBase):
    """The convolution stacks of `CostVolumeInitNet` (network/init_net.py:540-574, applied at :606-636): `depth_conv` on the
    quarter-resolution MVS depth (1 -> 32 channels) and `out_conv` on cat(image features, depth features) -> `ray_feats`.
    Same parameter names (`depth_conv.*`, `out_conv.*`), so the sub-state_dict of a reference `init_net` loads with strict=True.
    The image encoder (`res_net`) and the MVS depth network stay the caller's."""

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = {"use_wrap_padding": True, **(cfg or {})}
        self.wrap = bool(self.cfg["use_wrap_padding"])
        self.depth_conv = _stack_modules(1, 1, self.wrap)
        self.out_conv = _stack_modules(64, 1, self.wrap)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, ref_feats, depth):
        """ref_feats (N,32,h,w) = res_net output, depth (N,1,h,w) = the MVS depth after `extract_depth_for_init_impl` and the x0.25
        bilinear resize (init_net.py:620-627) -> ray_feats (N,32,h,w) fp32"""
        _lib.require_cuda(ref_feats, depth)
        lib = _lib.load()
        feats = ref_feats.detach().float().contiguous()
        d = depth.detach().float().contiguous()
        n, c, h, w = feats.shape
        if d.shape != (n, 1, h, w) or c != 32:
            raise RuntimeError(f"CostVolumeInitConvs expects (N,32,h,w) features and (N,1,h,w) depth, got {tuple(feats.shape)} and {tuple(d.shape)}")
        dev = feats.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            a = torch.empty((n, 1, h, w, 16), device=dev, dtype=torch.bfloat16)            # 1 channel padded to 16
            _lib.check(lib.pgrf_conv3d_to_bf16_cl(_lib.ptr(d), h * w, h * w, 0, w, 1, n, 1, 1, h, w, 16, _lib.ptr(a), st),
                       "pgrf_conv3d_to_bf16_cl")
            depth_feats = self._stack(lib, st, self.depth_conv, a, self.wrap)
            b = torch.empty((n, 1, h, w, 64), device=dev, dtype=torch.bfloat16)
            _lib.check(lib.pgrf_feats_to_bf16_cl(_lib.ptr(feats), 32, h, w, _lib.ptr(depth_feats), 32, n, h, w, _lib.ptr(b), st),
                       "pgrf_feats_to_bf16_cl")
            return self._stack(lib, st, self.out_conv, b, self.wrap)


name2vis_encoder = {"default": DefaultVisEncoder}
