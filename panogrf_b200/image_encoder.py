"""Image encoder `ResUNetLight` on the B200 tensor cores — SURVEY.md 8 f2 (the per-view CNN of the render path).

Drop-in for `network/ops.py:235-455` as the renderer builds it (`network/renderer.py:106`: `ResUNetLight(cfg, 3, [1, 2, 6, 4], 32,
inplanes=16, use_wrap_padding=...)`; `handle_distort` off, no autoencoder head): same constructor arguments, same module tree and
parameter names (a reference state_dict loads with `load_state_dict`), same construction order (a seeded construction draws the
reference's initial weights).  Forward through the C ABI only:
  conv1 (7x7, stride 2)      = patch gather (csrc/vis_encoder.cu) + pointwise tensor-core GEMM
  3x3 / 1x1 convolutions     = csrc/conv3d.cu with D = 1 (stride 2 = stride-1 result / input sub-sampled at the even pixels)
  InstanceNorm + ReLU / ELU  = statistics pass (fp64 accumulation) + apply pass, the BasicBlock skip added inside the apply pass
  upconv                     = x2 bilinear (align_corners=True) + convolution with ZERO padding (the reference's `upconv` never
                               forwards `use_wrap_padding`, ops.py:229), skip concatenations = the convolution's second operand.
bf16 activations between layers (the accuracy class of the bf16 render mode); inference only.
"""
import torch
from torch import nn

from . import _lib
from .regulariser import conv3d, pack_conv
from .vis_encoder import _Pad


def _conv3x3(cin, cout, stride, wrap):
    conv = nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=0 if wrap else 1, bias=False)
    return nn.Sequential(_Pad(), conv) if wrap else conv


def _inorm(dim):
    return nn.InstanceNorm2d(dim, track_running_stats=False, affine=True)


class _BasicBlock(nn.Module):
    """Parameter holder with BasicBlock's names and registration order (network/ops.py:126-197)."""

    def __init__(self, inplanes, planes, stride, downsample, wrap):
        super().__init__()
        self.conv1 = _conv3x3(inplanes, planes, stride, wrap)
        self.bn1 = _inorm(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv3x3(planes, planes, 1, wrap)
        self.bn2 = _inorm(planes)
        self.downsample = downsample
        self.stride = stride


class _Conv(nn.Module):
    """`conv` (network/ops.py:199-224): [WrapPadding] Conv2d(+bias) - InstanceNorm - ELU"""

    def __init__(self, cin, cout, wrap):
        super().__init__()
        c = nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=0 if wrap else 1)
        self.conv = nn.Sequential(_Pad(), c) if wrap else c
        self.bn = _inorm(cout)
        self.wrap = wrap


class _UpConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _Conv(cin, cout, False)            # ops.py:229: use_wrap_padding is not forwarded


class ResUNetLight(nn.Module):
    def __init__(self, cfg=None, in_dim=3, layers=(2, 3, 6, 3), out_dim=32, inplanes=32, use_wrap_padding=False, autoencoder=False):
        super().__init__()
        self.cfg = dict(cfg or {})
        if self.cfg.get("handle_distort") or self.cfg.get("handle_distort_input_all") or autoencoder:
            raise _lib.PanoGRFError("ResUNetLight: handle_distort / autoencoder variants are not built")
        wrap = bool(use_wrap_padding)
        self.use_wrap_padding = wrap
        self.in_dim = in_dim
        self.inplanes = inplanes
        c1 = nn.Conv2d(in_dim, inplanes, kernel_size=7, stride=2, padding=0 if wrap else 3, bias=False)
        self.conv1 = nn.Sequential(_Pad(), c1) if wrap else c1
        self.bn1 = _inorm(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(32, layers[0], wrap)
        self.layer2 = self._make_layer(64, layers[1], wrap)
        self.layer3 = self._make_layer(128, layers[2], wrap)
        self.upconv3 = _UpConv(128, 64)
        self.iconv3 = _Conv(64 + 64, 64, wrap)
        self.upconv2 = _UpConv(64, 32)
        self.iconv2 = _Conv(64, 32, wrap)
        self.out_conv = nn.Conv2d(32, out_dim, 1, 1)
        self.out_dim = out_dim
        self._packed = {}
        self._ws_cache = {}
        for p in self.parameters():
            p.requires_grad_(False)

    def _make_layer(self, planes, blocks, wrap):
        """ops.py:340-365: the downsample branch is constructed BEFORE its block (RNG order), every stage has stride 2"""
        downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=2, bias=False), _inorm(planes))
        layers = [_BasicBlock(self.inplanes, planes, 2, downsample, wrap)]
        self.inplanes = planes
        layers += [_BasicBlock(planes, planes, 1, None, wrap) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def invalidate_weight_cache(self):
        """REQUIRED after writing parameters through `.data` (not seen by the version counters the cache key uses)."""
        self._packed.clear()

    # ---- packed operands ----------------------------------------------------------------------------------------------------------
    def _pack(self, conv, ca, cb=0, kind="3x3"):
        ver = (conv.weight._version, conv.weight.data_ptr(), str(conv.weight.device), None if conv.bias is None else conv.bias._version)
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != ver:
            w = conv.weight.detach().float()
            co = w.shape[0]
            bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(co, device=w.device)
            if kind == "3x3":
                w3 = torch.zeros((co, ca + cb, 3, 3, 3), device=w.device, dtype=torch.float32)
                w3[:, :, 1] = w
                hit = (ver,) + pack_conv(w3, bias, ca, cb, ca, cb)
            else:                                        # pointwise: (co, K) with K padded to ca
                k = w.reshape(co, -1)
                w3 = torch.zeros((co, ca, 3, 3, 3), device=w.device, dtype=torch.float32)
                w3[:, :k.shape[1], 1, 1, 1] = k
                hit = (ver,) + pack_conv(w3, bias, ca, 0, ca, 0)
            self._packed[id(conv)] = hit
        return hit[1], hit[2]

    # ---- launches -----------------------------------------------------------------------------------------------------------------
    def _norm(self, lib, st, x, norm, act, res=None):
        n, _, h, w, c = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(2 * n * c, device=x.device, dtype=torch.float64)
        _lib.check(lib.pgrf_instnorm_act_fwd(_lib.ptr(x), n, h * w, c, _lib.ptr(norm.weight.detach().float().contiguous()),
                                             _lib.ptr(norm.bias.detach().float().contiguous()), float(norm.eps), _lib.ptr(stats),
                                             _lib.ptr(res) if res is not None else None, act, _lib.ptr(y), st), "pgrf_instnorm_act_fwd")
        return y

    def _conv3(self, lib, st, xa, xb, seq_or_conv, wrap_mod, wrap):
        conv = seq_or_conv[1] if wrap_mod else seq_or_conv
        n, _, h, w, ca = xa.shape
        cb = xb.shape[-1] if xb is not None else 0
        wk, bk = self._pack(conv, ca, cb)
        return conv3d(xa, xb, wk, bk, conv.weight.shape[0], (n, 1, h, w), 0, self._ws_cache, lib, st, act=False, wrap=wrap)

    def _pointwise(self, lib, st, x, conv, f32=False):
        n, _, h, w, ca = x.shape
        co = conv.weight.shape[0]
        wk, bk = self._pack(conv, ca, kind="1x1")
        if f32:
            y = torch.empty((n, co, 1, h, w), device=x.device, dtype=torch.float32)
        else:
            y = torch.empty((n, 1, h, w, co), device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.pgrf_conv3d_pointwise_fwd(_lib.ptr(x), ca, None, 0, _lib.ptr(wk), _lib.ptr(bk), None if f32 else _lib.ptr(y),
                                                 _lib.ptr(y) if f32 else None, co if f32 else 0, co, n, 1, h, w, 0, st),
                   "pgrf_conv3d_pointwise_fwd")
        return y

    def _subsample(self, lib, st, x):
        n, _, h, w, c = x.shape
        y = torch.empty((n, 1, (h + 1) // 2, (w + 1) // 2, c), device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.pgrf_subsample2_fwd(_lib.ptr(x), n, h, w, c, _lib.ptr(y), st), "pgrf_subsample2_fwd")
        return y

    def _block(self, lib, st, blk, x):
        wrap = self.use_wrap_padding
        t = self._conv3(lib, st, x, None, blk.conv1, wrap, wrap)
        if blk.stride == 2:
            t = self._subsample(lib, st, t)                  # stride 2 = the stride-1 result at the even pixels
        t = self._norm(lib, st, t, blk.bn1, 1)
        t = self._conv3(lib, st, t, None, blk.conv2, wrap, wrap)
        idn = x
        if blk.downsample is not None:
            idn = self._pointwise(lib, st, self._subsample(lib, st, x), blk.downsample[0])
            idn = self._norm(lib, st, idn, blk.downsample[1], 0)
        return self._norm(lib, st, t, blk.bn2, 1, res=idn)   # relu(bn2(conv2) + identity)

    def _conv_module(self, lib, st, mod, xa, xb=None):
        t = self._conv3(lib, st, xa, xb, mod.conv, mod.wrap, mod.wrap)
        return self._norm(lib, st, t, mod.bn, 2)

    def _upconv(self, lib, st, mod, x):
        n, _, h, w, c = x.shape
        up = torch.empty((n, 1, 2 * h, 2 * w, c), device=x.device, dtype=torch.bfloat16)
        _lib.check(lib.pgrf_upsample2d2_ac_fwd(_lib.ptr(x), n, h, w, c, _lib.ptr(up), st), "pgrf_upsample2d2_ac_fwd")
        return self._conv_module(lib, st, mod.conv, up)

    def forward(self, x):
        """x (N, in_dim, H, W) fp32 CUDA tensor, H and W multiples of 16 -> (N, out_dim, H/4, W/4) fp32"""
        _lib.require_cuda(x)
        lib = _lib.load()
        x = x.detach().float().contiguous()
        n, cin, H, W = x.shape
        if cin != self.in_dim:
            raise RuntimeError(f"ResUNetLight expects {self.in_dim} input channels, got {cin}")
        if H % 16 or W % 16:
            raise RuntimeError(f"ResUNetLight: H, W must be multiples of 16 (got {H}, {W}); the padded skip connections of "
                               "ops.py:370-381 are not built")
        dev = x.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            kpad = (cin * 49 + 15) // 16 * 16
            ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            patches = torch.empty((n, 1, ho, wo, kpad), device=dev, dtype=torch.bfloat16)
            _lib.check(lib.pgrf_patch7x7_s2_fwd(_lib.ptr(x), n, cin, H, W, kpad, 1 if self.use_wrap_padding else 0, _lib.ptr(patches), st),
                       "pgrf_patch7x7_s2_fwd")
            c1 = self.conv1[1] if self.use_wrap_padding else self.conv1
            x0 = self._norm(lib, st, self._pointwise(lib, st, patches, c1), self.bn1, 1)
            feats, cur = [], x0
            for layer in (self.layer1, self.layer2, self.layer3):
                for blk in layer:
                    cur = self._block(lib, st, blk, cur)
                feats.append(cur)
            x1, x2, x3 = feats
            t = self._upconv(lib, st, self.upconv3, x3)
            t = self._conv_module(lib, st, self.iconv3, t, x2)       # torch.cat([upsampled, skip], 1)
            t = self._upconv(lib, st, self.upconv2, t)
            t = self._conv_module(lib, st, self.iconv2, t, x1)
            out = self._pointwise(lib, st, t, self.out_conv, f32=True)
        return out.view(n, self.out_dim, out.shape[-2], out.shape[-1])
