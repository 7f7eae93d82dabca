"""3-D cost regulariser (`unet3d`) on the B200 tensor cores — SURVEY.md 8 f1.

Mirrors what the reference builds in models/test_models.py:81-146 (initialize_cost_volume_network: a UNet2 of Conv3DBlockv2 with
WrapPadding3D, LeakyReLU, AvgPool3d(2), trilinear x2 upsampling) and applies at network/omni_mvsnet/pipeline3_model.py:847-855
(`self.unet3d(cost_volume)`): same constructor meaning (`size`, `num_layer`), same parameter names
(`encoders.{i}.conv{1,2}.{weight,bias}`, `decoders.{i}.conv{1,2}.*`) so a reference state_dict loads with `load_state_dict`, and the
same module construction order so that a seeded construction draws the same initial weights as the reference's.

The forward pass runs through the C ABI only (csrc/conv3d.cu): activations bf16 channels-last, 3x3x3 convolutions as tcgen05
implicit GEMMs with the wrap/zero padding folded into the operand gather, fp32 accumulation, bias + LeakyReLU in the epilogue, the
U-Net's concatenations as a second operand pointer.  Inference only (the MVS depth network is frozen at render time,
pipeline3_model.py:647).
"""
import ctypes

import torch
from torch import nn

from . import _lib


def _pad16(c):
    return (c + 15) // 16 * 16


def chunk_size(ca, cb):
    """Channels per pipeline stage: <= 64, divides both inputs of a concatenation (csrc/conv3d.cu conv3d_plan)."""
    kc = min(ca + cb, 64)
    while ca % kc or cb % kc:
        kc //= 2
    return kc


def pack_conv(weight, bias, ca, cb, ca_pad, cb_pad):
    """(Co, ca+cb, 3,3,3) conv weights -> the kernel's operand blocks [n_tiles][27][n_cc][KC/8][NT][8] bf16 (+ padded fp32 bias).

    Input channel ci sits at position ci (first input) or ca_pad + ci - ca (second input) of the padded K axis."""
    co = weight.shape[0]
    co_pad = _pad16(co)
    nt = min(128, co_pad)
    assert co_pad % nt == 0, f"C_out={co} unsupported"
    kpad = ca_pad + cb_pad
    kc = chunk_size(ca_pad, cb_pad)
    w = weight.detach().float().reshape(co, ca + cb, 27).permute(0, 2, 1)            # (co, tap, ci), tap = kd*9 + kh*3 + kw
    wp = torch.zeros((co_pad, 27, kpad), device=w.device, dtype=torch.float32)
    wp[:co, :, :ca] = w[:, :, :ca]
    if cb:
        wp[:co, :, ca_pad:ca_pad + cb] = w[:, :, ca:]
    wp = wp.reshape(co_pad // nt, nt, 27, kpad // kc, kc // 8, 8).permute(0, 2, 3, 4, 1, 5)
    bp = torch.zeros(co_pad, device=w.device, dtype=torch.float32)
    bp[:co] = bias.detach().float()
    return wp.contiguous().to(torch.bfloat16), bp


def pack_conv_cout1(weight, ca, cb, ca_pad, cb_pad):
    """(1, ca+cb, 3,3,3) -> fp32 [27][ca_pad+cb_pad] for the single-output-channel kernel."""
    w = weight.detach().float().reshape(ca + cb, 27).t()
    wp = torch.zeros((27, ca_pad + cb_pad), device=w.device, dtype=torch.float32)
    wp[:, :ca] = w[:, :ca]
    if cb:
        wp[:, ca_pad:ca_pad + cb] = w[:, ca:]
    return wp.contiguous()


def conv3d(xa, xb, wpk, bias, co_pad, dims, f32_channels=0, ws_cache=None, lib=None, st=None, act=True, res=None, wrap=True):
    """One Conv3d(3x3x3) + WrapPadding3D + bias + LeakyReLU layer on [xa | xb] (bf16 channels-last, packed weights from `pack_conv`).
    Returns bf16 channels-last (B,D,H,W,co_pad), or fp32 (B,f32_channels,D,H,W) when f32_channels > 0.  `res`: bf16 channels-last
    residual added after bias / activation; `wrap=False`: zeros instead of the wrap along the width."""
    lib = lib or _lib.load()
    st = _lib.stream_ptr() if st is None else st
    B, D, H, W = dims
    ca_pad, cb_pad = xa.shape[-1], (xb.shape[-1] if xb is not None else 0)
    need = ctypes.c_longlong(0)
    _lib.check(lib.pgrf_conv3d_workspace(ca_pad, cb_pad, co_pad, B, D, H, W, ctypes.byref(need)), "pgrf_conv3d_workspace")
    ws = None
    if need.value:
        key = (xa.device, getattr(st, "value", st))             # stream-ordered reuse: one buffer per (device, stream)
        ws = ws_cache.get(key) if ws_cache is not None else None
        if ws is None or ws.numel() < need.value:
            ws = torch.empty(need.value, device=xa.device, dtype=torch.float32)
            if ws_cache is not None:
                ws_cache[key] = ws
    if f32_channels:
        y = torch.empty((B, f32_channels, D, H, W), device=xa.device, dtype=torch.float32)
    else:
        y = torch.empty((B, D, H, W, co_pad), device=xa.device, dtype=torch.bfloat16)
    _lib.check(lib.pgrf_conv3d_ex_fwd(_lib.ptr(xa), ca_pad, _lib.ptr(xb) if xb is not None else None, cb_pad, _lib.ptr(wpk), _lib.ptr(bias),
                                      None if f32_channels else _lib.ptr(y), _lib.ptr(y) if f32_channels else None, f32_channels, co_pad,
                                      B, D, H, W, 1 if act else 0, _lib.ptr(ws) if ws is not None else None, need.value,
                                      _lib.ptr(res) if res is not None else None, 1 if wrap else 0, st), "pgrf_conv3d_ex_fwd")
    return y


class _Block(nn.Module):
    """Parameter holder with Conv3DBlockv2's names (models/common_blocks.py:366-445)."""

    def __init__(self, cin, cout, pool):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=0)
        self.conv2 = nn.Conv3d(cout, cout, kernel_size=(3, 3, 3), padding=0)
        self.pool = pool


class CostRegulariser3D(nn.Module):
    """`unet3d` of the reference: (B, 2^(size+1), D, H, W) cost volume -> (B, 1, D, H, W) regularised cost."""

    def __init__(self, size=4, num_layer=3):
        super().__init__()
        # construction order == models/test_models.py:113-146: the first decoder is created before the encoders
        enc, dec = [], [_Block(2 ** (size + 3), 1, False)]
        for i in range(num_layer):
            ch = 2 ** (i + size + 1)
            enc.append(_Block(ch, 2 * ch, True))
            if i > 0:
                dec.append(_Block(4 * ch, ch, False))
        enc.append(_Block(2 ** (num_layer + size + 1), 2 ** (num_layer + size + 2), False))
        self.encoders = nn.ModuleList(enc)
        self.decoders = nn.ModuleList(dec)
        self.in_channels = 2 ** (size + 1)
        self._packed = {}
        self._ws_cache = {}          # split-K workspace per device (stream-ordered reuse)
        for p in self.parameters():
            p.requires_grad_(False)

    # ---- weights ----------------------------------------------------------------------------------------------------------------
    def invalidate_weight_cache(self):
        """Drop the packed operands.  REQUIRED after writing parameters through `.data` (such writes do not bump the version counter
        the cache key uses); optimizer steps, load_state_dict and .to()/.cuda() are detected automatically."""
        self._packed.clear()

    def _pack(self, conv, ca, cb, ca_pad, cb_pad, simt=False):
        key = id(conv)
        ver = (conv.weight._version, conv.bias._version, conv.weight.data_ptr(), str(conv.weight.device), ca, cb, simt)
        hit = self._packed.get(key)
        if hit is None or hit[0] != ver:
            if simt:
                hit = (ver, pack_conv_cout1(conv.weight, ca, cb, ca_pad, cb_pad), float(conv.bias.detach().float().item()))
            else:
                hit = (ver,) + pack_conv(conv.weight, conv.bias, ca, cb, ca_pad, cb_pad)
            self._packed[key] = hit
        return hit[1], hit[2]

    def _pack_scalar(self, conv):
        """(1, 1, 3,3,3) -> 27 host floats (passed by value to the stencil kernel) + bias"""
        ver = (conv.weight._version, conv.bias._version, conv.weight.data_ptr(), str(conv.weight.device), "scalar")
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != ver:
            w = (ctypes.c_float * 27)(*conv.weight.detach().float().reshape(-1).cpu().tolist())
            hit = (ver, w, float(conv.bias.detach().float().item()))
            self._packed[id(conv)] = hit
        return hit[1], hit[2]

    def _pack_head(self, conv, ca, cb, ca_pad, cb_pad):
        """(1, ca+cb, 3,3,3) head -> pointwise weights with the 27 taps as output channels (centre tap of a (32, ca+cb, 3,3,3) kernel)"""
        ver = (conv.weight._version, conv.bias._version, conv.weight.data_ptr(), str(conv.weight.device), ca, cb, "head")
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != ver:
            w = conv.weight.detach().float()
            wz = torch.zeros((32, ca + cb, 3, 3, 3), device=w.device, dtype=torch.float32)
            wz[:27, :, 1, 1, 1] = w[0].reshape(ca + cb, 27).t()
            wpk, bz = pack_conv(wz, torch.zeros(32, device=w.device), ca, cb, ca_pad, cb_pad)
            hit = (ver, wpk, bz, float(conv.bias.detach().float().item()))
            self._packed[id(conv)] = hit
        return hit[1], hit[2], hit[3]

    # ---- kernels ----------------------------------------------------------------------------------------------------------------
    def _conv(self, lib, st, xa, ca_pad, xb, cb_pad, wpk, bias, co_pad, dims, f32_channels=0):
        return conv3d(xa, xb, wpk, bias, co_pad, dims, f32_channels, self._ws_cache, lib, st)

    def _block(self, lib, st, blk, xa, ca, xb, cb, dims):
        """pad-conv1-lrelu-pad-conv2-lrelu of one block on [xa | xb]; returns the un-pooled bf16 activation and its channel count."""
        ca_pad, cb_pad = xa.shape[-1], (xb.shape[-1] if xb is not None else 0)
        co = blk.conv1.weight.shape[0]
        w1, b1 = self._pack(blk.conv1, ca, cb, ca_pad, cb_pad)
        t = self._conv(lib, st, xa, ca_pad, xb, cb_pad, w1, b1, _pad16(co), dims)
        w2, b2 = self._pack(blk.conv2, co, 0, _pad16(co), 0)
        return self._conv(lib, st, t, _pad16(co), None, 0, w2, b2, _pad16(co), dims), co

    def forward(self, cost_volume):
        """cost_volume (B, C, D, H, W) fp32 CUDA tensor with ANY strides (the permuted views the sweep returns are consumed in place)."""
        _lib.require_cuda(cost_volume)
        lib = _lib.load()
        x = cost_volume.detach()
        B, C, D, H, W = x.shape
        # bf16 channels-last storage (what calculate_cost_volume_erp(out_layout="bdhwc_bf16").permute(0, 4, 1, 2, 3) is): consumed in place
        direct = (x.dtype == torch.bfloat16 and C % 16 == 0 and x.stride() == (D * H * W * C, 1, H * W * C, W * C, C)
                  and x.data_ptr() % 16 == 0)
        if not direct and x.dtype != torch.float32:
            x = x.float()
        n_enc = len(self.encoders)
        f = 2 ** (n_enc - 1)
        if C != self.in_channels:
            raise RuntimeError(f"unet3d expects {self.in_channels} input channels, got {C}")       # the reference's conv would raise
        if D % f or H % f or W % f:
            raise RuntimeError(f"unet3d: D, H, W must be multiples of {f} (got {D}, {H}, {W}); the reference's skip concatenation fails "
                               "on other sizes")
        dev = x.device
        if next(self.parameters()).device != dev:
            raise RuntimeError("CostRegulariser3D parameters and input are on different devices; call .to(device) first")
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            if direct:
                a = x.permute(0, 2, 3, 4, 1)                      # the (B,D,H,W,C) storage itself
            else:
                cpad = _pad16(C)
                a = torch.empty((B, D, H, W, cpad), device=dev, dtype=torch.bfloat16)
                sb, sc, sd, sh, sw = x.stride()
                _lib.check(lib.pgrf_conv3d_to_bf16_cl(_lib.ptr(x), sb, sc, sd, sh, sw, B, C, D, H, W, cpad, _lib.ptr(a), st),
                           "pgrf_conv3d_to_bf16_cl")
            dims = (B, D, H, W)
            skips = []
            ch = C
            for i, blk in enumerate(self.encoders):
                u, ch = self._block(lib, st, blk, a, ch, None, 0, dims)
                skips.append((u, ch, dims))
                if blk.pool:
                    b_, d_, h_, w_ = dims
                    a = torch.empty((b_, d_ // 2, h_ // 2, w_ // 2, u.shape[-1]), device=dev, dtype=torch.bfloat16)
                    _lib.check(lib.pgrf_avgpool3d2_fwd(_lib.ptr(u), b_, d_, h_, w_, u.shape[-1], _lib.ptr(a), st), "pgrf_avgpool3d2_fwd")
                    dims = (b_, d_ // 2, h_ // 2, w_ // 2)
                else:
                    a = u
            n_dec = len(self.decoders)
            for i in range(n_dec - 1, -1, -1):
                b_, d_, h_, w_ = dims
                up = torch.empty((b_, 2 * d_, 2 * h_, 2 * w_, a.shape[-1]), device=dev, dtype=torch.bfloat16)
                _lib.check(lib.pgrf_upsample3d2_fwd(_lib.ptr(a), b_, d_, h_, w_, a.shape[-1], _lib.ptr(up), st), "pgrf_upsample3d2_fwd")
                dims = (b_, 2 * d_, 2 * h_, 2 * w_)
                if i == n_dec - 1:
                    xb, cb = None, 0
                else:
                    xb, cb, sdims = skips[i]
                    assert sdims == dims
                blk = self.decoders[i]
                if blk.conv1.weight.shape[0] > 1:
                    a, ch = self._block(lib, st, blk, up, ch, xb, cb, dims)
                    continue
                # last decoder, conv1: (ch + cb) -> 1.  Evaluated as a pointwise GEMM with the 27 taps as output channels (every voxel's
                # channels are read once, not once per tap row) + a 27-tap stencil over the fp32 scalar planes; then 1 -> 1 on the
                # fp32 pipes
                ca_pad, cb_pad = up.shape[-1], (xb.shape[-1] if xb is not None else 0)
                wz, bz, b1 = self._pack_head(blk.conv1, ch, cb, ca_pad, cb_pad)
                z = torch.empty((dims[0], 27) + tuple(dims[1:]), device=dev, dtype=torch.float32)
                _lib.check(lib.pgrf_conv3d_pointwise_fwd(_lib.ptr(up), ca_pad, _lib.ptr(xb) if xb is not None else None, cb_pad,
                                                         _lib.ptr(wz), _lib.ptr(bz), None, _lib.ptr(z), 27, 32, *dims, 0, st),
                           "pgrf_conv3d_pointwise_fwd")
                t = torch.empty(dims, device=dev, dtype=torch.float32)
                _lib.check(lib.pgrf_conv3d_tapsum_fwd(_lib.ptr(z), b1, *dims, 1, _lib.ptr(t), st), "pgrf_conv3d_tapsum_fwd")
                w2, b2 = self._pack_scalar(blk.conv2)
                out = torch.empty(dims, device=dev, dtype=torch.float32)
                _lib.check(lib.pgrf_conv3d_scalar_fwd(_lib.ptr(t), w2, b2, *dims, 1, _lib.ptr(out), st), "pgrf_conv3d_scalar_fwd")
                return out.unsqueeze(1)
        raise RuntimeError("unet3d: the first decoder must have one output channel (models/test_models.py:113)")


class _Block2D(nn.Module):
    """Parameter holder with ConvBlock2's names (models/common_blocks.py:96-184)."""

    def __init__(self, cin, cout, upscale, act):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, kernel_size=3, padding=0)
        self.conv2 = nn.Conv2d(cout, cout, kernel_size=3, padding=0)
        self.upscale, self.act = upscale, act


class _Conv1x1(nn.Module):
    """Parameter holder with ConvBlock's name `conv` (models/common_blocks.py:10-94)."""

    def __init__(self, cin):
        super().__init__()
        self.conv = nn.Conv2d(cin, 1, kernel_size=1)


class CostDecoders2D(nn.Module):
    """`decoders1` / `decoders2` of the reference's MVS network (models/test_models.py:147-205), applied as in
    network/omni_mvsnet/pipeline3_model.py:866-905: `depth_d1(cost_reg)` = 1x1 convolution over the depth axis + x4 bilinear +
    rectification; `forward(image_features)` = the three ConvBlock2 of the mono-stereo fusion (x2 bilinear, WrapPadding, 3x3
    convolutions on the tensor cores as D == 1 volumes) -> (B, 1, 4H, 4W) fp32.  `out_channels` = 1 (no `mvs_uncertainty` head)."""

    def __init__(self, size=4, cost_volume_channels=64, wo_mono_feat=False, with_sin=False):
        super().__init__()
        self.decoders1 = _Conv1x1(cost_volume_channels)
        in_dim = cost_volume_channels + (0 if wo_mono_feat else 2 ** (size + 1)) + (1 if with_sin else 0)
        self.decoders2 = nn.ModuleList([_Block2D(in_dim, 2 ** (size + 1), True, True), _Block2D(2 ** (size + 1), 2 ** size, True, True),
                                        _Block2D(2 ** size, 1, False, False)])
        self._packed = {}
        self._ws_cache = {}
        for p in self.parameters():
            p.requires_grad_(False)

    def _pack(self, conv, ca_pad, simt=False):
        ver = (conv.weight._version, conv.bias._version, conv.weight.data_ptr(), str(conv.weight.device), ca_pad, simt)
        hit = self._packed.get(id(conv))
        if hit is None or hit[0] != ver:
            co, ci = conv.weight.shape[:2]
            w3 = torch.zeros((co, ci, 3, 3, 3), device=conv.weight.device, dtype=torch.float32)
            w3[:, :, 1] = conv.weight.detach().float()                 # a 2-D kernel is the kd == 1 slice of a 3-D one
            if simt:
                hit = (ver, pack_conv_cout1(w3, ci, 0, ca_pad, 0), float(conv.bias.detach().float().item()))
            else:
                hit = (ver,) + pack_conv(w3, conv.bias, ci, 0, ca_pad, 0)
            self._packed[id(conv)] = hit
        return hit[1], hit[2]

    def depth_d1(self, cost_reg, out_type="depth"):
        """cost_reg (B, D, H, W) fp32, any strides (e.g. `unet3d(x)[:, 0]`) -> raw or rectified (B, 4H, 4W, 1)
        (`out_type` in {"raw", "depth", "disparity"}; pipeline3_model.py:866-879)."""
        _lib.require_cuda(cost_reg)
        lib = _lib.load()
        x = cost_reg.detach().float()
        B, C, H, W = x.shape
        conv = self.decoders1.conv
        if C != conv.weight.shape[1]:
            raise RuntimeError(f"decoders1 expects {conv.weight.shape[1]} channels, got {C}")
        w = conv.weight.detach().float().reshape(-1).contiguous()
        out = torch.empty((B, 4 * H, 4 * W, 1), device=x.device, dtype=torch.float32)
        mode = {"raw": 0, "depth": 1, "disparity": 2}[out_type]
        with torch.cuda.device(x.device):
            _lib.check(lib.pgrf_channel_dot_upsample_fwd(_lib.ptr(x), *x.stride(), B, C, H, W, _lib.ptr(w), float(conv.bias.item()), 4, mode,
                                                         _lib.ptr(out), _lib.stream_ptr()), "pgrf_channel_dot_upsample_fwd")
        return out

    def forward(self, image_features):
        """image_features (B, C, H, W) fp32 = cat(regularised cost, mono features) -> (B, 1, 4H, 4W) fp32 (pipeline3_model.py:902-905)."""
        _lib.require_cuda(image_features)
        lib = _lib.load()
        x = image_features.detach().float()
        B, C, H, W = x.shape
        if C != self.decoders2[0].conv1.weight.shape[1]:
            raise RuntimeError(f"decoders2 expects {self.decoders2[0].conv1.weight.shape[1]} channels, got {C}")
        dev = x.device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            cpad = _pad16(C)
            a = torch.empty((B, 1, H, W, cpad), device=dev, dtype=torch.bfloat16)
            sb, sc, sh, sw = x.stride()
            _lib.check(lib.pgrf_conv3d_to_bf16_cl(_lib.ptr(x), sb, sc, 0, sh, sw, B, C, 1, H, W, cpad, _lib.ptr(a), st), "pgrf_conv3d_to_bf16_cl")
            for blk in self.decoders2[:2]:
                up = torch.empty((B, 1, 2 * H, 2 * W, a.shape[-1]), device=dev, dtype=torch.bfloat16)
                _lib.check(lib.pgrf_upsample2d2_fwd(_lib.ptr(a), B, H, W, a.shape[-1], _lib.ptr(up), st), "pgrf_upsample2d2_fwd")
                H, W = 2 * H, 2 * W
                co_pad = _pad16(blk.conv1.weight.shape[0])
                w1, b1 = self._pack(blk.conv1, up.shape[-1])
                t = conv3d(up, None, w1, b1, co_pad, (B, 1, H, W), 0, self._ws_cache, lib, st)
                w2, b2 = self._pack(blk.conv2, co_pad)
                a = conv3d(t, None, w2, b2, co_pad, (B, 1, H, W), 0, self._ws_cache, lib, st)
            blk = self.decoders2[2]                                     # 2^size -> 1 -> 1, no activation, fp32
            w1, b1 = self._pack(blk.conv1, a.shape[-1])
            t = conv3d(a, None, w1, b1, 16, (B, 1, H, W), 1, self._ws_cache, lib, st, act=False)           # (B,1,1,H,W) fp32
            w2, b2 = self._pack(blk.conv2, 1, simt=True)
            out = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
            _lib.check(lib.pgrf_conv3d_cout1_fwd(None, 0, None, 0, _lib.ptr(t), _lib.ptr(w2), b2, B, 1, H, W, 0, _lib.ptr(out), st),
                       "pgrf_conv3d_cout1_fwd")
        return out
