"""GPU: the tensor-core 3-D cost regulariser (csrc/conv3d.cu through the C ABI) against the reference-generated goldens, and each of
its kernels against the oracle's blocks on identically rounded operands."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402
from test_oracle_regulariser import golden_weights  # noqa: E402

from oracle import regulariser as oreg  # noqa: E402

pytestmark = pytest.mark.gpu

#: bf16 operands / activations through 14 convolution layers, fp32 accumulation (the reference's cuDNN path is TF32-class): deviation
#: of the regularised cost relative to its range
UNET_TOL = 2e-2
#: one layer on identically rounded operands: accumulation order + one bf16 rounding of the output (2^-9)
LAYER_TOL = 4e-3


def _cl(x, cpad):
    """(B,C,D,H,W) fp32 -> bf16 channels-last with zero padded channels"""
    B, C, D, H, W = x.shape
    y = torch.zeros((B, D, H, W, cpad), device=x.device, dtype=torch.bfloat16)
    y[..., :C] = x.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return y


@pytest.mark.parametrize("name", list(cases.UNET3D_CASES))
def test_unet3d_matches_reference_golden(name):
    from panogrf_b200 import regulariser as reg
    g = load_golden(name)
    W = golden_weights(name, g)
    size, _ = cases.UNET3D_CASES[name]
    net = reg.CostRegulariser3D(size)
    net.load_state_dict(W)
    net = net.cuda()
    y = net(g["x"].cuda()).cpu()
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    scale = float(g["y"].abs().max())
    err = float((y - g["y"]).abs().max())
    print(f"{name}: max err {err:.3e} of range {scale:.3e} ({err / scale:.2e})")
    assert err <= UNET_TOL * scale
    # the strided (permuted) view the cost-volume sweep returns is consumed in place
    xs = g["x"].cuda().permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4)
    assert not xs.is_contiguous()
    assert torch.equal(net(xs).cpu(), y)


@pytest.mark.parametrize("ca,cb,co,dims", [(16, 0, 16, (1, 4, 8, 16)), (32, 0, 64, (2, 4, 6, 10)), (64, 64, 64, (1, 4, 8, 8)),
                                           (128, 0, 256, (1, 2, 4, 8)), (256, 0, 128, (1, 2, 5, 7)), (4, 0, 8, (1, 3, 5, 9)),
                                           (48, 16, 32, (1, 2, 4, 8)),
                                           # W % 128 == 0: the row variant (one gather serves the three kw taps)
                                           (32, 0, 64, (1, 3, 3, 128)), (64, 64, 16, (2, 2, 2, 256)), (128, 0, 128, (1, 2, 2, 128)),
                                           (16, 0, 32, (1, 2, 3, 128))])
def test_conv3d_layer_vs_oracle_block(ca, cb, co, dims):
    from panogrf_b200 import _lib
    from panogrf_b200 import regulariser as reg
    lib = _lib.load()
    B, D, H, W = dims
    torch.manual_seed(ca + cb + co)
    xa = torch.randn(B, ca, D, H, W, device="cuda")
    xb = torch.randn(B, cb, D, H, W, device="cuda") if cb else None
    w = torch.randn(co, ca + cb, 3, 3, 3, device="cuda") / (27 * (ca + cb)) ** 0.5
    b = torch.randn(co, device="cuda")
    ca_pad, cb_pad, co_pad = reg._pad16(ca), reg._pad16(cb) if cb else 0, reg._pad16(co)
    a_cl, b_cl = _cl(xa, ca_pad), (_cl(xb, cb_pad) if cb else None)
    wpk, bp = reg.pack_conv(w, b, ca, cb, ca_pad, cb_pad)
    y = reg.conv3d(a_cl, b_cl, wpk, bp, co_pad, (B, D, H, W))
    torch.cuda.synchronize()
    xin = torch.cat([xa] + ([xb] if cb else []), 1).to(torch.bfloat16).float().cpu()
    ref = F.leaky_relu(F.conv3d(oreg.wrap_pad3d(xin), w.to(torch.bfloat16).float().cpu(), b.cpu()), 0.01)
    got = y[..., :co].float().permute(0, 4, 1, 2, 3).cpu()
    assert float(y[..., co:].float().abs().sum()) == 0
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= LAYER_TOL * scale


def test_cout1_pool_upsample_convert_vs_torch():
    from panogrf_b200 import _lib
    from panogrf_b200 import regulariser as reg
    lib = _lib.load()
    st = _lib.stream_ptr()
    torch.manual_seed(3)
    B, C, D, H, W = 2, 24, 4, 6, 8
    x = torch.randn(B, C, D, H, W, device="cuda")
    # strided fp32 -> bf16 channels-last (planar (B,D,C,H,W) storage, as the sweep returns it)
    xs = x.permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4)
    cpad = 32
    a = torch.empty((B, D, H, W, cpad), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.pgrf_conv3d_to_bf16_cl(_lib.ptr(xs), *xs.stride(), B, C, D, H, W, cpad, _lib.ptr(a), st), "to_bf16_cl")
    assert torch.equal(a, _cl(x, cpad))
    xr = a[..., :C].float().permute(0, 4, 1, 2, 3)
    # AvgPool3d(2)
    p = torch.empty((B, D // 2, H // 2, W // 2, cpad), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.pgrf_avgpool3d2_fwd(_lib.ptr(a), B, D, H, W, cpad, _lib.ptr(p), st), "avgpool")
    ref = F.avg_pool3d(xr, 2)
    assert float((p[..., :C].float().permute(0, 4, 1, 2, 3) - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    # trilinear x2
    u = torch.empty((B, 2 * D, 2 * H, 2 * W, cpad), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.pgrf_upsample3d2_fwd(_lib.ptr(a), B, D, H, W, cpad, _lib.ptr(u), st), "upsample")
    ref = F.interpolate(xr, scale_factor=2, mode="trilinear", align_corners=False)
    assert float((u[..., :C].float().permute(0, 4, 1, 2, 3) - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    # single-output-channel convolutions (bf16 pair input, then fp32 single channel)
    xb = _cl(torch.randn(B, 16, D, H, W, device="cuda"), 16)
    w = torch.randn(1, C + 16, 3, 3, 3, device="cuda") / 30
    w1 = reg.pack_conv_cout1(w, C, 16, cpad, 16)
    t = torch.empty((B, D, H, W), device="cuda")
    _lib.check(lib.pgrf_conv3d_cout1_fwd(_lib.ptr(a), cpad, _lib.ptr(xb), 16, None, _lib.ptr(w1), 0.25, B, D, H, W, 1, _lib.ptr(t), st),
               "cout1")
    xin = torch.cat([xr, xb.float().permute(0, 4, 1, 2, 3)], 1).cpu()
    ref = F.leaky_relu(F.conv3d(oreg.wrap_pad3d(xin), w.cpu(), torch.tensor([0.25])), 0.01)[:, 0]
    assert torch.allclose(t.cpu(), ref, rtol=1e-4, atol=1e-5)
    w2 = torch.randn(1, 1, 3, 3, 3, device="cuda")
    o = torch.empty_like(t)
    _lib.check(lib.pgrf_conv3d_cout1_fwd(None, 0, None, 0, _lib.ptr(t), _lib.ptr(reg.pack_conv_cout1(w2, 1, 0, 1, 0)), -0.5, B, D, H, W,
                                         1, _lib.ptr(o), st), "cout1 fp32")
    ref2 = F.leaky_relu(F.conv3d(oreg.wrap_pad3d(ref[:, None]), w2.cpu(), torch.tensor([-0.5])), 0.01)[:, 0]
    assert torch.allclose(o.cpu(), ref2, rtol=1e-4, atol=1e-5)


def test_row_variant_equals_per_tap_variant():
    """Same layer through both kernels (debug switch): identical operands and accumulation order per tap -> equal results."""
    from panogrf_b200 import _lib
    from panogrf_b200 import regulariser as reg
    lib = _lib.load()
    torch.manual_seed(11)
    B, D, H, W, ca, co = 1, 4, 4, 128, 64, 64
    a = _cl(torch.randn(B, ca, D, H, W, device="cuda"), ca)
    wpk, bp = reg.pack_conv(torch.randn(co, ca, 3, 3, 3, device="cuda") / 40, torch.randn(co, device="cuda"), ca, 0, ca, 0)
    outs = {}
    try:
        for row in (1, 0):
            for splits in (1, 3, 9):
                _lib.check(lib.pgrf_debug_set(b"conv_row", row), "debug_set")
                _lib.check(lib.pgrf_debug_set(b"conv_splits", splits), "debug_set")
                outs[row, splits] = reg.conv3d(a, None, wpk, bp, co, (B, D, H, W))
    finally:
        lib.pgrf_debug_set(b"conv_row", 1)
        lib.pgrf_debug_set(b"conv_splits", 0)
    for splits in (1, 3, 9):
        assert torch.equal(outs[1, splits], outs[0, splits])
    # the non-persistent kernels (one CTA per tile) accumulate in the same order: bit-identical
    try:
        _lib.check(lib.pgrf_debug_set(b"conv_persist", 0), "debug_set")
        for row in (1, 0):
            _lib.check(lib.pgrf_debug_set(b"conv_row", row), "debug_set")
            _lib.check(lib.pgrf_debug_set(b"conv_splits", 1), "debug_set")
            assert torch.equal(reg.conv3d(a, None, wpk, bp, co, (B, D, H, W)), outs[1, 1])
    finally:
        lib.pgrf_debug_set(b"conv_persist", 1)
        lib.pgrf_debug_set(b"conv_row", 1)
        lib.pgrf_debug_set(b"conv_splits", 0)
    # split-K changes the fp32 summation order of the 9 tap rows: equal to bf16 rounding, not bitwise
    ref = outs[1, 1].float()
    for splits in (3, 9):
        assert float((outs[1, splits].float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())


def test_unet3d_rejects_bad_shapes():
    from panogrf_b200 import regulariser as reg
    net = reg.CostRegulariser3D(1).cuda()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 4, 8, 8, 12, device="cuda"))       # W not a multiple of 8
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 8, 8, 8, 16, device="cuda"))       # wrong channel count


@pytest.mark.parametrize("name", list(cases.DEC2D_CASES))
def test_decoders_2d_match_reference_golden(name):
    """decoders1 (fp32: 1x1 convolution + x4 bilinear + rectification) and decoders2 (bf16 tensor-core 3x3 convolutions as D == 1
    volumes, bilinear x2 upscaling) against outputs of the reference's ConvBlock / ConvBlock2 classes."""
    from panogrf_b200 import regulariser as reg
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    size, D, _, out_type = cases.DEC2D_CASES[name]
    net = reg.CostDecoders2D(size, D)
    net.load_state_dict(W)
    net = net.cuda()
    # decoders1 on the strided view unet3d returns ((B,1,D,H,W)[:, 0])
    cost5 = g["cost_reg"].cuda()[:, None]
    raw = net.depth_d1(cost5[:, 0], "raw").cpu()
    assert torch.allclose(raw, g["raw_d1"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(net.depth_d1(cost5[:, 0], out_type).cpu(), g["depth_d1"], rtol=1e-3, atol=1e-5)
    feats = net(torch.cat((g["cost_reg"], g["mono"]), 1).cuda()).cpu()
    assert feats.shape == g["feats"].shape
    scale = float(g["feats"].abs().max())
    err = float((feats - g["feats"]).abs().max())
    print(f"{name}: decoders2 max err {err:.3e} of range {scale:.3e} ({err / scale:.2e})")
    assert err <= UNET_TOL * scale
