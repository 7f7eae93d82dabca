"""Times the per-view CNNs of the render path (image encoder ResUNetLight, DefaultVisEncoder) at the benched size and compares the
image encoder with the same network through torch's library convolutions (the reference's path).  python tools/time_encoders.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from panogrf_b200 import _lib  # noqa: E402
from panogrf_b200.image_encoder import ResUNetLight  # noqa: E402
from time_regulariser import time_fn  # noqa: E402


def torch_resunet(net, x):
    """the same parameters through F.conv2d / F.instance_norm (cuDNN), as the reference's nn.Module runs them"""
    wrap = net.use_wrap_padding

    def pad(t, p, w):
        if w:
            t = F.pad(t, (0, 0, p, p))
            return torch.cat([t[..., -p:], t, t[..., :p]], -1)
        return F.pad(t, (p, p, p, p))

    inorm = lambda t, m: F.instance_norm(t, weight=m.weight, bias=m.bias, eps=m.eps)
    c3 = lambda t, m, s, w: F.conv2d(pad(t, 1, w), (m[1] if w else m).weight, (m[1] if w else m).bias, stride=s)
    x0 = F.relu(inorm(F.conv2d(pad(x, 3, wrap), (net.conv1[1] if wrap else net.conv1).weight, stride=2), net.bn1))
    feats, cur = [], x0
    for layer in (net.layer1, net.layer2, net.layer3):
        for blk in layer:
            t = F.relu(inorm(c3(cur, blk.conv1, blk.stride, wrap), blk.bn1))
            t = inorm(c3(t, blk.conv2, 1, wrap), blk.bn2)
            idn = cur if blk.downsample is None else inorm(F.conv2d(cur, blk.downsample[0].weight, stride=2), blk.downsample[1])
            cur = F.relu(t + idn)
        feats.append(cur)
    x1, x2, x3 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    cm = lambda m, t: F.elu(inorm(c3(t, m.conv, 1, m.wrap), m.bn))
    t = cm(net.iconv3, torch.cat([cm(net.upconv3.conv, up(x3)), x2], 1))
    t = cm(net.iconv2, torch.cat([cm(net.upconv2.conv, up(t)), x1], 1))
    return F.conv2d(t, net.out_conv.weight, net.out_conv.bias)


if __name__ == "__main__":
    torch.manual_seed(0)
    net = ResUNetLight({}, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=True).cuda()
    x = torch.rand(2, 3, 512, 1024, device="cuda")
    l0 = _lib.launch_count()
    y = net(x)
    launches = _lib.launch_count() - l0
    ms = time_fn(lambda: net(x), 5)
    with torch.no_grad():
        ref = torch_resunet(net, x)
        ms_lib = time_fn(lambda: torch_resunet(net, x), 5)
    err = float((y - ref).abs().max() / ref.abs().max())
    rms = float((y - ref).pow(2).mean().sqrt() / ref.abs().max())
    print(f"ResUNetLight 2x3x512x1024 -> {tuple(y.shape)}: {ms:.3f} ms ({launches} launches); torch/cuDNN fp32 (TF32): {ms_lib:.3f} ms; "
          f"max err {err:.2e}, rms {rms:.2e} of range")
    from panogrf_b200.graphs import GraphedForward
    from panogrf_b200.vis_encoder import DefaultVisEncoder
    from panogrf_b200 import regulariser as reg
    g = GraphedForward(net)
    yg = g(x)
    print(f"  replayed from a CUDA graph: {time_fn(lambda: g(x), 5):.3f} ms (bit-identical: {bool(torch.equal(yg, y))})")
    vis = DefaultVisEncoder({"use_wrap_padding": True}).cuda()
    rf, imf = torch.randn(2, 32, 64, 128, device="cuda"), torch.randn(2, 32, 128, 256, device="cuda")
    gv = GraphedForward(vis)
    print(f"DefaultVisEncoder 2 x (32 + 32) x 128x256: eager {time_fn(lambda: vis(rf, imf), 5):.3f} ms, graph {time_fn(lambda: gv(rf, imf), 5):.3f} ms "
          f"(bit-identical: {bool(torch.equal(gv(rf, imf), vis(rf, imf)))})")
    unet = reg.CostRegulariser3D(4).cuda()
    vol = torch.randn(1, 32, 64, 64, 128, device="cuda")
    gu = GraphedForward(unet)
    print(f"CostRegulariser3D 1x32x64x64x128: eager {time_fn(lambda: unet(vol), 5):.3f} ms, graph {time_fn(lambda: gu(vol), 5):.3f} ms "
          f"(bit-identical: {bool(torch.equal(gu(vol), unet(vol)))})")
