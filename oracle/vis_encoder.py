"""CPU restatement of the reference's DefaultVisEncoder (TEST INFRASTRUCTURE, never imported by the product).

Reference: network/vis_encoder.py:6-33 — cat(F.interpolate(img_feats, size of ray_feats, 'bilinear'), ray_feats) -> conv3x3 (no bias)
-> 2 x ResidualBlock(InstanceNorm2d(affine), ReLU, conv3x3, InstanceNorm2d, ReLU, conv3x3; + skip) -> conv1x1 (no bias), with
WrapPadding (zeros along height, wrap along width; models... network/ops.py:6-29, 61-115) or plain zero padding.
Pinned by tests/golden/visenc_*.npz: outputs of the reference class with seeded weights (tests/golden/make_golden_visenc.py).
Weights are a state_dict with the reference's names (`out_conv.0.1.weight` with wrap padding, `out_conv.0.weight` without, ...).
"""
import torch
import torch.nn.functional as F


def _pad(x, wrap):
    if wrap:
        x = F.pad(x, (0, 0, 1, 1))
        return torch.cat([x[..., -1:], x, x[..., :1]], -1)
    return F.pad(x, (1, 1, 1, 1))


def _inorm_relu(x, g, b):
    mean = x.mean((2, 3), keepdim=True)
    var = x.var((2, 3), unbiased=False, keepdim=True)
    return F.relu((x - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1))


def vis_encoder(W, ray_feats, img_feats, wrap=True):
    """W: state_dict of DefaultVisEncoder; ray_feats (N,32,h,w), img_feats (N,32,hi,wi) -> (N,32,h,w)"""
    if img_feats.shape[2:] != ray_feats.shape[2:]:
        img_feats = F.interpolate(img_feats, ray_feats.shape[2:], mode="bilinear")
    x = torch.cat([img_feats, ray_feats], 1)
    c0 = "out_conv.0.1.weight" if wrap else "out_conv.0.weight"
    x = F.conv2d(_pad(x, wrap), W[c0])
    for blk in (1, 2):
        i = (0, 3, 4, 7) if wrap else (0, 2, 3, 5)          # Sequential indices: norm, conv, norm, conv (WrapPadding modules in between)
        p = f"out_conv.{blk}.conv."
        t = _inorm_relu(x, W[p + f"{i[0]}.weight"], W[p + f"{i[0]}.bias"])
        t = F.conv2d(_pad(t, wrap), W[p + f"{i[1]}.weight"])
        t = _inorm_relu(t, W[p + f"{i[2]}.weight"], W[p + f"{i[2]}.bias"])
        t = F.conv2d(_pad(t, wrap), W[p + f"{i[3]}.weight"])
        x = x + t
    return F.conv2d(x, W["out_conv.3.weight"])


def conv_stack(W, prefix, x, n_blocks, wrap=True):
    """`conv3x3 -> n x ResidualBlock -> conv1x1` (network/init_net.py:540-574) with the reference's state_dict names under `prefix`."""
    c0 = f"{prefix}.0.1.weight" if wrap else f"{prefix}.0.weight"
    x = F.conv2d(_pad(x, wrap), W[c0])
    i = (0, 3, 4, 7) if wrap else (0, 2, 3, 5)
    for blk in range(1, n_blocks + 1):
        p = f"{prefix}.{blk}.conv."
        t = _inorm_relu(x, W[p + f"{i[0]}.weight"], W[p + f"{i[0]}.bias"])
        t = F.conv2d(_pad(t, wrap), W[p + f"{i[1]}.weight"])
        t = _inorm_relu(t, W[p + f"{i[2]}.weight"], W[p + f"{i[2]}.bias"])
        t = F.conv2d(_pad(t, wrap), W[p + f"{i[3]}.weight"])
        x = x + t
    return F.conv2d(x, W[f"{prefix}.{n_blocks + 1}.weight"])


def init_net_convs(W, ref_feats, depth, wrap=True):
    """network/init_net.py:629-636: depth_feats = depth_conv(depth); ray_feats = out_conv(cat(ref_feats, depth_feats))"""
    depth_feats = conv_stack(W, "depth_conv", depth, 1, wrap)
    return conv_stack(W, "out_conv", torch.cat([ref_feats, depth_feats], 1), 1, wrap)
