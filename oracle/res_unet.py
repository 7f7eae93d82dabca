"""CPU restatement of the reference's image encoder ResUNetLight (TEST INFRASTRUCTURE, never imported by the product).

Reference: network/ops.py:235-455 as built by network/renderer.py:106 (`ResUNetLight(cfg, 3, [1, 2, 6, 4], 32, inplanes=16,
use_wrap_padding=...)`, handle_distort off, no autoencoder): conv1 (7x7, stride 2) - InstanceNorm - ReLU, three stages of
BasicBlocks (ops.py:126-197; the first block of a stage has stride 2 and a 1x1 stride-2 downsample branch), upconv3 / iconv3 /
upconv2 / iconv2 (`conv` = conv3x3 + bias + InstanceNorm + ELU, ops.py:199-224; `upconv` = x2 bilinear align_corners=True + `conv`
WITHOUT wrap padding, ops.py:226-233), skip concatenations, out_conv 1x1.
Pinned by tests/golden/resunet_*.npz (outputs of the reference class with seeded weights, tests/golden/make_golden_resunet.py).
"""
import torch
import torch.nn.functional as F


def _pad(x, p, wrap):
    if wrap:
        x = F.pad(x, (0, 0, p, p))
        return torch.cat([x[..., -p:], x, x[..., :p]], -1)
    return F.pad(x, (p, p, p, p))


def _inorm(x, W, prefix):
    mean = x.mean((2, 3), keepdim=True)
    var = x.var((2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-5) * W[prefix + ".weight"].view(1, -1, 1, 1) + W[prefix + ".bias"].view(1, -1, 1, 1)


def _conv3x3(W, name, x, stride, wrap):
    key = f"{name}.1.weight" if wrap else f"{name}.weight"
    return F.conv2d(_pad(x, 1, wrap), W[key], stride=stride)


def _basic_block(W, p, x, stride, wrap):
    out = F.relu(_inorm(_conv3x3(W, p + ".conv1", x, stride, wrap), W, p + ".bn1"))
    out = _inorm(_conv3x3(W, p + ".conv2", out, 1, wrap), W, p + ".bn2")
    identity = x
    if (p + ".downsample.0.weight") in W:
        identity = _inorm(F.conv2d(x, W[p + ".downsample.0.weight"], stride=stride), W, p + ".downsample.1")
    return F.relu(out + identity)


def _conv_module(W, p, x, wrap):
    """`conv` (ops.py:199-224): [WrapPadding] conv3x3 + bias, InstanceNorm, ELU"""
    key = p + ".conv.1" if wrap else p + ".conv"
    y = F.conv2d(_pad(x, 1, wrap), W[key + ".weight"], W[key + ".bias"])
    return F.elu(_inorm(y, W, p + ".bn"))


def res_unet_light(W, x, layers=(1, 2, 6), wrap=True):
    """x (N,3,H,W) -> (N,32,H/4,W/4) (H, W multiples of 16 so that the skip connections need no padding)"""
    key = "conv1.1.weight" if wrap else "conv1.weight"
    x0 = F.relu(_inorm(F.conv2d(_pad(x, 3, wrap), W[key], stride=2), W, "bn1"))
    feats = [x0]
    cur = x0
    for li, nb in enumerate(layers, 1):
        for b in range(nb):
            cur = _basic_block(W, f"layer{li}.{b}", cur, 2 if b == 0 else 1, wrap)
        feats.append(cur)
    _, x1, x2, x3 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    x = _conv_module(W, "upconv3.conv", up(x3), False)                 # upconv never passes use_wrap_padding on (ops.py:229)
    assert x.shape[2:] == x2.shape[2:], "skipconnect padding (ops.py:370-381) is not restated: use sizes divisible by 16"
    x = _conv_module(W, "iconv3", torch.cat([x, x2], 1), wrap)
    x = _conv_module(W, "upconv2.conv", up(x), False)
    assert x.shape[2:] == x1.shape[2:]
    x = _conv_module(W, "iconv2", torch.cat([x, x1], 1), wrap)
    return F.conv2d(x, W["out_conv.weight"], W["out_conv.bias"])
