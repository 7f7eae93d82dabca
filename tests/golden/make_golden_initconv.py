"""Golden fixtures of CostVolumeInitNet's convolution stacks (`depth_conv`, `out_conv`: network/init_net.py:540-574, applied at
:629-636), built from the REFERENCE's own layer constructors (network/ops.py: conv3x3, ResidualBlock, conv1x1) with seeded weights.
python tests/golden/make_golden_initconv.py"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import _refimport  # noqa: E402

_refimport.install()
import cases  # noqa: E402


class Holder(nn.Module):
    """the two nn.Sequential of CostVolumeInitNet.__init__ (init_net.py:540-574), verbatim constructor calls"""

    def __init__(self, use_wrap_padding):
        super().__init__()
        from network.ops import ResidualBlock, conv1x1, conv3x3
        norm_layer = lambda dim: nn.InstanceNorm2d(dim, track_running_stats=False, affine=True)
        in_dim, depth_dim = 32 + 32, 32
        self.depth_conv = nn.Sequential(
            conv3x3(1, depth_dim, use_wrap_padding=use_wrap_padding),
            ResidualBlock(depth_dim, depth_dim, norm_layer=norm_layer, use_wrap_padding=use_wrap_padding),
            conv1x1(depth_dim, depth_dim, use_wrap_padding=use_wrap_padding))
        self.out_conv = nn.Sequential(
            conv3x3(in_dim, 32, use_wrap_padding=use_wrap_padding),
            ResidualBlock(32, 32, norm_layer=norm_layer, use_wrap_padding=use_wrap_padding),
            conv1x1(32, 32, use_wrap_padding=use_wrap_padding))


if __name__ == "__main__":
    for name, (wrap, n, hw) in cases.INITCONV_CASES.items():
        torch.manual_seed(sum(map(ord, name)))
        net = Holder(wrap).eval()
        with torch.no_grad():
            for k, p in net.named_parameters():
                if p.dim() == 1:
                    p.copy_(torch.randn_like(p) * 0.3 + (1.0 if k.endswith("weight") else 0.0))
        ref_feats, depth = cases.make_initconv_inputs(name)
        with torch.no_grad():
            depth_feats = net.depth_conv(depth)                                   # init_net.py:629
            ray_feats = net.out_conv(torch.cat([ref_feats, depth_feats], 1))      # :636
        blob = {"ref_feats": ref_feats.numpy(), "depth": depth.numpy(), "depth_feats": depth_feats.numpy(), "ray_feats": ray_feats.numpy()}
        for k, v in net.state_dict().items():
            blob["w." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, tuple(ref_feats.shape), "->", tuple(ray_feats.shape))
