"""Golden fixtures of the 3-D cost regulariser, produced by the REFERENCE classes (models/common_blocks.py: UNet2, Conv3DBlockv2)
wired exactly as models/test_models.py:81-146 does, with seeded weights, on CPU.   python tests/golden/make_golden_unet3d.py"""
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


def build_reference_unet3d(size, num_layer=3, use_wrap_padding=True, use_v_input=False):
    """models/test_models.py:81-146 (the non-`use_new_reg3dnet` branch) with the reference's own block classes."""
    from models.common_blocks import Conv3DBlockv2, UNet2
    kw = dict(kernel_size=(3, 3, 3), stride=(1, 1, 1), padding=(1, 1, 1), use_batch_norm=False, use_wrap_padding=use_wrap_padding,
              use_v_input=use_v_input)
    enc, dec = [], [Conv3DBlockv2(in_channels=2 ** (size + 3), out_channels=1, pooling=nn.Identity(), **kw)]
    for i in range(num_layer):
        ch = 2 ** (i + size + 1)
        enc.append(Conv3DBlockv2(in_channels=ch, out_channels=2 * ch, **kw))
        if i > 0:
            dec.append(Conv3DBlockv2(in_channels=4 * ch, out_channels=ch, pooling=nn.Identity(), **kw))
    enc.append(Conv3DBlockv2(in_channels=2 ** (num_layer + size + 1), out_channels=2 ** (num_layer + size + 2), pooling=nn.Identity(), **kw))
    return UNet2(nn.ModuleList(enc), nn.ModuleList(dec), interpolation="trilinear", name="unet3d")


if __name__ == "__main__":
    for name in cases.UNET3D_CASES:
        size, shape = cases.UNET3D_CASES[name]
        torch.manual_seed(sum(map(ord, name)))
        net = build_reference_unet3d(size).eval()
        x = cases.make_unet3d_input(name)
        with torch.no_grad():
            y = net(x)
        blob = {"x": x.numpy(), "y": y.numpy()}
        if cases.UNET3D_STORE_WEIGHTS[name]:
            for k, v in net.state_dict().items():
                blob["w." + k] = v.numpy()
        else:   # large case: the weights are re-created from the seed by the same construction order; checksums make a mismatch loud
            for k, v in net.state_dict().items():
                blob["sum." + k] = np.asarray([float(v.double().sum()), float(v.double().abs().sum())])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, tuple(x.shape), "->", tuple(y.shape), "params", sum(v.numel() for v in net.state_dict().values()))
