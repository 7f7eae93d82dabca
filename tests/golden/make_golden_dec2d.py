"""Golden fixtures of the 2-D heads that follow the 3-D regulariser, produced by the REFERENCE classes (models/common_blocks.py:
ConvBlock, ConvBlock2) built exactly as models/test_models.py:147-205 does and applied as pipeline3_model.py:866-905 does, with
seeded weights, on CPU.   python tests/golden/make_golden_dec2d.py"""
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


def build_reference_decoders(size, cost_volume_channels, out_channels=1, use_wrap_padding=True, use_v_input=False):
    """models/test_models.py:147-205 (wo_mono_feat False, with_sin False)"""
    from models.common_blocks import ConvBlock, ConvBlock2
    decoders1 = ConvBlock(cost_volume_channels, 1, kernel_size=1, padding=0, stride=1, upscale=False, gate=False,
                          use_wrap_padding=False, use_batch_norm=False, use_activation=False)
    in_dim = cost_volume_channels + 2 ** (size + 1)
    kw = dict(kernel_size=3, stride=1, padding=1, use_wrap_padding=use_wrap_padding, use_residual=False, pooling=False,
              use_v_input=use_v_input)
    decoders2 = nn.ModuleList([
        ConvBlock2(in_channels=in_dim, out_channels=2 ** (size + 1), use_activation=True, upscale=True, **kw),
        ConvBlock2(in_channels=2 ** (size + 1), out_channels=2 ** size, use_activation=True, upscale=True, **kw),
        ConvBlock2(in_channels=2 ** size, out_channels=out_channels, use_activation=False, upscale=False, **kw)])
    return decoders1, decoders2


if __name__ == "__main__":
    for name, (size, D, _, out_type) in cases.DEC2D_CASES.items():
        torch.manual_seed(sum(map(ord, name)))
        d1, d2 = build_reference_decoders(size, D)
        cost_reg, mono = cases.make_dec2d_inputs(name)
        with torch.no_grad():
            # pipeline3_model.py:866-879
            raw = nn.functional.interpolate(d1(cost_reg), scale_factor=4, mode="bilinear", align_corners=False).permute((0, 2, 3, 1))
            depth_d1 = 1.0 / (torch.clamp(raw, min=0) + 1e-10) if out_type == "disparity" else torch.clamp(raw, min=0)
            # :884-905 (contain_dnet / UniFuse branch: cat(cost, mono features)), decoders2 loop
            feats = torch.cat((cost_reg, mono), 1)
            for blk in d2:
                feats, _ = blk(feats)
            pred = feats[:, :1]
            depth = (1.0 / (torch.clamp(pred, min=0) + 1e-10) if out_type == "disparity" else torch.clamp(pred, min=0)).permute((0, 2, 3, 1))
        blob = {"cost_reg": cost_reg.numpy(), "mono": mono.numpy(), "raw_d1": raw.numpy(), "depth_d1": depth_d1.numpy(),
                "feats": feats.numpy(), "depth": depth.numpy()}
        for k, v in d1.state_dict().items():
            blob["w.decoders1." + k] = v.numpy()
        for k, v in d2.state_dict().items():
            blob["w.decoders2." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, tuple(cost_reg.shape), "->", tuple(feats.shape), sorted(k for k in blob if k.startswith("w."))[:4])
