"""Golden fixtures of the image encoder, produced by the REFERENCE class (network/ops.py: ResUNetLight, constructed like
network/renderer.py:106) with seeded weights on CPU.   python tests/golden/make_golden_resunet.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import _refimport  # noqa: E402

_refimport.install()
import cases  # noqa: E402

if __name__ == "__main__":
    from network.ops import ResUNetLight
    for name, (wrap, n, hw) in cases.RESUNET_CASES.items():
        torch.manual_seed(sum(map(ord, name)))
        cfg = {"handle_distort": False, "handle_distort_input_all": False}
        net = ResUNetLight(cfg, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=wrap).eval()
        with torch.no_grad():
            for k, p in net.named_parameters():          # non-trivial affine parameters of the instance norms
                if p.dim() == 1 and ("bn" in k or "downsample.1" in k):
                    p.copy_(torch.randn_like(p) * 0.3 + (1.0 if k.endswith("weight") else 0.0))
        x = cases.make_resunet_input(name)
        with torch.no_grad():
            y = net(x)
        blob = {"x": x.numpy(), "y": y.numpy()}
        # 2 M parameters: not stored.  The product's container re-creates them from the seed with the same construction order and the
        # same affine-parameter loop (tests/test_oracle_res_unet.py); per-tensor checksums make any divergence loud.
        for k, v in net.state_dict().items():
            blob["sum." + k] = np.asarray([float(v.double().sum()), float(v.double().abs().sum())])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, tuple(x.shape), "->", tuple(y.shape), "params", sum(v.numel() for v in net.state_dict().values()))
