"""Golden fixtures of DefaultVisEncoder, produced by the REFERENCE class (network/vis_encoder.py) with seeded weights on CPU.
python tests/golden/make_golden_visenc.py"""
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
    from network.vis_encoder import DefaultVisEncoder
    for name, (wrap, n, (h, w), (hi, wi)) in cases.VISENC_CASES.items():
        torch.manual_seed(sum(map(ord, name)))
        net = DefaultVisEncoder({"use_wrap_padding": wrap}).eval()
        with torch.no_grad():
            for k, p in net.named_parameters():          # non-trivial affine parameters of the instance norms
                if p.dim() == 1:
                    p.copy_(torch.randn_like(p) * 0.3 + (1.0 if k.endswith("weight") else 0.0))
        ray, img = cases.make_visenc_inputs(name)
        with torch.no_grad():
            out = net(ray, img)
        blob = {"ray_feats": ray.numpy(), "img_feats": img.numpy(), "out": out.numpy()}
        for k, v in net.state_dict().items():
            blob["w." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, tuple(ray.shape), tuple(img.shape), "->", tuple(out.shape), sorted(net.state_dict())[:3])
