"""Golden fixture of MixtureLogisticsDistDecoder.compute_prob with is_ref=False (network/dist_decoder.py:6-51,109-140), produced by the
REFERENCE class on CPU.   python tests/golden/make_golden_probque.py"""
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

if __name__ == "__main__":
    from network.dist_decoder import MixtureLogisticsDistDecoder
    g = torch.Generator().manual_seed(77)
    qn, rn, dn = 2, 37, 16
    blob = {}
    for tag, use_vis, per_sample in (("a", True, False), ("b", False, True)):
        dec = MixtureLogisticsDistDecoder({"use_vis": use_vis})
        depth = torch.sort(0.6 + 9.0 * torch.rand(qn, rn, dn, generator=g), -1)[0]
        depth_range = torch.tensor([[0.5, 10.0], [0.4, 12.0]])
        inv = (-1 / depth + 1 / depth_range[:, 0, None, None]) / (-1 / depth_range[:, 1, None, None] + 1 / depth_range[:, 0, None, None])
        interval = torch.cat([inv[..., 1:] - inv[..., :-1], torch.full((qn, rn, 1), 0.02)], -1)
        k = dn if per_sample else 1
        mean = torch.rand(qn, rn, k, 2, generator=g)
        var = 2.0 + 30.0 * torch.rand(qn, rn, k, 2, generator=g)
        vis = torch.rand(qn, rn, k, 1, generator=g)
        aw = torch.rand(qn, rn, k, 1, generator=g)
        with torch.no_grad():
            alpha, visibility, hit = dec.compute_prob(depth, interval, mean, var, vis, aw, False, depth_range)
        for n, v in dict(depth=depth, interval=interval, mean=mean, var=var, vis=vis, aw=aw, depth_range=depth_range, alpha=alpha,
                         visibility=visibility, hit_prob=hit).items():
            blob[f"{tag}.{n}"] = v.numpy()
        blob[f"{tag}.use_vis"] = np.asarray(use_vis)
    np.savez_compressed(os.path.join(HERE, "prob_que.npz"), **blob)
    print("prob_que", {k: v.shape for k, v in blob.items() if k.endswith("alpha")})
