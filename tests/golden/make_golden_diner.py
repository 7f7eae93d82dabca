"""Golden fixtures for the depth-prior sample placement, produced by the REAL reference on CPU.

`NeuralRayBaseRenderer.render_impl` is run with `diner_depth_guided_sampling` (renderer.py:570-600).  The
reference's two random draws (original_depth_guided_sample.py:271 `randn_like`, :356 `rand_like`) are
redirected to explicit tables so that the result is reproducible: the patched functions look up the caller's
`ray_mask` / `missing_iray, missing_isample` locals and return the table entries of exactly those slots.
"""
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
from make_golden_render import build_reference_renderer, hot_weights  # noqa: E402


class TableRNG:
    def __init__(self, fill_rand, gauss):
        self.fill_rand, self.gauss = fill_rand, gauss

    def __enter__(self):
        self._rand_like, self._randn_like = torch.rand_like, torch.randn_like

        def rand_like(t, *a, **k):
            loc = sys._getframe(1).f_locals
            if "missing_iray" in loc:
                return self.fill_rand[loc["missing_iray"], loc["missing_isample"]].to(t.dtype)
            return self._rand_like(t, *a, **k)

        def randn_like(t, *a, **k):
            loc = sys._getframe(1).f_locals
            if "ray_mask" in loc and "gauss_samples" in loc:
                return self.gauss[loc["ray_mask"]].to(t.dtype)
            return self._randn_like(t, *a, **k)

        torch.rand_like, torch.randn_like = rand_like, randn_like
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.randn_like = self._rand_like, self._randn_like


def gen_diner():
    for name in cases.DINER_CASES:
        cfg, que, ref, fill_rand, gauss = cases.make_diner_inputs(name)
        net = build_reference_renderer(cfg, seed=sum(map(ord, name)))
        with torch.no_grad(), TableRNG(fill_rand, gauss):
            out = net.render_impl(dict(que), dict(ref), False)
        W = hot_weights(net)
        blob = {"fill_rand": fill_rand.numpy(), "gauss": gauss.numpy()}
        for k, v in que.items():
            blob["que." + k] = v.numpy()
        for k, v in ref.items():
            blob["ref." + k] = v.numpy()
        for k, v in W.items():
            blob["w." + k] = v.numpy()
        for k, v in out.items():
            blob["out." + k] = v.float().numpy() if v.dtype != torch.bool else v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    gen_diner()
