"""Generate the committed golden fixtures by running the REAL reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py [cv|render|all]
Outputs: tests/golden/*.npz  (inputs + reference outputs, fp32).  The GPU box never runs this.
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


def gen_cv():
    from models.spherical_cost_volume import calculate_cost_volume_erp as ref_cv
    from models.spherical_cost_volume_mv import calculate_cost_volume_erp_multiview as ref_mv
    for name in cases.CV_CASES:
        inp = cases.make_cv_inputs(name)
        kw = dict(depth_volume=inp["depth_volume"], cost_type=inp["cost_type"])
        with torch.no_grad():
            if inp["mv"]:
                out = ref_mv(inp["args"], inp["images"], inp["depths"], inp["trans"], inp["rots"],
                             curr_idx=inp["curr_idx"], **kw)
            else:
                out = ref_cv(inp["args"], inp["images"], inp["depths"], inp["trans"], inp["rots"], **kw)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            images=inp["images"].numpy(), depths=inp["depths"].numpy(), trans=inp["trans"].numpy(),
            rots=inp["rots"].numpy(),
            depth_volume=(inp["depth_volume"].numpy() if inp["depth_volume"] is not None else np.zeros(0, np.float32)),
            out=out.contiguous().numpy())
        print(name, tuple(out.shape))


def gen_cv_bwd():
    """d(sum(out * weight))/d(images) through the reference's own autograd graph (grid_sample + abs / mul)."""
    from models.spherical_cost_volume import calculate_cost_volume_erp as ref_cv
    from models.spherical_cost_volume_mv import calculate_cost_volume_erp_multiview as ref_mv
    for name in cases.CV_BWD_CASES:
        inp = cases.make_cv_inputs(name)
        images = inp["images"].clone().requires_grad_(True)
        kw = dict(depth_volume=inp["depth_volume"], cost_type=inp["cost_type"])
        if inp["mv"]:
            out = ref_mv(inp["args"], images, inp["depths"], inp["trans"], inp["rots"], curr_idx=inp["curr_idx"], **kw)
        else:
            out = ref_cv(inp["args"], images, inp["depths"], inp["trans"], inp["rots"], **kw)
        weight = cases.cv_bwd_weight(name, out.shape)
        (out * weight).sum().backward()
        np.savez_compressed(os.path.join(HERE, name + "_bwd.npz"), grad_images=images.grad.numpy(),
                            weight_sum=np.float64(weight.double().sum().item()))
        print(name, "grad", tuple(images.grad.shape), float(images.grad.abs().mean()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("cv", "all"):
        gen_cv()
    if what in ("cv_bwd", "all"):
        gen_cv_bwd()
    if what in ("render", "all"):
        try:
            from make_golden_render import gen_render
        except ImportError:
            gen_render = None
        if gen_render:
            gen_render()
