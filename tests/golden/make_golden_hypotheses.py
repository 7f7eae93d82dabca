"""Golden vectors of the depth-hypothesis builder produced by EXECUTING THE REFERENCE'S OWN SOURCE LINES.

The builder is not a function of its own in the reference: it is two statement blocks inside `forward_thru_unet`
(network/omni_mvsnet/pipeline3_model.py:717-733 and :774-821) of a model class that cannot be constructed here (checkpoints).
This script reads those line ranges from /root/reference at generation time, dedents them and `exec`s them in a namespace that
provides the variables the blocks read (`self`, `args`, `ref_gmms`, `min_depth`, `max_depth`); `.cuda()` is made a no-op for the
run (no GPU in the build container).  Nothing of the reference is copied into the repo: only its OUTPUTS are committed
(tests/golden/hyp_*.npz).   python tests/golden/make_golden_hypotheses.py
"""
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

SRC = "/root/reference/network/omni_mvsnet/pipeline3_model.py"


def k_list_reference(n_samples, sampling_range):
    """pipeline3_model.py:537-545 executed from the reference source (scipy erf / norm.ppf)."""
    lines = open(SRC).read().split("\n")
    body = textwrap.dedent("\n".join(lines[536:545]))            # def depth_sampling(self): ... return list(k_list)
    ns = {"np": np}
    exec(body, ns)
    return ns["depth_sampling"](types.SimpleNamespace(n_samples=n_samples, sampling_range=sampling_range))


def run_reference_blocks(case):
    lines = open(SRC).read().split("\n")
    # the blocks keep their own indentation (comment lines in between are indented less): run them as the body of `if True:`
    block1 = "if True:\n" + "\n".join(lines[716:733])            # :717-733  mono-guided hypotheses
    block2 = "if True:\n" + "\n".join(lines[773:821])            # :774-821  d_centers, concatenation, per-pixel sort
    args = dict(case["args"])
    k_list = k_list_reference(case["n_samples"], case["sampling_range"]) if case["n_samples"] > 0 else []
    self = types.SimpleNamespace(args=args, n_samples=case["n_samples"], k_list=k_list, contain_dnet=case["contain_dnet"],
                                 cost_volume_channels=case["cost_volume_channels"], weighting="CW5")
    ns = {"torch": torch, "self": self, "args": args, "ref_gmms": case["ref_gmms"], "min_depth": args["min_depth"],
          "max_depth": args["max_depth"], "depth_volume": None, "d_centers": None}
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda t, *a, **k: t
    try:
        exec(block1, ns)
        exec(block2, ns)
    finally:
        torch.Tensor.cuda = orig
    return ns.get("depth_volume"), ns.get("d_centers"), k_list


if __name__ == "__main__":
    for name in cases.HYP_CASES:
        case = cases.make_hyp_inputs(name)
        vol, centers, k_list = run_reference_blocks(case)
        out = {"ref_gmms": case["ref_gmms"].numpy(), "k_list": np.asarray(k_list, dtype=np.float64)}
        if vol is not None:
            out["depth_volume"] = vol.numpy()
        if centers is not None:
            out["d_centers"] = centers.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, None if vol is None else tuple(vol.shape), None if centers is None else tuple(centers.shape))
