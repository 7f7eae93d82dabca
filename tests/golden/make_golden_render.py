"""Golden fixtures for the per-ray render path, produced by the REAL reference renderer on CPU.

The reference `NeuralRayBaseRenderer` is constructed with random-init weights (seeded); its own
`render_impl` (network/renderer.py:567-633) is run on pre-encoded feature maps (the CNN encoders
are outside the hot path).  Stored: inputs, the hot-path weights by state_dict name,
and every output of render_impl incl. hit_prob and intermediate prj tensors of the coarse pass.
"""
import copy
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


def build_reference_renderer(cfg, seed):
    from network.renderer import NeuralRayBaseRenderer
    torch.manual_seed(seed)
    net = NeuralRayBaseRenderer(copy.deepcopy(cfg)).eval()
    # biases are zero-initialised by weights_init; randomise them so bias handling is exercised
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.startswith(cases.RENDER_WEIGHT_PREFIXES) and n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
            if "layer_norm.weight" in n:
                p.copy_(1 + torch.randn(p.shape, generator=gen) * 0.1)
    return net


def hot_weights(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items() if k.startswith(cases.RENDER_WEIGHT_PREFIXES)}


def gen_render(only=None):
    for name in cases.RENDER_CASES:
        if only and name not in only:
            continue
        cfg, que, ref = cases.make_render_inputs(name)
        net = build_reference_renderer(cfg, seed=sum(map(ord, name)))
        with torch.no_grad():
            out = net.render_impl(dict(que), dict(ref), False)
        W = hot_weights(net)
        blob = {}
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


def gen_perspec(only=None):
    for name in cases.PERSPEC_CASES:
        if only and name not in only:
            continue
        cfg, que, ref = cases.make_perspec_inputs(name)
        net = build_reference_renderer(cfg, seed=sum(map(ord, name)))
        with torch.no_grad():
            out = net.render_impl(dict(que), dict(ref), False, is_perspec=True)
        blob = {}
        for k, v in que.items():
            blob["que." + k] = v.numpy()
        for k, v in ref.items():
            blob["ref." + k] = v.numpy()
        for k, v in hot_weights(net).items():
            blob["w." + k] = v.numpy()
        for k, v in out.items():
            blob["out." + k] = v.float().numpy() if v.dtype != torch.bool else v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob)
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    gen_render(sys.argv[1:] or None)
    gen_perspec(sys.argv[1:] or None)      # optional: names of the cases to (re)generate
