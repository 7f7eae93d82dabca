"""CPU: the ResUNetLight oracle reproduces the reference class' outputs; the product container re-creates the reference's seeded weights
(same construction order, same parameter names)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402

from oracle import res_unet as oru  # noqa: E402


def golden_resunet(name, g):
    """container with the golden's weights: seeded construction + the generator's affine-parameter loop, checked against checksums"""
    from panogrf_b200.image_encoder import ResUNetLight
    wrap = cases.RESUNET_CASES[name][0]
    torch.manual_seed(sum(map(ord, name)))
    net = ResUNetLight({}, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=wrap)
    with torch.no_grad():
        for k, p in net.named_parameters():
            if p.dim() == 1 and ("bn" in k or "downsample.1" in k):
                p.copy_(torch.randn_like(p) * 0.3 + (1.0 if k.endswith("weight") else 0.0))
    sums = {k[4:]: v for k, v in g.items() if k.startswith("sum.")}
    W = net.state_dict()
    assert set(sums) == set(W), sorted(set(sums) ^ set(W))[:6]
    for k, v in W.items():
        got = torch.tensor([float(v.double().sum()), float(v.double().abs().sum())], dtype=torch.float64)
        assert torch.allclose(got, sums[k].double(), rtol=1e-9, atol=1e-9), f"seeded weights differ from the reference's: {k}"
    return net


@pytest.mark.parametrize("name", list(cases.RESUNET_CASES))
def test_res_unet_oracle_matches_reference_class(name):
    g = load_golden(name)
    net = golden_resunet(name, g)
    W = {k: v.detach().clone() for k, v in net.state_dict().items()}
    y = oru.res_unet_light(W, g["x"], wrap=cases.RESUNET_CASES[name][0])
    assert y.shape == g["y"].shape
    assert float((y - g["y"]).abs().max()) <= 1e-5 * max(1.0, float(g["y"].abs().max()))
