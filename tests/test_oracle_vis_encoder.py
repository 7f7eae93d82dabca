"""CPU: the DefaultVisEncoder oracle reproduces the reference class' outputs; the product container has the reference's parameter names."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402

from oracle import vis_encoder as ovis  # noqa: E402


@pytest.mark.parametrize("name", list(cases.VISENC_CASES))
def test_vis_encoder_oracle_matches_reference_class(name):
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    out = ovis.vis_encoder(W, g["ray_feats"], g["img_feats"], wrap=cases.VISENC_CASES[name][0])
    assert out.shape == g["out"].shape
    assert float((out - g["out"]).abs().max()) <= 1e-5 * max(1.0, float(g["out"].abs().max()))


@pytest.mark.parametrize("name", ["visenc_wrap", "visenc_zero"])
def test_container_has_the_reference_parameter_names(name):
    from panogrf_b200.vis_encoder import DefaultVisEncoder
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    net = DefaultVisEncoder({"use_wrap_padding": cases.VISENC_CASES[name][0]})
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in W.items()}
    net.load_state_dict(W)


@pytest.mark.parametrize("name", list(cases.INITCONV_CASES))
def test_init_net_convs_oracle_and_container(name):
    from panogrf_b200.vis_encoder import CostVolumeInitConvs
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    wrap = cases.INITCONV_CASES[name][0]
    out = ovis.init_net_convs(W, g["ref_feats"], g["depth"], wrap)
    assert float((out - g["ray_feats"]).abs().max()) <= 1e-5 * max(1.0, float(g["ray_feats"].abs().max()))
    net = CostVolumeInitConvs({"use_wrap_padding": wrap})
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in W.items()}
    net.load_state_dict(W)
