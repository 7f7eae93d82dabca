"""GPU: the image encoder ResUNetLight on the tensor cores (patch gather + pointwise GEMM, D = 1 convolutions, instance norms) against
outputs of the reference class with the same seeded weights."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402
from test_oracle_res_unet import golden_resunet  # noqa: E402

pytestmark = pytest.mark.gpu

#: bf16 activations through 27 convolutions and 30 instance norms: deviation relative to the output range (worst element / rms)
RESUNET_TOL, RESUNET_RMS = 4e-2, 8e-3


@pytest.mark.parametrize("name", list(cases.RESUNET_CASES))
def test_res_unet_matches_reference_golden(name):
    g = load_golden(name)
    net = golden_resunet(name, g).cuda()
    y = net(g["x"].cuda()).cpu()
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    scale = float(g["y"].abs().max())
    err = float((y - g["y"]).abs().max())
    rms = float((y - g["y"]).pow(2).mean().sqrt())
    print(f"{name}: max err {err:.3e}, rms {rms:.3e} of range {scale:.3e} ({err / scale:.2e}, {rms / scale:.2e})")
    assert err <= RESUNET_TOL * scale and rms <= RESUNET_RMS * scale


def test_res_unet_rejects_bad_sizes():
    from panogrf_b200.image_encoder import ResUNetLight
    net = ResUNetLight({}, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=True).cuda()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 40, 64, device="cuda"))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 4, 32, 64, device="cuda"))
