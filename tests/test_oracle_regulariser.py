"""CPU: the 3-D cost regulariser oracle reproduces the reference classes' outputs (tests/golden/unet3d_*.npz), and the product's
parameter container draws the reference's seeded initial weights / packs them as the kernel expects."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402

from oracle import regulariser as oreg  # noqa: E402
from panogrf_b200 import regulariser as reg  # noqa: E402


def golden_weights(name, g):
    """Weights of a golden case: stored (`w.*`) or re-created from the case's seed through the product's container, whose
    construction order is the reference's; the stored per-tensor checksums make any divergence loud."""
    if cases.UNET3D_STORE_WEIGHTS[name]:
        return {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    size, _ = cases.UNET3D_CASES[name]
    torch.manual_seed(cases.unet3d_seed(name))
    W = {k: v.detach().clone() for k, v in reg.CostRegulariser3D(size).state_dict().items()}
    sums = {k[4:]: v for k, v in g.items() if k.startswith("sum.")}
    assert set(sums) == set(W)
    for k, v in W.items():
        got = torch.tensor([float(v.double().sum()), float(v.double().abs().sum())], dtype=torch.float64)
        assert torch.allclose(got, sums[k].double(), rtol=1e-9, atol=1e-9), f"seeded weights differ from the reference's: {k}"
    return W


@pytest.mark.parametrize("name", list(cases.UNET3D_CASES))
def test_unet3d_oracle_matches_reference_classes(name):
    g = load_golden(name)
    W = golden_weights(name, g)
    y = oreg.unet3d(W, g["x"])
    assert y.shape == g["y"].shape
    assert float((y - g["y"]).abs().max()) <= 1e-5 * max(1.0, float(g["y"].abs().max()))


def test_container_has_the_reference_parameter_names():
    g = load_golden("unet3d_s1")
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    net = reg.CostRegulariser3D(cases.UNET3D_CASES["unet3d_s1"][0])
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in W.items()}
    net.load_state_dict(W)          # strict


@pytest.mark.parametrize("co,ca,cb", [(8, 4, 0), (16, 16, 8), (256, 128, 128), (64, 96, 32)])
def test_pack_conv_layout(co, ca, cb):
    """Every weight lands at [n_tile][tap][chunk][k8][n][8] of the padded K axis (what csrc/conv3d.cu's descriptors walk)."""
    torch.manual_seed(co + ca + cb)
    w = torch.randn(co, ca + cb, 3, 3, 3)
    b = torch.randn(co)
    ca_pad, cb_pad = (ca + 15) // 16 * 16, (cb + 15) // 16 * 16
    wp, bp = reg.pack_conv(w, b, ca, cb, ca_pad, cb_pad)
    kc = reg.chunk_size(ca_pad, cb_pad)
    assert (ca_pad % kc, cb_pad % kc) == (0, 0) and kc >= 16
    n_tiles, taps, n_cc, k8, nt, eight = wp.shape
    assert (taps, eight, k8 * 8, n_cc * kc) == (27, 8, kc, ca_pad + cb_pad)
    assert torch.equal(bp[:co], b) and float(bp[co:].abs().sum()) == 0
    dense = wp.float().permute(0, 4, 1, 2, 3, 5).reshape(n_tiles * nt, 27, ca_pad + cb_pad)     # (co_pad, tap, kpos)
    ref = w.reshape(co, ca + cb, 27).permute(0, 2, 1).to(torch.bfloat16).float()
    assert torch.equal(dense[:co, :, :ca], ref[:, :, :ca])
    assert torch.equal(dense[:co, :, ca_pad:ca_pad + cb], ref[:, :, ca:])
    assert float(dense[co:].abs().sum()) == 0 and float(dense[:, :, ca:ca_pad].abs().sum()) == 0


@pytest.mark.parametrize("name", list(cases.DEC2D_CASES))
def test_decoders_oracle_matches_reference_classes(name):
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    out_type = cases.DEC2D_CASES[name][3]
    raw, depth_d1 = oreg.decoders1(W, g["cost_reg"], out_type)
    assert torch.allclose(raw, g["raw_d1"], rtol=1e-5, atol=1e-6) and torch.allclose(depth_d1, g["depth_d1"], rtol=1e-4, atol=1e-6)
    feats = oreg.decoders2(W, torch.cat((g["cost_reg"], g["mono"]), 1))
    assert torch.allclose(feats, g["feats"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(oreg.rectify(feats[:, :1], out_type).permute(0, 2, 3, 1), g["depth"], rtol=1e-4, atol=1e-6)


def test_decoders_container_has_the_reference_parameter_names():
    g = load_golden("dec2d_s1")
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    size, D, _, _ = cases.DEC2D_CASES["dec2d_s1"]
    net = reg.CostDecoders2D(size, D)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in W.items()}
    net.load_state_dict(W)
