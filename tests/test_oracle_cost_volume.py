"""CPU: the cost-volume oracle reproduces the reference outputs stored in tests/golden/ (HOT 1)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import assert_close, load_golden  # noqa: E402

from oracle import cost_volume as ocv  # noqa: E402


def _run_oracle(name, g):
    c = cases.CV_CASES[name]
    args = {"dataset_name": c["dataset"], "contain_dnet": c["per_pixel"], "mono_uncertainty": False}
    dv = g["depth_volume"] if c["per_pixel"] else None
    if c.get("mv"):
        return ocv.calculate_cost_volume_erp_multiview(args, g["images"], g["depths"], g["trans"], g["rots"],
                                                       depth_volume=dv, cost_type=c["cost_type"],
                                                       curr_idx=c["curr_idx"])
    return ocv.calculate_cost_volume_erp(args, g["images"], g["depths"], g["trans"], g["rots"],
                                         depth_volume=dv, cost_type=c["cost_type"])


@pytest.mark.parametrize("name", list(cases.CV_CASES))
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    # fixtures store their inputs; they must still equal what the seeded generator produces
    inp = cases.make_cv_inputs(name)
    assert torch.equal(inp["images"], g["images"]) and torch.equal(inp["rots"], g["rots"])
    out = _run_oracle(name, g)
    # white-noise features: error = (feature gradient ~ O(1..4)) x (sub-1e-4 pixel coordinate error)
    assert_close(out, g["out"], atol=2e-4, max_bad_frac=2e-4, what=name)


def test_unknown_cost_type_raises():
    inp = cases.make_cv_inputs("cv_m3d_scalar")
    with pytest.raises(ValueError):
        ocv.calculate_cost_volume_erp(inp["args"], inp["images"], inp["depths"], inp["trans"], inp["rots"],
                                      cost_type="ssd")


def test_identity_pose_gives_zero_cost_at_pixel_centres():
    """Same pose, same image: every voxel samples its own pixel up to the reference's half-pixel
    inconsistency ((j+0.5)/W forward vs align_corners=True backward), so on a constant image the
    abs_diff cost is exactly 0 — a size-independent property also used at full size on the GPU."""
    B, H, W, C, D = 1, 8, 16, 4, 3
    images = torch.ones(B, 2, H, W, C) * 0.75
    rots = torch.eye(3).expand(B, 2, 3, 3).contiguous()
    trans = torch.zeros(B, 2, 3)
    out = ocv.calculate_cost_volume_erp({"dataset_name": "m3d", "contain_dnet": False}, images,
                                        torch.linspace(1, 5, D), trans, rots)
    assert float(out.abs().max()) < 1e-6


def test_group_mean_and_hypotheses():
    x = torch.arange(2 * 3 * 2 * 2 * 8, dtype=torch.float32).reshape(2, 3, 2, 2, 8)
    g = ocv.group_mean(x, 4)
    assert g.shape == (2, 4, 3, 2, 2)
    assert torch.allclose(g[0, 1, 2, 1, 0], x[0, 2, 1, 0, 2:4].mean())
    ks = ocv.magnet_k_list(5, 3)
    assert len(ks) == 5 and abs(ks[2]) < 1e-9 and ks[0] < 0 < ks[-1]
    mu = torch.rand(1, 1, 4, 4) * 8 + 1
    vol = ocv.mono_guided_hypotheses(mu, ks, 0.5, 0.1, 10.0, 59)
    assert vol.shape == (1, 64, 4, 4) and bool((vol[:, 1:] >= vol[:, :-1]).all())


@pytest.mark.parametrize("name", cases.CV_BWD_CASES)
def test_oracle_gradient_matches_reference_autograd(name):
    """d(sum(out*w))/d(images) of the oracle (torch autograd through its gather restatement) == the reference's."""
    inp = cases.make_cv_inputs(name)
    images = inp["images"].clone().requires_grad_(True)
    fn = ocv.calculate_cost_volume_erp_multiview if inp["mv"] else ocv.calculate_cost_volume_erp
    kw = dict(depth_volume=inp["depth_volume"], cost_type=inp["cost_type"])
    if inp["mv"]:
        kw["curr_idx"] = inp["curr_idx"]
    out = fn(inp["args"], images, inp["depths"], inp["trans"], inp["rots"], **kw)
    w = cases.cv_bwd_weight(name, out.shape)
    (out * w).sum().backward()
    gold = load_golden(name + "_bwd")
    assert abs(float(w.double().sum()) - float(gold["weight_sum"])) < 1e-6
    assert_close(images.grad, gold["grad_images"], rtol=1e-4, atol=1e-4, max_bad_frac=2e-3, what=name)


@pytest.mark.parametrize("name", list(cases.HYP_CASES))
def test_hypothesis_builder_oracle_matches_reference_lines(name):
    """a5: the restated builder equals the outputs of the reference's own statements (pipeline3_model.py:717-733, 774-821,
    executed by tests/golden/make_golden_hypotheses.py) bit for bit; the k list agrees with scipy's (fp64)."""
    case = cases.make_hyp_inputs(name)
    g = load_golden(name)
    k_ref = g["k_list"].double()
    k_list = ocv.magnet_k_list(case["n_samples"], case["sampling_range"]) if case["n_samples"] > 0 else []
    assert len(k_list) == k_ref.numel() and all(abs(a - float(b)) < 1e-13 for a, b in zip(k_list, k_ref))
    vol, cen = ocv.depth_hypotheses(case["args"], g["ref_gmms"], [float(k) for k in k_ref], case["cost_volume_channels"],
                                    case["contain_dnet"])
    if "depth_volume" in g:
        assert torch.equal(vol, g["depth_volume"])
    else:
        assert vol is None
    if "d_centers" in g and not case["args"]["wo_hdh"]:
        assert torch.equal(cen.reshape(g["d_centers"].shape) if cen.dim() == 1 else cen.expand_as(g["d_centers"]) if case["args"]["revise_range"] else cen, g["d_centers"])
