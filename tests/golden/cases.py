"""Seeded synthetic inputs shared by the golden generator, the oracle tests and the GPU parity tests.

Everything is derived from (name, seed) so a fixture file only has to store the *reference output*
plus the inputs (kept too, so the fixtures stay valid if torch's RNG stream ever changes).
"""
import numpy as np
import torch


def small_rotations(gen, B, S, deg):
    ang = torch.randn(B, S, 3, generator=gen) * np.deg2rad(deg)
    K = torch.zeros(B, S, 3, 3)
    K[..., 0, 1], K[..., 0, 2] = -ang[..., 2], ang[..., 1]
    K[..., 1, 0], K[..., 1, 2] = ang[..., 2], -ang[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -ang[..., 1], ang[..., 0]
    return torch.matrix_exp(K)


def smooth(x, passes=2):
    """3x3 box blur over (H,W) of a channels-last (...,H,W,C) tensor with ERP wrap in W."""
    for _ in range(passes):
        x = (torch.roll(x, 1, -2) + x + torch.roll(x, -1, -2)) / 3
        up = torch.cat([x[..., :1, :, :], x[..., :-1, :, :]], -3)
        dn = torch.cat([x[..., 1:, :, :], x[..., -1:, :, :]], -3)
        x = (up + x + dn) / 3
    return x


# name -> dict(dataset, B,S,H,W,C,D, per_pixel, cost_type, mv, curr_idx)
CV_CASES = {
    "cv_m3d_scalar":      dict(dataset="m3d", B=2, S=2, H=16, W=32, C=8, D=6, per_pixel=False, cost_type="abs_diff"),
    "cv_m3d_volume":      dict(dataset="m3d", B=1, S=2, H=16, W=64, C=32, D=5, per_pixel=True, cost_type="abs_diff"),
    "cv_m3d_dot":         dict(dataset="m3d", B=1, S=2, H=12, W=40, C=16, D=4, per_pixel=False, cost_type="dot"),
    "cv_m3d_none":        dict(dataset="m3d", B=1, S=2, H=12, W=40, C=4, D=4, per_pixel=True, cost_type="none"),
    "cv_replica_scalar":  dict(dataset="replica_test", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=False, cost_type="abs_diff"),
    "cv_residential_vol": dict(dataset="residential", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=True, cost_type="abs_diff"),
    "cv_coffee_scalar":   dict(dataset="CoffeeArea", B=1, S=2, H=16, W=32, C=8, D=4, per_pixel=False, cost_type="abs_diff"),
    "cv_mv4_scalar":      dict(dataset="m3d", B=1, S=4, H=16, W=32, C=8, D=5, per_pixel=False, cost_type="abs_diff", mv=True, curr_idx=0),
    "cv_mv5_volume":      dict(dataset="m3d", B=2, S=5, H=8, W=32, C=32, D=3, per_pixel=True, cost_type="abs_diff", mv=True, curr_idx=1),
}


def make_cv_inputs(name, seed=0, smooth_feats=False):
    c = CV_CASES[name]
    gen = torch.Generator().manual_seed(seed + sum(map(ord, name)))
    B, S, H, W, C, D = c["B"], c["S"], c["H"], c["W"], c["C"], c["D"]
    images = torch.randn(B, S, H, W, C, generator=gen)
    if smooth_feats:
        images = smooth(images)
    rots = small_rotations(gen, B, S, 5.0)
    trans = torch.randn(B, S, 3, generator=gen) * 0.3
    depths = torch.linspace(0.5, 10.0, D)
    depth_volume = None
    if c["per_pixel"]:
        depth_volume = torch.sort(torch.rand(B, D, H, W, generator=gen) * 9.5 + 0.5, dim=1)[0]
    args = {"dataset_name": c["dataset"], "contain_dnet": c["per_pixel"], "mono_uncertainty": False}
    return dict(args=args, images=images, depths=depths, trans=trans, rots=rots, depth_volume=depth_volume,
                cost_type=c["cost_type"], mv=c.get("mv", False), curr_idx=c.get("curr_idx", 0))
