"""CPU: driver-side helpers against the reference's own functions, imported from /root/reference when it exists
(build container); on the GPU box the reference is absent and the checks fall back to closed-form properties."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panogrf_b200 import driver  # noqa: E402

HAVE_REF = os.path.isdir("/root/reference/network")


def _ref_modules():
    from oracle import _refimport
    _refimport.install()
    from utils.imgs_info import build_render_imgs_info
    from network.metrics import WSPSNR
    return build_render_imgs_info, WSPSNR


def test_build_render_imgs_info():
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    pose = np.concatenate([q, rng.normal(size=(3, 1))], 1)
    got = driver.build_render_imgs_info(pose, (6, 10), (0.5, 15.0))
    assert got["coords"].shape == (1, 60, 2) and got["coords"][0, 11].tolist() == [1.0, 1.0] and got["shape"] == (6, 10)
    # c2w inverts w2c
    full = np.concatenate([got["w2c"][0], [[0, 0, 0, 1]]], 0) @ np.concatenate([got["c2w"][0], [[0, 0, 0, 1]]], 0)
    assert np.abs(full - np.eye(4)).max() < 1e-5
    if HAVE_REF:
        ref_fn, _ = _ref_modules()
        want = ref_fn(pose, (6, 10), (0.5, 15.0))
        assert set(want) == set(got)
        for k in want:
            if isinstance(want[k], np.ndarray):
                assert want[k].dtype == got[k].dtype and np.array_equal(want[k], got[k]), k
            else:
                assert want[k] == got[k]


def test_ws_psnr():
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(2, 16, 32, 3, generator=g), torch.rand(2, 16, 32, 3, generator=g)
    m = driver.WSPSNR()
    v = m.ws_psnr(a, b)
    assert v.shape == (2,) and bool(torch.isfinite(v).all())
    # uniform error e -> psnr = -20 log10 e whatever the weights
    assert abs(float(m.ws_psnr(a, a + 0.1)[0]) - 20.0) < 1e-4
    # a polar error counts less than the same error at the equator
    e_pole, e_eq = a.clone(), a.clone()
    e_pole[:, 0] += 0.5
    e_eq[:, 8] += 0.5
    assert float(m.ws_psnr(e_pole, a)[0]) > float(m.ws_psnr(e_eq, a)[0])
    if HAVE_REF:
        _, Ref = _ref_modules()
        assert torch.equal(Ref().ws_psnr(a, b), v)


def test_pose_loop_with_stub_renderer():
    calls = []

    def renderer(data):
        q = data["que_imgs_info"]
        assert data["eval"] is True and "ref_imgs_info" in data and q["coords"].shape[1] == 4 * 8
        calls.append(q["c2w"].clone())
        n = q["coords"].shape[1]
        return {"pixel_colors_nr_fine": torch.full((1, n, 3), 0.5), "render_depth_fine": torch.full((1, n), 2.0)}

    poses = [np.concatenate([np.eye(3), [[0.0], [0.0], [float(i)]]], 1) for i in range(3)]
    imgs, depths = driver.render_poses(renderer, {"imgs": torch.zeros(2, 3, 4, 8)}, poses, [(4, 8)] * 3, [(0.5, 15.0)] * 3, device="cpu")
    assert len(imgs) == 3 and imgs[0].shape == (4, 8, 3) and imgs[0].dtype == np.uint8 and int(imgs[0][0, 0, 0]) == 127
    want = int(np.uint8((1 / 2.0 - 1 / 0.5) / (1 / 15.0 - 1 / 0.5) * 255))
    assert depths[0].shape == (4, 8) and int(depths[0][0, 0]) == want
    assert float(calls[2][0, 2, 3]) == -2.0                     # c2w of pose t=(0,0,2)
