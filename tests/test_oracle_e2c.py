"""CPU: the equirect -> cubemap oracle vs scipy's map_coordinates and vs the goldens produced by the reference class."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402

from oracle import e2c as oe2c  # noqa: E402


def test_wrap_bilinear_equals_scipy():
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(0)
    img = rng.normal(size=(11, 17)).astype(np.float32)
    cy = rng.uniform(-14, 25, size=(40, 60)).astype(np.float32)
    cx = rng.uniform(-20, 40, size=(40, 60)).astype(np.float32)
    cy[0, :5] = [-0.5, 0.0, 10.0, 10.5, -10.0]
    cx[0, :5] = [16.0, -0.5, 16.5, 0.0, 32.0]
    want = map_coordinates(img, [cy, cx], order=1, mode="wrap")
    got = oe2c.sample_wrap_bilinear(img, cy, cx)
    assert np.abs(got - want).max() <= 1.2e-7 * np.abs(want).max()


@pytest.mark.parametrize("name", list(cases.E2C_CASES))
def test_e2c_oracle_matches_reference_class(name):
    g = load_golden(name)
    got = oe2c.e2c(g["equ"].numpy(), cases.E2C_CASES[name][2])
    want = g["cube"].numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1.2e-7


def test_product_tables_equal_oracle_and_reference():
    """the host-side coordinate tables of panogrf_b200.e2c (no GPU needed to build them) are bit-identical to the oracle's and,
    where /root/reference exists, to the reference class's."""
    import os
    import sys
    from panogrf_b200 import e2c
    for h, w, f in cases.E2C_CASES.values():
        inst = e2c.Equirec2Cube(h, w, f)
        cx, cy = oe2c.cube_tables(h, w, f)
        assert np.array_equal(inst.coor_x[..., 0], cx) and np.array_equal(inst.coor_y[..., 0], cy)
        ref_dir = "/root/reference/UniFuse-Unidirectional-Fusion/UniFuse"
        if os.path.isdir(ref_dir):
            sys.path.insert(0, ref_dir)
            from datasets.util import Equirec2Cube as Ref
            r = Ref(h, w, f)
            assert np.array_equal(r.coor_x, inst.coor_x) and np.array_equal(r.coor_y, inst.coor_y)
