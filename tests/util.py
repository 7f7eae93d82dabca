"""Comparison helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for fp32 colours / depths / cost values
RTOL = 1e-4


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def assert_close(actual, expected, rtol=RTOL, atol=1e-4, max_bad_frac=0.0, what=""):
    """|a-e| <= rtol*|e| + atol for all but `max_bad_frac` of the elements.

    `max_bad_frac` > 0 is only used where the reference itself is discontinuous in its inputs
    (the longitude seam of the ERP mapping: `fmod(theta + pi/2 + 2pi, 2pi)` jumps between the first
    and last column under a 1-ulp change of atan2), so that two correct fp32 implementations may
    pick different sides for a handful of voxels.
    """
    a = torch.as_tensor(actual).detach().cpu().double()
    e = torch.as_tensor(expected).detach().cpu().double()
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    bad = (a - e).abs() > (rtol * e.abs() + atol)
    n_bad = int(bad.sum())
    frac = n_bad / max(1, bad.numel())
    if frac > max_bad_frac:
        idx = torch.nonzero(bad)[:5].tolist()
        raise AssertionError(
            f"{what}: {n_bad}/{bad.numel()} elements ({frac:.3e}) outside rtol={rtol} atol={atol}; "
            f"max abs err {float((a - e).abs().max()):.3e}; first bad idx {idx}")
