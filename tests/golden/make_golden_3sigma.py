"""Golden fixture of sample_3sigma / sample_pdf (network/sample_utils.py:6-60, det=True) produced by the REFERENCE functions on CPU.
python tests/golden/make_golden_3sigma.py"""
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

if __name__ == "__main__":
    from network.sample_utils import sample_3sigma
    g = torch.Generator().manual_seed(31)
    blob = {}
    for tag, n, near, far in (("a", 64, 0.5, 15.0), ("b", 16, 0.1, 10.0), ("c", 33, 0.5, 15.0)):
        R = 200
        mu = near - 0.5 + (far - near + 1.0) * torch.rand(R, generator=g)          # some intervals stick out of [near, far]
        half = 0.02 + 2.0 * torch.rand(R, generator=g)
        low, high = mu - half, mu + half
        with torch.no_grad():
            z = sample_3sigma(low, high, n, True, near, far)
        blob.update({f"{tag}.low": low.numpy(), f"{tag}.high": high.numpy(), f"{tag}.n": np.asarray(n), f"{tag}.near": np.asarray(near),
                     f"{tag}.far": np.asarray(far), f"{tag}.z": z.numpy()})
    np.savez_compressed(os.path.join(HERE, "sample_3sigma.npz"), **blob)
    print({k: v.shape for k, v in blob.items() if k.endswith(".z")})
