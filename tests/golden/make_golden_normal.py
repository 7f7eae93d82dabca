"""Golden fixtures of depth2normal (network/orig_diner_depth2normal.py:7-110), produced by the REAL reference on CPU.
python tests/golden/make_golden_normal.py"""
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
import cases  # noqa: E402

if __name__ == "__main__":
    from network.orig_diner_depth2normal import depth2normal
    from network.spt_utils import Utils
    for name in cases.NORMAL_CASES:
        cfg, dmap = cases.make_normal_inputs(name)
        n = depth2normal({"mvs_depth": dmap}, Utils(cfg))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), mvs_depth=dmap.numpy(), normal=n.numpy())
        print(name, tuple(n.shape), "nan", int(torch.isnan(n).sum()))
