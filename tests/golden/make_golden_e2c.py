"""Golden fixtures of the equirect -> cubemap resampling, produced by the reference's own `Equirec2Cube` (UniFuse util.py) on CPU.
python tests/golden/make_golden_e2c.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/UniFuse-Unidirectional-Fusion/UniFuse")
import cases  # noqa: E402

if __name__ == "__main__":
    from datasets.util import Equirec2Cube
    for name, (h, w, f) in cases.E2C_CASES.items():
        img = cases.make_e2c_input(name)
        cube = Equirec2Cube(h, w, f).run(img)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), equ=img, cube=cube)
        print(name, cube.shape, cube.dtype)
