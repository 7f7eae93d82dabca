"""CPU restatement of the equirectangular -> cubemap resampling on the MVS path (TEST INFRASTRUCTURE, never imported by the product).

Reference: `Equirec2Cube` in UniFuse-Unidirectional-Fusion/UniFuse/datasets/util.py:7-100 (based on py360convert), called per batch
item and per view from `e2c_process` (network/omni_mvsnet/pipeline3_model.py:262-283) on the CPU through scipy's
`map_coordinates(order=1, mode='wrap')` — two GPU -> CPU -> GPU hops per call.

Third-party arithmetic: scipy.ndimage.map_coordinates (not under /root/reference).  Its published behaviour for order = 1 and the
legacy 'wrap' mode, restated in `sample_wrap_bilinear`: a coordinate outside [0, n-1] is wrapped with period n-1 (first and last
sample coincide), then linearly interpolated between floor(c) and floor(c)+1 in double precision.  Pinned: tests compare this file
with scipy itself and with tests/golden/e2c_*.npz (outputs of the reference class).
"""
import numpy as np


def cube_tables(equ_h, equ_w, face_w):
    """util.py:26-72: xyz of the unit cube in [F R B L U D] order and the equirectangular sampling coordinates (float32)."""
    xyz = np.zeros((face_w, face_w * 6, 3), np.float32)
    rng = np.linspace(-0.5, 0.5, num=face_w, dtype=np.float32)
    grid = np.stack(np.meshgrid(rng, -rng), -1)
    f = face_w
    xyz[:, 0 * f:1 * f, [0, 1]] = grid
    xyz[:, 0 * f:1 * f, 2] = 0.5
    xyz[:, 1 * f:2 * f, [2, 1]] = grid[:, ::-1]
    xyz[:, 1 * f:2 * f, 0] = 0.5
    xyz[:, 2 * f:3 * f, [0, 1]] = grid[:, ::-1]
    xyz[:, 2 * f:3 * f, 2] = -0.5
    xyz[:, 3 * f:4 * f, [2, 1]] = grid
    xyz[:, 3 * f:4 * f, 0] = -0.5
    xyz[:, 4 * f:5 * f, [0, 2]] = grid[::-1, :]
    xyz[:, 4 * f:5 * f, 1] = 0.5
    xyz[:, 5 * f:6 * f, [0, 2]] = grid
    xyz[:, 5 * f:6 * f, 1] = -0.5
    x, y, z = np.split(xyz, 3, axis=-1)
    lon = np.arctan2(x, z)
    c = np.sqrt(x ** 2 + z ** 2)
    lat = np.arctan2(y, c)
    coor_x = (lon / (2 * np.pi) + 0.5) * equ_w - 0.5
    coor_y = (-lat / np.pi + 0.5) * equ_h - 0.5
    return coor_x[..., 0], coor_y[..., 0]


def _wrap(c, n):
    """scipy NI_EXTEND_WRAP (legacy 'wrap'): period n-1."""
    c = np.asarray(c, dtype=np.float64).copy()
    sz = n - 1
    lo = c < 0
    c[lo] += sz * (np.trunc(-c[lo] / sz) + 1)
    hi = c > n - 1
    c[hi] -= sz * np.trunc(c[hi] / sz)
    return c


def sample_wrap_bilinear(img, coor_y, coor_x):
    """map_coordinates(img, [coor_y, coor_x], order=1, mode='wrap') for a 2-D float image."""
    h, w = img.shape
    cy, cx = _wrap(coor_y, h), _wrap(coor_x, w)
    y0, x0 = np.floor(cy).astype(np.int64), np.floor(cx).astype(np.int64)
    ty, tx = cy - y0, cx - x0
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)        # c = n-1 exactly: weight 0 on the clamped neighbour
    a = img.astype(np.float64)
    out = (a[y0, x0] * (1 - ty) * (1 - tx) + a[y0, x1] * (1 - ty) * tx + a[y1, x0] * ty * (1 - tx) + a[y1, x1] * ty * tx)
    return out.astype(img.dtype)


def e2c(equ_img, face_w):
    """Equirec2Cube.run (util.py:86-100) without the resize branch: equ_img (H,W,C) -> cube (face_w, 6 face_w, C)."""
    h, w = equ_img.shape[:2]
    coor_x, coor_y = cube_tables(h, w, face_w)
    chans = []
    for i in range(equ_img.shape[2]):
        e = equ_img[..., i]
        pad_u = np.roll(e[[0]], w // 2, 1)
        pad_d = np.roll(e[[-1]], w // 2, 1)
        e = np.concatenate([e, pad_d, pad_u], 0)                          # util.py:76-78: both pads are appended at the bottom
        chans.append(sample_wrap_bilinear(e, coor_y, coor_x))
    return np.stack(chans, axis=-1)
