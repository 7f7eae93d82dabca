"""Equirectangular -> cubemap on the GPU: drop-in for `Equirec2Cube` (UniFuse-Unidirectional-Fusion/UniFuse/datasets/util.py:7-100)
and for `e2c_process` (network/omni_mvsnet/pipeline3_model.py:262-283), which in the reference copy every panorama to the host,
resample it with scipy and copy it back.  The sampling-coordinate tables are built once per size with the
reference's numpy expressions; the resampling itself is one CUDA launch for all (batch, view) images.
"""
import numpy as np
import torch

from . import _lib


class Equirec2Cube:
    """Same constructor and `run` semantics as the reference class (no depth branch: the hot path never passes `equ_dep`)."""

    def __init__(self, equ_h, equ_w, face_w):
        self.equ_h, self.equ_w, self.face_w = int(equ_h), int(equ_w), int(face_w)
        self._xyzcube()
        self._xyz2coor()
        self._dev = {}

    def _xyzcube(self):
        """util.py:26-60: xyz of the unit cube in [F R B L U D] order."""
        f = self.face_w
        self.xyz = np.zeros((f, f * 6, 3), np.float32)
        rng = np.linspace(-0.5, 0.5, num=f, dtype=np.float32)
        self.grid = np.stack(np.meshgrid(rng, -rng), -1)
        faces = ((0, [0, 1], self.grid, 2, 0.5), (1, [2, 1], self.grid[:, ::-1], 0, 0.5), (2, [0, 1], self.grid[:, ::-1], 2, -0.5),
                 (3, [2, 1], self.grid, 0, -0.5), (4, [0, 2], self.grid[::-1, :], 1, 0.5), (5, [0, 2], self.grid, 1, -0.5))
        for k, axes, g, fixed, val in faces:
            self.xyz[:, k * f:(k + 1) * f, axes] = g
            self.xyz[:, k * f:(k + 1) * f, fixed] = val

    def _xyz2coor(self):
        """util.py:62-72."""
        x, y, z = np.split(self.xyz, 3, axis=-1)
        lon = np.arctan2(x, z)
        c = np.sqrt(x ** 2 + z ** 2)
        lat = np.arctan2(y, c)
        self.coor_x = (lon / (2 * np.pi) + 0.5) * self.equ_w - 0.5
        self.coor_y = (-lat / np.pi + 0.5) * self.equ_h - 0.5

    def _tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(np.ascontiguousarray(self.coor_x[..., 0])).to(device),
                              torch.from_numpy(np.ascontiguousarray(self.coor_y[..., 0])).to(device))
        return self._dev[key]

    def run(self, equ_img):
        """equ_img: CUDA tensor (..., H, W, C) channels-last, any leading dims -> (..., face_w, 6*face_w, C).  The reference resizes
        inputs of another size with cv2 first; here the size must match (that is how e2c_process calls it)."""
        _lib.require_cuda(equ_img)
        *lead, h, w, c = equ_img.shape
        if h != self.equ_h or w != self.equ_w:
            raise _lib.PanoGRFError(f"Equirec2Cube built for {self.equ_h}x{self.equ_w}, got {h}x{w} (resize outside the kernel)")
        x = equ_img.detach().float().contiguous()
        n = int(np.prod(lead)) if lead else 1
        cx, cy = self._tables(x.device)
        out = torch.empty(*lead, self.face_w, 6 * self.face_w, c, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            rc = _lib.load().pgrf_e2c_fwd(_lib.ptr(x), n, h, w, c, _lib.ptr(cx), _lib.ptr(cy), self.face_w, _lib.ptr(out),
                                          _lib.stream_ptr())
        _lib.check(rc, "pgrf_e2c_fwd")
        return out


def e2c_process(panos_small, e2c_instance):
    """pipeline3_model.py:262-283: (B, S, H, W, 3) panoramas -> (B, S, face_w, 6 face_w, 3) cube maps, without leaving the GPU."""
    return e2c_instance.run(panos_small)
