"""Import recipe for the *real* reference (thucz/PanoGRF) in the build container.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists only in the build container, never on the
GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py` may import this module.  It is used by
`tests/golden/make_golden.py` (fixture generation) and by the `not gpu` tests that pin the oracle
restatement against the reference when the reference tree happens to be present.

The reference imports a number of packages that are absent here and are *not* on the arithmetic
path: they are replaced by inert stub modules.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("PANOGRF_REFERENCE", "/root/reference")

_STUB_ROOTS = {
    "matplotlib", "kornia", "easydict", "inplace_abn", "skimage", "imageio", "lmdb", "h5py",
    "plyfile", "transforms3d", "sklearn", "habitat", "habitat_sim", "lpips", "tensorboardX", "ipdb",
}


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return None


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _Stub(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "network"))


def install():
    """Make `import network.renderer`, `import models.spherical_cost_volume` work on CPU."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.meta_path.insert(0, _StubFinder())
    for name in ("data_readers.habitat_data_neuray_ft", "data_readers.habitat_data_neuray_ft_lmdb"):
        mod = types.ModuleType(name)
        mod.HabitatImageGeneratorFT = object
        mod.HabitatImageGeneratorFT_LMDB = object
        sys.modules[name] = mod
    # the vendored UniFuse `datasets` package must shadow HuggingFace `datasets`
    sys.path.insert(0, os.path.join(REF_ROOT, "UniFuse-Unidirectional-Fusion", "UniFuse"))
    sys.path.insert(0, REF_ROOT)
    _installed = True
    import torch
    if not torch.cuda.is_available():
        # network/ibrnet.py:312 hard-codes .to("cuda:0") for the positional table
        import numpy as np
        from network import ibrnet

        def _posenc(self, d_hid, n_samples):
            pos = np.arange(n_samples)[:, None].astype(np.float64)
            j = np.arange(d_hid)[None, :]
            table = pos / np.power(10000, 2 * (j // 2) / d_hid)
            table[:, 0::2] = np.sin(table[:, 0::2])
            table[:, 1::2] = np.cos(table[:, 1::2])
            return torch.from_numpy(table).float().unsqueeze(0)

        ibrnet.IBRNetWithNeuRay.posenc = _posenc
