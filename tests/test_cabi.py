"""CPU: the C-ABI library builds, loads, and exports every symbol include/panogrf_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from panogrf_b200 import build as b
    b.build()
    from panogrf_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "panogrf_b200.h")).read()
    return sorted(set(re.findall(r"PGRF_API[^;(]*?\b(pgrf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from panogrf_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 5
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_string(lib):
    assert lib.pgrf_version() >= 100
    assert isinstance(lib.pgrf_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    """Shape validation happens before any CUDA call, so it is checkable on a CPU-only box."""
    import ctypes
    from panogrf_b200 import _lib
    views = (ctypes.c_int * 1)(0)
    dummy = ctypes.c_void_p(256)
    rc = lib.pgrf_cost_volume_fwd(dummy, 1, 2, 8, 16, 5, dummy, None, 4, dummy, dummy, 1, views, 1, 0.0,
                                  0, 0, 0, 0, dummy, dummy, None)
    assert rc == _lib.PGRF_EINVAL and b"C=5" in lib.pgrf_last_error()
    rc = lib.pgrf_cost_volume_fwd(dummy, 1, 2, 8, 16, 8, dummy, None, 4, dummy, dummy, 1, views, 1, 0.0,
                                  0, 7, 0, 0, dummy, dummy, None)
    assert rc == _lib.PGRF_EINVAL and lib.pgrf_last_error() == b"Unknown cost type"
    # sample_3sigma: sizes, row stride of the prior table, coarse depths only in selection mode
    rc = lib.pgrf_sample_3sigma_fwd(dummy, 3, 0.5, dummy, dummy, 1, 0.5, 15.0, None, 0, 0, 1, 8, dummy, None)
    assert rc == _lib.PGRF_EINVAL and b"n=1" in lib.pgrf_last_error()
    rc = lib.pgrf_sample_3sigma_fwd(dummy, 2, 0.5, dummy, dummy, 16, 0.5, 15.0, None, 0, 0, 1, 8, dummy, None)
    assert rc == _lib.PGRF_EINVAL
    rc = lib.pgrf_sample_3sigma_fwd(dummy, 2, 0.5, dummy, dummy, 16, 0.5, 15.0, dummy, 0, 16, 0, 8, dummy, None)
    assert rc == _lib.PGRF_EINVAL and b"selection mode" in lib.pgrf_last_error()
    rc = lib.pgrf_sample_3sigma_fwd(None, 3, 0.5, dummy, dummy, 16, 0.5, 15.0, None, 0, 0, 1, 8, dummy, None)
    assert rc == _lib.PGRF_EINVAL and b"null pointer" in lib.pgrf_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from panogrf_b200 import calculate_cost_volume_erp
    from panogrf_b200._lib import PanoGRFError
    from panogrf_b200.render_ops import sample_3sigma
    with pytest.raises(PanoGRFError):
        calculate_cost_volume_erp({"dataset_name": "m3d", "contain_dnet": False}, torch.zeros(1, 2, 8, 16, 8),
                                  torch.ones(3), torch.zeros(1, 2, 3), torch.eye(3).expand(1, 2, 3, 3))
    with pytest.raises(PanoGRFError):
        sample_3sigma(torch.ones(4), 2 * torch.ones(4), 16, True, 0.5, 15.0)


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirrors of pgrf_render_args / pgrf_render_view_args / pgrf_diner_args have the size and field offsets the C
    compiler gives the header's structs (a silent mismatch would shift every pointer after the first wrong field)."""
    import ctypes
    import re
    import shutil
    import subprocess
    from panogrf_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    header = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "panogrf_b200.h")
    structs = {"pgrf_render_args": _lib.RenderArgs, "pgrf_render_view_args": _lib.RenderViewArgs, "pgrf_diner_args": _lib.DinerArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{header}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            cfield = "pass" if fname == "pass_" else fname
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {cfield}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for m in re.finditer(r"(\w+) (\w+) (\d+)", out):
        cname, fname, val = m.group(1), m.group(2), int(m.group(3))
        cls = structs[cname]
        if fname == "size":
            assert ctypes.sizeof(cls) == val, (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == val, (cname, fname, getattr(cls, fname).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_conv3d_plan_and_validation_without_gpu(lib):
    """The tile / split-K plan of the regulariser's convolution is host code: workspace sizes and argument checks on a CPU box."""
    import ctypes
    from panogrf_b200 import _lib
    need = ctypes.c_longlong(-1)
    # a coarse U-Net level (8 voxel tiles x 4 channel tiles = 32 CTAs) splits K over all nine tap rows: 9 partial volumes
    assert lib.pgrf_conv3d_workspace(512, 0, 512, 1, 8, 8, 16, ctypes.byref(need)) == _lib.PGRF_OK
    assert need.value == 9 * 1024 * 512
    # a full-resolution level fills the GPU by itself: no workspace
    assert lib.pgrf_conv3d_workspace(64, 0, 64, 1, 64, 64, 128, ctypes.byref(need)) == _lib.PGRF_OK and need.value == 0
    # D == 1 is a 2-D convolution: only the three kd == 1 tap rows exist, so at most 3 splits
    assert lib.pgrf_conv3d_workspace(32, 0, 32, 1, 1, 32, 64, ctypes.byref(need)) == _lib.PGRF_OK
    assert need.value == 3 * 2048 * 32
    # channel counts must be multiples of 16; a concatenation needs a common chunk size
    assert lib.pgrf_conv3d_workspace(24, 0, 32, 1, 8, 8, 16, ctypes.byref(need)) == _lib.PGRF_EINVAL
    assert b"multiples of 16" in lib.pgrf_last_error()
    dummy = ctypes.c_void_p(256)
    rc = lib.pgrf_conv3d_fwd(dummy, 32, None, 16, dummy, dummy, dummy, None, 0, 32, 1, 8, 8, 16, 1, None, 0, None)
    assert rc == _lib.PGRF_EINVAL and b"second input" in lib.pgrf_last_error()
    rc = lib.pgrf_conv3d_fwd(dummy, 512, None, 0, dummy, dummy, dummy, None, 0, 512, 1, 8, 8, 16, 1, None, 0, None)
    assert rc == _lib.PGRF_EINVAL and b"workspace" in lib.pgrf_last_error()
    rc = lib.pgrf_conv3d_fwd(dummy, 32, None, 0, dummy, dummy, dummy, dummy, 1, 32, 1, 8, 8, 16, 1, None, 0, None)
    assert rc == _lib.PGRF_EINVAL and b"exactly one of y / yf" in lib.pgrf_last_error()
