"""ctypes binding of `libpanogrf_b200.so` (the C ABI declared in include/panogrf_b200.h).

There is NO fallback: if the shared library is missing or a symbol cannot be resolved, importing
the op raises.  Torch is used only to own device memory and streams; tensors cross the boundary
as raw device pointers.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpanogrf_b200.so")

DATASET_IDS = {"m3d": 0, "replica_test": 1, "residential": 2, "CoffeeArea": 3}
COST_IDS = {"abs_diff": 0, "dot": 1, "none": 2}
CV_LAYOUT_IDS = {"bdchw": 0, "bdhwc": 1, "bcdhw": 2}

PGRF_OK, PGRF_EINVAL, PGRF_ECUDA, PGRF_ERANGE = 0, -1, -2, -3

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float

# name -> (restype, argtypes); must list every symbol of include/panogrf_b200.h
SIGNATURES = {
    "pgrf_last_error": (_c.c_char_p, []),
    "pgrf_version": (_I, []),
    "pgrf_launch_count": (_c.c_int64, []),
    "pgrf_cost_volume_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _F, _I, _I, _I, _I, _P, _P, _P]),
    "pgrf_cost_volume_host": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _F, _I, _I, _I, _I, _P]),
}

_lib = None


class PanoGRFError(RuntimeError):
    pass


def load():
    """Load the library (once). Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PanoGRFError(
            f"{LIB_PATH} is missing: build it with `python -m panogrf_b200.build` "
            "(panogrf_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().pgrf_last_error().decode()


def check(rc, what):
    if rc == PGRF_OK:
        return
    msg = last_error()
    if rc == PGRF_ERANGE:
        raise AssertionError(msg)
    if rc == PGRF_EINVAL and msg == "Unknown cost type":
        raise ValueError(msg)
    raise PanoGRFError(f"{what}: {msg} (code {rc})")


def launch_count():
    return int(load().pgrf_launch_count())


def ptr(t):
    """Raw pointer of a torch tensor (or None)."""
    return None if t is None else _c.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return _c.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise PanoGRFError("panogrf_b200 ops need CUDA tensors (no CPU fallback exists)")
