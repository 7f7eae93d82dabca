"""Drop-in for the reference's `models/spherical_cost_volume.py` / `models/spherical_cost_volume_mv.py`.

Same function names, argument meaning and error behaviour
(`calculate_cost_volume_erp` models/spherical_cost_volume.py:231-341,
 `calculate_cost_volume_erp_multiview` models/spherical_cost_volume_mv.py:219-347), but the D
(x views) python iterations, each a chain of ~25 elementwise launches + grid_sample + a host sync,
run as ONE fused sm_100a kernel (csrc/cost_volume.cu) through the C ABI.
"""
import ctypes

import torch

from . import _lib

#: How the reference's `assert uv in [-1, 1]` (one host sync per depth, :191) is honoured:
#:   "deferred" (default)  the kernel ORs a device flag; the flag travels to pinned host memory asynchronously and is looked at
#:                         — without ever blocking the stream — at the next cost-volume call or at `check_pending()`; a violation
#:                         raises the reference's AssertionError there (asynchronous error reporting, like CUDA's own);
#:   "sync"                read the flag back before returning (one host sync per call);
#:   "off"                 never look at it.
UV_CHECK = "deferred"
_CHECK_UV = True          # legacy switch (False == "off")
_RING = 256


class _UvChecks:
    """Per-device ring of device flags + pinned host mirrors + events: deferred, non-blocking range assertion."""

    def __init__(self):
        self.dev = {}

    def slot(self, device):
        st = self.dev.get(device)
        if st is None:
            st = self.dev[device] = {"flags": torch.zeros(_RING, device=device, dtype=torch.int32),
                                     "host": torch.zeros(_RING, dtype=torch.int32).pin_memory(), "next": 0, "pending": []}
        if len(st["pending"]) >= _RING - 1:
            self.drain(device, block=True)
        i = st["next"]
        st["next"] = (i + 1) % _RING
        return st, i

    def queue(self, st, i):
        st["host"][i:i + 1].copy_(st["flags"][i:i + 1], non_blocking=True)
        st["flags"][i:i + 1].zero_()                       # stream-ordered after the copy: the slot is clean when reused
        ev = torch.cuda.Event()
        ev.record()
        st["pending"].append((ev, i))

    def drain(self, device=None, block=False):
        for d, st in self.dev.items():
            if device is not None and d != device:
                continue
            keep, bad = [], False
            for ev, i in st["pending"]:
                if block:
                    ev.synchronize()
                if ev.query():
                    bad = bad or int(st["host"][i]) != 0
                else:
                    keep.append((ev, i))
            st["pending"] = keep
            if bad:
                raise AssertionError("Wrong UV mapping, UV must be in [-1, 1]!")


_uv = _UvChecks()


def check_pending(block=True):
    """Look at the deferred uv-range flags of earlier cost-volume calls (block=True waits for them first); raises the
    reference's AssertionError if any call produced a uv outside [-1, 1]."""
    _uv.drain(None, block)


class _SweepFn(torch.autograd.Function):
    """torch.autograd.Function around the two C-ABI kernels (forward sweep, feature-map gradient)."""

    @staticmethod
    def forward(ctx, images, depth_t, rots, trans, meta):
        lib = _lib.load()
        B, S, H, W, C = images.shape
        dev = images.device
        use_volume, D = meta["use_volume"], meta["D"]
        groups, out_layout = meta["groups"], meta["out_layout"]
        OC = groups if groups > 0 else C
        if out_layout == "bdchw":
            store = torch.empty((B, D, OC, H, W), device=dev, dtype=torch.float32)
        elif out_layout == "bdhwc":
            store = torch.empty((B, D, H, W, OC), device=dev, dtype=torch.float32)
        elif out_layout == "bdhwc_bf16":
            store = torch.empty((B, D, H, W, OC), device=dev, dtype=torch.bfloat16)
        else:
            store = torch.empty((B, OC, D, H, W), device=dev, dtype=torch.float32)
        mode = UV_CHECK if _CHECK_UV else "off"
        if mode == "deferred":
            _uv.drain(dev, block=False)                    # raise for an EARLIER call whose flag has arrived
            st, slot = _uv.slot(dev)
            err = st["flags"][slot:slot + 1]
        else:
            err = torch.zeros(1, device=dev, dtype=torch.int32)
        views = (ctypes.c_int * len(meta["src_views"]))(*meta["src_views"])
        with torch.cuda.device(dev):
            rc = lib.pgrf_cost_volume_fwd(
                _lib.ptr(images), B, S, H, W, C, None if use_volume else _lib.ptr(depth_t), _lib.ptr(depth_t) if use_volume else None,
                D, _lib.ptr(rots), _lib.ptr(trans), meta["ref_idx"], views, len(meta["src_views"]), float(meta["divisor"]),
                meta["dataset"], meta["cost"], _lib.CV_LAYOUT_IDS[out_layout], groups,
                _lib.ptr(store), _lib.ptr(err), _lib.stream_ptr())
        _lib.check(rc, "pgrf_cost_volume_fwd")
        if mode == "deferred":
            _uv.queue(st, slot)
        elif mode == "sync" and int(err.item()) != 0:
            raise AssertionError("Wrong UV mapping, UV must be in [-1, 1]!")
        ctx.meta = meta
        ctx.save_for_backward(images, depth_t, rots, trans)
        return store

    @staticmethod
    def backward(ctx, grad_store):
        images, depth_t, rots, trans = ctx.saved_tensors
        meta = ctx.meta
        lib = _lib.load()
        B, S, H, W, C = images.shape
        groups, out_layout, D = meta["groups"], meta["out_layout"], meta["D"]
        if out_layout == "bdhwc_bf16":
            raise RuntimeError("the bf16 channels-last cost volume feeds the inference-only tensor-core regulariser; use an fp32 layout "
                               "for training")
        g = grad_store.float()
        if groups > 0:                                             # (B,G,D,H,W): mean over C/G channels
            cpg = C // groups
            g = (g / cpg).repeat_interleave(cpg, dim=1)
        if out_layout == "bdchw":
            g = g.permute(0, 1, 3, 4, 2)
        elif out_layout == "bcdhw":
            g = g.permute(0, 2, 3, 4, 1)
        g = g.contiguous()                                         # (B,D,H,W,C)
        grad_images = torch.zeros_like(images)
        views = (ctypes.c_int * len(meta["src_views"]))(*meta["src_views"])
        use_volume = meta["use_volume"]
        with torch.cuda.device(images.device):
            rc = lib.pgrf_cost_volume_bwd(
                _lib.ptr(g), _lib.ptr(images), B, S, H, W, C, None if use_volume else _lib.ptr(depth_t),
                _lib.ptr(depth_t) if use_volume else None, D, _lib.ptr(rots), _lib.ptr(trans), meta["ref_idx"], views,
                len(meta["src_views"]), float(meta["divisor"]), meta["dataset"], meta["cost"], _lib.ptr(grad_images),
                _lib.stream_ptr())
        _lib.check(rc, "pgrf_cost_volume_bwd")
        return grad_images, None, None, None, None


def _sweep(args, images, depths, trans, rots, depth_volume, cost_type, ref_idx, src_views, divisor,
           out_layout="bdchw", groups=0):
    name = args["dataset_name"]
    if name not in _lib.DATASET_IDS:
        raise Exception(f"unknown dataset_name {name!r}")          # reference: bare `raise Exception` (:189,:294)
    if cost_type not in _lib.COST_IDS:
        raise ValueError("Unknown cost type")                      # reference :217
    if out_layout not in _lib.CV_LAYOUT_IDS:
        raise ValueError(f"unknown out_layout {out_layout!r}")
    _lib.require_cuda(images, trans, rots)
    B, S, H, W, C = images.shape
    images = images.contiguous().float()
    rots = rots.reshape(B, S, 3, 3).contiguous().float().detach()
    trans = trans.reshape(B, S, 3).contiguous().float().detach()
    dev = images.device
    use_volume = bool(args["contain_dnet"])
    if use_volume:
        # hypotheses carry no gradient (built under no_grad / detached, pipeline3_model.py:647,671)
        depth_t = depth_volume.detach().to(device=dev, dtype=torch.float32).contiguous()
        D = depth_t.shape[1]
        assert depth_t.shape == (B, D, H, W), "depth_volume must be (B,D,H,W)"
    else:
        depth_t = torch.as_tensor(depths, dtype=torch.float32, device=dev).detach().reshape(-1).contiguous()
        D = depth_t.numel()
    meta = dict(use_volume=use_volume, D=D, groups=groups, out_layout=out_layout, src_views=list(src_views), ref_idx=ref_idx,
                divisor=divisor, dataset=_lib.DATASET_IDS[name], cost=_lib.COST_IDS[cost_type])
    store = _SweepFn.apply(images, depth_t, rots, trans, meta)
    if groups > 0:
        return store                                               # (B,G,D,H,W)
    if out_layout == "bdchw":
        return store.permute(0, 1, 3, 4, 2)                        # same strides as the reference's :340
    if out_layout == "bcdhw":
        return store.permute(0, 2, 3, 4, 1)
    return store


def calculate_cost_volume_erp(args, images, depths, trans, rots, depth_volume=None, cost_type="abs_diff",
                              ref_gmms=None, nghbr_gmms=None, thres=None, direction="up",
                              out_layout="bdchw", groups=0):
    """(B,2,H,W,C) channels-last features [0]=source,[1]=reference -> (B,D,H,W,C) cost volume.

    `ref_gmms`, `nghbr_gmms`, `thres`, `direction` are accepted and ignored exactly like the
    reference.  Extensions (keyword-only in spirit): `out_layout` picks the physical layout
    ("bdchw" = the reference's strides, "bdhwc" channels-last, "bcdhw" the regulariser's, "bdhwc_bf16" = channels-last
    bf16, the operand layout of `regulariser.CostRegulariser3D`, which then consumes `cv.permute(0, 4, 1, 2, 3)` in place), and
    `groups>0` fuses the group-wise mean of pipeline3_model.py:849-853, returning (B,G,D,H,W).
    """
    if groups > 0:
        out_layout = "bcdhw"
    return _sweep(args, images, depths, trans, rots, depth_volume, cost_type, 1, [0], 0.0, out_layout, groups)


def calculate_cost_volume_erp_multiview(args, images, depths, trans, rots, depth_volume=None,
                                        cost_type="abs_diff", ref_gmms=None, nghbr_gmms=None, thres=None,
                                        direction="up", curr_idx=0, out_layout="bdhwc", groups=0):
    """(B,S,H,W,C) -> mean over views {0..S-2}\\{curr_idx} of the 2-view volume, each /(S-2).

    The LAST view is skipped on purpose, like the reference (:312 "exclude the last").  The
    reference returns a contiguous (B,D,H,W,C) tensor here (sum of permuted views), hence the
    channels-last default.
    """
    S = images.shape[1]
    assert S > 2
    views = [v for v in range(S - 1) if v != curr_idx]
    if groups > 0:
        out_layout = "bcdhw"
    return _sweep(args, images, depths, trans, rots, depth_volume, cost_type, curr_idx, views, float(S - 2),
                  out_layout, groups)
