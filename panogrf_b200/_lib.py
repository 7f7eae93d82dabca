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
CV_LAYOUT_IDS = {"bdchw": 0, "bdhwc": 1, "bcdhw": 2, "bdhwc_bf16": 3}

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
    "pgrf_debug_set": (_I, [_c.c_char_p, _I]),
    "pgrf_umma_selftest": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "pgrf_cost_volume_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _F, _I, _I, _I, _I, _P, _P, _P]),
    "pgrf_cost_volume_host": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _F, _I, _I, _I, _I, _P]),
    "pgrf_cost_volume_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _F, _I, _I, _P, _P]),
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
    # experiment knobs: PGRF_DEBUG="key=value,key=value" -> pgrf_debug_set (tuning only, never changes results)
    for kv in filter(None, os.environ.get("PGRF_DEBUG", "").split(",")):
        k, v = kv.split("=")
        if lib.pgrf_debug_set(k.encode(), int(v)) != PGRF_OK:
            raise PanoGRFError(f"PGRF_DEBUG: {lib.pgrf_last_error().decode()}")
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


class RenderArgs(ctypes.Structure):
    """Mirror of `pgrf_render_args` (include/panogrf_b200.h) — field order and types must match."""
    _fields_ = [
        ("dataset", _I), ("H", _I), ("W", _I), ("rfn", _I), ("rn", _I), ("dn", _I), ("use_vis", _I),
        ("bias_val", _F),
        ("coords", _P), ("depth", _P), ("depth_ray_stride", _I),
        ("que_c2w", _P), ("que_near", _F), ("que_far", _F),
        ("ref_w2c", _P), ("ref_depth_range", _P),
        ("imgs_cl", _P), ("img_h", _I), ("img_w", _I),
        ("img_feats_cl", _P), ("if_h", _I), ("if_w", _I),
        ("ray_feats_cl", _P), ("rf_h", _I), ("rf_w", _I),
        ("weights", _P), ("f1", _P), ("f2", _P),
        ("pixel_colors", _P), ("render_depth", _P), ("hit_prob", _P), ("density", _P), ("colors", _P),
        ("fine_depth", _P), ("fine_dn", _I), ("fine_u", _P), ("fine_use_all", _I), ("use_disp", _I),
        ("fine_inds", _P), ("prob_dbg", _P), ("prj_dbg", _P), ("feat_dbg", _P),
        ("stage_mask", _I), ("mlp_bf16", _I), ("sched", _P), ("weights16", _P),
        ("prj_in", _P), ("feat_in", _P), ("prob_in", _P), ("que_dir_in", _P), ("interval_in", _P), ("dec_dbg", _P),
        ("wo_geometry", _I), ("wo_appearance", _I), ("ray_dirs", _P),
    ]


class RenderViewArgs(ctypes.Structure):
    """Mirror of `pgrf_render_view_args`."""
    _fields_ = [
        ("pass_", RenderArgs), ("hierarchical", _I), ("weights_fine", _P), ("weights16_fine", _P), ("bias_val_fine", _F),
        ("rays_per_launch", _I), ("fine_depth_ws", _P),
        ("pixel_colors_fine", _P), ("render_depth_fine", _P), ("hit_prob_fine", _P), ("density_fine", _P),
        ("colors_fine", _P), ("que_depth_fine", _P),
    ]


class DinerArgs(ctypes.Structure):
    """Mirror of `pgrf_diner_args` (depth-prior sample placement)."""
    _fields_ = [
        ("dataset", _I), ("H", _I), ("W", _I), ("rfn", _I), ("rn", ctypes.c_longlong),
        ("n_candidates", _I), ("n_samples", _I), ("n_gaussian", _I), ("n_uniform", _I),
        ("include_norm", _I), ("sigma_is_var", _I), ("diner_sigma", _F), ("cand_step", _F),
        ("min_depth", _F), ("max_depth", _F), ("depth_diff_max", _F),
        ("coords", _P), ("cand_depth", _P), ("cand_ray_stride", ctypes.c_longlong),
        ("que_c2w", _P), ("ref_w2c", _P), ("mvs_depth", _P), ("mvs_uncert", _P), ("mvs_normal", _P),
        ("map_h", _I), ("map_w", _I), ("img_h", _I), ("img_w", _I),
        ("fill_rand", _P), ("gauss", _P), ("uniform_depth", _P), ("out_depth", _P), ("likelihood", _P),
        ("prj_mu", _P), ("prj_uncert", _P), ("prj_depth", _P), ("prj_normal", _P), ("que_dir", _P),
    ]


_PI = ctypes.POINTER(_I)
_PLL = ctypes.POINTER(ctypes.c_longlong)
SIGNATURES.update({
    "pgrf_depth_guided_sample_fwd": (_I, [ctypes.POINTER(DinerArgs), _P]),
    "pgrf_project_gather_diner_fwd": (_I, [_P, ctypes.c_longlong, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I,
                                           _P, _P, _P, _P, _P, _P]),
    "pgrf_render_pass_fwd": (_I, [ctypes.POINTER(RenderArgs), _P]),
    "pgrf_agg_mlp_fwd": (_I, [ctypes.POINTER(RenderArgs), _P]),
    "pgrf_compute_prob_que_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, ctypes.c_longlong, _I, _P, _P, _P, _P]),
    "pgrf_compute_prob_fwd": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _I, ctypes.c_longlong, _I, _P, _P, _P, _P]),
    "pgrf_interpolate_feature_map_fwd": (_I, [_P, _I, _I, _I, _I, _P, ctypes.c_longlong, _I, _I, _P, _P]),
    "pgrf_composite_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "pgrf_interpolate_feature_map_bwd": (_I, [_P, _I, _I, _I, _I, _P, ctypes.c_longlong, _I, _I, _P, _P]),
    "pgrf_e2c_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "pgrf_depth2normal_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "pgrf_conv3d_to_bf16_cl": (_I, [_P, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _I, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_conv3d_workspace": (_I, [_I, _I, _I, _I, _I, _I, _I, _PLL]),
    "pgrf_conv3d_fwd": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, ctypes.c_longlong, _P]),
    "pgrf_conv3d_ex_fwd": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, ctypes.c_longlong, _P, _I, _P]),
    "pgrf_feats_to_bf16_cl": (_I, [_P, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P]),
    "pgrf_instnorm_relu_fwd": (_I, [_P, _I, _I, _I, _P, _P, _F, _P, _P, _P]),
    "pgrf_instnorm_act_fwd": (_I, [_P, _I, _I, _I, _P, _P, _F, _P, _P, _I, _P, _P]),
    "pgrf_patch7x7_s2_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_subsample2_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "pgrf_upsample2d2_ac_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "pgrf_conv3d_pointwise_fwd": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pgrf_conv3d_tapsum_fwd": (_I, [_P, _F, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_conv3d_scalar_fwd": (_I, [_P, _P, _F, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_conv3d_cout1_fwd": (_I, [_P, _I, _P, _I, _P, _P, _F, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_avgpool3d2_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_upsample3d2_fwd": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "pgrf_upsample2d2_fwd": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "pgrf_channel_dot_upsample_fwd": (_I, [_P, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _I, _I, _I, _I,
                                           _P, _F, _I, _I, _P, _P]),
    "pgrf_depth2points_fwd": (_I, [_P, _P, _I, _P, _I, _I, _I, ctypes.c_longlong, _I, _P, _P, _P]),
    "pgrf_render_workspace": (_I, [_I, ctypes.c_longlong, _PLL, _PLL]),
    "pgrf_render_view_fwd": (_I, [ctypes.POINTER(RenderViewArgs), _P]),
    "pgrf_render_view_host": (_I, [ctypes.POINTER(RenderViewArgs)]),
    "pgrf_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "pgrf_project_gather_fwd": (_I, [_P, ctypes.c_longlong, _P, _I, _I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _I,
                                      _P, _P, _P, _P, _P, _P, _P]),
    "pgrf_composite_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "pgrf_fine_sample_fwd": (_I, [_P, _I, _P, _P, _F, _F, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pgrf_sample_3sigma_fwd": (_I, [_P, _I, _F, _P, _P, _I, _F, _F, _P, _I, _I, _I, _I, _P, _P]),
    "pgrf_depth_hypotheses_fwd": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _F, _F, _P, _P]),
    "pgrf_depth_hypotheses2_fwd": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _F, _F, _P, _I, _I, _F, _F, _F, _I, _P, _P]),
    "pgrf_weight_blob_floats": (_I, []),
    "pgrf_weight_num_layers": (_I, []),
    "pgrf_weight_layer_info": (_I, [_I, ctypes.c_char_p, _I, _PI, _PI, _PI, _PI, _PI, _PI, _PI]),
    "pgrf_weight_aux_offsets": (_I, [_PI, _PI, _PI]),
    "pgrf_w16_blob_bytes": (_I, []),
    "pgrf_w16_num_layers": (_I, []),
    "pgrf_w16_layer_info": (_I, [_I, ctypes.c_char_p, _I, _PI, _PI, _PI, _PI, _PI, _PI, _PI]),
    "pgrf_w16_layer_info2": (_I, [_I, _PI, _PI, _PI]),
})

BIAS_F32, BIAS_CHUNK, BIAS_INLINE, BIAS_NONE = 0, 1, 2, 3


def weight_layers():
    """[(name, K, N, Npad, has_bias, k_begin, w_offset, b_offset)] straight from the C side."""
    lib = load()
    out = []
    for i in range(lib.pgrf_weight_num_layers()):
        name = ctypes.create_string_buffer(128)
        vals = [_I() for _ in range(7)]
        check(lib.pgrf_weight_layer_info(i, name, 128, *[ctypes.byref(v) for v in vals]), "pgrf_weight_layer_info")
        out.append((name.value.decode(),) + tuple(v.value for v in vals))
    return out


def weight_aux_offsets():
    lib = load()
    a, b, c = _I(), _I(), _I()
    check(lib.pgrf_weight_aux_offsets(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "pgrf_weight_aux_offsets")
    return a.value, b.value, c.value


def w16_layers():
    """[(name, Kpad, Npad, w_offset_bytes, b_offset_bytes, kmap, nmap, is_small, bias_kind, in_ln2, out_log2e)] of the bf16
    tensor-core blob (kmap: -1 = zero column, -2 / -3 = bias_hi / bias_lo column)."""
    lib = load()
    out = []
    for i in range(lib.pgrf_w16_num_layers()):
        name = ctypes.create_string_buffer(128)
        kp, np_, wo, bo, small = _I(), _I(), _I(), _I(), _I()
        kmap = (_I * 256)()
        nmap = (_I * 64)()
        check(lib.pgrf_w16_layer_info(i, name, 128, ctypes.byref(kp), ctypes.byref(np_), ctypes.byref(wo), ctypes.byref(bo),
                                      kmap, nmap, ctypes.byref(small)), "pgrf_w16_layer_info")
        bk, il, ol = _I(), _I(), _I()
        check(lib.pgrf_w16_layer_info2(i, ctypes.byref(bk), ctypes.byref(il), ctypes.byref(ol)), "pgrf_w16_layer_info2")
        out.append((name.value.decode(), kp.value, np_.value, wo.value, bo.value, list(kmap[:kp.value]), list(nmap[:np_.value]),
                    bool(small.value), bk.value, bool(il.value), bool(ol.value)))
    return out
