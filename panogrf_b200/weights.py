"""Packing of the hot-path parameters into the blob layout of csrc/render_layout.cuh.

Parameters keep the reference's `state_dict()` names (SURVEY.md §5 "Checkpoint / resume"):
`[fine_]dist_decoder.{mean,var,aw,vis}_decoder.{0,2,4}`, `[fine_]agg_net.prob_embed.{0,2}`,
`[fine_]agg_net.agg_impl.*` — so a reference `model.pth` loads unchanged.
"""
import numpy as np
import torch

from . import _lib


def posenc_table(d_hid, n_samples):
    """Sinusoid table of IBRNetWithNeuRay.posenc (network/ibrnet.py:305-313), fp64 -> fp32."""
    pos = np.arange(n_samples)[:, None].astype(np.float64)
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table).float()


def pack_blob(state, fine, n_samples, device, allow_missing=False):
    """state: mapping name -> tensor (a state_dict). Returns a (blob_floats,) fp32 tensor on `device`.

    `n_samples` is the length of the aggregation net's positional table (its cfg `sample_num`).
    `allow_missing`: layers absent from `state` stay zero (a stand-alone dist decoder or aggregation net whose
    counterpart is not evaluated: module-level API of panogrf_b200.renderer).
    """
    lib = _lib.load()
    dd = "fine_dist_decoder" if fine else "dist_decoder"
    agg = "fine_agg_net" if fine else "agg_net"
    blob = torch.zeros(lib.pgrf_weight_blob_floats(), dtype=torch.float32)
    for name, K, N, Npad, has_bias, k_begin, w_off, b_off in _lib.weight_layers():
        key = name.replace("{dd}", dd).replace("{agg}", agg)
        if key.endswith("ray_attention.qkv"):
            base = key[:-len(".qkv")]
            if allow_missing and base + ".w_qs.weight" not in state:
                continue
            w = torch.cat([state[base + ".w_qs.weight"], state[base + ".w_ks.weight"], state[base + ".w_vs.weight"]], 0)
            b = None
        elif key.endswith("ray_attention.fc"):
            if allow_missing and key + ".weight" not in state:
                continue
            w, b = state[key + ".weight"], None
        else:
            if key + ".weight" not in state:
                if ".vis_decoder." in key or allow_missing:   # use_vis == False: decoder absent, slots stay zero
                    continue
                raise KeyError(f"missing parameter {key}.weight")
            w = state[key + ".weight"]
            b = state.get(key + ".bias")
        w = w.detach().float().cpu()
        assert w.shape[0] == N and w.shape[1] >= k_begin + K, (key, tuple(w.shape), K, N)
        wt = torch.zeros(K, Npad)
        wt[:, :N] = w[:, k_begin:k_begin + K].t()
        blob[w_off:w_off + K * Npad] = wt.reshape(-1)
        if has_bias:
            assert b is not None, key
            blob[b_off:b_off + N] = b.detach().float().cpu()
    ln_off, pe_off, max_samples = _lib.weight_aux_offsets()
    lnk = agg + ".agg_impl.ray_attention.layer_norm"
    if lnk + ".weight" in state or not allow_missing:
        blob[ln_off:ln_off + 16] = state[lnk + ".weight"].detach().float().cpu()
        blob[ln_off + 16:ln_off + 32] = state[lnk + ".bias"].detach().float().cpu()
    if n_samples > max_samples:
        raise _lib.PanoGRFError(f"sample_num={n_samples} exceeds the kernel limit {max_samples}")
    blob[pe_off:pe_off + n_samples * 16] = posenc_table(16, n_samples).reshape(-1)
    return blob.to(device)


def _bf16_hi_lo(x):
    """fp32 vector -> (hi, lo) fp32 tensors holding bf16-representable values with hi + lo ~= x to ~2^-17 relative."""
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def pack_blob16(state, fine, device):
    """bf16 tensor-core blob (csrc/render_layout16.cuh): per layer the padded / permuted weight as a tcgen05 B
    operand [Kpad/8][Npad][8] bf16.  Section 0 (fused MLP kernel): the bias is a bf16 (hi, lo) pair, either in its own
    [Npad][8] chunk or in two spare K columns of the layer (kmap -2 / -3); layers flagged `out_log2e` are stored
    multiplied by log2(e) (weights and bias), layers flagged `in_ln2` have their weights multiplied by ln 2 (the
    producing ELU is evaluated on pre-scaled values).  Section 1 (rays kernel): fp32 biases.
    Returns a uint8 tensor of pgrf_w16_blob_bytes() bytes."""
    lib = _lib.load()
    dd = "fine_dist_decoder" if fine else "dist_decoder"
    agg = "fine_agg_net" if fine else "agg_net"
    LOG2E, LN2 = 1.4426950408889634, 0.6931471805599453
    blob = torch.zeros(lib.pgrf_w16_blob_bytes(), dtype=torch.uint8)
    for name, Kpad, Npad, w_off, b_off, kmap, nmap, is_small, bias_kind, in_ln2, out_log2e in _lib.w16_layers():
        key = name.replace("{dd}", dd).replace("{agg}", agg)
        if key.endswith("ray_attention.qkv"):
            base = key[:-len(".qkv")]
            w = torch.cat([state[base + ".w_qs.weight"], state[base + ".w_ks.weight"], state[base + ".w_vs.weight"]], 0)
            w, b = w.detach().float().cpu(), None
        else:
            if key + ".weight" not in state:
                if ".vis_decoder." in key:
                    continue
                raise KeyError(f"missing parameter {key}.weight")
            w = state[key + ".weight"].detach().float().cpu()
            b = state.get(key + ".bias")
        if b is not None:
            b = b.detach().float().cpu()
        scale = (LN2 if in_ln2 else 1.0) * (LOG2E if out_log2e else 1.0)
        if scale != 1.0:
            w = (w.double() * scale).float()
        if out_log2e and b is not None:
            b = (b.double() * LOG2E).float()
        if is_small:      # fp32 W[N][K] row-major + bias[N], evaluated as a register GEMV in the previous epilogue
            blob[w_off:w_off + Kpad * Npad * 4] = w[:Npad, :Kpad].contiguous().view(torch.uint8).reshape(-1)
            if b is not None:
                blob[b_off:b_off + Npad * 4] = b.contiguous().view(torch.uint8).reshape(-1)
            continue
        km = torch.tensor(kmap)
        nm = torch.tensor(nmap)
        wp = torch.zeros(Npad, Kpad)
        rows = torch.nonzero(nm >= 0).flatten()
        cols = torch.nonzero(km >= 0).flatten()
        wp[rows[:, None], cols[None, :]] = w[nm[rows][:, None], km[cols][None, :]]
        bp = torch.zeros(Npad)
        if b is not None:
            bp[rows] = b[nm[rows]]
        if bias_kind == _lib.BIAS_INLINE:     # two spare K columns carry (bias_hi, bias_lo); the kernel writes ones there
            hi, lo = _bf16_hi_lo(bp)
            wp[:, int(torch.nonzero(km == -2).flatten()[0])] = hi
            wp[:, int(torch.nonzero(km == -3).flatten()[0])] = lo
        # B operand: element (n, k) at [(k/8)][n][k%8]
        op = wp.reshape(Npad, Kpad // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)
        blob[w_off:w_off + Kpad * Npad * 2] = op.view(torch.uint8).reshape(-1)
        if bias_kind == _lib.BIAS_CHUNK:      # [Npad][8] bf16 = (hi, lo, 0 x6): B operand of the "ones" K-step
            hi, lo = _bf16_hi_lo(bp)
            ch = torch.zeros(Npad, 8)
            ch[:, 0] = hi
            ch[:, 1] = lo
            blob[b_off:b_off + Npad * 16] = ch.to(torch.bfloat16).view(torch.uint8).reshape(-1)
        elif bias_kind == _lib.BIAS_F32:
            blob[b_off:b_off + Npad * 4] = bp.view(torch.uint8).reshape(-1)
    return blob.to(device)
