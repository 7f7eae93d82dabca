"""CPU: weight-blob layouts exported by the C side and the Python packers agree (no GPU needed)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402


def _net():
    from panogrf_b200.renderer import NeuralRayBaseRenderer
    torch.manual_seed(0)
    cfg = cases.render_cfg(use_vis=True)
    net = NeuralRayBaseRenderer(cfg)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn_like(p))
    return net


def test_fp32_blob_roundtrip():
    from panogrf_b200 import _lib
    from panogrf_b200.weights import pack_blob
    net = _net()
    sd = net.state_dict()
    blob = pack_blob(sd, False, 16, "cpu")
    assert blob.numel() == _lib.load().pgrf_weight_blob_floats()
    seen = 0
    for name, K, N, Npad, has_bias, k_begin, w_off, b_off in _lib.weight_layers():
        key = name.replace("{dd}", "dist_decoder").replace("{agg}", "agg_net")
        if key.endswith(".qkv"):
            base = key[:-4]
            w = torch.cat([sd[base + ".w_qs.weight"], sd[base + ".w_ks.weight"], sd[base + ".w_vs.weight"]], 0)
        else:
            w = sd[key + ".weight"]
        wt = blob[w_off:w_off + K * Npad].reshape(K, Npad)
        assert torch.equal(wt[:, :N], w[:, k_begin:k_begin + K].t()), key
        assert float(wt[:, N:].abs().sum()) == 0
        if has_bias:
            assert torch.equal(blob[b_off:b_off + N], sd[key + ".bias"]), key
        seen += 1
    assert seen == 34


def test_bf16_blob_layout_and_roundtrip():
    """Every slot of the bf16 blob decodes back to the reference parameter it was packed from: permutations, zero padding,
    the log2(e) / ln 2 activation pre-scaling, and the three bias encodings (fp32 vector, bf16 hi/lo chunk, inline hi/lo
    K columns)."""
    from panogrf_b200 import _lib
    from panogrf_b200.weights import pack_blob16
    net = _net()
    sd = net.state_dict()
    blob = pack_blob16(sd, False, "cpu")
    assert blob.numel() == _lib.load().pgrf_w16_blob_bytes()
    LOG2E, LN2 = 1.4426950408889634, 0.6931471805599453
    spans = []
    n_chunk = n_inline = 0
    for name, Kpad, Npad, w_off, b_off, kmap, nmap, small, bias_kind, in_ln2, out_log2e in _lib.w16_layers():
        key = name.replace("{dd}", "dist_decoder").replace("{agg}", "agg_net")
        if key.endswith(".qkv"):
            base = key[:-4]
            w = torch.cat([sd[base + ".w_qs.weight"], sd[base + ".w_ks.weight"], sd[base + ".w_vs.weight"]], 0)
            b = None
        else:
            w, b = sd[key + ".weight"], sd.get(key + ".bias")
        scale = (LN2 if in_ln2 else 1.0) * (LOG2E if out_log2e else 1.0)
        w = (w.double() * scale).float()
        if b is not None and out_log2e:
            b = (b.double() * LOG2E).float()
        if small:
            got = blob[w_off:w_off + Kpad * Npad * 4].view(torch.float32).reshape(Npad, Kpad)
            assert torch.equal(got, w), key
            assert torch.equal(blob[b_off:b_off + Npad * 4].view(torch.float32), b), key
            spans.append((w_off, w_off + Kpad * Npad * 4))
            spans.append((b_off, b_off + Npad * 4))
            continue
        assert Kpad % 8 == 0 and Npad % 16 == 0 and w_off % 16 == 0, key
        ks = [k for k in kmap if k >= 0]
        ns = [n for n in nmap if n >= 0]
        assert len(set(ks)) == len(ks) == w.shape[1] and len(set(ns)) == len(ns) == w.shape[0], key   # a permutation + padding
        op = blob[w_off:w_off + Kpad * Npad * 2].view(torch.bfloat16).reshape(Kpad // 8, Npad, 8)
        dense = op.permute(1, 0, 2).reshape(Npad, Kpad).float()               # [n][k]
        for n_slot in (0, Npad // 2, Npad - 1):
            for k_slot in (0, 7, Kpad // 2, Kpad - 1):
                n, k = nmap[n_slot], kmap[k_slot]
                if k < -1:
                    continue
                expect = float(w[n, k].bfloat16()) if n >= 0 and k >= 0 else 0.0
                assert float(dense[n_slot, k_slot]) == expect, (key, n_slot, k_slot)
        spans.append((w_off, w_off + Kpad * Npad * 2))
        bexp = torch.zeros(Npad)
        if b is not None:
            for n_slot, n in enumerate(nmap):
                if n >= 0:
                    bexp[n_slot] = b[n]
        if bias_kind == _lib.BIAS_F32:
            assert torch.equal(blob[b_off:b_off + Npad * 4].view(torch.float32), bexp), key
            spans.append((b_off, b_off + Npad * 4))
        elif bias_kind == _lib.BIAS_CHUNK:
            ch = blob[b_off:b_off + Npad * 16].view(torch.bfloat16).reshape(Npad, 8).float()
            assert float(ch[:, 2:].abs().sum()) == 0, key
            assert float((ch[:, 0] + ch[:, 1] - bexp).abs().max()) <= 2e-5 * float(bexp.abs().max()) + 1e-12, key
            spans.append((b_off, b_off + Npad * 16))
            n_chunk += 1
        elif bias_kind == _lib.BIAS_INLINE:
            assert b_off == -1, key
            hi, lo = kmap.index(-2), kmap.index(-3)
            assert float((dense[:, hi] + dense[:, lo] - bexp).abs().max()) <= 2e-5 * float(bexp.abs().max()) + 1e-12, key
            n_inline += 1
    assert n_chunk == 16 and n_inline == 3
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "overlapping regions in the bf16 blob"


def test_render_argument_validation_without_gpu():
    """Shape validation of the render entry points happens before any CUDA call."""
    from panogrf_b200 import _lib
    lib = _lib.load()
    a = _lib.RenderArgs()
    a.dataset, a.rfn, a.dn, a.rn = 0, 9, 64, 10
    rc = lib.pgrf_render_pass_fwd(ctypes.byref(a), None)
    assert rc == _lib.PGRF_EINVAL and b"source views" in lib.pgrf_last_error()
    a.rfn, a.dn = 2, 500
    rc = lib.pgrf_render_pass_fwd(ctypes.byref(a), None)
    assert rc == _lib.PGRF_EINVAL and b"samples per ray" in lib.pgrf_last_error()
    f1, f2 = ctypes.c_longlong(), ctypes.c_longlong()
    assert lib.pgrf_render_workspace(2, 64 * 100, ctypes.byref(f1), ctypes.byref(f2)) == 0
    # f2 covers both hand-off formats: fp32 [68][T] tiles, and the bf16 path's operand tiles (9 x 2 KB per <= 128 samples,
    # > 64 of them used) + one float4 per sample
    assert f1.value == 100 * 76 * 128 and f2.value == max(100 * 68 * 64, ((6400 // 65 + 2) * 9 * 2048 + 6400 * 16 + 3) // 4)
    assert lib.pgrf_render_workspace(7, 64, ctypes.byref(f1), ctypes.byref(f2)) == _lib.PGRF_EINVAL


def test_checkpoint_names_match_reference_goldens():
    """The parameter containers expose exactly the reference's state_dict names for the hot path."""
    from panogrf_b200.renderer import NeuralRayBaseRenderer
    from util import load_golden
    g = load_golden("render_m3d_vis_nodisp")
    ref_names = sorted(k[2:] for k in g if k.startswith("w."))
    cfg, _, _ = cases.make_render_inputs("render_m3d_vis_nodisp")
    ours = sorted(NeuralRayBaseRenderer(cfg).state_dict().keys())
    assert ours == ref_names
