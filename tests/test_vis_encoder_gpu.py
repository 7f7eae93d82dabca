"""GPU: DefaultVisEncoder on the tensor cores (csrc/conv3d.cu D = 1 mode + csrc/vis_encoder.cu) against outputs of the reference class,
and its element-wise kernels against torch on identically rounded operands."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import cases  # noqa: E402
from util import load_golden  # noqa: E402

pytestmark = pytest.mark.gpu

#: bf16 activations through 6 convolutions and 4 instance norms (fp32 accumulation): deviation relative to the output range
VISENC_TOL = 2e-2


@pytest.mark.parametrize("name", list(cases.VISENC_CASES))
def test_vis_encoder_matches_reference_golden(name):
    from panogrf_b200.vis_encoder import DefaultVisEncoder
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    net = DefaultVisEncoder({"use_wrap_padding": cases.VISENC_CASES[name][0]})
    net.load_state_dict(W)
    net = net.cuda()
    out = net(g["ray_feats"].cuda(), g["img_feats"].cuda()).cpu()
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    scale = float(g["out"].abs().max())
    err = float((out - g["out"]).abs().max())
    rms = float((out - g["out"]).pow(2).mean().sqrt())
    print(f"{name}: max err {err:.3e}, rms {rms:.3e} of range {scale:.3e} ({err / scale:.2e})")
    assert err <= VISENC_TOL * scale and rms <= 3e-3 * scale


def test_resize_concat_and_instance_norm_kernels_vs_torch():
    from panogrf_b200 import _lib
    lib = _lib.load()
    st = _lib.stream_ptr()
    torch.manual_seed(5)
    n, h, w, hi, wi = 2, 12, 20, 24, 40
    ray = torch.randn(n, 32, h, w, device="cuda")
    img = torch.randn(n, 32, hi, wi, device="cuda")
    a = torch.empty((n, h, w, 64), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.pgrf_feats_to_bf16_cl(_lib.ptr(img), 32, hi, wi, _lib.ptr(ray), 32, n, h, w, _lib.ptr(a), st), "feats_to_bf16_cl")
    ref = torch.cat([F.interpolate(img, (h, w), mode="bilinear"), ray], 1).permute(0, 2, 3, 1)
    assert float((a.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    # InstanceNorm2d(affine) + ReLU on the bf16 map
    x = a[..., :32].contiguous()
    gamma, beta = torch.randn(32, device="cuda"), torch.randn(32, device="cuda")
    y = torch.empty_like(x)
    stats = torch.empty(2 * n * 32, device="cuda", dtype=torch.float64)
    _lib.check(lib.pgrf_instnorm_relu_fwd(_lib.ptr(x), n, h * w, 32, _lib.ptr(gamma), _lib.ptr(beta), 1e-5, _lib.ptr(stats), _lib.ptr(y), st),
               "instnorm_relu")
    xr = x.float().permute(0, 3, 1, 2)
    ref = F.relu(F.instance_norm(xr, weight=gamma, bias=beta, eps=1e-5)).permute(0, 2, 3, 1)
    assert float((y.float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())


def test_renderer_runs_the_attached_vis_encoder():
    """network/renderer.py:639-642: without 'img_feats' in ref_imgs_info the renderer encodes first; the attached DefaultVisEncoder's
    parameters carry the reference's state_dict names (`vis_encoder.out_conv...`)."""
    import panogrf_b200 as pg
    from panogrf_b200.vis_encoder import DefaultVisEncoder
    from test_oracle_render import split_golden
    name = "render_m3d_2src"
    cfg, _, _ = cases.make_render_inputs(name)
    que, ref, W, _ = split_golden(load_golden(name))
    net = pg.NeuralRayBaseRenderer({**cfg, "mlp_dtype": "bf16"}).cuda().eval()
    net.load_state_dict(W, strict=False)
    net.vis_encoder = DefaultVisEncoder({"use_wrap_padding": True}).cuda()
    assert any(k.startswith("vis_encoder.out_conv.0.1.weight") for k in net.state_dict())
    cuda = lambda d: {k: v.cuda() for k, v in d.items()}
    q, r = cuda(que), cuda(ref)
    feats = r.pop("img_feats")
    net.image_encoder = lambda imgs: feats                      # stands in for ResUNetLight
    out = net.render(q, dict(r), False)
    r2 = dict(r)
    r2["img_feats"] = feats
    r2["ray_feats"] = net.vis_encoder(r["ray_feats"], feats)
    ref_out = net.render(q, r2, False)
    assert torch.equal(out["pixel_colors_nr_fine"], ref_out["pixel_colors_nr_fine"])
    assert torch.isfinite(out["pixel_colors_nr_fine"]).all()


@pytest.mark.parametrize("name", list(cases.INITCONV_CASES))
def test_init_net_convs_match_reference_golden(name):
    """CostVolumeInitNet's depth_conv / out_conv stacks (init_net.py:540-574, 629-636) vs the reference's own layers"""
    from panogrf_b200.vis_encoder import CostVolumeInitConvs
    g = load_golden(name)
    W = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    net = CostVolumeInitConvs({"use_wrap_padding": cases.INITCONV_CASES[name][0]})
    net.load_state_dict(W)
    net = net.cuda()
    out = net(g["ref_feats"].cuda(), g["depth"].cuda()).cpu()
    scale = float(g["ray_feats"].abs().max())
    err = float((out - g["ray_feats"]).abs().max())
    rms = float((out - g["ray_feats"]).pow(2).mean().sqrt())
    print(f"{name}: max err {err:.3e}, rms {rms:.3e} of range {scale:.3e} ({err / scale:.2e})")
    assert err <= VISENC_TOL * scale and rms <= 3e-3 * scale
