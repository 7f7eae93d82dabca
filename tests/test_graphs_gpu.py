"""GPU: CUDA-graph replay of the per-view CNNs (panogrf_b200/graphs.py) == the eager launches, bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(graphed, eager, make_inputs):
    a = make_inputs(0)
    y0 = eager(*a)
    assert torch.equal(graphed(*a), y0)                       # capture + first replay
    b = make_inputs(1)
    yb = graphed(*b)                                          # replay with new data copied into the static buffers
    assert torch.equal(yb, eager(*b))
    assert not torch.equal(yb, y0)
    assert torch.equal(graphed(*a), y0)                       # outputs are copies: earlier results stay valid


def test_graphed_image_encoder():
    from panogrf_b200.graphs import GraphedForward
    from panogrf_b200.image_encoder import ResUNetLight
    torch.manual_seed(0)
    net = ResUNetLight({}, 3, [1, 2, 6, 4], 32, inplanes=16, use_wrap_padding=True).cuda()
    g = GraphedForward(net)
    mk = lambda s: (torch.rand(2, 3, 64, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(s)),)
    _check(g, net, mk)
    assert len(g._graphs) == 1
    g(torch.rand(1, 3, 32, 64, device="cuda"))                # another signature: another graph
    assert len(g._graphs) == 2
    # a parameter update drops the capture (the packed bf16 weights are part of it)
    with torch.no_grad():
        net.out_conv.bias.add_(0.5)
    x = mk(2)[0]
    assert torch.equal(g(x), net(x))


def test_graphed_vis_encoder_and_regulariser():
    from panogrf_b200 import regulariser as reg
    from panogrf_b200.graphs import GraphedForward
    from panogrf_b200.vis_encoder import DefaultVisEncoder
    torch.manual_seed(1)
    vis = DefaultVisEncoder({"use_wrap_padding": True}).cuda()
    gen = lambda s: torch.Generator(device="cuda").manual_seed(s)
    _check(GraphedForward(vis), vis, lambda s: (torch.randn(2, 32, 16, 32, device="cuda", generator=gen(s)),
                                                torch.randn(2, 32, 32, 64, device="cuda", generator=gen(s + 10))))
    unet = reg.CostRegulariser3D(1).cuda()
    _check(GraphedForward(unet), unet, lambda s: (torch.randn(1, 4, 8, 16, 16, device="cuda", generator=gen(s)),))


def test_graphed_forward_refuses_cpu_tensors():
    from panogrf_b200._lib import PanoGRFError
    from panogrf_b200.graphs import GraphedForward
    with pytest.raises(PanoGRFError):
        GraphedForward(lambda x: x)(torch.zeros(4))
