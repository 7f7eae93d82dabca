"""Times the 3-D cost regulariser at the shipped size (size=4: 32 -> 64 -> 128 -> 256 -> 512 channels) on a synthetic cost volume and
reports tensor-core throughput per layer group.   python tools/time_regulariser.py [D H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panogrf_b200 import _lib  # noqa: E402
from panogrf_b200 import regulariser as reg  # noqa: E402


def unet_flops(size, D, H, W, B=1):
    """2 * MACs of every convolution of the U-Net (27 taps)."""
    tot = 0
    vox = B * D * H * W
    layers = []
    for i in range(3):
        ch = 2 ** (i + size + 1)
        layers += [(ch, 2 * ch, vox // 8 ** i), (2 * ch, 2 * ch, vox // 8 ** i)]
    ch = 2 ** (3 + size + 1)
    layers += [(ch, 2 * ch, vox // 512), (2 * ch, 2 * ch, vox // 512)]
    layers += [(2 ** (size + 5), 2 ** (size + 3), vox // 64), (2 ** (size + 3), 2 ** (size + 3), vox // 64)]      # decoders.2
    layers += [(2 ** (size + 4), 2 ** (size + 2), vox // 8), (2 ** (size + 2), 2 ** (size + 2), vox // 8)]        # decoders.1
    layers += [(2 ** (size + 3), 1, vox), (1, 1, vox)]                                                            # decoders.0
    for ci, co, n in layers:
        tot += 2 * 27 * ci * co * n
    return tot, layers


def torch_unet3d(net, x):
    """The same network through torch's library convolutions (cuDNN), as the reference runs it: wrap padding by concatenation,
    F.conv3d, F.avg_pool3d, F.interpolate — the library baseline the tensor-core kernels are compared with."""
    import torch.nn.functional as F

    def pad(t):
        t = F.pad(t, (0, 0, 1, 1, 1, 1))
        return torch.cat([t[..., -1:], t, t[..., :1]], -1)

    def block(blk, t):
        for conv in (blk.conv1, blk.conv2):
            t = F.leaky_relu(F.conv3d(pad(t), conv.weight.to(t.dtype), conv.bias.to(t.dtype)), 0.01)
        return t

    skips = []
    for blk in net.encoders:
        u = block(blk, x)
        skips.append(u)
        x = F.avg_pool3d(u, 2) if blk.pool else u
    n_dec = len(net.decoders)
    for i in range(n_dec - 1, -1, -1):
        x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        if i < n_dec - 1:
            x = torch.cat((x, skips[i]), 1)
        x = block(net.decoders[i], x)
    return x


def time_fn(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


if __name__ == "__main__":
    D, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 64, 128)
    torch.manual_seed(0)
    net = reg.CostRegulariser3D(4).cuda()
    x = torch.rand(1, 32, D, H, W, device="cuda")
    for _ in range(3):
        y = net(x)
    torch.cuda.synchronize()
    n = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    l0 = _lib.launch_count()
    ev[0].record()
    for _ in range(n):
        y = net(x)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    fl, layers = unet_flops(4, D, H, W)
    print(f"unet3d size=4 on 1x32x{D}x{H}x{W}: {ms:.3f} ms, {fl / 1e9:.1f} GFLOP -> {fl / ms / 1e9:.1f} TFLOP/s, "
          f"{(_lib.launch_count() - l0) // n} launches, finite={bool(torch.isfinite(y).all())}")
    if os.environ.get("PGRF_TIME_TORCH", "1") != "0":
        with torch.no_grad():
            ref = torch_unet3d(net, x)
            err = float((y - ref).abs().max() / ref.abs().max())
            t32 = time_fn(lambda: torch_unet3d(net, x), 5)
            xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
            t16 = time_fn(lambda: torch_unet3d(net, xb), 5)
        print(f"library baseline (torch/cuDNN, same network): fp32 (TF32 convolutions, the reference's default) {t32:.3f} ms, "
              f"bf16 channels_last_3d {t16:.3f} ms; ours vs fp32 library output: max err {err:.2e} of range")
