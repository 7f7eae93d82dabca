"""CPU restatement of the 3-D cost regulariser `unet3d` (TEST INFRASTRUCTURE, never imported by the product).

Reference: the UNet2 of Conv3DBlockv2 built by models/test_models.py:81-146 (initialize_cost_volume_network: encoders
2^(i+size+1) -> 2^(i+size+2) with AvgPool3d(2), a last un-pooled encoder, decoders on [upsampled | skip]), the blocks in
models/common_blocks.py:366-503 (Conv3DBlockv2 = pad-conv-lrelu-pad-conv-lrelu-pool, WrapPadding3D = zeros in depth / height,
wrap in width) and the U-Net wiring in models/common_blocks.py:187-242 (UNet2.forward, trilinear x2 upsampling,
align_corners=False).  Consumed by network/omni_mvsnet/pipeline3_model.py:847-855 (`self.unet3d(cost_volume)[:, 0]`).
Pinned by tests/golden/unet3d_*.npz: outputs of the reference classes with seeded weights (tests/golden/make_golden_unet3d.py).
Weights are a state_dict with the reference's names: `encoders.{i}.conv{1,2}.{weight,bias}`, `decoders.{i}.conv{1,2}.*`.
"""
import torch
import torch.nn.functional as F


def wrap_pad3d(x, pad=(1, 1, 1)):
    """WrapPadding3D (common_blocks.py:448-503): zeros along depth and height, wrap along width."""
    pd, ph, pw = pad
    if pd:
        x = F.pad(x, (0, 0, 0, 0, pd, pd))
    if ph:
        x = F.pad(x, (0, 0, ph, ph, 0, 0))
    if pw:
        x = torch.cat([x[..., -pw:], x, x[..., :pw]], -1)
    return x


def conv_block(W, prefix, x, pool):
    """Conv3DBlockv2.forward (common_blocks.py:427-445) with use_wrap_padding, no batch norm, LeakyReLU(0.01)."""
    for k in ("conv1", "conv2"):
        x = F.leaky_relu(F.conv3d(wrap_pad3d(x), W[f"{prefix}.{k}.weight"], W[f"{prefix}.{k}.bias"]), 0.01)
    return (F.avg_pool3d(x, 2) if pool else x), x


def unet3d(W, x, n_enc=4):
    """UNet2.forward (common_blocks.py:211-242): x (B,C,D,H,W) -> (B,1,D,H,W)."""
    skips = []
    for i in range(n_enc):
        x, unpooled = conv_block(W, f"encoders.{i}", x, pool=i < n_enc - 1)
        skips.append(unpooled)
    n_dec = n_enc - 1
    x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
    x, _ = conv_block(W, f"decoders.{n_dec - 1}", x, pool=False)
    for i in range(n_dec - 2, -1, -1):
        x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        x = torch.cat((x, skips[i]), dim=1)
        x, _ = conv_block(W, f"decoders.{i}", x, pool=False)
    return x


# ---- 2-D heads after the regulariser (models/test_models.py:147-205, pipeline3_model.py:866-905) -----------------------------------
def wrap_pad2d(x):
    """WrapPadding (common_blocks.py:258-293): zeros along height, wrap along width."""
    x = F.pad(x, (0, 0, 1, 1))
    return torch.cat([x[..., -1:], x, x[..., :1]], -1)


def rectify(x, out_type):
    """pipeline3_model.py:875-879: disparity -> depth, or clamp."""
    return 1.0 / (torch.clamp(x, min=0) + 1e-10) if out_type == "disparity" else torch.clamp(x, min=0)


def decoders1(W, cost_reg, out_type):
    """ConvBlock(kernel 1, no norm / activation) + x4 bilinear + rectification -> raw (B,4H,4W,1), depth (B,4H,4W,1)"""
    y = F.conv2d(cost_reg, W["decoders1.conv.weight"], W["decoders1.conv.bias"])
    raw = F.interpolate(y, scale_factor=4, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    return raw, rectify(raw, out_type)


def decoders2(W, feats):
    """three ConvBlock2 (common_blocks.py:96-184): [x2 bilinear] pad-conv-lrelu-pad-conv-lrelu; the last one without upscale / activation"""
    for i, (up, act) in enumerate(((True, True), (True, True), (False, False))):
        if up:
            feats = F.interpolate(feats, scale_factor=2, mode="bilinear", align_corners=False)
        for k in ("conv1", "conv2"):
            feats = F.conv2d(wrap_pad2d(feats), W[f"decoders2.{i}.{k}.weight"], W[f"decoders2.{i}.{k}.bias"])
            if act:
                feats = F.leaky_relu(feats, 0.01)
    return feats
