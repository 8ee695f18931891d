"""Per-kernel CPU restatements (torch fp32) — TEST INFRASTRUCTURE, see oracle/__init__.py.

One function per hand-written kernel family of flair_b200/csrc, each the plain-PyTorch statement of the
reference call the kernel replaces, used by tests/test_gpu_kernels.py on the 16-bit-rounded operands."""
from __future__ import annotations

import torch
import torch.nn.functional as F
import torchvision

from .unet_blur import BlurUNetOracle


def conv_cl(x, w, bias=None, *, stride=1, act=None, residual=None):
    """Channels-last [B,T,H,W,C] conv with (kt,kh,kw) kernel, zero padding k//2, stride in H/W
    (nn.Conv2d / nn.Conv3d of the torso, e.g. guided_diffusion/unet_new.py:240-244,271-276)."""
    kt, kh, kw = w.shape[2:]
    y = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), None if bias is None else bias.float(),
                 stride=(1, stride, stride), padding=(kt // 2, kh // 2, kw // 2))
    if act == "silu":
        y = F.silu(y)
    elif act == "relu":
        y = F.relu(y)
    elif act == "lrelu":
        y = F.leaky_relu(y, 0.1)
    y = y.permute(0, 2, 3, 4, 1)
    return y if residual is None else y + residual.float()


def group_norm_cl(x, gamma, beta, groups=32, *, scale=None, shift=None, silu=False, eps=1e-5):
    """GroupNorm32 over (C/G, T, H, W) of a [B,T,H,W,C] map (+ FiLM, + SiLU): nn_new.py:17-19 inside
    LazyReshaper3D (nn.py:359-367), unet_new.py:249-254,310-325.  scale/shift: [B*T, C]."""
    B, T, H, W, C = x.shape
    y = F.group_norm(x.float().permute(0, 4, 1, 2, 3), groups, gamma.float(), beta.float(), eps)
    y = y.permute(0, 2, 3, 4, 1)
    if scale is not None:
        y = y * (1 + scale.float().reshape(B, T, 1, 1, C)) + shift.float().reshape(B, T, 1, 1, C)
    return F.silu(y) if silu else y


def qkv_attention_legacy(qkv, heads):
    """QKVAttentionLegacy (unet_new.py:540-570) on a channels-last [B,T,H,W,3C] map with head-major (H,3,d)
    channels; attention over the H*W tokens of each frame; fp32 softmax; scale d^-1/4 on q and k."""
    B, T, H, W, C3 = qkv.shape
    d = C3 // (3 * heads)
    t = qkv.float().reshape(B * T, H * W, heads, 3, d)
    q, k, v = t[..., 0, :], t[..., 1, :], t[..., 2, :]
    s = d ** -0.25
    w = torch.softmax(torch.einsum("nlhd,nmhd->nhlm", q * s, k * s), dim=-1)
    return torch.einsum("nhlm,nmhd->nlhd", w, v).reshape(B, T, H, W, heads * d)


def flow_warp_cl(x, flow):
    """mmedit flow_warp (bilinear, zeros padding, align_corners=True; unet_new.py:706) on [N,H,W,C] maps."""
    return BlurUNetOracle.flow_warp(x.float().permute(0, 3, 1, 2), flow.permute(0, 2, 3, 1)).permute(0, 2, 3, 1)


def deform_align_core(xa, xb, o, flow_1, flow_2, weight, bias, mrm=10.0):
    """SecondOrderDeformableAlignment.forward after the offset net (unet_new.py:874-898): o is the raw 27*dg-channel
    output in REFERENCE channel order, NCHW; xa/xb NCHW; returns NCHW."""
    o1, o2, mask = torch.chunk(o.float(), 3, dim=1)
    offset = mrm * torch.tanh(torch.cat((o1, o2), dim=1))
    off1, off2 = torch.chunk(offset, 2, dim=1)
    off1 = off1 + flow_1.flip(1).repeat(1, off1.size(1) // 2, 1, 1)
    off2 = off2 + flow_2.flip(1).repeat(1, off2.size(1) // 2, 1, 1)
    x = torch.cat([xa, xb], 1).float()
    return torchvision.ops.deform_conv2d(x, torch.cat([off1, off2], dim=1), weight.float(),
                                         None if bias is None else bias.float(), 1, 1, 1, torch.sigmoid(mask))
