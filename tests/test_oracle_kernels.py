"""CPU cross-checks of oracle/kernels.py: the per-kernel restatements that lean on library calls (torchvision
deform_conv2d, F.grid_sample, F.group_norm) are compared with naive loop implementations of the published rules on tiny
inputs, so that the GPU parity tests of tests/test_gpu_kernels.py do not rest on a library's reading of an operator
alone.  (The mmedit / mmcv pieces of the reference are not vendored: SURVEY 8c.)"""
import math

import torch

from oracle import kernels as K


def _bilinear_zero(img, y, x):
    """torchvision deform_conv2d `bilinear_interpolate`: 0 outside (-1, H) x (-1, W); corners outside the map add 0."""
    H, W = img.shape[-2:]
    if y <= -1 or y >= H or x <= -1 or x >= W:
        return torch.zeros(img.shape[:-2])
    y0, x0 = math.floor(y), math.floor(x)
    ly, lx = y - y0, x - x0
    out = torch.zeros(img.shape[:-2])
    for yy, wy in ((y0, 1 - ly), (y0 + 1, ly)):
        for xx, wx in ((x0, 1 - lx), (x0 + 1, lx)):
            if 0 <= yy <= H - 1 and 0 <= xx <= W - 1:
                out = out + wy * wx * img[..., yy, xx]
    return out


def test_deform_align_core_matches_naive_loops():
    """Second-order deformable alignment after the offset net (unet_new.py:874-898): 10*tanh residual + flipped flow,
    sigmoid mask, 16 deform groups over cat(xa, xb), 3x3 taps, zero padding."""
    g = torch.Generator().manual_seed(0)
    C, H, W, dg, mrm = 16, 5, 6, 16, 10.0
    xa, xb = torch.randn(1, C, H, W, generator=g), torch.randn(1, C, H, W, generator=g)
    o = torch.randn(1, 27 * dg, H, W, generator=g) * 0.4
    f1, f2 = torch.randn(1, 2, H, W, generator=g) * 1.5, torch.randn(1, 2, H, W, generator=g) * 2.5
    weight, bias = torch.randn(C, 2 * C, 3, 3, generator=g) * 0.2, torch.randn(C, generator=g)
    got = K.deform_align_core(xa, xb, o, f1, f2, weight, bias, mrm)

    x = torch.cat([xa, xb], 1)[0]                      # (2C, H, W)
    cpg = 2 * C // dg
    ref = torch.zeros(C, H, W)
    for h in range(H):
        for w in range(W):
            col = torch.zeros(2 * C, 9)
            for grp in range(dg):
                flow = f1 if grp < dg // 2 else f2      # o1 (groups 0..7) + flow_1.flip, o2 (groups 8..15) + flow_2.flip
                for tap in range(9):
                    dy = mrm * math.tanh(float(o[0, (grp * 9 + tap) * 2, h, w])) + float(flow[0, 1, h, w])
                    dx = mrm * math.tanh(float(o[0, (grp * 9 + tap) * 2 + 1, h, w])) + float(flow[0, 0, h, w])
                    m = 1.0 / (1.0 + math.exp(-float(o[0, 18 * dg + grp * 9 + tap, h, w])))
                    sy, sx = h + tap // 3 - 1 + dy, w + tap % 3 - 1 + dx
                    col[grp * cpg:(grp + 1) * cpg, tap] = m * _bilinear_zero(x[grp * cpg:(grp + 1) * cpg], sy, sx)
            ref[:, h, w] = (weight.reshape(C, 2 * C, 9) * col[None]).sum((1, 2)) + bias
    assert float((got[0] - ref).abs().max()) < 1e-4


def test_flow_warp_matches_naive_loops():
    """mmedit flow_warp: bilinear, zeros padding, align_corners=True == sample at (w + fx, h + fy) in pixel units."""
    g = torch.Generator().manual_seed(1)
    N, H, W, C = 1, 5, 7, 8
    x = torch.randn(N, H, W, C, generator=g)
    flow = torch.randn(N, 2, H, W, generator=g) * 2.5
    got = K.flow_warp_cl(x, flow)
    img = x[0].permute(2, 0, 1)
    for h in range(H):
        for w in range(W):
            sx, sy = w + float(flow[0, 0, h, w]), h + float(flow[0, 1, h, w])
            y0, x0 = math.floor(sy), math.floor(sx)
            ref = torch.zeros(C)
            for yy, wy in ((y0, 1 - (sy - y0)), (y0 + 1, sy - y0)):
                for xx, wx in ((x0, 1 - (sx - x0)), (x0 + 1, sx - x0)):
                    if 0 <= yy < H and 0 <= xx < W:
                        ref = ref + wy * wx * img[:, yy, xx]
            assert float((got[0, h, w] - ref).abs().max()) < 1e-5


def test_group_norm_cl_matches_definition():
    g = torch.Generator().manual_seed(2)
    B, T, H, C, G = 2, 3, 4, 16, 4
    x = torch.randn(B, T, H, H, C, generator=g) * 1.7 + 0.4
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    scale, shift = torch.randn(B * T, C, generator=g) * 0.3, torch.randn(B * T, C, generator=g)
    got = K.group_norm_cl(x, gamma, beta, G, scale=scale, shift=shift, silu=True)
    xg = x.reshape(B, T * H * H, G, C // G)
    mean = xg.mean((1, 3), keepdim=True)                     # statistics over (T, H, W, C/G) of one batch element
    var = ((xg - mean) ** 2).mean((1, 3), keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + 1e-5)).reshape(B, T, H, H, C) * gamma + beta
    y = y * (1 + scale.reshape(B, T, 1, 1, C)) + shift.reshape(B, T, 1, 1, C)
    ref = y * torch.sigmoid(y)
    assert float((got - ref).abs().max()) < 1e-5


def test_qkv_attention_legacy_matches_definition():
    """QKVAttentionLegacy (unet_new.py:540-570): channels are head-major (H, 3, d); scale d^-1/4 on q and on k."""
    g = torch.Generator().manual_seed(3)
    heads, d, L = 2, 64, 3
    qkv = torch.randn(1, 1, L, L, heads * 3 * d, generator=g)
    got = K.qkv_attention_legacy(qkv, heads).reshape(L * L, heads, d)
    t = qkv.reshape(L * L, heads, 3, d)
    for hd in range(heads):
        q, k, v = t[:, hd, 0], t[:, hd, 1], t[:, hd, 2]
        w = torch.softmax((q * d ** -0.25) @ (k * d ** -0.25).t(), dim=-1)
        assert float((got[:, hd] - w @ v).abs().max()) < 1e-5
