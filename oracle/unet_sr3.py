"""SR3 bicubic denoiser forward on the CPU (torch fp32, functional, reference-named state dict) —
restates guided_diffusion/sr3.py: PositionalEncoding :45-60, FeatureWiseAffine :63-83, Block :113-124,
ResnetBlock :127-161, TemporalWrapper2 :197-226, ResnetBlocWithAttn :229-315, UNet.__init__/forward
:318-525, with the cross-frame modules of guided_diffusion/unet.py (ResBlock (3,1,1) :113-254,
TemporalAttention F=7 :664-758, BasicVSRPP :313-595 incl. its antialiased resize of `lqs` and per-module
SPyNet call).  mmedit / flash-attn pieces as in oracle/unet_blur.py (PARITY UNPINNED for those)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .unet_blur import BlurUNetOracle


def default_config(image_size=256):
    """MODEL_CONFIG["x8_bicubic"] of scripts/video_sample.py:78-96 with image_size overridable."""
    return dict(image_size=image_size, in_channel=6, out_channel=3, inner_channel=64, norm_groups=16,
                channel_mults=(1, 2, 4, 8, 16), attn_res=(64, 32), vsrpp_res=(512, 256), temporal_attn=True,
                res_blocks=1, cross_frame_module=True, num_frames=7, head_dim=64)


def plan(cfg):
    """(downs, mid, ups): per layer ('conv_in'|'block'|'down'|'up', dim_in, dim_out, temp_attn, vsrpp)."""
    ic, mults, rb = cfg["inner_channel"], cfg["channel_mults"], cfg["res_blocks"]
    cf = cfg["cross_frame_module"]
    pre, feat, now = ic, [ic], cfg["image_size"]
    downs = [("conv_in",)]
    for ind, m in enumerate(mults):
        ta = now in cfg["attn_res"] and cfg["temporal_attn"] and cf
        vs = now in cfg["vsrpp_res"] and cf
        ch = ic * m
        for _ in range(rb):
            downs.append(("block", pre, ch, ta, vs))
            feat.append(ch)
            pre = ch
        if ind != len(mults) - 1:
            downs.append(("down", pre))
            feat.append(pre)
            now //= 2
    mid = [("block", pre, pre, cfg["temporal_attn"] and cf, False)] * 2
    ups = []
    for ind in reversed(range(len(mults))):
        ta = now in cfg["attn_res"] and cfg["temporal_attn"] and cf
        vs = now in cfg["vsrpp_res"] and cf
        ch = ic * mults[ind]
        for _ in range(rb + 1):
            ups.append(("block", pre + feat.pop(), ch, ta, vs))
            pre = ch
        if ind >= 1:
            ups.append(("up", pre))
            now *= 2
    return downs, mid, ups


class SR3UNetOracle(BlurUNetOracle):
    def __init__(self, cfg, state_dict):
        self.cfg = cfg
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.downs, self.mid, self.ups = plan(cfg)
        self.heads_dim = cfg["head_dim"]
        self.cross = cfg["cross_frame_module"]

    def gn_g(self, x, pre, groups):
        y = F.group_norm(x.permute(0, 2, 1, 3, 4), groups, self.p(pre + ".wrapped_module.weight"),
                         self.p(pre + ".wrapped_module.bias"), eps=1e-5)
        return y.permute(0, 2, 1, 3, 4)

    def block(self, x, pre):  # Block: GN -> Swish -> conv3x3
        return self.conv2d(F.silu(self.gn_g(x, pre + ".block.0", self.cfg["norm_groups"])), pre + ".block.3.wrapped_module")

    def resnet_block(self, x, t, pre):
        h = self.block(x, pre + ".block1")
        e = F.linear(t, self.p(pre + ".noise_func.noise_func.0.weight"), self.p(pre + ".noise_func.noise_func.0.bias"))
        h = h + e.view(x.shape[0], x.shape[1], -1, 1, 1)
        h = self.block(h, pre + ".block2")
        if pre + ".res_conv.wrapped_module.weight" in self.sd:
            x = self.conv2d(x, pre + ".res_conv.wrapped_module")
        return h + x

    def temporal_resblock(self, x, t, pre):  # unet.ResBlock dims=3, kernel (3,1,1), h + emb conditioning
        def conv(v, name):
            w = self.p(name + ".weight")
            return F.conv3d(v.permute(0, 2, 1, 3, 4), w, self.p(name + ".bias"), padding=(1, 0, 0)).permute(0, 2, 1, 3, 4)
        h = conv(F.silu(self.gn(x, pre + ".in_layers.0")), pre + ".in_layers.2.wrapped_module")
        e = F.linear(F.silu(t), self.p(pre + ".emb_layers.1.weight"), self.p(pre + ".emb_layers.1.bias"))
        h = h + e.reshape(x.shape[0], x.shape[1], -1)[..., None, None]
        h = conv(F.silu(self.gn(h, pre + ".out_layers.0")), pre + ".out_layers.3.wrapped_module")
        return x + h

    def gated(self, x, t, pre, fn):  # TemporalWrapper2
        out = fn(pre + ".wrapped_module")
        g = torch.sigmoid(F.linear(F.silu(t), self.p(pre + ".emb_layers.1.weight"), self.p(pre + ".emb_layers.1.bias")))
        g = g.view(x.shape[0], x.shape[1], -1, 1, 1)
        return (1 - g) * x + g * out

    def vsrpp(self, hidden, lqs, weight, pre):  # unet.BasicVSRPP.forward
        h, w = hidden.shape[-2:]
        if lqs.shape[-2] != h or lqs.shape[-1] != w:
            lqs = F.interpolate(lqs.flatten(0, 1), size=(h, w), mode="bilinear", align_corners=False,
                                antialias=True).unflatten(0, lqs.shape[:2])
        assert lqs.size(3) >= 64 and lqs.size(4) >= 64
        self._sp = pre + ".spynet"
        ff, fb = self.compute_flow(lqs)
        return self.basicvsrpp5(hidden, ff, fb, weight, pre)

    def conv(self, x, pre, **kw):  # SPyNet lives under each BasicVSR++ module in sr3 state dicts
        if pre.startswith("spynet."):
            pre = self._sp + pre[len("spynet"):]
        return super().conv(x, pre, **kw)

    def p(self, name):
        if name.startswith("spynet."):
            name = self._sp + name[len("spynet"):]
        return self.sd[name]

    def basicvsrpp5(self, hidden, ff, fb, weight, pre):
        real = self.deform_align
        self.deform_align = lambda x, e, f1, f2, p_, mrm=5: real(x, e, f1, f2, p_, mrm=5)  # max_residue_magnitude=5
        try:
            return self.basicvsrpp(hidden, ff, fb, weight, pre)
        finally:
            self.deform_align = real

    def run_block(self, layer, pre, x, t, lqs, weights, cross):
        x = self.resnet_block(x, t, pre + ".res_block")
        if not cross or not self.cross:
            return x
        x = self.gated(x, t, pre + ".conv_3d", lambda p_: self.temporal_resblock(x, t, p_))
        if layer[3]:
            x = self.gated(x, t, pre + ".temp_attn", lambda p_: self.temporal_attention(x, p_, frames=self.cfg["num_frames"]))
        if layer[4]:
            x = self.gated(x, t, pre + ".vsrpp", lambda p_: self.vsrpp(x, lqs, weights, p_))
        return x

    def forward(self, x, noise_level, low_res_input, num_frames, rnn_input=None, enable_cross_frames=True,
                vsrpp_weights=None):
        rnn_input = low_res_input if rnn_input is None else rnn_input
        x = x.reshape(-1, num_frames, *x.shape[1:])
        x = torch.cat((low_res_input, x), dim=2)
        count = self.cfg["inner_channel"] // 2
        step = torch.arange(count, dtype=noise_level.dtype) / count
        enc = noise_level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
        enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
        t = F.linear(F.silu(F.linear(enc, self.p("noise_level_mlp.1.weight"), self.p("noise_level_mlp.1.bias"))),
                     self.p("noise_level_mlp.3.weight"), self.p("noise_level_mlp.3.bias"))
        feats = []
        for i, layer in enumerate(self.downs):
            if layer[0] == "conv_in":
                x = self.conv2d(x, f"downs.{i}.wrapped_module")
            elif layer[0] == "down":
                w = self.p(f"downs.{i}.wrapped_module.conv.weight")
                x = F.conv2d(x.flatten(0, 1), w, self.p(f"downs.{i}.wrapped_module.conv.bias"), stride=2,
                             padding=1).unflatten(0, x.shape[:2])
            else:
                x = self.run_block(layer, f"downs.{i}", x, t, rnn_input, vsrpp_weights, enable_cross_frames)
            feats.append(x)
        for i, layer in enumerate(self.mid):
            x = self.run_block(layer, f"mid.{i}", x, t, rnn_input, vsrpp_weights, enable_cross_frames)
        for i, layer in enumerate(self.ups):
            if layer[0] == "up":
                up = F.interpolate(x.flatten(0, 1), scale_factor=2, mode="nearest")
                x = F.conv2d(up, self.p(f"ups.{i}.wrapped_module.conv.weight"), self.p(f"ups.{i}.wrapped_module.conv.bias"),
                             padding=1).unflatten(0, x.shape[:2])
            else:
                x = self.run_block(layer, f"ups.{i}", torch.cat((x, feats.pop()), dim=2), t, rnn_input, vsrpp_weights,
                                   enable_cross_frames)
        return self.block(x, "final_conv").flatten(0, 1)
