"""Blur/JPEG denoiser forward on the CPU (torch fp32, functional, driven by a reference-named
state dict) — restates guided_diffusion/unet_new.py:

  UNetModel.__init__/forward :932-1222,1311-1362   ResBlock 2-D / 3-D :198-329
  AttentionBlock / AttentionbottleBlock :332-429    QKVAttentionLegacy :540-570
  TemporalAttention :432-517                        BasicVSRPP :608-832
  SecondOrderDeformableAlignment :835-898           TimestepEmbedSequential dispatch :106-133
  GroupNorm32 / timestep_embedding nn_new.py:17-19,103-121;  LazyReshaper2D/3D nn.py:350-367

plus, from the un-vendored mmedit 0.12 / mmcv 1.4.8 (restated from their published behaviour —
PARITY UNPINNED for these): SPyNet, flow_warp, ResidualBlocksWithInputConv; flash-attn's
softmax(q k^T / sqrt(d)) v for the 1 x (F-1) temporal window (nn.py:370-394).

Tensors are (B, T, C, H, W) throughout, exactly like the reference's internal layout."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
import torchvision


def default_config(image_size=256):
    """MODEL_CONFIG["gaussian"] / ["jpeg"] of scripts/video_sample.py:116-155 (image_size overridable)."""
    return dict(image_size=image_size, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
                attention_resolutions=(16, 32, 64), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4),
                num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True, temporal_block=True)


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def plan(cfg):
    """Layer list per block, mirroring the constructor loops (unet_new.py:989-1222)."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    temporal = cfg["temporal_block"]

    def level_layers(cin, cout, ds):
        ls = [("res2d", cin, cout, None)]
        if temporal:
            ls.append(("res3d", cout, cout, None))
        if ds in cfg["attention_resolutions"]:
            ls.append(("attn", cout))
            if temporal:
                ls.append(("tattn", cout))
        if ds in cfg["rnn_resolutions"] and temporal:
            ls.append(("vsr", cout))
        return ls

    ch = int(mult[0] * mc)
    inputs, chans, ds = [[("conv_in", cfg["in_channels"], ch)]], [ch], 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            inputs.append(level_layers(ch, int(m * mc), ds))
            ch = int(m * mc)
            chans.append(ch)
        if level != len(mult) - 1:
            inputs.append([("res2d", ch, ch, "down")])
            chans.append(ch)
            ds *= 2
    middle = [("res2d", ch, ch, None)] + ([("res3d", ch, ch, None)] if temporal else [("id",)]) + \
             [("attn_bottle", ch)] + ([("tattn", ch)] if temporal else [("id",)]) + \
             [("res2d", ch, ch, None)] + ([("res3d", ch, ch, None)] if temporal else [("id",)])
    outputs = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            ls = level_layers(ch + ich, int(mc * m), ds)
            ch = int(mc * m)
            if level and i == nrb:
                ls.append(("res2d", ch, ch, "up"))
                ds //= 2
            outputs.append(ls)
    return inputs, middle, outputs


class BlurUNetOracle:
    def __init__(self, cfg, state_dict):
        self.cfg = cfg
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.inputs, self.middle, self.outputs = plan(cfg)
        self.heads_dim = cfg["num_head_channels"]

    # ------------------------------------------------------------------ primitives
    def p(self, name):
        return self.sd[name]

    def gn(self, x, pre):  # LazyReshaper3D(GroupNorm32(32, C)): statistics over (C/G, T, H, W)
        y = F.group_norm(x.permute(0, 2, 1, 3, 4), 32, self.p(pre + ".wrapped_module.weight"),
                         self.p(pre + ".wrapped_module.bias"), eps=1e-5)
        return y.permute(0, 2, 1, 3, 4)

    def conv2d(self, x, pre):
        B, T = x.shape[:2]
        w = self.p(pre + ".weight")
        y = F.conv2d(x.flatten(0, 1), w, self.sd.get(pre + ".bias"), padding=w.shape[-1] // 2)
        return y.unflatten(0, (B, T))

    def conv3d(self, x, pre):
        w = self.p(pre + ".weight")
        y = F.conv3d(x.permute(0, 2, 1, 3, 4), w, self.sd.get(pre + ".bias"), padding=1)
        return y.permute(0, 2, 1, 3, 4)

    # ------------------------------------------------------------------ blocks
    def resblock(self, x, emb, pre, dims, updown):
        conv = self.conv2d if dims == 2 else self.conv3d
        h = F.silu(self.gn(x, pre + ".in_layers.0"))
        if updown == "up":
            up = lambda v: F.interpolate(v.flatten(0, 1), scale_factor=2, mode="nearest").unflatten(0, v.shape[:2])
            h, x = up(h), up(x)
        elif updown == "down":
            dn = lambda v: F.avg_pool2d(v.flatten(0, 1), 2, 2).unflatten(0, v.shape[:2])
            h, x = dn(h), dn(x)
        h = conv(h, pre + ".in_layers.2.wrapped_module")
        e = F.linear(F.silu(emb), self.p(pre + ".emb_layers.1.weight"), self.p(pre + ".emb_layers.1.bias"))
        e = e.reshape(x.shape[0], x.shape[1], -1)[..., None, None]
        scale, shift = torch.chunk(e, 2, dim=2)
        h = self.gn(h, pre + ".out_layers.0") * (1 + scale) + shift
        h = conv(F.silu(h), pre + ".out_layers.3.wrapped_module")
        skip = pre + ".skip_connection.wrapped_module"
        if skip + ".weight" in self.sd:
            x = conv(x, skip) if dims == 3 else self.conv2d(x, skip)
        return x + h

    def attention(self, x, emb, pre, bottle):
        B, T, C, H, W = x.shape
        qkv = F.conv1d(self.gn(x, pre + ".norm").flatten(0, 1).flatten(2), self.p(pre + ".qkv.weight"),
                       self.p(pre + ".qkv.bias"))
        nh = C // self.heads_dim
        bs, width, L = qkv.shape
        ch = width // (3 * nh)
        q, k, v = qkv.reshape(bs * nh, ch * 3, L).split(ch, dim=1)  # head-major (H, 3, d)
        s = 1 / math.sqrt(math.sqrt(ch))
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
        a = torch.einsum("bts,bcs->bct", w, v).reshape(bs, -1, L)
        if bottle:
            a = a + F.linear(F.silu(emb), self.p(pre + ".emb_layers.1.weight"),
                             self.p(pre + ".emb_layers.1.bias"))[..., None]
        a = F.conv1d(a, self.p(pre + ".proj_out.weight"), self.p(pre + ".proj_out.bias"))
        return x + a.reshape(B, T, C, H, W)

    def temporal_attention(self, h, pre, frames=5):
        B, T, C, H, W = h.shape
        nh = C // self.heads_dim
        x = self.gn(h, pre + ".norm")
        pad = frames // 2
        xp = torch.cat([x[:, :1].repeat(1, pad, 1, 1, 1), x, x[:, -1:].repeat(1, pad, 1, 1, 1)], dim=1)
        win = xp.unfold(1, frames, 1)  # b t c h w f
        win = win.permute(0, 1, 3, 4, 5, 2).reshape(-1, frames, C)  # (b t h w) f c
        pe = timestep_embedding(torch.arange(frames) - pad, C)
        rest = [i for i in range(frames) if i != pad]
        lin = lambda v, n: F.linear(v, self.p(f"{pre}.{n}.weight"), self.p(f"{pre}.{n}.bias"))
        q = lin(win[:, [pad]] + pe[None, [pad]], "q_linear").reshape(-1, 1, nh, C // nh)
        k = lin(win[:, rest] + pe[None, rest], "k_linear").reshape(-1, frames - 1, nh, C // nh)
        v = lin(win[:, rest], "v_linear").reshape(-1, frames - 1, nh, C // nh)
        s = torch.einsum("bqhd,bkhd->bhqk", q, k) / math.sqrt(C // nh)  # flash_attn default scale
        a = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v)
        a = a.reshape(B, T, H, W, C).permute(0, 1, 4, 2, 3)
        return self.conv2d(a, pre + ".proj.wrapped_module") + h

    # ------------------------------------------------------------------ BasicVSR++ (mmedit restated)
    @staticmethod
    def flow_warp(x, flow, padding_mode="zeros"):
        _, _, h, w = x.shape
        gy, gx = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
        grid = torch.stack((gx, gy), 2).type_as(x) + flow
        gxn = 2.0 * grid[..., 0] / max(w - 1, 1) - 1.0
        gyn = 2.0 * grid[..., 1] / max(h - 1, 1) - 1.0
        return F.grid_sample(x, torch.stack((gxn, gyn), dim=3), mode="bilinear", padding_mode=padding_mode,
                             align_corners=True)

    def conv(self, x, pre, **kw):  # plain (n,c,h,w) conv
        w = self.p(pre + ".weight")
        return F.conv2d(x, w, self.p(pre + ".bias"), padding=w.shape[-1] // 2, **kw)

    def res_blocks_with_input_conv(self, x, pre):
        x = F.leaky_relu(self.conv(x, pre + ".main.0"), 0.1)
        return x + self.conv(F.relu(self.conv(x, pre + ".main.2.0.conv1")), pre + ".main.2.0.conv2")

    def deform_align(self, x, extra, flow_1, flow_2, pre, mrm=10):
        o = torch.cat([extra, flow_1, flow_2], dim=1)
        for i in (0, 2, 4):
            o = F.leaky_relu(self.conv(o, f"{pre}.conv_offset.{i}"), 0.1)
        o = self.conv(o, pre + ".conv_offset.6")
        from .kernels import deform_align_core  # (kernels imports this module: late import)
        return deform_align_core(x[:, :x.shape[1] // 2], x[:, x.shape[1] // 2:], o, flow_1, flow_2, self.p(pre + ".weight"),
                                 self.p(pre + ".bias"), mrm)

    def basicvsrpp(self, hidden, flows_forward, flows_backward, weight, pre):
        n, t, c, h, w = hidden.shape
        spatial = [hidden[:, i] for i in range(t)]
        if weight is None:
            weight = torch.ones(n, t, 1, 1, 1)
        elif isinstance(weight, float):
            weight = torch.ones(n, t, 1, 1, 1) * weight
        elif weight.shape[-2] != h or weight.shape[-1] != w:
            weight = F.interpolate(weight.flatten(0, 1), size=(h, w), mode="nearest").unflatten(0, (n, t))
        feats = {"spatial": spatial}
        for name in ("backward_1", "forward_1"):
            feats[name] = []
            flows = flows_backward if "backward" in name else flows_forward
            frame_idx = list(range(0, t))
            flow_idx = list(range(-1, t - 1))
            if "backward" in name:
                frame_idx = frame_idx[::-1]
                flow_idx = frame_idx
            prop = hidden.new_zeros(n, c, h, w)
            for i, idx in enumerate(frame_idx):
                cur = spatial[idx]
                if i > 0:
                    f1 = flows[:, flow_idx[i]]
                    cond1 = self.flow_warp(prop, f1.permute(0, 2, 3, 1))
                    feat2, f2, cond2 = torch.zeros_like(prop), torch.zeros_like(f1), torch.zeros_like(cond1)
                    if i > 1:
                        feat2 = feats[name][-2]
                        f2 = flows[:, flow_idx[i - 1]]
                        f2 = f1 + self.flow_warp(f2, f1.permute(0, 2, 3, 1))
                        cond2 = self.flow_warp(feat2, f2.permute(0, 2, 3, 1))
                    prop = self.deform_align(torch.cat([prop, feat2], dim=1), torch.cat([cond1, cur, cond2], dim=1),
                                             f1, f2, f"{pre}.deform_align.{name}")
                other = [feats[k][idx] for k in feats if k not in ("spatial", name)]
                prop = prop + self.res_blocks_with_input_conv(torch.cat([cur] + other + [prop], dim=1),
                                                              f"{pre}.backbone.{name}")
                prop = prop * weight[:, idx]  # in-place `*=` after append: the stored feature is scaled too
                feats[name].append(prop)
            if "backward" in name:
                feats[name] = feats[name][::-1]
        rec = [self.res_blocks_with_input_conv(torch.cat([spatial[i], feats["backward_1"][i], feats["forward_1"][i]],
                                                         dim=1), pre + ".reconstruction") for i in range(t)]
        rec = torch.stack(rec, dim=1)
        return self.conv(rec.flatten(0, 1), pre + ".conv_last").unflatten(0, (n, t)) + hidden

    def spynet(self, ref, supp):
        def basic(x, level):
            for k in range(5):
                x = self.conv(x, f"spynet.basic_module.{level}.basic_module.{k}.conv")
                if k < 4:
                    x = F.relu(x)
            return x

        def compute(ref, supp):
            n, _, h, w = ref.shape
            mean, std = self.p("spynet.mean"), self.p("spynet.std")
            ref, supp = [(ref - mean) / std], [(supp - mean) / std]
            for _ in range(5):
                ref.append(F.avg_pool2d(ref[-1], 2, 2, count_include_pad=False))
                supp.append(F.avg_pool2d(supp[-1], 2, 2, count_include_pad=False))
            ref, supp = ref[::-1], supp[::-1]
            flow = ref[0].new_zeros(n, 2, h // 32, w // 32)
            for lv in range(6):
                up = flow if lv == 0 else F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
                flow = up + basic(torch.cat([ref[lv], self.flow_warp(supp[lv], up.permute(0, 2, 3, 1), "border"), up], 1), lv)
            return flow

        h, w = ref.shape[2:4]
        wu = w if w % 32 == 0 else 32 * (w // 32 + 1)
        hu = h if h % 32 == 0 else 32 * (h // 32 + 1)
        ref = F.interpolate(ref, size=(hu, wu), mode="bilinear", align_corners=False)
        supp = F.interpolate(supp, size=(hu, wu), mode="bilinear", align_corners=False)
        flow = F.interpolate(compute(ref, supp), size=(h, w), mode="bilinear", align_corners=False)
        flow[:, 0] *= float(w) / float(wu)
        flow[:, 1] *= float(h) / float(hu)
        return flow

    def compute_flow(self, lqs):  # unet_new.py:1283-1309
        lqs = ((lqs + 1) / 2).clamp(0, 1)
        n, t, c, h, w = lqs.shape
        a, b = lqs[:, :-1].reshape(-1, c, h, w), lqs[:, 1:].reshape(-1, c, h, w)
        return self.spynet(b, a).view(n, t - 1, 2, h, w), self.spynet(a, b).view(n, t - 1, 2, h, w)

    # ------------------------------------------------------------------ forward
    def run_block(self, layers, pre, h, emb, flows, weights, cross):
        for j, layer in enumerate(layers):
            kind, name = layer[0], f"{pre}.{j}"
            if kind == "conv_in":
                h = self.conv2d(h, name + ".wrapped_module")
            elif kind == "res2d":
                h = self.resblock(h, emb, name, 2, layer[3])
            elif kind == "res3d" and cross:
                h = self.resblock(h, emb, name + ".wrapped_module", 3, None)
            elif kind in ("attn", "attn_bottle"):
                h = self.attention(h, emb, name, kind == "attn_bottle")
            elif kind == "tattn" and cross:
                h = self.temporal_attention(h, name + ".wrapped_module")
            elif kind == "vsr" and cross:
                ff, fb = flows[h.shape[-1]]
                h = self.basicvsrpp(h, ff, fb, weights, name + ".wrapped_module")
        return h

    def forward(self, x, timesteps, low_res_input, num_frames, rnn_input=None, enable_cross_frames=True,
                vsrpp_weights=None):
        cfg = self.cfg
        x = x.reshape(-1, num_frames, *x.shape[1:])
        x = torch.cat([x, low_res_input], dim=2)
        rnn_input = low_res_input if rnn_input is None else rnn_input
        flows = {}
        if enable_cross_frames and cfg["temporal_block"]:
            for s in cfg["rnn_resolutions"]:
                res = cfg["image_size"] // s
                fi = rnn_input
                if rnn_input.shape[-1] != res:
                    fi = F.interpolate(rnn_input.flatten(0, 1), (res, res), mode="bicubic").unflatten(
                        0, rnn_input.shape[:2])
                flows[res] = self.compute_flow(fi)
        e = timestep_embedding(timesteps, cfg["model_channels"])
        emb = F.linear(F.silu(F.linear(e, self.p("time_embed.0.weight"), self.p("time_embed.0.bias"))),
                       self.p("time_embed.2.weight"), self.p("time_embed.2.bias"))
        hs, h = [], x
        for i, layers in enumerate(self.inputs):
            h = self.run_block(layers, f"input_blocks.{i}", h, emb, flows, vsrpp_weights, enable_cross_frames)
            hs.append(h)
        h = self.run_block(self.middle, "middle_block", h, emb, flows, vsrpp_weights, enable_cross_frames)
        for i, layers in enumerate(self.outputs):
            h = torch.cat([h, hs.pop()], dim=2)
            h = self.run_block(layers, f"output_blocks.{i}", h, emb, flows, vsrpp_weights, enable_cross_frames)
        h = F.silu(self.gn(h, "out.0"))
        return self.conv2d(h, "out.2.wrapped_module").flatten(0, 1)
