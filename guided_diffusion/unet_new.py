"""Blur / JPEG video denoiser (`UNetModel`) behind the reference API, executed by sm_100a kernels.

Module tree, constructor arguments, state-dict names and `forward` signature follow the reference's
guided_diffusion/unet_new.py (UNetModel :901-1362, ResBlock :198-329, AttentionBlock :332-377,
AttentionbottleBlock :380-429, TemporalAttention :432-517, BasicVSRPP :608-832,
SecondOrderDeformableAlignment :835-898, TimestepEmbedSequential :106-133, TemporalWrapper :50-59),
so reference checkpoints load unchanged.  What runs is different:

  * activations are channels-last [B,T,H,W,C] 16-bit maps from the first kernel to the last; the
    reference's (b n) folds / b t c h w <-> b c t h w permutes do not exist;
  * every Conv2d / Conv3d / Conv1d(k=1) / Linear of the torso is one launch of the tcgen05 + TMA
    implicit-GEMM kernel (flair_conv_igemm) with bias / activation / residual fused in the epilogue;
  * GroupNorm32 + SiLU + scale-shift + 2x resampling is a statistics pass plus one apply pass;
  * the first conv consumes an im2col-packed 64-channel map built directly from the two fp32 NCHW
    inputs, the last conv writes fp32 NCHW directly;
  * all ResBlock `emb_layers` Linears are evaluated by ONE fp32 launch per forward;
  * spatial attention is a flash-style kernel, temporal attention a gather kernel over per-frame
    q/k/v projections with the positional terms folded into constants;
  * BasicVSR++: warps, the deformable-conv sampling and every conv are native kernels; SPyNet flows
    depend only on `rnn_input`, which is constant over the 100 sampling steps, so they are computed
    once per distinct input and cached (the reference recomputes them every step, :1334-1348).

Weights are re-packed ([tap][Cout][Cin] 16-bit, K-major) lazily on first use after any parameter
change.  `compute_dtype` (fp16 by default like the reference's torso, bf16 selectable) is the operand type of the GEMMs;
accumulation, GroupNorm statistics, softmax and the conditioning path are fp32.
"""
from __future__ import annotations

import math
from abc import abstractmethod

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from flair_b200 import _lib as L
from flair_b200 import ops

from .nn import FalshAttn, LazyReshaper2D, LazyReshaper3D
from .nn_new import avg_pool_nd, conv_nd, linear, normalization, timestep_embedding, zero_module


# --------------------------------------------------------------------------------------------------
# per-forward context
# --------------------------------------------------------------------------------------------------
class _Ctx:
    """dtype: GEMM operand type (bf16/fp16); sdtype: storage type of the residual stream."""
    __slots__ = ("emb_all", "flows", "weights", "cross", "dtype", "T", "sdtype", "emb_plain", "gates", "gn_next", "gn_tail")

    def __init__(self, emb_all, flows, weights, cross, dtype, T, sdtype=None):
        self.emb_all, self.flows, self.weights, self.cross, self.dtype, self.T = emb_all, flows, weights, cross, dtype, T
        self.sdtype = sdtype or dtype
        self.emb_plain = self.gates = None  # sr3: un-activated noise embeddings / sigmoid gates
        # gn_next: the tensor a module is about to produce is consumed by a GroupNorm next (its producing conv then
        # leaves the statistics, ops.conv(gn_groups=...)); gn_tail: the same for the last module of a block
        self.gn_next, self.gn_tail = False, True

    def operand(self, x):
        """A 16-bit GEMM-operand copy of a residual-stream map (no-op when the stream is already 16-bit)."""
        return x if x.dtype == self.dtype else ops.gn_apply(x, None, out_dtype=self.dtype)


class _Packed:
    """Lazily packed kernel-side weights of one module, invalidated when parameters change."""

    def stamp(self, dtype):
        return (dtype,) + tuple((p.data_ptr(), p._version, p.dtype) for p in self.parameters(recurse=False)) + \
            tuple(c.stamp(dtype) if isinstance(c, _Packed) else
                  tuple((p.data_ptr(), p._version, p.dtype) for p in c.parameters()) for c in self.children())

    def packed(self, dtype):
        st = self.stamp(dtype)
        if getattr(self, "_pk_stamp", None) != st:
            self._pk = self._pack(dtype)
            self._pk_stamp = st
        return self._pk


def _w(conv, dtype):
    return ops.pack_conv_weight(conv.weight.detach(), dtype)


def _f(p):
    return None if p is None else p.detach().float().contiguous()


def convert_module_to_f16(l):
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d, SecondOrderDeformableAlignment)):
        l.weight.data = l.weight.data.half()
        if l.bias is not None:
            l.bias.data = l.bias.data.half()


def convert_module_to_f32(l):
    if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d, SecondOrderDeformableAlignment)):
        l.weight.data = l.weight.data.float()
        if l.bias is not None:
            l.bias.data = l.bias.data.float()


class TemporalWrapper(nn.Module):
    """Cross-frame module switch (reference :50-59): identity when enable_cross_frames is False."""

    def __init__(self, module):
        super().__init__()
        self.wrapped_module = module


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        """Apply the module given timestep conditioning."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """Sequential with type-based argument routing (reference :106-133)."""

    def forward(self, x, ctx):
        run = []
        for layer in self:
            if isinstance(layer, TemporalWrapper):
                if ctx.cross:
                    run.append(layer.wrapped_module)
            elif not isinstance(layer, nn.Identity):
                run.append(layer)
        for i, layer in enumerate(run):
            # every module but BasicVSR++ (and the first conv) starts with a GroupNorm of its input
            ctx.gn_next = ctx.gn_tail if i + 1 == len(run) else not isinstance(run[i + 1], BasicVSRPP)
            x = layer(x, ctx)
        ctx.gn_next = False
        return x


class Upsample(nn.Module):
    """Nearest x2 (conv-less variant only is used by FLAIR: resblock_updown, reference :136-165)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=1)


class Downsample(nn.Module):
    """2x2 average pooling (conv-less variant only is used by FLAIR, reference :168-195)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        stride = 2 if dims != 3 else (1, 2, 2)
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=stride, padding=1)
        else:
            self.op = avg_pool_nd(dims, kernel_size=stride, stride=stride)


class ResBlock(TimestepBlock, _Packed):
    """GN-SiLU-conv, FiLM'd GN-SiLU-conv, plus skip (reference :198-329); dims=3 -> 3x3x3 convs."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False, kernel_size=3,
                 padding=1):
        super().__init__()
        k3 = (kernel_size,) * dims if isinstance(kernel_size, int) else tuple(kernel_size)
        self._ks = (1,) * (3 - len(k3)) + k3  # (kt, kh, kw) as flair_conv_igemm wants it
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.dims = dims
        wrap = LazyReshaper2D if dims == 2 else LazyReshaper3D
        self.in_layers = nn.Sequential(
            LazyReshaper3D(normalization(channels)), nn.SiLU(),
            wrap(conv_nd(dims, channels, self.out_channels, kernel_size, padding=padding)))
        self.updown = up or down
        self._resample = 1 if up else (2 if down else 0)
        if up:
            self.h_upd = LazyReshaper2D(Upsample(channels, False, dims))
            self.x_upd = LazyReshaper2D(Upsample(channels, False, dims))
        elif down:
            self.h_upd = LazyReshaper2D(Downsample(channels, False, dims))
            self.x_upd = LazyReshaper2D(Downsample(channels, False, dims))
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(
            nn.SiLU(), linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels))
        self.out_layers = nn.Sequential(
            LazyReshaper3D(normalization(self.out_channels)), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(wrap(conv_nd(dims, self.out_channels, self.out_channels, kernel_size, padding=padding))))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = wrap(conv_nd(dims, channels, self.out_channels, 3, padding=1))
        else:
            self.skip_connection = wrap(conv_nd(dims, channels, self.out_channels, 1))
        self._emb_slot = None  # (offset, width) into the model-wide emb_layers output

    def _pack(self, dtype):
        c1, c2 = self.in_layers[2].wrapped_module, self.out_layers[3].wrapped_module
        n1, n2 = self.in_layers[0].wrapped_module, self.out_layers[0].wrapped_module
        pk = dict(w1=_w(c1, dtype), b1=_f(c1.bias), w2=_w(c2, dtype), b2=_f(c2.bias),
                  g1=_f(n1.weight), be1=_f(n1.bias), g2=_f(n2.weight), be2=_f(n2.bias), ws=None)
        if not isinstance(self.skip_connection, nn.Identity):
            sc = self.skip_connection.wrapped_module
            pk.update(ws=_w(sc, dtype), bs=_f(sc.bias), ks=tuple(sc.kernel_size))
        return pk

    def forward(self, x, ctx, gate=None):
        """gate: optional per-(frame, channel) sigmoid gate of sr3.TemporalWrapper2, fused into the last conv:
        (1-s) x + s (x + branch) == x + s * branch."""
        pk = self.packed(ctx.dtype)
        cout = self.out_channels
        ks = self._ks
        off, width = self._emb_slot
        emb = ctx.emb_all[:, off:off + width]
        a1 = ops.gn_apply(x, ops.gn_stats(x), pk["g1"], pk["be1"], silu=True, resample=self._resample,
                          out_dtype=ctx.dtype)
        if self.use_scale_shift_norm:
            h1 = ops.conv(a1, pk["w1"], cout, ks, bias=pk["b1"], gn_groups=32)   # statistics of out_layers' norm
            a2 = ops.gn_apply(h1, ops.gn_stats(h1), pk["g2"], pk["be2"], scale=emb[:, :cout], shift=emb[:, cout:],
                              silu=True)
        else:  # h + emb_out, then norm (reference :326-328)
            h1 = ops.conv(a1, pk["w1"], cout, ks, bias=pk["b1"], rowbias=emb)
            a2 = ops.gn_apply(h1, ops.gn_stats(h1), pk["g2"], pk["be2"], silu=True)
        if pk["ws"] is not None:  # skip conv consumes a 16-bit operand copy of the (resampled) stream
            xs = ops.gn_apply(x, None, resample=self._resample, out_dtype=ctx.dtype) \
                if (self.updown or x.dtype != ctx.dtype) else x
            kk = pk["ks"]
            xs = ops.conv(xs, pk["ws"], cout, (1, 1, 1) if kk[-1] == 1 else ks, bias=pk["bs"], out_dtype=ctx.sdtype)
        else:
            xs = ops.gn_apply(x, None, resample=self._resample) if self.updown else x
        return ops.conv(a2, pk["w2"], cout, ks, bias=pk["b2"], residual=xs, rowscale=gate, out_dtype=ctx.sdtype,
                        gn_groups=32 if ctx.gn_next else None)


class _AttnBase(_Packed):
    def _init_attn(self, channels, num_heads, num_head_channels, use_checkpoint, use_new_attention_order):
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            if channels % num_head_channels != 0:
                raise AssertionError(
                    f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}")
            self.num_heads = channels // num_head_channels
        if channels // self.num_heads != 64:
            raise NotImplementedError("flair_attn_spatial is built for 64-channel heads (FLAIR: num_head_channels=64)")
        if use_new_attention_order:
            raise NotImplementedError("FLAIR uses the legacy head-major qkv order (unet_new.py:365)")
        self.use_checkpoint = use_checkpoint
        self.norm = LazyReshaper3D(normalization(channels))
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def _pack(self, dtype):
        n = self.norm.wrapped_module
        return dict(g=_f(n.weight), b=_f(n.bias), wqkv=_w(self.qkv, dtype), bqkv=_f(self.qkv.bias),
                    wp=_w(self.proj_out, dtype), bp=_f(self.proj_out.bias))

    _gn_next_hint = False

    def _attend(self, x, rowbias, dtype):
        pk = self.packed(dtype)
        c = self.channels
        a = ops.gn_apply(x, ops.gn_stats(x), pk["g"], pk["b"], out_dtype=dtype)
        qkv = ops.conv(a, pk["wqkv"], 3 * c, (1, 1, 1), bias=pk["bqkv"])
        att = ops.attn_spatial(qkv, self.num_heads, rowbias=rowbias)
        return ops.conv(att, pk["wp"], c, (1, 1, 1), bias=pk["bp"], residual=x, out_dtype=x.dtype,
                        gn_groups=32 if self._gn_next_hint else None)


class AttentionBlock(nn.Module, _AttnBase):
    """Per-frame spatial self-attention over H*W tokens (reference :332-377)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False):
        super().__init__()
        self._init_attn(channels, num_heads, num_head_channels, use_checkpoint, use_new_attention_order)

    def forward(self, x, ctx):
        self._gn_next_hint = ctx.gn_next
        return self._attend(x, None, ctx.dtype)


class AttentionbottleBlock(TimestepBlock, _AttnBase):
    """Spatial attention whose value path also receives Linear(SiLU(emb)) (reference :380-429)."""

    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(512, 512))
        self._init_attn(channels, num_heads, num_head_channels, use_checkpoint, use_new_attention_order)
        self._emb_slot = None

    def forward(self, x, ctx):
        off, width = self._emb_slot
        self._gn_next_hint = ctx.gn_next
        return self._attend(x, ctx.emb_all[:, off:off + width], ctx.dtype)


class QKVAttentionLegacy(nn.Module):
    """Head-major (H, 3, d) qkv attention (reference :540-570) — executed by flair_attn_spatial."""

    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads


class TemporalAttention(nn.Module, _Packed):
    """Each pixel's centre frame attends to its F-1 neighbours (replicate-padded), reference :432-517."""

    def __init__(self, channels, num_frames, num_heads=1, num_head_channels=-1, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        if num_frames % 2 != 1:
            raise AssertionError("num_frames must be odd")
        if channels % self.num_heads != 0 or channels // self.num_heads != 64:
            raise NotImplementedError("flair_attn_temporal is built for 64-channel heads (FLAIR: num_head_channels=64); "
                                      f"got channels={channels}, heads={self.num_heads}")
        if num_frames not in (5, 7):
            raise NotImplementedError("flair_attn_temporal supports temporal windows of 5 (blur UNet) or 7 (SR3) frames")
        self.num_frames, self.num_head_channels, self.use_checkpoint = num_frames, num_head_channels, use_checkpoint
        self.qk_scale = (channels // num_heads) ** -0.5
        self.q_linear = linear(channels, channels)
        self.k_linear = linear(channels, channels)
        self.v_linear = linear(channels, channels)
        self.attn = FalshAttn()
        self.proj = zero_module(LazyReshaper2D(conv_nd(2, channels, channels, 1)))
        self.norm = LazyReshaper3D(normalization(channels))

    def _pack(self, dtype):
        dev = self.q_linear.weight.device
        fr, half, c = self.num_frames, self.num_frames // 2, self.channels
        freqs = th.exp(-math.log(10000) * th.arange(0, c // 2, dtype=th.float32) / (c // 2))
        args = (th.arange(fr, dtype=th.float32) - half)[:, None] * freqs[None]
        pe = th.cat([th.cos(args), th.sin(args)], dim=-1).to(dev)  # timestep_embedding(arange(F)-F//2, C)
        rest = [i for i in range(fr) if i != half]
        wq, wk, wv = (m.weight.detach().float() for m in (self.q_linear, self.k_linear, self.v_linear))
        n = self.norm.wrapped_module
        pj = self.proj.wrapped_module
        return dict(
            g=_f(n.weight), b=_f(n.bias),
            wqkv=ops.pack_conv_weight(th.cat([wq, wk, wv], 0), dtype),
            cq=(pe[half] @ wq.t() + self.q_linear.bias.detach().float()).contiguous(),
            ck=(pe[rest] @ wk.t() + self.k_linear.bias.detach().float()).contiguous(),
            bv=_f(self.v_linear.bias), wp=_w(pj, dtype), bp=_f(pj.bias))

    def forward(self, h, ctx, gate=None):
        pk = self.packed(ctx.dtype)
        c = self.channels
        x = ops.gn_apply(h, ops.gn_stats(h), pk["g"], pk["b"], out_dtype=ctx.dtype)
        qkv = ops.conv(x, pk["wqkv"], 3 * c, (1, 1, 1))
        att = ops.attn_temporal(qkv, pk["cq"], pk["ck"], pk["bv"], self.num_frames)
        return ops.conv(att, pk["wp"], c, (1, 1, 1), bias=pk["bp"], residual=h, rowscale=gate, out_dtype=h.dtype,
                        gn_groups=32 if ctx.gn_next else None)


# --------------------------------------------------------------------------------------------------
# BasicVSR++ (parameter containers named like mmedit's / mmcv's so reference checkpoints load)
# --------------------------------------------------------------------------------------------------
class _ResidualBlockNoBN(nn.Module):
    def __init__(self, mid):
        super().__init__()
        self.conv1 = nn.Conv2d(mid, mid, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(mid, mid, 3, 1, 1, bias=True)
        self.relu = nn.ReLU(inplace=True)
        for m in (self.conv1, self.conv2):  # mmedit default_init_weights(scale=0.1)
            nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in", nonlinearity="relu")
            m.weight.data *= 0.1
            nn.init.constant_(m.bias, 0)


class ResidualBlocksWithInputConv(nn.Module, _Packed):
    """conv3x3 + LeakyReLU(0.1) + n x (conv-ReLU-conv + identity)   (mmedit 0.12, restated)."""

    def __init__(self, in_channels, out_channels=64, num_blocks=30):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.main = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, 1, 1, bias=True),
                                  nn.LeakyReLU(negative_slope=0.1, inplace=True),
                                  nn.Sequential(*[_ResidualBlockNoBN(out_channels) for _ in range(num_blocks)]))

    def _pack(self, dtype):
        blocks = [(_w(b.conv1, dtype), _f(b.conv1.bias), _w(b.conv2, dtype), _f(b.conv2.bias)) for b in self.main[2]]
        w0 = self.main[0].weight.detach()
        c = self.out_channels
        pk = dict(w0=_w(self.main[0], dtype), b0=_f(self.main[0].bias), blocks=blocks)
        if self.in_channels > c and self.in_channels % c == 0:
            # input conv split by linearity: [static slices (known for all frames) | last slice (recurrent)]
            pk["w0_static"] = ops.pack_conv_weight(w0[:, :-c], dtype)
            pk["w0_last"] = ops.pack_conv_weight(w0[:, -c:], dtype)
        return pk

    def _tail(self, x, pk, c, extra_residual, out, out2):
        for i, (w1, b1, w2, b2) in enumerate(pk["blocks"]):
            last = i == len(pk["blocks"]) - 1
            t = ops.conv(x, w1, c, (1, 3, 3), bias=b1, act=L.ACT_RELU)
            x = ops.conv(t, w2, c, (1, 3, 3), bias=b2, residual=x, residual2=extra_residual if last else None,
                         out=out if last else None, out2=out2 if last else None,
                         out2_neighbor=ops.pair_neighbor(c, x.shape[3]))
        return x

    def run(self, feat, dtype, extra_residual=None, out=None, out2=None):
        """feat: [1,N,H,W,Cin] map.  Returns main(feat) (+ extra_residual fused into the last conv).
        out2: optional pair-plane copy of the result ([8][pixels][2][C/8], flair_deform_conv's source layout)."""
        pk = self.packed(dtype)
        c = self.out_channels
        x = ops.conv(feat, pk["w0"], c, (1, 3, 3), bias=pk["b0"], act=L.ACT_LRELU01)
        return self._tail(x, pk, c, extra_residual, out, out2)

    def static_part(self, feat_static, dtype):
        """Input conv over the slices of the concat that are known for ALL frames before the recurrence starts
        (reference :729-735: [feat_current | other branches]); batched over T, bias included, no activation."""
        pk = self.packed(dtype)
        return ops.conv(feat_static, pk["w0_static"], self.out_channels, (1, 3, 3), bias=pk["b0"])

    def run_split(self, last_slice, static_part, dtype, extra_residual=None, out=None, out2=None):
        """main(cat(static slices, last_slice)) with the static slices' share of the input conv precomputed
        (`static_part`): the recurrent launch contracts K = C instead of 2C / 3C."""
        pk = self.packed(dtype)
        c = self.out_channels
        x = ops.conv(last_slice, pk["w0_last"], c, (1, 3, 3), preadd=static_part, act=L.ACT_LRELU01)
        return self._tail(x, pk, c, extra_residual, out, out2)


class ModulatedDeformConv2d(nn.Module):
    """Parameter container of mmcv's ModulatedDeformConv2d (weight (out, in, k, k), bias)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = (kernel_size,) * 2, (stride,) * 2, (padding,) * 2, (dilation,) * 2
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(th.empty(out_channels, in_channels // groups, kernel_size, kernel_size))
        self.bias = nn.Parameter(th.zeros(out_channels)) if bias else None
        bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
        self.weight.data.uniform_(-bound, bound)


class SecondOrderDeformableAlignment(ModulatedDeformConv2d, _Packed):
    """Flow-guided second-order deformable alignment (reference :835-898)."""

    def __init__(self, *args, **kwargs):
        self.max_residue_magnitude = kwargs.pop("max_residue_magnitude", 10)
        super().__init__(*args, **kwargs)
        oc = self.out_channels
        self.conv_offset = nn.Sequential(
            nn.Conv2d(3 * oc + 4, oc, 3, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            nn.Conv2d(oc, oc, 3, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            nn.Conv2d(oc, oc, 3, 1, 1), nn.LeakyReLU(negative_slope=0.1, inplace=True),
            nn.Conv2d(oc, 27 * self.deform_groups, 3, 1, 1))
        nn.init.constant_(self.conv_offset[-1].weight, 0)
        nn.init.constant_(self.conv_offset[-1].bias, 0)

    def _pack(self, dtype):
        co = self.conv_offset
        oc = self.out_channels
        # deformable weight (C, 2C, 3, 3) -> 1x1 GEMM weight over K = tap*2C + ci
        wd = self.weight.detach().float().permute(0, 2, 3, 1).reshape(self.out_channels, -1)
        off = [(_w(co[i], dtype), _f(co[i].bias)) for i in (0, 2, 4)]
        w6, b6 = co[6].weight.detach(), co[6].bias.detach()
        if self.fused:  # tap-major offset channels: what flair_deform_conv stages per tap with one TMA box
            perm = ops.deform_offset_perm(self.deform_groups).to(w6.device)
            w6, b6 = w6[perm], b6[perm]
        off.append((ops.pack_conv_weight(w6, dtype), b6.float().contiguous()))
        # first offset conv split by linearity over its input slices [warp(prop) | cur | warp(prev2) | f1 f2]
        # (reference :874-879): the `cur` and flow slices are known for all frames up front
        w0 = co[0].weight.detach()
        split = dict(rec2=ops.pack_conv_weight(th.cat([w0[:, :oc], w0[:, 2 * oc:3 * oc]], 1), dtype),
                     rec1=ops.pack_conv_weight(w0[:, :oc], dtype),
                     static=ops.pack_conv_weight(th.cat([w0[:, oc:2 * oc], w0[:, 3 * oc:]], 1), dtype))
        # fused kernel: channel-block-major K (ops.deform_weight_kperm); generic im2col + GEMM path: tap-major K
        wd_pk = ops.pack_deform_weight(wd, dtype) if self.fused else ops.pack_conv_weight(wd, dtype)
        return dict(off=off, wd=wd_pk, bd=_f(self.bias), split=split)

    @property
    def fused(self):
        """One-launch gather + tcgen05 GEMM (flair_deform_conv): the FLAIR shapes (16 deform groups, C = 64 / 128)."""
        return self.deform_groups == 16 and self.out_channels in (64, 128)

    def static_part(self, cur_flows, dtype):
        """conv_offset[0] over the [cur | f1 f2] slices of its input for ALL frames (bias included, no activation)."""
        pk = self.packed(dtype)
        return ops.conv(cur_flows, pk["split"]["static"], self.out_channels, (1, 3, 3), bias=pk["off"][0][1])

    def run(self, xa, xb, cond, flow_1, flow_2, dtype, out, xa_g=None, xb_g=None, static_part=None):
        """xa/xb: [N,H,W,C] (feat_prop, feat_n2), xa_g/xb_g their pair-plane copies [8][N*H*W][2][C/8].
        cond: with `static_part` (the precomputed [cur | flows] share of conv_offset[0], [1,N,H,W,C]) the warped
        features only: [1,N,H,W,2C] = [warp(prop) | warp(prev2)] or [1,N,H,W,C] = [warp(prop)] (first-order step);
        without it the full offset-net input [1,N,H,W,3C+4]."""
        pk = self.packed(dtype)
        oc = self.out_channels
        if static_part is not None:
            w0 = pk["split"]["rec2" if cond.shape[-1] == 2 * oc else "rec1"]
            o = ops.conv(cond, w0, oc, (1, 3, 3), preadd=static_part, act=L.ACT_LRELU01)
        else:
            o = ops.conv(cond, pk["off"][0][0], oc, (1, 3, 3), bias=pk["off"][0][1], act=L.ACT_LRELU01)
        for i, (w, b) in enumerate(pk["off"][1:3]):
            o = ops.conv(o, w, oc, (1, 3, 3), bias=b, act=L.ACT_LRELU01)
        w, b = pk["off"][3]
        om = ops.conv(o, w, 27 * self.deform_groups, (1, 3, 3), bias=b, out_dtype=th.float16)
        if self.fused:
            return ops.deform_conv(xa_g, xb_g, om[0], flow_1, flow_2, pk["wd"], pk["bd"], self.max_residue_magnitude,
                                   out=out[0])
        cols = ops.deform_im2col(xa, xb, om[0], flow_1, flow_2, self.deform_groups, self.max_residue_magnitude)
        return ops.conv(cols, pk["wd"], oc, (1, 1, 1), bias=pk["bd"], out=out)


class BasicVSRPP(nn.Module, _Packed):
    """Bidirectional second-order propagation with deformable alignment (reference :608-832)."""

    def stamp(self, dtype):  # only conv_last is packed here; the sub-modules pack themselves
        return (dtype,) + tuple((p.data_ptr(), p._version, p.dtype) for p in self.conv_last.parameters())

    def _pack(self, dtype):
        return (_w(self.conv_last, dtype), _f(self.conv_last.bias))

    def __init__(self, mid_channels=64, max_residue_magnitude=10, use_checkpoint=False):
        super().__init__()
        self.mid_channels, self.use_checkpoint = mid_channels, use_checkpoint
        self.deform_align = nn.ModuleDict()
        self.backbone = nn.ModuleDict()
        for i, name in enumerate(["backward_1", "forward_1"]):
            # the reference only builds this "if th.cuda.is_available()" (:650); a B200 is mandatory here
            self.deform_align[name] = SecondOrderDeformableAlignment(
                2 * mid_channels, mid_channels, 3, padding=1, deform_groups=16,
                max_residue_magnitude=max_residue_magnitude)
            self.backbone[name] = ResidualBlocksWithInputConv((2 + i) * mid_channels, mid_channels, 1)
        self.reconstruction = ResidualBlocksWithInputConv(3 * mid_channels, mid_channels, 1)
        self.conv_last = zero_module(nn.Conv2d(mid_channels, mid_channels, 1, 1))

    def forward(self, hidden, ctx, gate=None):
        if hidden.shape[0] > 1:  # windows are independent: run them one by one (the demo uses B = 1)
            outs = []
            T_ = hidden.shape[1]
            for b in range(hidden.shape[0]):
                w = ctx.weights
                sub = _Ctx(ctx.emb_all, {hidden.shape[3]: tuple(f[b:b + 1] for f in ctx.flows[hidden.shape[3]])},
                           w[b:b + 1] if th.is_tensor(w) else w, ctx.cross, ctx.dtype, ctx.T, ctx.sdtype)
                outs.append(self.forward(hidden[b:b + 1], sub, None if gate is None else gate[b * T_:(b + 1) * T_]))
            return th.cat(outs, 0)
        gn_next = ctx.gn_next         # (the sub-module calls below do not go through a Sequential)
        stream = hidden               # residual-stream copy (may be fp32): only the final add reads it
        hidden = ctx.operand(hidden)  # 16-bit operand copy: features / warps / convs
        _, T, H, W, C = hidden.shape
        dt, dev = hidden.dtype, hidden.device
        f12 = {"backward_1": ctx.flows[W][3][0], "forward_1": ctx.flows[W][2][0]}  # [T,4,H,W]: f1 | composed f2
        weight = ctx.weights
        wmap = None
        if weight is not None and not isinstance(weight, float):
            if weight.shape[-2] != H or weight.shape[-1] != W:
                weight = F.interpolate(weight.flatten(0, 1), size=(H, W), mode="nearest").unflatten(0, (1, T))
            wmap = weight.reshape(T, H, W).float().contiguous()
        elif isinstance(weight, float) and weight != 1.0:
            wmap = th.full((T, H, W), weight, dtype=th.float32, device=dev)
        wmap8 = wmap8n = None
        if wmap is not None:  # per-plane weight maps for the pair-plane copies (entry p: pixels p and p+1)
            wmap8 = wmap[:, None].expand(T, 8, H, W).contiguous()
            # slot 1 of entry p is pixel p + neighbor (1: horizontal pairs at C = 128, W: vertical pairs at C = 64)
            wmap8n = th.roll(wmap.reshape(T, H * W), -ops.pair_neighbor(C, W), 1).reshape(T, 1, H, W).expand(T, 8, H, W).contiguous()
        frames = hidden[0]  # [T,H,W,C]
        # reconstruction input for ALL frames: [spatial | backward feature | forward feature]; the two
        # propagation passes write their outputs straight into its channel slices
        rec_cat = th.empty(1, T, H, W, 3 * C, dtype=dt, device=dev)
        ops.copy_channels_into(frames[None], rec_cat, 0)
        # [cur | f1 f2] per direction: the non-recurrent slices of the offset net's input
        cf = C + 8  # C + 4 flow channels, channel stride padded to a multiple of 8 (the pad is never read)
        for name in ("backward_1", "forward_1"):
            fwd = name == "forward_1"
            so = 2 * C if fwd else C
            order = list(range(T)) if fwd else list(range(T - 1, -1, -1))
            da, bb = self.deform_align[name], self.backbone[name]
            # ---- everything that does not depend on the recurrence, once for all T frames (batched launches):
            # a convolution is linear in the channel slices of its input, so the `cur` / flow slices of the first
            # offset conv (reference :874-879) and the `cur` (+ backward) slices of the backbone's first conv
            # (:729-735) are convolved here and enter the per-frame launches as a pre-activation addend.
            cur_fl = th.empty(1, T, H, W, cf, dtype=dt, device=dev)
            ops.copy_channels_into(frames[None], cur_fl, 0)
            ops.planes_to_cl(f12[name], cur_fl[0], C)
            p_off = da.static_part(cur_fl[..., :C + 4], ctx.dtype)                 # [1,T,H,W,C]
            p_bb = bb.static_part(rec_cat[..., :so], ctx.dtype)                    # [1,T,H,W,C]
            warps = th.empty(1, T, H, W, 2 * C, dtype=dt, device=dev)              # [warp(prop) | warp(prev2)]
            aligned_all = th.empty(1, T, H, W, C, dtype=dt, device=dev)
            prop = prev2 = prop_g = prev2_g = None
            fused = da.fused
            # pair-plane copies of the propagated features (written by the backbone's last conv): the layout the
            # fused deformable conv gathers from; slot 1 of each plane's last entry / last row is never written -> the
            # buffer is persistent and zeroed once (it is read with weight 0, it only has to stay finite)
            gm_all = self._persistent(("gm", name, T, H, W, C, dt, str(dev)),
                                      lambda: th.zeros(T, 8, H * W, 2, C // 8, dtype=dt, device=dev)) if fused else None
            for i, idx in enumerate(order):
                if i == 0:
                    aligned = self._zeros(aligned_all[0, :1])  # [1,H,W,C] of zeros (cached)
                else:
                    aligned = aligned_all[0, idx:idx + 1]
                    wb = warps[:, idx:idx + 1]
                    f1, f2 = f12[name][idx:idx + 1, 0:2], f12[name][idx:idx + 1, 2:4]
                    if i > 1:
                        ops.flow_warp2(prop, f1, wb[0, ..., :C], prev2, f2, wb[0, ..., C:])
                        cond, xb, xb_g = wb, prev2, prev2_g
                    else:  # first-order step: the second-order slices are zeros -> contract K = C only
                        ops.flow_warp(prop, f1, out=wb[0, ..., :C])
                        cond, xb = wb[..., :C], self._zeros(prop)
                        xb_g = self._zeros(gm_all[0]) if fused else None
                    da.run(prop, xb, cond, f1, f2, ctx.dtype, out=aligned[None], xa_g=prop_g, xb_g=xb_g,
                           static_part=p_off[:, idx:idx + 1])
                new = rec_cat[0, idx:idx + 1, :, :, so:so + C]
                new_g = gm_all[idx] if fused else None
                bb.run_split(aligned[None], p_bb[:, idx:idx + 1], ctx.dtype, extra_residual=aligned[None], out=new[None],
                             out2=new_g)
                if wmap is not None:
                    ops.scale_pixels_(new, wmap[idx:idx + 1])
                    if fused:  # slot 0 of entry p is pixel p, slot 1 is pixel p+1
                        ops.scale_pixels_(new_g[:, :, 0].unflatten(1, (H, W)), wmap8[idx])
                        ops.scale_pixels_(new_g[:, :, 1].unflatten(1, (H, W)), wmap8n[idx])
                prev2, prop, prev2_g, prop_g = prop, new, prop_g, new_g
        # reconstruction + zero-init 1x1 + residual: not recurrent -> one batched launch chain for all frames
        pk_last = self.packed(ctx.dtype)
        rec = self.reconstruction.run(rec_cat, ctx.dtype)
        return ops.conv(rec, pk_last[0], C, (1, 1, 1), bias=pk_last[1], residual=stream, rowscale=gate,
                        out_dtype=stream.dtype, gn_groups=32 if gn_next else None)

    def _persistent(self, key, make):
        """Module-owned buffers that outlive a forward (allocated during the eager warm-up that precedes every
        CUDA-graph capture, so captured graphs see stable addresses and no fill kernels)."""
        bufs = self.__dict__.setdefault("_persist_bufs", {})
        if key not in bufs:  # never evicted: captured graphs hold these addresses
            bufs[key] = make()
        return bufs[key]

    def _zeros(self, like):
        bufs = self.__dict__.setdefault("_zero_bufs", {})
        key = (tuple(like.shape), like.dtype, like.device)
        if key not in bufs:
            bufs[key] = th.zeros(like.shape, dtype=like.dtype, device=like.device)
        return bufs[key]


# --------------------------------------------------------------------------------------------------
# SPyNet (flow estimator; input is constant over the sampling loop -> outside the per-step hot path)
# --------------------------------------------------------------------------------------------------
class _ConvModule(nn.Module):
    def __init__(self, cin, cout, act):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 7, 1, 3)
        self.act = act

    def forward(self, x):
        x = self.conv(x)
        return F.relu(x) if self.act else x


class _SPyNetBasicModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.basic_module = nn.Sequential(_ConvModule(8, 32, True), _ConvModule(32, 64, True),
                                          _ConvModule(64, 32, True), _ConvModule(32, 16, True),
                                          _ConvModule(16, 2, False))

    def forward(self, x):
        return self.basic_module(x)


def _grid_warp(x, flow_nhw2, padding_mode):
    _, _, h, w = x.shape
    gy, gx = th.meshgrid(th.arange(0, h, device=x.device), th.arange(0, w, device=x.device), indexing="ij")
    g = th.stack((gx, gy), 2).type_as(x) + flow_nhw2
    g = th.stack((2.0 * g[..., 0] / max(w - 1, 1) - 1.0, 2.0 * g[..., 1] / max(h - 1, 1) - 1.0), dim=3)
    return F.grid_sample(x, g, mode="bilinear", padding_mode=padding_mode, align_corners=True)


class SPyNet(nn.Module):
    """mmedit 0.12 SPyNet restated (6-level pyramid of 5 x conv7x7).  Plain PyTorch on purpose: it is the
    auxiliary flow network BASELINE.json leaves as a reference PyTorch call outside the timed path."""

    def __init__(self, pretrained=None):
        super().__init__()
        self.basic_module = nn.ModuleList([_SPyNetBasicModule() for _ in range(6)])
        self.register_buffer("mean", th.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", th.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def compute_flow(self, ref, supp):
        n, _, h, w = ref.size()
        ref, supp = [(ref - self.mean) / self.std], [(supp - self.mean) / self.std]
        for _ in range(5):
            ref.append(F.avg_pool2d(ref[-1], 2, 2, count_include_pad=False))
            supp.append(F.avg_pool2d(supp[-1], 2, 2, count_include_pad=False))
        ref, supp = ref[::-1], supp[::-1]
        flow = ref[0].new_zeros(n, 2, h // 32, w // 32)
        for lv in range(len(ref)):
            up = flow if lv == 0 else F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
            flow = up + self.basic_module[lv](
                th.cat([ref[lv], _grid_warp(supp[lv], up.permute(0, 2, 3, 1), "border"), up], 1))
        return flow

    def forward(self, ref, supp):
        h, w = ref.shape[2:4]
        wu = w if w % 32 == 0 else 32 * (w // 32 + 1)
        hu = h if h % 32 == 0 else 32 * (h // 32 + 1)
        ref = F.interpolate(ref, size=(hu, wu), mode="bilinear", align_corners=False)
        supp = F.interpolate(supp, size=(hu, wu), mode="bilinear", align_corners=False)
        flow = F.interpolate(self.compute_flow(ref, supp), size=(h, w), mode="bilinear", align_corners=False)
        flow[:, 0, :, :] *= float(w) / float(wu)
        flow[:, 1, :, :] *= float(h) / float(hu)
        return flow


# --------------------------------------------------------------------------------------------------
class UNetModel(nn.Module):
    """Video-conditional UNet of the gaussian / jpeg tasks (reference :901-1362)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, rnn_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True,
                 dims=2, num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, temporal_block=False):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if dims != 2 or num_classes is not None or not resblock_updown or in_channels != 6:
            raise NotImplementedError("only the FLAIR configuration family is supported: dims=2, 6 input channels, "
                                      "resblock_updown=True, no class conditioning (scripts/video_sample.py:116-155)")
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.rnn_resolutions = attention_resolutions, rnn_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.num_classes, self.use_checkpoint = num_classes, use_checkpoint
        self.dtype = th.float16 if use_fp16 else th.float32
        # GEMM operand type and residual-stream storage type.  fp16 is the reference's torso dtype
        # (use_fp16=True, scripts/video_sample.py:129) and has the same tcgen05 rate as bf16; with fp32
        # accumulation a forward is ~1.5e-3 from the fp32 reference.  bf16 operands are supported
        # (`compute_dtype = th.bfloat16`); keep the stream in fp32/fp16 then, or the forward drifts past
        # 1e-2 (measured: DESIGN.md, "numerics").
        self.compute_dtype = th.float16
        self.stream_dtype = th.float16
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.need_flows_res = [image_size // s for s in rnn_resolutions]

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))
        self.spynet = SPyNet(pretrained=None)

        def res(cin, cout, **kw):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=2, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, **kw)

        def res3d(c):
            return TemporalWrapper(ResBlock(c, emb_dim, dropout, use_scale_shift_norm=use_scale_shift_norm, dims=3,
                                            use_checkpoint=use_checkpoint))

        def stage(cin, cout, ds, heads):
            layers = [res(cin, cout)]
            if temporal_block:
                layers.append(res3d(cout))
            if ds in attention_resolutions:
                layers.append(AttentionBlock(cout, use_checkpoint=use_checkpoint, num_heads=heads,
                                             num_head_channels=num_head_channels,
                                             use_new_attention_order=use_new_attention_order))
                if temporal_block:
                    layers.append(TemporalWrapper(TemporalAttention(cout, 5, heads, num_head_channels, use_checkpoint)))
            if ds in rnn_resolutions and temporal_block:
                layers.append(TemporalWrapper(BasicVSRPP(mid_channels=cout, use_checkpoint=use_checkpoint)))
            return layers

        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(LazyReshaper2D(conv_nd(dims, in_channels, ch, 3, padding=1)))])
        self._feature_size = ch
        skip_chans, ds = [ch], 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                cout = int(mult * model_channels)
                self.input_blocks.append(TimestepEmbedSequential(*stage(ch, cout, ds, num_heads)))
                ch = cout
                self._feature_size += ch
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(res(ch, ch, down=True)))
                skip_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        ident = nn.Identity
        self.middle_block = TimestepEmbedSequential(
            res(ch, ch), res3d(ch) if temporal_block else ident(),
            AttentionbottleBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads,
                                 num_head_channels=num_head_channels,
                                 use_new_attention_order=use_new_attention_order),
            TemporalWrapper(TemporalAttention(ch, 5, num_heads, num_head_channels, use_checkpoint))
            if temporal_block else ident(),
            res(ch, ch), res3d(ch) if temporal_block else ident())
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                cout = int(model_channels * mult)
                layers = stage(ch + skip_chans.pop(), cout, ds, num_heads_upsample)
                ch = cout
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, up=True))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(LazyReshaper3D(normalization(ch)), nn.SiLU(),
                                 zero_module(LazyReshaper2D(conv_nd(dims, input_ch, out_channels, 3, padding=1))))
        self._flow_cache = {}
        self._emb_pack = None
        self._graphs = {}
        self.use_cuda_graph = True

    # ------------------------------------------------------------------ dtype helpers (reference :1224-1254)
    def _torso_apply(self, fn, lin_dtype):
        self.input_blocks.apply(fn)
        self.middle_block.apply(fn)
        self.output_blocks.apply(fn)
        for m in self.modules():
            if isinstance(m, TemporalAttention):
                for lin in (m.q_linear, m.k_linear, m.v_linear):
                    for p in lin.parameters():
                        p.data = p.data.to(lin_dtype)

    def convert_to_fp16(self):
        """Parameter storage like the reference (so fp16 checkpoints load bit-exactly); the kernels always
        consume `compute_dtype` packed copies."""
        self._torso_apply(convert_module_to_f16, th.float16)

    def convert_to_fp32(self):
        self._torso_apply(convert_module_to_f32, th.float32)

    # ------------------------------------------------------------------ conditioning
    def _emb_weights(self):
        """time_embed + every emb_layers Linear concatenated, fp32, [K, N] layout."""
        mods = [m for m in self.modules() if isinstance(m, (ResBlock, AttentionbottleBlock))]
        stamp = tuple((m.emb_layers[1].weight.data_ptr(), m.emb_layers[1].weight._version) for m in mods) + \
            tuple((p.data_ptr(), p._version) for p in self.time_embed.parameters())
        if self._emb_pack is None or self._emb_pack[0] != stamp:
            off, ws, bs = 0, [], []
            for m in mods:
                lin = m.emb_layers[1]
                m._emb_slot = (off, lin.out_features)
                off += lin.out_features
                ws.append(lin.weight.detach().float())
                bs.append(lin.bias.detach().float())
            te0, te2 = self.time_embed[0], self.time_embed[2]
            pack = dict(w0=te0.weight.detach().float().t().contiguous(), b0=_f(te0.bias),
                        w2=te2.weight.detach().float().t().contiguous(), b2=_f(te2.bias),
                        wall=th.cat(ws, 0).t().contiguous(), ball=th.cat(bs, 0).contiguous())
            self._emb_pack = (stamp, pack)
        return self._emb_pack[1]

    @th.no_grad()
    def compute_flow(self, lqs):
        """(flows_forward, flows_backward), each (n, t-1, 2, h, w) fp32 (reference :1283-1309)."""
        lqs = ((lqs + 1) / 2).clamp(0, 1)
        n, t, c, h, w = lqs.size()
        a = lqs[:, :-1].reshape(-1, c, h, w)
        b = lqs[:, 1:].reshape(-1, c, h, w)
        return self.spynet(b, a).view(n, t - 1, 2, h, w), self.spynet(a, b).view(n, t - 1, 2, h, w)

    @staticmethod
    def _flow_pack(flows, forward):
        """[B,T,4,H,W] per propagation direction, indexed by frame: channels 0:2 = the flow that aligns the
        previous feature to this frame (f1, unet_new.py:704), 2:4 = the composed second-order flow
        f1 + warp(f2, f1) (:716-718).  Depends on the flows only -> computed once per window."""
        B, tm1, _, H, W = flows.shape
        T = tm1 + 1
        out = th.zeros(B, T, 4, H, W, dtype=th.float32, device=flows.device)
        order = list(range(T)) if forward else list(range(T - 1, -1, -1))
        flow_idx = list(range(-1, T - 1)) if forward else order
        for i, idx in enumerate(order):
            if i > 0:
                f1 = flows[:, flow_idx[i]].contiguous()
                out[:, idx, 0:2] = f1
                if i > 1:
                    out[:, idx, 2:4] = ops.flow_compose(flows[:, flow_idx[i - 1]].contiguous(), f1)
        return out

    def _flows(self, rnn_input, num_frames):
        key = (rnn_input.data_ptr(), tuple(rnn_input.shape), rnn_input._version,
               tuple((p.data_ptr(), p._version) for p in self.spynet.parameters()))
        if self._flow_cache.get("key") != key:
            flows = {}
            for res in self.need_flows_res:
                fi = rnn_input
                if rnn_input.shape[-1] != res:
                    fi = F.interpolate(rnn_input.flatten(0, 1).float(), (res, res), mode="bicubic").unflatten(
                        0, rnn_input.shape[:2])
                ff, fb = (f.float().contiguous() for f in self.compute_flow(fi.float()))
                flows[res] = (ff, fb, self._flow_pack(ff, True), self._flow_pack(fb, False))
            self._flow_cache = {"key": key, "flows": flows, "src": rnn_input}
        return self._flow_cache["flows"]

    # ------------------------------------------------------------------ forward
    @th.no_grad()
    def flows_for(self, x_shape, low_res_input=None, rnn_input=None, num_frames=None, enable_cross_frames=True):
        """The SPyNet flow pack a forward with these arguments uses ({} when no BasicVSR++ module runs).  Depends on
        `rnn_input` / `low_res_input` only, i.e. it is constant over the sampling steps of a window."""
        if enable_cross_frames and self.need_flows_res and any(isinstance(m, BasicVSRPP) for m in self.modules()):
            return self._flows(low_res_input if rnn_input is None else rnn_input, int(num_frames))
        return {}

    def forward(self, x, timesteps, low_res_input=None, num_frames=None, rnn_input=None, enable_cross_frames=True,
                vsrpp_weights=None, _static_flows=None, **kwargs):
        """x (B*T,3,H,W) fp32, timesteps (B*T,), low_res_input (B,T,3,H,W) -> (B*T,out_channels,H,W) fp32.

        With `use_cuda_graph` (default) the ~5000 kernel launches of one forward are captured once per
        input signature and replayed: inputs are copied into static buffers, flows into static flow
        buffers, and the result is returned as a fresh tensor."""
        if not x.is_cuda:
            raise RuntimeError("guided_diffusion.unet_new.UNetModel runs on a B200 only (no CPU fallback)")
        T = int(num_frames)
        cross = bool(enable_cross_frames)
        # _static_flows: flow pack precomputed by the caller (the graphed sampling step keeps it in static buffers)
        flows = _static_flows if _static_flows is not None else \
            self.flows_for(x.shape, low_res_input, rnn_input, T, cross)
        if not self.use_cuda_graph or th.cuda.is_current_stream_capturing():
            return self._forward_impl(x, timesteps, low_res_input, T, flows, cross, vsrpp_weights)
        return self._forward_graphed(x, timesteps, low_res_input, T, flows, cross, vsrpp_weights)

    def _param_stamp(self):
        return hash(tuple((p.data_ptr(), p._version) for p in self.parameters()))

    def _forward_graphed(self, x, timesteps, low, T, flows, cross, weights):
        wkey = ("t", tuple(weights.shape)) if th.is_tensor(weights) else ("s", weights)
        key = (tuple(x.shape), str(x.device), T, cross, wkey, self.compute_dtype, self.stream_dtype,
               tuple(sorted(flows)), timesteps.dtype)
        stamp = self._param_stamp()
        if self._graphs.get("stamp") != stamp:
            self._graphs = {"stamp": stamp}
        g = self._graphs.get(key)
        if g is None:
            st = dict(x=x.float().clone(), t=timesteps.clone(), low=low.float().clone(),
                      flows={r: tuple(f.clone() for f in ff) for r, ff in flows.items()},
                      w=weights.float().clone() if th.is_tensor(weights) else weights)
            run = lambda: self._forward_impl(st["x"], st["t"], st["low"], T, st["flows"], cross, st["w"])
            side = th.cuda.Stream()
            side.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(side):  # warm-up: packs weights, sets kernel attributes, sizes the pool
                run()
            th.cuda.current_stream().wait_stream(side)
            graph = th.cuda.CUDAGraph()
            n0 = L.LAUNCHES[0]
            with th.cuda.graph(graph):
                st["out"] = run()
            g = dict(graph=graph, st=st, flow_src=None, launches=L.LAUNCHES[0] - n0)
            if len(self._graphs) > 8:
                self._graphs = {"stamp": stamp}
            self._graphs[key] = g
        st = g["st"]
        st["x"].copy_(x)
        st["t"].copy_(timesteps)
        st["low"].copy_(low)
        if th.is_tensor(weights):
            st["w"].copy_(weights)
        if flows and g["flow_src"] is not self._flow_cache.get("flows"):
            for r, ff in flows.items():
                for dst, src in zip(st["flows"][r], ff):
                    dst.copy_(src)
            g["flow_src"] = self._flow_cache.get("flows")
        g["graph"].replay()
        L.LAUNCHES[0] += g["launches"]
        return st["out"].clone()

    def _forward_impl(self, x, timesteps, low_res_input, T, flows, cross, vsrpp_weights):
        N, _, H, W = x.shape
        B = N // T
        dt = self.compute_dtype
        ew = self._emb_weights()
        emb0 = ops.linear_f32(timestep_embedding(timesteps, self.model_channels), ew["w0"], ew["b0"], silu_out=True)
        emb = ops.linear_f32(emb0, ew["w2"], ew["b2"])
        emb_all = ops.linear_f32(emb, ew["wall"], ew["ball"], silu_in=True)
        ctx = _Ctx(emb_all, flows, vsrpp_weights, cross, dt, T, self.stream_dtype)

        conv_in = self.input_blocks[0][0].wrapped_module
        st = (dt, conv_in.weight.data_ptr(), conv_in.weight._version)
        if getattr(self, "_in_stamp", None) != st:  # 3x3x6 -> K = tap*6 + c (matches flair_pack_im2col6)
            w = conv_in.weight.detach().float().permute(0, 2, 3, 1).reshape(conv_in.out_channels, 54)
            self._in_pk = (ops.pack_conv_weight(w, dt), _f(conv_in.bias))
            self._in_stamp = st
        packed = ops.pack_im2col6(x, low_res_input.reshape(N, 3, H, W), dt).view(B, T, H, W, 64)
        h = ops.conv(packed, self._in_pk[0], conv_in.out_channels, (1, 1, 1), bias=self._in_pk[1],
                     out_dtype=self.stream_dtype, gn_groups=32)
        ctx.gn_tail = True
        hs = [h]
        for block in list(self.input_blocks)[1:]:
            h = block(h, ctx)
            hs.append(h)
        h = self.middle_block(h, ctx)
        for i, block in enumerate(self.output_blocks):
            ctx.gn_tail = i + 1 == len(self.output_blocks)   # other outputs go through a channel concat first
            h = block(ops.concat_channels(h, hs.pop()), ctx)
        n_out, c_out = self.out[0].wrapped_module, self.out[2].wrapped_module
        st = (dt, c_out.weight.data_ptr(), c_out.weight._version)
        if getattr(self, "_out_stamp", None) != st:
            self._out_pk = (_w(c_out, dt), _f(c_out.bias), _f(n_out.weight), _f(n_out.bias))
            self._out_stamp = st
        a = ops.gn_apply(h, ops.gn_stats(h), self._out_pk[2], self._out_pk[3], silu=True, out_dtype=dt)
        return ops.conv(a, self._out_pk[0], self.out_channels, (1, 3, 3), bias=self._out_pk[1], nchw_out=True)
