"""Thin torch-tensor wrappers over the C ABI (raw pointers + sizes + stream).

Feature maps are channels-last `[B, T, H, W, C]` tensors (bf16 unless noted).
Nothing here computes on the host or through PyTorch ops: every function is a
single launch of a hand-written sm_100a kernel on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_DT = {torch.bfloat16: L.BF16, torch.float32: L.F32, torch.float16: L.F16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _cstride(t: torch.Tensor) -> int:
    """Per-pixel channel stride of a channels-last [B,T,H,W,C] view."""
    assert t.dim() == 5 and t.stride(4) == 1, "expected channels-last [B,T,H,W,C]"
    cs = t.stride(3)
    B, T, H, W, _ = t.shape
    assert t.stride(2) == cs * W and t.stride(1) == cs * W * H and t.stride(0) == cs * W * H * T, (
        "feature map must be pixel-contiguous (only the channel stride may be padded)"
    )
    return cs


def pack_conv_weight(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """[Cout, Cin, (kt,) kh, kw] or [Cout, Cin] -> packed [taps][Cout_pad][Cin_pad] 16-bit.

    Setup-time only (runs once per layer at model load; plain torch ops on
    whatever device `w` lives on)."""
    if w.dim() == 2:
        w = w[:, :, None, None, None]
    elif w.dim() == 3:  # Conv1d k=1
        assert w.shape[2] == 1
        w = w[:, :, :, None, None]
    elif w.dim() == 4:
        w = w[:, :, None]
    cout, cin, kt, kh, kw = w.shape
    cin_pad = (cin + 63) // 64 * 64
    cout_pad = (cout + 15) // 16 * 16
    out = torch.zeros(kt * kh * kw, cout_pad, cin_pad, dtype=dtype, device=w.device)
    out[:, :cout, :cin] = w.permute(2, 3, 4, 0, 1).reshape(kt * kh * kw, cout, cin).to(dtype)
    return out.contiguous()


def conv(x, wpk, cout, ksize=(1, 3, 3), *, bias=None, rowbias=None, residual=None, out=None,
         out_dtype=None, act=L.ACT_NONE, stride=1, nchw_out=False, out_scale=1.0):
    """Implicit-GEMM convolution (see flair_conv_igemm in include/flair_b200.h).

    x: [B,T,H,W,Cin] channels-last 16-bit; wpk: pack_conv_weight(...) output."""
    B, T, H, W, cin = x.shape
    kt, kh, kw = ksize
    assert wpk.shape[0] == kt * kh * kw and wpk.dtype == x.dtype
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    if out is None:
        if nchw_out:
            out = torch.empty(B * T, cout, Ho, Wo, dtype=torch.float32, device=x.device)
        else:
            out = torch.empty(B, T, Ho, Wo, cout, dtype=out_dtype or x.dtype, device=x.device)
    p = L.ConvParams()
    p.x = _ptr(x); p.B, p.T, p.H, p.W, p.Cin = B, T, H, W, cin
    p.x_cstride = _cstride(x)
    p.wgt = _ptr(wpk); p.Cout = cout
    p.kt, p.kh, p.kw = kt, kh, kw
    p.stride_hw = stride
    p.bias = _ptr(bias)
    p.rowbias = _ptr(rowbias)
    p.rowbias_stride = 0 if rowbias is None else rowbias.stride(0)
    if residual is not None:
        p.residual = _ptr(residual)
        p.residual_dtype = _DT[residual.dtype]
        p.residual_cstride = _cstride(residual)
    p.out = _ptr(out)
    p.out_dtype = _DT[out.dtype]
    p.out_layout = L.OUT_NCHW if nchw_out else L.OUT_NHWC
    p.out_cstride = 0 if nchw_out else _cstride(out)
    p.act = act
    p.in_dtype = _DT[x.dtype]
    p.out_scale = out_scale
    L.check(L.lib().flair_conv_igemm(C.byref(p), _stream()))
    return out
