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


# ----------------------------------------------------------------------------------------------
# fp32 sampler / data-consistency kernels (NCHW planes)
# ----------------------------------------------------------------------------------------------
def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "flair_b200 ops run on CUDA tensors only (no CPU fallback)"
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.float().contiguous()
    return t


def sampler_update(x_t, coef, *, model_out=None, noise, t=0, t_arr=None, gamma_arr=None, R=None, q_lr=None,
                   up_taps=None, sf=4, pre_stride=1, prev=None, frames_per_window=1, rho=0.35, x0_in=None,
                   clip_denoised=True, want_pred_xstart=True):
    """Fused p_sample tail (flair_sampler_update_f32).  Returns (sample, pred_xstart|None)."""
    import math
    x_t = _f32c(x_t)
    N, _, H, W = x_t.shape
    sample = torch.empty_like(x_t)
    x0 = torch.empty_like(x_t) if want_pred_xstart else None
    p = L.UpdateParams()
    p.x_t = _ptr(x_t)
    if model_out is not None:
        model_out = _f32c(model_out)
        p.model_out = _ptr(model_out); p.model_ch = model_out.shape[1]
    noise = _f32c(noise)
    p.noise = _ptr(noise)
    keep = [x_t, model_out, noise]
    if R is not None:
        R = _f32c(R); p.R = _ptr(R); keep.append(R)
    if q_lr is not None:
        q_lr = _f32c(q_lr); p.q_lr = _ptr(q_lr); p.up_taps = _ptr(up_taps)
        p.up_k = up_taps.shape[-1]; p.sf = sf; p.pre_stride = pre_stride
    if prev is not None:
        prev = _f32c(prev); p.prev = _ptr(prev); p.prev_k = prev.shape[1]
        p.frames_per_window = frames_per_window
        keep.append(prev)
    if x0_in is not None:
        x0_in = _f32c(x0_in); p.x0_in = _ptr(x0_in); keep.append(x0_in)
    p.coef = _ptr(coef)
    if t_arr is not None:
        assert t_arr.dtype == torch.int64 and t_arr.is_cuda and t_arr.numel() == N
        p.t_arr = _ptr(t_arr)
    if gamma_arr is not None:
        gamma_arr = _f32c(gamma_arr); p.gamma_arr = _ptr(gamma_arr); keep.append(gamma_arr)
    p.t = int(t)
    # python floats multiplied into fp32 tensors by the reference (gaussian_diffusion.py:514)
    p.sqrt_one_minus_rho = float(torch.tensor(math.sqrt(1 - rho), dtype=torch.float64).float())
    p.sqrt_rho = float(torch.tensor(math.sqrt(rho), dtype=torch.float64).float())
    p.sample = _ptr(sample); p.pred_xstart = _ptr(x0)
    p.N, p.H, p.W = N, H, W
    p.clip_denoised = 1 if clip_denoised else 0
    L.check(L.lib().flair_sampler_update_f32(C.byref(p), _stream()))
    return sample, x0


def pred_xstart(x_t, model_out, coef, *, t=0, t_arr=None, clip_denoised=True):
    x_t, model_out = _f32c(x_t), _f32c(model_out)
    N, _, H, W = x_t.shape
    out = torch.empty_like(x_t)
    L.check(L.lib().flair_pred_xstart_f32(_ptr(x_t), _ptr(model_out), model_out.shape[1], _ptr(coef),
                                          _ptr(t_arr), int(t), _ptr(out), N, H, W, int(clip_denoised), _stream()))
    return out


def dc_apply(x0, R, *, gamma=1.0, gamma_arr=None, clip_denoised=True):
    x0, R = _f32c(x0), _f32c(R)
    N, _, H, W = x0.shape
    out = torch.empty_like(x0)
    L.check(L.lib().flair_dc_apply_f32(_ptr(x0), _ptr(R), _ptr(gamma_arr), float(gamma), _ptr(out), N, H, W,
                                       int(clip_denoised), _stream()))
    return out


def mean_variance(x_t, x0, model_out, tab, *, learned_range, t=0, t_arr=None):
    x_t, x0 = _f32c(x_t), _f32c(x0)
    N, _, H, W = x_t.shape
    mean, var, logvar = torch.empty_like(x_t), torch.empty_like(x_t), torch.empty_like(x_t)
    mo = _f32c(model_out) if learned_range else None
    L.check(L.lib().flair_mean_variance_f32(_ptr(x_t), _ptr(x0), _ptr(mo), 6 if learned_range else 3,
                                            int(learned_range), _ptr(tab), _ptr(t_arr), int(t), _ptr(mean),
                                            _ptr(var), _ptr(logvar), N, H, W, _stream()))
    return mean, var, logvar


def axpby(x, y, alpha, beta):
    x, y = _f32c(x), _f32c(y)
    out = torch.empty_like(x)
    assert x.numel() % 4 == 0
    L.check(L.lib().flair_axpby_f32(_ptr(x), _ptr(y), float(alpha), float(beta), _ptr(out), x.numel(), _stream()))
    return out


def blur_down(x, taps, sf, pre):
    x = _f32c(x)
    N, Cc, H, W = x.shape
    out = torch.empty(N, Cc, H // sf, W // sf, dtype=torch.float32, device=x.device)
    L.check(L.lib().flair_blur_down_f32(_ptr(x), _ptr(out), _ptr(taps), taps.shape[-1], sf, pre, N * Cc, H, W,
                                        _stream()))
    return out


def filter_same(x, taps, sub=None):
    x = _f32c(x)
    N, Cc, H, W = x.shape
    out = torch.empty_like(x)
    if sub is not None:
        sub = _f32c(sub)
    L.check(L.lib().flair_filter_same_f32(_ptr(x), _ptr(sub), _ptr(out), _ptr(taps), taps.shape[-1], N * Cc, H, W,
                                          _stream()))
    return out


def blur_up(lr, taps, sf, pre):
    lr = _f32c(lr)
    N, Cc, h, w = lr.shape
    out = torch.empty(N, Cc, h * sf, w * sf, dtype=torch.float32, device=lr.device)
    L.check(L.lib().flair_blur_up_f32(_ptr(lr), _ptr(out), _ptr(taps), taps.shape[-1], sf, pre, N * Cc, h * sf,
                                      w * sf, _stream()))
    return out


def jpeg(mode, tables, x=None, planes=None):
    """mode 0 encode -> [luma, chroma]; 1 decode(planes) -> img; 2 decode(encode(x)) -> img."""
    dct, idct, q1, q2 = tables
    if mode == 1:
        luma, chroma = _f32c(planes[0]), _f32c(planes[1])
        N, _, h, w = luma.shape
        dev = luma.device
    else:
        x = _f32c(x)
        N, _, h, w = x.shape
        dev = x.device
        luma = chroma = None
        if mode == 0:
            luma = torch.empty(N, 1, h, w, dtype=torch.float32, device=dev)
            chroma = torch.empty(N, 2, h // 2, w // 2, dtype=torch.float32, device=dev)
    out = torch.empty(N, 3, h, w, dtype=torch.float32, device=dev) if mode != 0 else None
    L.check(L.lib().flair_jpeg_f32(mode, _ptr(x), _ptr(luma), _ptr(chroma), _ptr(out), _ptr(dct), _ptr(idct),
                                   _ptr(q1), _ptr(q2), N, h, w, _stream()))
    return [luma, chroma] if mode == 0 else out


def sandwich(Lm, X, Rm, sub=None):
    """out[pl] = Lm @ X[pl] @ Rm (- sub[pl]);  X: (planes, q, r) fp32."""
    X = _f32c(X)
    planes, q, r = X.shape
    p_, s_ = Lm.shape[0], Rm.shape[1]
    assert Lm.shape[1] == q and Rm.shape[0] == r
    ws = torch.empty(planes, p_, r, dtype=torch.float32, device=X.device)
    out = torch.empty(planes, p_, s_, dtype=torch.float32, device=X.device)
    if sub is not None:
        sub = _f32c(sub)
    L.check(L.lib().flair_sandwich_f32(_ptr(Lm), _ptr(X), _ptr(Rm), _ptr(sub), _ptr(out), planes, p_, q, r, s_,
                                       _ptr(ws), _stream()))
    return out
