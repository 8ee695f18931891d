"""Thin torch-tensor wrappers over the C ABI (raw pointers + sizes + stream).

Feature maps are channels-last `[B, T, H, W, C]` tensors (bf16 unless noted).
Nothing here computes on the host or through PyTorch ops: every function is a
single launch of a hand-written sm_100a kernel on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import os

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
    B, T, H, W, _ = t.shape
    cs = t.stride(3) if W > 1 else (t.stride(2) if H > 1 else t.shape[4])
    ok = (H == 1 or t.stride(2) == cs * W) and (T == 1 or t.stride(1) == cs * W * H) and \
        (B == 1 or t.stride(0) == cs * W * H * T)  # strides of size-1 dims are meaningless
    assert ok, "feature map must be pixel-contiguous (only the channel stride may be padded)"
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


def conv(x, wpk, cout, ksize=(1, 3, 3), *, bias=None, rowbias=None, rowscale=None, residual=None, residual2=None, out=None,
         out_dtype=None, act=L.ACT_NONE, stride=1, nchw_out=False, out_scale=1.0, out2=None, preadd=None, out2_neighbor=1,
         gn_groups=None):
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
    if kt == 1:  # frames are independent: let M-tiles span them (matters at 4x4 .. 16x16 maps)
        B, T = 1, B * T
    p.x = _ptr(x); p.B, p.T, p.H, p.W, p.Cin = B, T, H, W, cin
    p.x_cstride = _cstride(x)
    p.wgt = _ptr(wpk); p.Cout = cout
    p.kt, p.kh, p.kw = kt, kh, kw
    p.stride_hw = stride
    p.bias = _ptr(bias)
    p.rowbias = _ptr(rowbias)
    p.rowbias_stride = 0 if rowbias is None else rowbias.stride(0)
    p.rowscale = _ptr(rowscale)
    p.rowscale_stride = 0 if rowscale is None else rowscale.stride(0)
    if residual is not None:
        p.residual = _ptr(residual)
        p.residual_dtype = _DT[residual.dtype]
        p.residual_cstride = _cstride(residual)
    if residual2 is not None:
        p.residual2 = _ptr(residual2)
        p.residual2_dtype = _DT[residual2.dtype]
        p.residual2_cstride = _cstride(residual2)
    if preadd is not None:  # pre-activation addend (a partial convolution computed elsewhere), geometry of `out`
        assert preadd.shape[-1] == cout and preadd.shape[:4].numel() == B * T * Ho * Wo
        p.preadd = _ptr(preadd)
        p.preadd_dtype = _DT[preadd.dtype]
        p.preadd_cstride = _cstride(preadd)
    p.out = _ptr(out)
    p.out_dtype = _DT[out.dtype]
    p.out_layout = L.OUT_NCHW if nchw_out else L.OUT_NHWC
    p.out_cstride = 0 if nchw_out else _cstride(out)
    p.act = act
    p.in_dtype = _DT[x.dtype]
    p.out_scale = out_scale
    if out2 is not None:  # pair planes [groups][pixels][2][channels per group] for the deformable gather
        assert out2.dim() == 4 and out2.shape[2] == 2 and out2.stride(3) == 1 and out2.stride(2) == out2.shape[3] \
            and out2.stride(1) == 2 * out2.shape[3] and out2.dtype == out.dtype
        assert out2.shape[0] * out2.shape[3] == cout and out2.shape[1] == B * T * Ho * Wo
        p.out2 = _ptr(out2); p.out2_group_channels = out2.shape[3]; p.out2_group_stride = out2.stride(0)
        p.out2_neighbor = int(out2_neighbor)
    gn = None
    if gn_groups and FUSED_GN and not nchw_out and out.dtype != torch.float32 and cout % 16 == 0 and cout % gn_groups == 0:
        # GroupNorm statistics of the output from the conv's own epilogue (flair_conv_params.gn_partial): the next
        # norm then needs no statistics pass over the map.  Only when every M tile lies inside one batch element.
        mt, tpb, fpt = C.c_int(0), C.c_int(0), C.c_int(0)
        L.check(L.lib().flair_conv_gn_tiles(B, T, H, W, kh, kw, stride, C.byref(mt), C.byref(tpb), C.byref(fpt)))
        Bo = out.shape[0]
        # (2-D kernels run with the frames of all batch elements flattened: a tile of `fpt` consecutive frames stays
        # inside one batch element iff T % fpt == 0 — a rule that does not depend on the batch size, so a frame gets
        # the same bits whether it is run alone or in a batch)
        if kt == 3 or out.shape[1] % fpt.value == 0:
            gn = dict(groups=gn_groups, tpb=mt.value // Bo,
                      partial=torch.empty(mt.value * 4 * cout, dtype=torch.float32, device=x.device))
            p.gn_partial = _ptr(gn["partial"]); p.gn_groups = gn_groups
    rc = L.lib().flair_conv_igemm(C.byref(p), _stream())
    if rc == -3 and gn is not None:  # no compact epilogue for this launch (nothing ran): plain conv, separate statistics
        gn = None
        p.gn_partial = None; p.gn_groups = 0
        rc = L.lib().flair_conv_igemm(C.byref(p), _stream())
    L.check(rc)
    if gn is not None:
        Bo, To = out.shape[0], out.shape[1]
        nsplit = L.lib().flair_gn_finalize_splits(gn["tpb"])
        scratch = torch.empty(Bo * nsplit * cout, dtype=torch.float64, device=x.device)
        fin = torch.empty(Bo, gn_groups, 2, dtype=torch.float32, device=x.device)
        L.check(L.lib().flair_gn_finalize(_ptr(gn["partial"]), Bo, gn["tpb"], cout, gn_groups, To * Ho * Wo, _ptr(scratch),
                                          _ptr(_gn_counter(x.device, Bo)), _ptr(fin), 1e-5, _stream()))
        out._flair_gn = (gn_groups, out.data_ptr(), (fin, 0))
    elif hasattr(out, "_flair_gn"):
        del out._flair_gn   # `out` was a reused buffer: its old statistics are stale
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
    if gamma_arr is not None:  # the kernel indexes gamma_arr[frame]: a stride-0 expand() view would read its neighbours
        gamma_arr = _f32c(gamma_arr)
        assert gamma_arr.numel() == N, "gamma_arr must hold one weight per frame"
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


# ----------------------------------------------------------------------------------------------
# UNet building blocks on channels-last [B,T,H,W,C] maps
# ----------------------------------------------------------------------------------------------
_GN_COUNTERS = {}


def _gn_counter(device, B):
    """Zeroed ticket counters of flair_gn_stats (self-cleaning).  One buffer per device, allocated on the first eager
    call (before any CUDA-graph capture: every captured forward is preceded by an eager warm-up); all launches of a
    process are stream-ordered (one stream per process), which is what sharing the counter requires."""
    key = device
    c = _GN_COUNTERS.get(key)
    if c is None or c.numel() < B:
        c = torch.zeros(max(B, 64), dtype=torch.int32, device=device)
        _GN_COUNTERS[key] = c
    return c


# GroupNorm statistics from the producing conv's epilogue (flair_conv_params.gn_partial + flair_gn_finalize) are
# implemented and tested, but OFF by default: the warp-shuffle reduction in the epilogue shares the L1 data pipe with
# the tensor core's shared-memory operand reads, the 152 batched convs that carry it get ~20 us slower each (conv
# replay 32.7 -> 35.8 ms) and the forward is 56.2 ms with it against 55.4 ms with the separate (pipelined) statistics
# kernel (ABBA on one box, profiles/r02_summary.md).  FLAIR_FUSED_GN=1 turns it on.
FUSED_GN = os.environ.get("FLAIR_FUSED_GN", "0") == "1"


def gn_stats(x, groups=32, eps=1e-5):
    """GroupNorm statistics over (T,H,W,C/G) per batch element.  Returns (final, 0): final[b][g] = (mean, rstd),
    reduced deterministically by the last CTA of the launch (see flair_gn_stats); gn_apply takes the pair as is."""
    B, T, H, W, Cc = x.shape
    fused = getattr(x, "_flair_gn", None)   # left by ops.conv(gn_groups=...) on the tensor it produced
    if fused is not None and fused[0] == groups and fused[1] == x.data_ptr() and eps == 1e-5 and FUSED_GN:
        return fused[2]
    P = T * H * W
    n = L.lib().flair_gn_stats_chunks(P, Cc)
    counter = _gn_counter(x.device, B)
    part = torch.empty(B, n, groups, 2, dtype=torch.float32, device=x.device)
    fin = torch.empty(B, groups, 2, dtype=torch.float32, device=x.device)
    L.check(L.lib().flair_gn_stats(_ptr(x), _DT[x.dtype], B, P, Cc, _cstride(x), groups, _ptr(part), n,
                                   _ptr(counter), _ptr(fin), float(eps), _stream()))
    return fin, 0


def gn_apply(x, stats=None, gamma=None, beta=None, *, scale=None, shift=None, silu=False, resample=0,
             out_dtype=None, groups=32, out=None):
    """Fused normalise (+FiLM) (+SiLU) (+2x resample).  stats=None -> plain resample / cast."""
    B, T, H, W, Cc = x.shape
    Ho, Wo = {0: (H, W), 1: (2 * H, 2 * W), 2: (H // 2, W // 2)}[resample]
    if out is None:
        out = torch.empty(B, T, Ho, Wo, Cc, dtype=out_dtype or x.dtype, device=x.device)
    p = L.GNApplyParams()
    p.x = _ptr(x); p.in_dtype = _DT[x.dtype]; p.out = _ptr(out); p.out_dtype = _DT[out.dtype]
    if stats is not None:
        partial, n = stats
        p.partial = _ptr(partial); p.nchunks = n; p.gamma = _ptr(gamma); p.beta = _ptr(beta); p.norm = 1
        if scale is not None:
            assert scale.stride(0) == shift.stride(0) and scale.stride(1) == 1
            p.scale = _ptr(scale); p.shift = _ptr(shift); p.film_stride = scale.stride(0)
    p.B, p.T, p.H, p.W, p.C, p.groups = B, T, H, W, Cc, groups
    p.x_cstride = _cstride(x); p.out_cstride = _cstride(out)
    p.silu = int(silu); p.resample = resample; p.eps = 1e-5
    L.check(L.lib().flair_gn_apply(C.byref(p), _stream()))
    return out


def concat_channels(a, b):
    """th.cat([a, b], channel) for channels-last maps."""
    B, T, H, W, Ca = a.shape
    Cb = b.shape[-1]
    out = torch.empty(B, T, H, W, Ca + Cb, dtype=a.dtype, device=a.device)
    pix = B * T * H * W
    es = a.element_size()
    L.check(L.lib().flair_copy_channels(_ptr(a), _ptr(out), pix, Ca, es, _cstride(a), Ca + Cb, 0, _stream()))
    L.check(L.lib().flair_copy_channels(_ptr(b), _ptr(out), pix, Cb, es, _cstride(b), Ca + Cb, Ca, _stream()))
    return out


def copy_channels_into(src, dst, coffset):
    B, T, H, W, Cs = src.shape
    L.check(L.lib().flair_copy_channels(_ptr(src), _ptr(dst), B * T * H * W, Cs, src.element_size(), _cstride(src),
                                        _cstride(dst), coffset, _stream()))


def attn_spatial(qkv, heads, rowbias=None):
    """qkv [B,T,H,W,3C] head-major (H,3,64) -> [B,T,H,W,C]; attention over the H*W tokens of each frame."""
    B, T, H, W, C3 = qkv.shape
    Cc = C3 // 3
    out = torch.empty(B, T, H, W, Cc, dtype=qkv.dtype, device=qkv.device)
    L.check(L.lib().flair_attn_spatial(_ptr(qkv), _ptr(out), _ptr(rowbias), 0 if rowbias is None else rowbias.stride(0),
                                       B * T, H * W, heads, _cstride(qkv), Cc, _DT[qkv.dtype], _stream()))
    return out


def attn_temporal(qkv, cq, ck, bv, frames):
    B, T, H, W, C3 = qkv.shape
    Cc = C3 // 3
    assert qkv.is_contiguous()
    out = torch.empty(B, T, H, W, Cc, dtype=qkv.dtype, device=qkv.device)
    L.check(L.lib().flair_attn_temporal(_ptr(qkv), _ptr(out), _ptr(cq), _ptr(ck), _ptr(bv), B, T, H * W, Cc, frames,
                                        _DT[qkv.dtype], _stream()))
    return out


def timestep_embedding(t, freqs):
    t = t.float().contiguous()
    out = torch.empty(t.shape[0], 2 * freqs.shape[0], dtype=torch.float32, device=t.device)
    L.check(L.lib().flair_timestep_embedding_f32(_ptr(t), _ptr(freqs), _ptr(out), t.shape[0], 2 * freqs.shape[0], _stream()))
    return out


def linear_f32(x, Wt, bias=None, silu_in=False, silu_out=False, sigmoid_out=False):
    """y = act_out(bias + act_in(x) @ Wt), Wt [K,N] fp32 (small M)."""
    silu_out = 2 if sigmoid_out else int(silu_out)
    M, K = x.shape
    N = Wt.shape[1]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    L.check(L.lib().flair_linear_f32(_ptr(x), _ptr(Wt), _ptr(bias), _ptr(y), M, K, N, int(silu_in), int(silu_out), _stream()))
    return y


def pack_im2col6(a, b, dtype=torch.bfloat16):
    """cat([a, b], 1) of two (N,3,H,W) fp32 maps -> [1,N,H,W,64] 16-bit im2col map (k = tap*6 + c)."""
    a, b = _f32c(a), _f32c(b)
    N, _, H, W = a.shape
    out = torch.empty(1, N, H, W, 64, dtype=dtype, device=a.device)
    L.check(L.lib().flair_pack_im2col6(_ptr(a), _ptr(b), _ptr(out), N, H, W, _DT[dtype], _stream()))
    return out


# ----------------------------------------------------------------------------------------------
# BasicVSR++ support (per-frame channels-last maps [N,H,W,C]; flows fp32 [N,2,H,W])
# ----------------------------------------------------------------------------------------------
def _cs4(t):
    N, H, W, _ = t.shape
    cs = t.stride(2) if W > 1 else (t.stride(1) if H > 1 else t.shape[3])
    assert t.dim() == 4 and t.stride(3) == 1 and (H == 1 or t.stride(1) == cs * W) \
        and (N == 1 or t.stride(0) == cs * W * H), "expected pixel-contiguous [N,H,W,C] view"
    return cs


def flow_warp(x, flow, out=None):
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty(N, H, W, Cc, dtype=x.dtype, device=x.device)
    L.check(L.lib().flair_flow_warp(_ptr(x), _ptr(flow), _ptr(out), N, H, W, Cc, _cs4(x), _cs4(out), _DT[x.dtype], _stream()))
    return out


def flow_warp2(xa, flow_a, out_a, xb, flow_b, out_b):
    """out_a = warp(xa, flow_a), out_b = warp(xb, flow_b) in one launch (same shapes, same output channel stride)."""
    N, H, W, Cc = xa.shape
    assert xb.shape == xa.shape and _cs4(out_a) == _cs4(out_b)
    L.check(L.lib().flair_flow_warp2(_ptr(xa), _ptr(xb), _ptr(flow_a), _ptr(flow_b), _ptr(out_a), _ptr(out_b), N, H, W, Cc,
                                     _cs4(xa), _cs4(xb), _cs4(out_a), _DT[xa.dtype], _stream()))


def flow_compose(f2, f1):
    out = torch.empty_like(f1)
    N, _, H, W = f1.shape
    L.check(L.lib().flair_flow_compose_f32(_ptr(f2), _ptr(f1), _ptr(out), N, H, W, _stream()))
    return out


def planes_to_cl(src, dst, coffset):
    N, Cs, H, W = src.shape
    L.check(L.lib().flair_planes_to_cl(_ptr(src), _ptr(dst), N, Cs, H, W, _cs4(dst), coffset, _DT[dst.dtype], _stream()))


def deform_im2col(xa, xb, om, flow1, flow2, deform_groups, mrm):
    N, H, W, Cc = xa.shape
    cols = torch.empty(1, N, H, W, 18 * Cc, dtype=xa.dtype, device=xa.device)
    L.check(L.lib().flair_deform_im2col(_ptr(xa), _ptr(xb), _cs4(xa), _cs4(xb), _ptr(om), _cs4(om), _DT[om.dtype],
                                        _ptr(flow1), _ptr(flow2), _ptr(cols), N, H, W, Cc, deform_groups, float(mrm),
                                        _DT[xa.dtype], _stream()))
    return cols


def deform_offset_perm(deform_groups=16):
    """Row permutation of the last offset-conv weight/bias that flair_deform_conv expects:
    new channel tap*48 + quad*12 + kind*4 + gi  <-  reference channel (dy/dx: (g*9+tap)*2 + kind, mask: 288 + g*9 + tap),
    g = quad*4 + gi."""
    assert deform_groups == 16
    perm = []
    for tap in range(9):
        for quad in range(4):
            for kind in range(3):
                for gi in range(4):
                    g = quad * 4 + gi
                    perm.append((g * 9 + tap) * 2 + kind if kind < 2 else 288 + g * 9 + tap)
    return torch.tensor(perm, dtype=torch.long)


def deform_weight_kperm(C):
    """Column permutation of the (C, 9*2C) tap-major deformable weight (k = tap*2C + channel of cat(xa, xb)) into the
    channel-block-major K order flair_deform_conv walks: k' = (kbq*9 + tap)*64 + c, kbq = 64-channel block."""
    nkb = 2 * C // 64
    k = torch.arange(9 * 2 * C).reshape(9, nkb, 64)          # [tap][kbq][c] -> old index
    return k.permute(1, 0, 2).reshape(-1)


def pack_deform_weight(w, dtype):
    """w: (C, 2C, 3, 3) ModulatedDeformConv2d weight or its (C, 9*2C) tap-major matrix -> packed weight of
    flair_deform_conv (channel-block-major K)."""
    if w.dim() == 4:
        w = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    return pack_conv_weight(w.float()[:, deform_weight_kperm(w.shape[0]).to(w.device)], dtype)


def pair_planes(x, vertical=False):
    """[N,H,W,C] channels-last map -> pair planes [8][N*H*W][2][C/8].  Horizontal (what flair_deform_conv gathers from):
    entry p = pixels (p, p+1) of the row-major map; vertical: entry (y,x) = pixels ((y,x), (y+1,x)).  Torch ops: test /
    setup helper; in the model the conv epilogue writes this layout directly (ops.conv(out2=..., out2_neighbor=...))."""
    N, H, W, Cc = x.shape
    g = x.reshape(N * H * W, 8, Cc // 8).permute(1, 0, 2)
    d = W if vertical else 1
    nxt = torch.cat([g[:, d:], torch.zeros_like(g[:, :d])], 1)
    return torch.stack([g, nxt], 2).contiguous()


def pair_neighbor(C, W):
    """out2_neighbor of the pair planes flair_deform_conv expects for C feature channels (horizontal pairs)."""
    return 1


def deform_conv(xa, xb, om, flow1, flow2, wpk, bias, mrm, *, out=None):
    """Fused offsets + modulated deformable 3x3 conv over cat(xa, xb) (flair_deform_conv).

    xa / xb: pair planes [8][N*H*W][2][C/8] (see pair_planes / ops.conv(out2=...));
    om: [N,H,W,>=432] fp16 tap-major offset-net output (see deform_offset_perm); out: [N,H,W,C] view."""
    N, H, W = om.shape[:3]
    Cc = wpk.shape[1]
    assert om.dtype == torch.float16, "the offset map is fp16 (also with bf16 features)"
    if out is None:
        out = torch.empty(N, H, W, Cc, dtype=wpk.dtype, device=om.device)
    p = L.DeformConvParams()
    for name, t in (("xa", xa), ("xb", xb)):
        setattr(p, name, _ptr(t))
        assert t.shape == (8, N * H * W, 2, Cc // 8) and t.stride(3) == 1 and t.stride(2) == Cc // 8 \
            and t.stride(1) == Cc // 4, "expected pair planes [8][N*H*W][2][C/8]"
        gs, ps, ns = t.stride(0), t.stride(1), H * W * t.stride(1)
        setattr(p, name + "_gstride", gs); setattr(p, name + "_pstride", ps); setattr(p, name + "_nstride", ns)
    p.om = _ptr(om); p.om_cstride = _cs4(om)
    p.flow1 = _ptr(flow1); p.flow2 = _ptr(flow2)
    p.wgt = _ptr(wpk); p.bias = _ptr(bias)
    p.out = _ptr(out); p.out_cstride = _cs4(out)
    p.N, p.H, p.W, p.C = N, H, W, Cc
    p.deform_groups = 16
    p.max_residue_magnitude = float(mrm)
    p.dtype = _DT[wpk.dtype]
    L.check(L.lib().flair_deform_conv(C.byref(p), _stream()))
    return out


def scale_pixels_(x, wmap):
    N, H, W, Cc = x.shape
    L.check(L.lib().flair_scale_pixels(_ptr(x), _ptr(wmap), N * H * W, Cc, _cs4(x), _DT[x.dtype], _stream()))
    return x


# ---------------------------------------------------------------------------------------------------------
# aux face-prior warps: facelib/utils/face_restoration_helper.py:225-345 on the device
# ---------------------------------------------------------------------------------------------------------
def warp_affine_cubic(src, minv, out_hw, *, border=None, in_mode=0, out_mode=0):
    """src (N, C<=4, Hs, Ws) fp32 -> (N, C, Hd, Wd); `minv` (N, 6) float64 DEVICE tensor = destination->source maps.
    cv2.warpAffine(INTER_CUBIC, BORDER_CONSTANT) arithmetic; in_mode/out_mode 1 fuse the [-1,1] <-> [0,255] maps."""
    src = _f32c(src)
    N, Cc, Hs, Ws = src.shape
    assert minv.is_cuda and minv.dtype == torch.float64 and minv.is_contiguous() and minv.numel() == 6 * N
    Hd, Wd = out_hw
    dst = torch.empty(N, Cc, Hd, Wd, dtype=torch.float32, device=src.device)
    b = None
    if border is not None:
        assert len(border) == Cc
        b = (C.c_float * Cc)(*[float(v) for v in border])
    L.check(L.lib().flair_warp_affine_cubic_f32(_ptr(src), _ptr(dst), _ptr(minv), N, Cc, Hs, Ws, Hd, Wd, b,
                                                int(in_mode), int(out_mode), _stream()))
    return dst


def parse_mask(logits, lut_bits):
    """(N, classes, H, W) fp32 logits -> (N, 1, H, W) fp32 mask: 255 where bit argmax of `lut_bits` is set, else 0."""
    logits = _f32c(logits)
    N, K, H, W = logits.shape
    mask = torch.empty(N, 1, H, W, dtype=torch.float32, device=logits.device)
    L.check(L.lib().flair_parse_mask_f32(_ptr(logits), _ptr(mask), N, K, H, W, int(lut_bits), _stream()))
    return mask


def gaussian_blur101(x, taps, *, finish=False, thres=10, scale=255.0, out=None, tmp=None):
    """cv2.GaussianBlur(x, (101, 101), sigma) per (N, 1, H, W) plane; taps = fp32 cast of getGaussianKernel(101, sigma)
    on the device.  finish: clear a `thres` frame and divide by `scale` (inverse_faces :309-318)."""
    x = _f32c(x)
    N, H, W = x.shape[0] * x.shape[1], x.shape[2], x.shape[3]
    out = torch.empty_like(x) if out is None else out
    tmp = torch.empty_like(x) if tmp is None else tmp
    assert taps.is_cuda and taps.dtype == torch.float32 and taps.numel() == 101
    L.check(L.lib().flair_gaussian_blur_f32(_ptr(x), _ptr(out), _ptr(tmp), _ptr(taps), 101, N, H, W, int(finish),
                                            int(thres), float(scale), _stream()))
    return out


def aux_blend(x0, face, mask, w, *, clip_denoised=True):
    """w x0 + (1 - w) clamp(x0 (1 - mask) + face mask)   (gaussian_diffusion.py:488-496); mask (N, 1, H, W)."""
    x0, face, mask = _f32c(x0), _f32c(face), _f32c(mask)
    N, Cc, H, W = x0.shape
    assert face.shape == x0.shape and mask.shape == (N, 1, H, W)
    out = torch.empty_like(x0)
    L.check(L.lib().flair_aux_blend_f32(_ptr(x0), _ptr(face), _ptr(mask), _ptr(out), float(w), N, Cc, H, W,
                                        int(clip_denoised), _stream()))
    return out
