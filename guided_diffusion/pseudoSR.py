"""Blur + x4 down-sampling operator and its pseudo-inverse, as fp32 stencil kernels on the B200.

API of the reference's guided_diffusion/pseudoSR.py: `Get_pseudoSR_Conf` (:383-396), `pseudoSR`
(:47-171: host-side filter preparation incl. the FFT inverse of h^T h) and `pseudoSR_PyTorch`
(:174-295: `DownscaleOP`, `Conv_LR_with_Inv_hTh_OP`, `Upscale_OP`, `A_pinv`, `A`).  The three
`Filter_Layer`s (:15-44, depth-wise nn.Conv2d + ReplicationPad2d) become shared-memory stencil
kernels (flair_blur_down_f32 / flair_filter_same_f32 / flair_blur_up_f32); by linearity
`A_pinv(LR, x)` is evaluated as Up(InvhTh(codec(Down(x)) - LR)) so the full-resolution work is one
down pass and one polyphase up pass (or none at all when fused into the sampler update).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from scipy.signal import convolve2d as conv2

from flair_b200 import ops

from .imresize_pseudoSR import calc_strides, upscale_kernel_from_blur


def Get_pseudoSR_Conf(sf):
    class conf:
        scale_factor = sf
        avoid_skip_connections = False
        generate_HR_image = False
        pseudo_pseudoSR_supplement = False
        desired_inv_hTh_energy_portion = 1 - 1e-6
        filter_pertubation_limit = 1.1
        sigmoid_range_limit = False
        lower_magnitude_bound = 0.01  # floor of |FFT(h^T h)| before inversion
    return conf


def Aliased_Down_Sampling(array, factor):
    pre, _ = calc_strides(array, 1 / factor, align_center=True)
    return array[pre[0]::factor, pre[1]::factor]


def Return_kernel(ds_factor, upscale_kernel=None, kernel_indx=0):
    if not isinstance(upscale_kernel, np.ndarray) or kernel_indx < 8:
        raise NotImplementedError("only the explicit-blur-kernel path of the FLAIR demo is supported "
                                  "(scripts/video_sample.py:258-260: kernels[0,3], kernel_indx=10)")
    k, pre, post = upscale_kernel_from_blur(upscale_kernel, int(ds_factor))
    return np.rot90(k, 2).astype(np.float32) / (ds_factor ** 2), pre, post


class pseudoSR:
    """Host-side operator description: `ds_kernel` (9x9 for the demo kernel), `inv_hTh` (39x39)."""
    NFFT_add = 36

    def __init__(self, conf, upscale_kernel=None, kernel_indx=0):
        self.conf = conf
        self.ds_factor = np.array(conf.scale_factor, dtype=np.int32)
        if np.round(self.ds_factor) != self.ds_factor:
            raise AssertionError("Currently only supporting integer scale factors")
        self.ds_kernel, self.pre_stride, self.post_stride = Return_kernel(
            self.ds_factor, upscale_kernel=upscale_kernel, kernel_indx=kernel_indx)
        self.compute_inv_hTh()

    def compute_inv_hTh(self):
        """39x39 spatial filter ~ (h^T h)^-1 at LR (reference :123-171)."""
        sf = int(self.ds_factor)
        hTh = Aliased_Down_Sampling(conv2(self.ds_kernel, np.rot90(self.ds_kernel, 2)) * sf ** 2, sf)
        pad = int(self.NFFT_add / 2)
        F = np.fft.fft2(np.pad(hTh, ((pad, pad), (pad, pad)), mode="constant", constant_values=0))
        F = F * np.maximum(1, self.conf.lower_magnitude_bound / np.abs(F))
        inv = np.real(np.fft.ifft2(1 / F))
        n = inv.shape[0]
        r, c = np.argmax(inv) // n, np.mod(np.argmax(inv), n)
        if not np.all(np.equal(np.ceil(np.array(inv.shape) / 2), np.array([r, c]) - 1)):
            half = np.min([n - r - 1, n - c - 1, r, c])
            inv = inv[r - half:r + half + 1, c - half:c + half + 1]
        self.inv_hTh_invalidity_half_size = 26
        drop = inv.shape[0] // 2 - 26
        if drop > 0:
            inv = inv[drop:-drop, drop:-drop]
        self.inv_hTh = inv

    def WrapArchitecture_PyTorch(self, grayscale=False):
        self.loss_mask = None
        wrapped = pseudoSR_PyTorch(self, grayscale=grayscale)
        self.OP_names = ["DownscaleOP.Filter_OP", "Conv_LR_with_Inv_hTh_OP.Filter_OP", "Upscale_OP.Filter_OP"]
        return wrapped


class _StencilOP(nn.Module):
    """One of the reference's three Filter_Layers; the taps live in a buffer so `.to(device)` works.
    `Filter_OP.weight` keeps the (C,1,k,k) view the reference exposes."""

    class _W(nn.Module):
        def __init__(self, taps, channels):
            super().__init__()
            self.register_buffer("taps", taps)
            self.channels = channels
            self.filter_layer = True

        @property
        def weight(self):
            return self.taps[None, None].expand(self.channels, 1, *self.taps.shape)

    def __init__(self, taps, kind, sf, pre, channels):
        super().__init__()
        self.Filter_OP = _StencilOP._W(torch.from_numpy(np.ascontiguousarray(taps)).float(), channels)
        self.kind, self.sf, self.pre = kind, sf, pre

    def forward(self, x):
        t = self.Filter_OP.taps
        if self.kind == "down":
            return ops.blur_down(x, t, self.sf, self.pre)
        if self.kind == "up":
            return ops.blur_up(x, t, self.sf, self.pre)
        return ops.filter_same(x, t)


class pseudoSR_PyTorch(nn.Module):
    def __init__(self, host: pseudoSR, grayscale=False):
        super().__init__()
        ch = 1 if grayscale else 3
        self.ds_factor = host.ds_factor
        self.conf = host.conf
        sf = int(host.ds_factor)
        pre, post = calc_strides(None, host.ds_factor)
        self.Conv_LR_with_Inv_hTh_OP = _StencilOP(host.inv_hTh, "same", sf, int(pre[0]), ch)
        self.Upscale_OP = _StencilOP(host.ds_kernel * host.ds_factor ** 2, "up", sf, int(pre[0]), ch)
        self.DownscaleOP = _StencilOP(np.rot90(host.ds_kernel, 2), "down", sf, int(pre[0]), ch)
        self.ds_kernel = host.ds_kernel
        self.pre_stride, self.post_stride = pre, post

    # -- fused entry used by the sampler: LR-domain correction q with R = Upscale_OP(q)
    def lr_correction(self, LR, generated_image, jpeg_decode=None, jpeg_encode=None, inv_LR=None):
        """q = InvhTh(codec(Down(x))) - InvhTh(LR).  `inv_LR` = InvhTh(LR) precomputed by the caller (constant
        over the sampling steps of a window); else it is cached per LR tensor."""
        lr = self.DownscaleOP(generated_image)
        if jpeg_encode is not None and jpeg_decode is not None:
            lr = jpeg_decode(jpeg_encode(lr))
        if inv_LR is None and LR is not None:
            inv_LR = self._inv_of(LR)
        return ops.filter_same(lr, self.Conv_LR_with_Inv_hTh_OP.Filter_OP.taps, sub=inv_LR)

    def _inv_of(self, LR):
        """InvhTh(LR) — constant over the 100 steps of a window, so cached per LR tensor."""
        key = (LR.data_ptr(), tuple(LR.shape), tuple(LR.stride()), LR._version)
        if getattr(self, "_inv_key", None) != key:
            self._inv_cache = self.Conv_LR_with_Inv_hTh_OP(LR)
            self._inv_key = key
            self._inv_src = LR  # keeps the storage alive, so the pointer in the key cannot be recycled
        return self._inv_cache

    def A_pinv(self, LR, generated_image=None, jpeg_decode=None, jpeg_encode=None):
        """Reference :248-281.  With an image: Up(InvhTh(codec(Down(x)))) - Up(InvhTh(LR)); without:
        Up(InvhTh(LR))."""
        LR = LR[:, -3:, :, :]
        if generated_image is None:
            return self.Upscale_OP(self.Conv_LR_with_Inv_hTh_OP(LR))
        if np.any(np.mod(generated_image.size()[2:], int(self.ds_factor)) != 0):
            raise AssertionError("image size must be a multiple of the scale factor")
        if self.conf.sigmoid_range_limit:
            raise NotImplementedError("sigmoid_range_limit is disabled in FLAIR (scripts/video_sample.py:256)")
        return self.Upscale_OP(self.lr_correction(LR, generated_image, jpeg_decode, jpeg_encode))

    def A(self, HR, scale_factor=1.0, use_zero_padding=False, kk=None):
        raise NotImplementedError("pseudoSR_PyTorch.A (reflect-padded imresize_efficient, reference :283-295) "
                                  "is not on the FLAIR sampling path; use DownscaleOP")
