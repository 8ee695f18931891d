"""Host-side (numpy, float64) preparation of the blur x4 operator's filters — setup, not hot path.

Restates the pieces of the reference's guided_diffusion/imresize_pseudoSR.py that the demo reaches
(`calc_strides` :63-76, `Center_Mass` :121-157, `Return_Filter_Energy_Distribution` :159-161 and the
ndarray-kernel branch of `imresize(..., return_upscale_kernel=True)` :10-61).  The result is checked
bit-for-bit against filters dumped from the reference (tests/golden/pseudosr_taps.pt).  Unlike the
reference, nothing is written to the working directory (it drops `rot59.mat`, :59)."""
from __future__ import annotations

import numpy as np
from scipy.signal import convolve2d


def calc_strides(array, factor, align_center=False):
    """Samples before / after each kept sample when (de)interleaving by an integer factor."""
    f = np.maximum(factor, 1 / factor).astype(np.int32)
    if align_center:
        half = np.ceil(np.array(array.shape[:2]) / 2 * (factor if factor > 1 else 1))
        pre = np.mod(half, f)
        pre[pre == 0] = f
        pre = (pre - 1).astype(np.int32)
        return pre, f - pre - 1
    post = (np.floor(f / 2) * np.ones([2])).astype(np.int32)
    return (f - post - 1).astype(np.int32), post


def _energy_profile(k):
    """sqrt-energy kept after peeling 0,1,2,... border frames, relative to the whole filter."""
    e = [np.sqrt(np.sum(k ** 2))]
    e += [np.sqrt(np.sum(k[i:-i, i:-i] ** 2)) for i in range(1, int(np.ceil(k.shape[0] / 2)))]
    return e / e[0]


def _rint(v):
    return int(np.round(np.asarray(v).reshape(-1)[0]))


def Center_Mass(kernel, ds_factor):
    """Pad a square kernel so its centre of mass sits in the middle, trim negligible borders so that
    (size - 1 + [ds_factor even]) is a multiple of ds_factor, renormalise to sum 1."""
    if kernel.shape[0] != kernel.shape[1]:
        raise AssertionError("Currently supporting only square kernels")
    n = kernel.shape[0]
    gx, gy = np.meshgrid(np.arange(n), np.arange(n))
    cx = convolve2d(gx, kernel, mode="valid") + 1
    cy = convolve2d(gy, kernel, mode="valid") + 1
    x_pad, y_pad = 2 * (n / 2 - cx), 2 * (n / 2 - cy)
    diff = np.round(np.abs(y_pad)) - np.round(np.abs(x_pad))
    pre_x, post_x = np.maximum(0, -x_pad), np.maximum(0, x_pad)
    pre_y, post_y = np.maximum(0, -y_pad), np.maximum(0, y_pad)

    def widen(pre, post, extra):
        lean_right = np.round(post) - post - (np.round(pre) - pre)
        pre, post = _rint(pre), _rint(post)
        if lean_right > 0:
            return pre + int(np.floor(extra / 2)), post + int(np.ceil(extra / 2))
        return pre + int(np.ceil(extra / 2)), post + int(np.floor(extra / 2))

    if diff > 0:
        pre_y, post_y = _rint(pre_y), _rint(post_y)
        pre_x, post_x = widen(pre_x, post_x, diff)
    elif diff < 0:
        pre_x, post_x = _rint(pre_x), _rint(post_x)
        pre_y, post_y = widen(pre_y, post_y, -diff)
    kernel = np.pad(kernel, ((_rint(pre_y), _rint(post_y)), (_rint(pre_x), _rint(post_x))), mode="constant")
    if kernel.shape[0] != kernel.shape[1]:
        raise AssertionError("kernel stopped being square")
    drop = np.argwhere(_energy_profile(kernel) < 0.99)[0][0] * np.ones([2]).astype(np.int32)
    side = 0
    while np.mod(kernel.shape[0] - np.sum(drop) - 1 + np.mod(ds_factor + 1, 2), ds_factor) != 0:
        drop[side] -= 1
        side = (side + 1) % 2
    kernel = kernel[drop[0]:-drop[1], drop[0]:-drop[1]]
    return kernel / np.sum(kernel)


def upscale_kernel_from_blur(kernel, sf):
    """imresize(None, [sf, sf], return_upscale_kernel=True, kernel=<ndarray>, kernel_indx>=8):
    centre the supplied blur kernel, scale by sf^2 and pad for the uneven pre/post strides."""
    if abs(1 - np.sum(kernel)) >= np.finfo(np.float32).eps:
        raise AssertionError("Supplied non-default kernel does not sum to 1")
    pre, post = calc_strides(None, sf)
    pad_after, pad_before = np.maximum(0, pre - post), np.maximum(0, post - pre)
    k = Center_Mass(kernel, ds_factor=sf) * sf ** 2
    if np.any(np.mod(np.array(k.shape) + pad_after + pad_before - 1, sf) != 0):
        raise AssertionError("Convolution-invalidated size should be an integer multiplication of sf")
    k = np.pad(k, ((pad_before[0], pad_after[0]), (pad_before[1], pad_after[1])), mode="constant")
    return k, pre, post
