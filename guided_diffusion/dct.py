"""Host-side DCT matrices for the JPEG operator.

The reference builds its 8x8 transform as a `LinearDCT` layer whose weight is the FFT-based
`dct(I, norm="ortho")` transposed (guided_diffusion/dct.py:31-60,167-191).  A quantiser `round()`
follows, so the matrix is reproduced the same way (FFT route, fp32) rather than from an analytic
cosine table; tests compare it bit-for-bit with the reference's weight."""
from __future__ import annotations

import numpy as np
import torch


def dct(x, norm=None):
    """DCT-II over the last dimension via an N-point FFT of the even/odd-reordered signal."""
    shape = x.shape
    N = shape[-1]
    x = x.contiguous().view(-1, N)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)
    Vc = torch.view_as_real(torch.fft.fft(v, dim=1))
    k = -torch.arange(N, dtype=x.dtype, device=x.device)[None, :] * np.pi / (2 * N)
    V = Vc[:, :, 0] * torch.cos(k) - Vc[:, :, 1] * torch.sin(k)
    if norm == "ortho":
        V[:, 0] /= np.sqrt(N) * 2
        V[:, 1:] /= np.sqrt(N / 2) * 2
    return 2 * V.view(*shape)


def idct(X, norm=None):
    """Inverse of `dct` (scaled DCT-III) over the last dimension."""
    shape = X.shape
    N = shape[-1]
    Xv = X.contiguous().view(-1, N) / 2
    if norm == "ortho":
        Xv[:, 0] *= np.sqrt(N) * 2
        Xv[:, 1:] *= np.sqrt(N / 2) * 2
    k = torch.arange(N, dtype=X.dtype, device=X.device)[None, :] * np.pi / (2 * N)
    wr, wi = torch.cos(k), torch.sin(k)
    tr = Xv
    ti = torch.cat([Xv[:, :1] * 0, -Xv.flip([1])[:, :-1]], dim=1)
    V = torch.complex(tr * wr - ti * wi, tr * wi + ti * wr)
    v = torch.fft.irfft(V, n=N, dim=1)
    x = v.new_zeros(v.shape)
    x[:, ::2] += v[:, : N - (N // 2)]
    x[:, 1::2] += v.flip([1])[:, : N // 2]
    return x.view(*shape)


def linear_dct_weight(n, kind, norm="ortho"):
    """Weight of the reference's LinearDCT(n, kind, norm): y = x @ W^T."""
    eye = torch.eye(n)
    if kind == "dct":
        return dct(eye, norm=norm).t().contiguous()
    if kind == "idct":
        return idct(eye, norm=norm).t().contiguous()
    raise NotImplementedError(kind)
