"""Separable-SVD bicubic degradation operator (DDRM-style `SRConv`) on the B200.

API of the reference's guided_diffusion/restore_util.py: `A_functions` (:11-100) and `SRConv`
(:102-227) with `A`, `At`, `A_pinv`, `A_pinv_eta` acting on flattened (N, C*H*W) batches.  The
reference evaluates each as up to eight batched 512x512 matmuls plus scatter/gather index ops; here
the singular values are folded into two small dense factors at construction,

    A(x)      = M X M^T,        M = U diag(s) V1^T            (small x img)
    A^+(y)    = P Y P^T,        P = V1 diag(s^+) U^T          (img x small)
    A^T(y)    = M^T Y M

(V1 = first `small` columns of V; the reference's permutation only selects that block,
:142-160,188-197) and every product is one flair_sandwich_f32 launch.  The SVD itself is computed
with the same `torch.svd(some=False)` call on the host (:134-136) at construction."""
from __future__ import annotations

import torch

from flair_b200 import ops


class A_functions:
    """Interface of the reference (:11-100); only what FLAIR uses is implemented by SRConv."""

    def A(self, vec):
        raise NotImplementedError()

    def At(self, vec):
        raise NotImplementedError()

    def A_pinv(self, vec):
        raise NotImplementedError()

    def A_pinv_eta(self, vec, eta):
        raise NotImplementedError()


class SRConv(A_functions):
    def __init__(self, kernel, channels, img_dim, device, stride=1):
        self.img_dim, self.channels, self.ratio = img_dim, channels, stride
        small = img_dim // stride
        self.y_dim = small
        taps = kernel.detach().float().cpu()
        half = taps.shape[0] // 2
        A_small = torch.zeros(small, img_dim)
        for i in range(stride // 2, img_dim + stride // 2, stride):  # reflect-without-repeat borders (:121-132)
            for j in range(i - half, i + half):
                je = -j - 1 if j < 0 else ((img_dim - 1) - (j - img_dim) if j >= img_dim else j)
                A_small[i // stride, je] += taps[j - i + half]
        U, s, V = torch.svd(A_small, some=False)
        s[s < 3e-2] = 0
        self.U_small, self.singulars_small, self.V_small = U.to(device), s.to(device), V.to(device)
        self._singulars = torch.outer(s, s).reshape(small ** 2).to(device)
        s_inv = torch.where(s == 0, torch.zeros_like(s), 1.0 / s)
        V1 = V[:, :small]
        self._M = ((U * s[None, :]) @ V1.t()).contiguous().to(device)          # small x img
        self._Mt = self._M.t().contiguous()
        self._P = ((V1 * s_inv[None, :]) @ U.t()).contiguous().to(device)      # img x small
        self._Pt = self._P.t().contiguous()
        self._eta_cache = {}

    def _img(self, vec, dim):
        return vec.reshape(vec.shape[0] * self.channels, dim, dim)

    def singulars(self):
        return self._singulars.repeat_interleave(3).reshape(-1)

    def A(self, vec):
        return ops.sandwich(self._M, self._img(vec, self.img_dim), self._Mt).reshape(vec.shape[0], -1)

    def At(self, vec):
        return ops.sandwich(self._Mt, self._img(vec, self.y_dim), self._M).reshape(vec.shape[0], -1)

    def A_pinv(self, vec):
        return ops.sandwich(self._P, self._img(vec, self.y_dim), self._Pt).reshape(vec.shape[0], -1)

    def A_pinv_eta(self, vec, eta):
        """V1 (s_i s_j / ((s_i s_j)^2 + eta)) .* (U^T Y U) V1^T — not separable, so the spectral weights
        are applied between the two sandwiches (reference :86-96)."""
        spec = ops.sandwich(self.U_small.t().contiguous(), self._img(vec, self.y_dim), self.U_small)
        ss = self._singulars.reshape(self.y_dim, self.y_dim)
        spec = spec * (ss / (ss * ss + eta))
        V1 = self.V_small[:, : self.y_dim].contiguous()
        return ops.sandwich(V1, spec, V1.t().contiguous()).reshape(vec.shape[0], -1)

    def restore(self, x, y):
        """bicubic_restore of scripts/video_sample.py:177-181, A^+(A x - y), as two launches pairs."""
        z = ops.sandwich(self._M, x.reshape(-1, self.img_dim, self.img_dim), self._Mt,
                         sub=y.reshape(-1, self.y_dim, self.y_dim))
        return ops.sandwich(self._P, z, self._Pt).reshape(x.shape)
