"""Data-consistency operators on the CPU (torch fp32).

G2  blur + x4 pseudo-inverse  — guided_diffusion/pseudoSR.py:180-281 (three depth-wise
    `Filter_Layer`s with replication padding, A_pinv) called through
    scripts/video_sample.py:183-193 (gaussian_restore).
J1  DCT-domain JPEG          — guided_diffusion/jpeg.py:7-167, dct.py:167-202.
B2  separable bicubic SRConv — guided_diffusion/restore_util.py:54-82,102-227 called
    through scripts/video_sample.py:177-181 (bicubic_restore).

The filter taps / DCT matrix / SVD factors are *inputs* here (they are host-side
setup): tests feed the ones dumped from the reference
(tests/golden) and, separately, check the product's own host-side preparation
against those dumps.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- G2


def _dw(x, k):
    """depth-wise cross-correlation with a (kh,kw) fp32 kernel, no padding (pseudoSR.py:15-44)."""
    c = x.shape[1]
    return F.conv2d(x, k[None, None].expand(c, 1, *k.shape).contiguous(), groups=c)


def blur_down(x, ds_kernel, sf=4):
    """DownscaleOP (pseudoSR.py:226-244): replicate-pad, 9x9 with rot180(ds_kernel), keep phase pre_stride."""
    k = torch.flip(ds_kernel, (0, 1))
    p = ds_kernel.shape[0] // 2
    y = _dw(F.pad(x, (p, p, p, p), mode="replicate"), k)
    pre = sf - sf // 2 - 1  # calc_strides, imresize_pseudoSR.py:72-74 -> pre_stride = 1 for sf 4
    return y[:, :, pre::sf, pre::sf]


def inv_hth(y, inv_k):
    """Conv_LR_with_Inv_hTh_OP (pseudoSR.py:180-195)."""
    p = inv_k.shape[0] // 2
    return _dw(F.pad(y, (p, p, p, p), mode="replicate"), inv_k)


def up_blur(y, ds_kernel, sf=4):
    """Upscale_OP (pseudoSR.py:196-225): zero-insert with the sample at phase pre_stride, replicate pad,
    9x9 with ds_kernel * sf^2."""
    n, c, h, w = y.shape
    pre = sf - sf // 2 - 1
    z = torch.zeros(n, c, h * sf, w * sf, dtype=y.dtype)
    z[:, :, pre::sf, pre::sf] = y
    p = ds_kernel.shape[0] // 2
    return _dw(F.pad(z, (p, p, p, p), mode="replicate"), ds_kernel * float(sf * sf))


def blur_restore(x, y_lr, ds_kernel, inv_k, sf=4, jpeg_qf=-1, dct_mat=None):
    """A_pinv(LR, x) = Up(InvhTh(codec(Down(x)))) - Up(InvhTh(LR))   (pseudoSR.py:248-277)."""
    lr = blur_down(x, ds_kernel, sf)
    if jpeg_qf != -1:
        lr = jpeg_decode(jpeg_encode(lr, jpeg_qf, dct_mat), jpeg_qf, dct_mat)
    return up_blur(inv_hth(lr, inv_k), ds_kernel, sf) - up_blur(inv_hth(y_lr, inv_k), ds_kernel, sf)


# ----------------------------------------------------------------------------- J1

_Q_LUMA = [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
           14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
           49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99]
_Q_CHROMA = [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
             47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32
# jpeg.py:9-11 and :19-25 (the decode matrix is the numerical inverse of the rounded encode matrix)
_RGB2YCC = [[0.299, 0.587, 0.114], [-0.1687, -0.3313, 0.5], [0.5, -0.4187, -0.0813]]
_YCC2RGB = [[1.00000000e00, -3.68199903e-05, 1.40198758e00],
            [1.00000000e00, -3.44113281e-01, -7.14103821e-01],
            [1.00000000e00, 1.77197812e00, -1.34583413e-04]]


def quant_tables(qf):
    """jpeg.py:35-65."""
    s = (5000 / qf) if qf < 50 else (200 - 2 * qf)
    out = []
    for q in (_Q_LUMA, _Q_CHROMA):
        t = torch.floor((s * torch.tensor(q) + 50) / 100)
        t[t <= 0] = 1
        t[t > 255] = 255
        out.append(t.reshape(8, 8))
    return out


def dct8_matrix():
    """The 8x8 ortho DCT-II matrix exactly as LinearDCT builds it (dct.py:31-60,180-191): FFT-based
    dct() of the identity, transposed.  D[k, n]; forward = D @ x."""
    N = 8
    x = torch.eye(N)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)
    Vc = torch.view_as_real(torch.fft.fft(v, dim=1))
    k = -torch.arange(N, dtype=x.dtype)[None, :] * np.pi / (2 * N)
    V = Vc[:, :, 0] * torch.cos(k) - Vc[:, :, 1] * torch.sin(k)
    V[:, 0] /= np.sqrt(N) * 2
    V[:, 1:] /= np.sqrt(N / 2) * 2
    return (2 * V).t().contiguous()  # weight = dct(I).t(); linear(x) = x @ weight.T


def _blocks(p):  # (N,C,H,W) -> (N,C,H/8,W/8,8,8)
    n, c, h, w = p.shape
    return p.reshape(n, c, h // 8, 8, w // 8, 8).permute(0, 1, 2, 4, 3, 5)


def _unblocks(b):
    n, c, bh, bw, _, _ = b.shape
    return b.permute(0, 1, 2, 4, 3, 5).reshape(n, c, bh * 8, bw * 8)


def jpeg_encode(x, qf, D=None, idct_mat=None):
    """jpeg.py:72-112: quantised DCT coefficient planes [luma (N,1,h,w), chroma (N,2,h/2,w/2)]."""
    D = dct8_matrix() if D is None else D
    x = (x + 1) / 2 * 255
    m = torch.tensor(_RGB2YCC)
    ycc = torch.einsum("nchw,kc->nkhw", x, m).clone()
    ycc[:, 1:] += 128
    planes = [ycc[:, :1], ycc[:, 1:, ::2, ::2]]
    out = []
    for p, q in zip(planes, quant_tables(qf)):
        b = _blocks(p) - 128
        # apply_linear_2d (dct.py:194-202): rows then columns with weight D: X = D b D^T
        c = torch.matmul(torch.matmul(b, D.t()).transpose(-1, -2), D.t()).transpose(-1, -2)
        out.append(_unblocks(torch.round(c / q)))
    return out


def idct8_matrix():
    """LinearDCT(8,'idct','ortho') weight (dct.py:63-104,188-189) = idct(I).t(); the inverse of dct8."""
    N = 8
    X = torch.eye(N)
    X_v = X.clone() / 2
    X_v[:, 0] *= np.sqrt(N) * 2
    X_v[:, 1:] *= np.sqrt(N / 2) * 2
    k = torch.arange(N, dtype=X.dtype)[None, :] * np.pi / (2 * N)
    W_r, W_i = torch.cos(k), torch.sin(k)
    V_t_r = X_v
    V_t_i = torch.cat([X_v[:, :1] * 0, -X_v.flip([1])[:, :-1]], dim=1)
    V_r = V_t_r * W_r - V_t_i * W_i
    V_i = V_t_r * W_i + V_t_i * W_r
    V = torch.complex(V_r, V_i)
    v = torch.fft.irfft(V, n=N, dim=1)
    x = v.new_zeros(v.shape)
    x[:, ::2] += v[:, : N - (N // 2)]
    x[:, 1::2] += v.flip([1])[:, : N // 2]
    return x.t().contiguous()


def jpeg_decode(planes, qf, D=None, Di=None):
    """jpeg.py:117-167."""
    Di = idct8_matrix() if Di is None else Di
    rec = []
    for p, q in zip(planes, quant_tables(qf)):
        b = _blocks(p) * q
        c = torch.matmul(torch.matmul(b, Di.t()).transpose(-1, -2), Di.t()).transpose(-1, -2)
        rec.append(_unblocks(c + 128))
    luma, chroma = rec
    chroma = chroma.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)  # nearest up
    ycc = torch.cat([luma, chroma], dim=1).clone()
    ycc[:, 1:] -= 128
    rgb = torch.einsum("nchw,kc->nkhw", ycc, torch.tensor(_YCC2RGB))
    return rgb / 255 * 2 - 1


# ----------------------------------------------------------------------------- B2


def bicubic_taps(factor):
    """scripts/video_sample.py:208-223: 4f-tap 1-D bicubic kernel (a=-0.5), normalised twice."""
    def cub(x, a=-0.5):
        x = abs(x)
        if x <= 1:
            return (a + 2) * x ** 3 - (a + 3) * x ** 2 + 1
        if 1 < x < 2:
            return a * x ** 3 - 5 * a * x ** 2 + 8 * a * x - 4 * a
        return 0
    k = np.zeros(factor * 4)
    for i in range(factor * 4):
        k[i] = cub((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
    k = k / np.sum(k)
    k = torch.from_numpy(k).float()
    return k / k.sum()


def srconv_matrix(taps, img_dim, stride):
    """restore_util.py:113-132: 1-D strided conv matrix with reflect-without-repeat borders."""
    small = img_dim // stride
    A = torch.zeros(small, img_dim)
    half = taps.shape[0] // 2
    for i in range(stride // 2, img_dim + stride // 2, stride):
        for j in range(i - half, i + half):
            je = j
            if je < 0:
                je = -je - 1
            if je >= img_dim:
                je = (img_dim - 1) - (je - img_dim)
            A[i // stride, je] += taps[j - i + half]
    return A


def srconv_restore(x, y, U, S, V):
    """bicubic_restore = A_pinv(A x - y) (scripts/video_sample.py:177-181) following the SVD route of
    restore_util.py:54-82,162-227 per channel image X (img x img), Y (small x small):
       A X      = U diag(S) V1^T X V1 diag(S) U^T          (V1 = first `small` columns of V)
       A^+ Z    = V1 diag(1/S) U^T Z U diag(1/S) V1^T
    (the permutation / zero padding of the reference only selects the top-left small x small block).
    x: (N,3,img,img), y: (N,3,small,small); U (small,small), S (small,), V (img,img)."""
    small = U.shape[0]
    V1 = V[:, :small]
    ss = S[:, None] * S[None, :]
    t = V1.t() @ x @ V1                  # Vt + permutation
    ax = U @ (ss * t) @ U.t()            # singulars * temp, then U
    z = U.t() @ (ax - y) @ U             # Ut
    inv = torch.where(ss == 0, torch.zeros_like(ss), 1.0 / ss)
    return V1 @ (z * inv) @ V1.t()       # add_zeros + V
