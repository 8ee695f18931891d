"""DCT-domain JPEG (encode to quantised coefficient planes / decode) as one fused kernel per call.

API of the reference's guided_diffusion/jpeg.py: `jpeg_encode(x, qf)` (:72-112), `jpeg_decode(x, qf)`
(:117-167), `general_quant_matrix` / `quantization_matrix` (:35-69).  The reference runs ~20 small
ops per call (Unfold/Fold, two LinearDCT matmuls, table rebuilds); here colour transform, 4:2:0
decimation, 8x8 DCT, quantise+round, dequantise, IDCT, chroma up-sampling and the inverse colour
transform happen in shared memory, one CTA per 16x16 macroblock (flair_jpeg_f32)."""
from __future__ import annotations

import torch

from flair_b200 import ops

from .dct import linear_dct_weight

_LUMA = [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
         14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
         49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99]
_CHROMA = [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
           47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32


def general_quant_matrix(qf=10):
    """IJG quality scaling of the standard tables (reference :35-65); returns flat (64,) tensors."""
    s = (5000 / qf) if qf < 50 else (200 - 2 * qf)
    tabs = []
    for base in (_LUMA, _CHROMA):
        q = torch.floor((s * torch.tensor(base) + 50) / 100)
        q[q <= 0] = 1
        q[q > 255] = 255
        tabs.append(q)
    return tabs[0], tabs[1]


def quantization_matrix(qf):
    return general_quant_matrix(qf)


_tables = {}


def _device_tables(qf, device):
    key = (qf, str(device))
    if key not in _tables:
        q1, q2 = general_quant_matrix(qf)
        _tables[key] = tuple(t.float().contiguous().to(device) for t in (
            linear_dct_weight(8, "dct"), linear_dct_weight(8, "idct"), q1, q2))
    return _tables[key]


def jpeg_encode(x, qf):
    """(N,3,h,w) in [-1,1] -> [luma (N,1,h,w), chroma (N,2,h/2,w/2)] quantised DCT coefficients."""
    return ops.jpeg(0, _device_tables(qf, x.device), x=x)


def jpeg_decode(x, qf):
    """[luma, chroma] coefficient planes -> (N,3,h,w) in [-1,1]."""
    return ops.jpeg(1, _device_tables(qf, x[0].device), planes=x)


def jpeg_roundtrip(x, qf):
    """jpeg_decode(jpeg_encode(x, qf), qf) without materialising the coefficient planes."""
    return ops.jpeg(2, _device_tables(qf, x.device), x=x)
