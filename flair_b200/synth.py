"""Deterministic synthetic inputs and weights.

The Google-Drive checkpoints and any real face video are unavailable offline,
so tests, the bench and the golden-vector generator all draw from here.  Every
tensor is generated on the CPU from an explicit seed, so the reference (CPU),
the oracle (CPU) and the CUDA path see bit-identical inputs.

Weights are a pure function of (seed, state-dict key, shape): they do not depend
on module construction order, so the same call produces the same state dict for
the reference model, the oracle and this repo's model.  Zero-initialised
tensors of the reference (`zero_module`, nn_new.py:68-74) are re-randomised as
well, otherwise a fresh UNet outputs exactly 0 and parity tests are vacuous
(SURVEY D10).
"""
from __future__ import annotations

import zlib

import torch
import torch.nn.functional as F


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synthetic_tensor(key: str, shape, seed: int = 1234) -> torch.Tensor:
    """Variance-preserving random tensor for one state-dict entry."""
    shape = tuple(shape)
    if "spynet." in key:
        # sr3 registers ONE shared SPyNet under every BasicVSR++ module (sr3.py:340,382): all copies of a
        # key must hold the same values, like in a real checkpoint
        key = key[key.index("spynet."):]
    g = _gen(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("spynet.mean"):
        return torch.tensor([0.485, 0.456, 0.406]).view(shape)
    if key.endswith("spynet.std"):
        return torch.tensor([0.229, 0.224, 0.225]).view(shape)
    if len(shape) == 1:
        v = torch.randn(shape, generator=g)
        if leaf == "weight":  # GroupNorm gamma
            return 1.0 + 0.1 * v
        return 0.05 * v  # biases / GroupNorm beta
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 1.0
    # keep the residual branches and the flow/offset heads tame
    if ".out_layers." in key or "proj_out" in key or ".proj." in key or "conv_last" in key:
        gain = 0.5
    if "conv_offset.6" in key or "spynet" in key:
        gain = 0.25
    return torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)


def synthetic_state_dict(model: torch.nn.Module, seed: int = 1234) -> dict:
    """State dict with every entry replaced by `synthetic_tensor` (dtype preserved)."""
    sd = {}
    for k, v in model.state_dict().items():
        sd[k] = synthetic_tensor(k, v.shape, seed).to(v.dtype)
    return sd


def synthetic_clip(n_frames: int, size: int = 256, seed: int = 1) -> torch.Tensor:
    """(N,3,size,size) fp32 in [0,1]: low-pass noise translated by (2t, t) px per frame."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    base = torch.rand(1, 3, size, size, generator=g)
    box = torch.ones(3, 1, 9, 9) / 81.0
    for _ in range(2):
        base = F.conv2d(F.pad(base, (4, 4, 4, 4), mode="circular"), box, groups=3)
    base = (base - base.amin()) / (base.amax() - base.amin())
    frames = [torch.roll(base[0], shifts=(t, 2 * t), dims=(1, 2)) for t in range(n_frames)]
    return torch.stack(frames, 0).contiguous()


def noise_tape(shape, n_steps: int, seed: int = 2) -> torch.Tensor:
    """(n_steps+1, *shape) standard normals: entry 0 is the q_sample noise, entry 1+i the
    noise drawn at the i-th executed sampling step."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n_steps + 1, *shape, generator=g)
