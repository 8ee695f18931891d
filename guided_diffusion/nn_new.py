"""Primitive layers of the blur/JPEG UNet as *parameter holders* (reference guided_diffusion/
nn_new.py).  The reference's modules compute with ATen; here they only own the tensors under the
reference's state-dict names — the arithmetic runs in flair_b200 kernels driven by unet_new.py."""
from __future__ import annotations

import math

import torch as th
import torch.nn as nn

from flair_b200 import ops


class SiLU(nn.Module):
    """Marker only: SiLU is fused into the GroupNorm-apply kernel."""


class GroupNorm32(nn.GroupNorm):
    """fp32-statistics GroupNorm (reference :17-19); executed by flair_gn_stats / flair_gn_apply."""


def conv_nd(dims, *args, **kwargs):
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    if dims == 2:
        return nn.Conv2d(*args, **kwargs)
    if dims == 3:
        return nn.Conv3d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims}")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def avg_pool_nd(dims, *args, **kwargs):
    return {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d}[dims](*args, **kwargs)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def scale_module(module, scale):
    for p in module.parameters():
        p.detach().mul_(scale)
    return module


def mean_flat(tensor):
    return tensor.mean(dim=list(range(1, len(tensor.shape))))


def normalization(channels):
    return GroupNorm32(32, channels)


_freq_cache = {}


def _freqs(half, max_period, device):
    key = (half, max_period, str(device))
    if key not in _freq_cache:
        # built on the host with the reference's exact fp32 expression (:114-116), uploaded once
        f = th.exp(-math.log(max_period) * th.arange(start=0, end=half, dtype=th.float32) / half)
        _freq_cache[key] = f.to(device)
    return _freq_cache[key]


def timestep_embedding(timesteps, dim, max_period=10000):
    """[N] timesteps -> [N, dim] (cos | sin) embeddings (reference :103-121)."""
    emb = ops.timestep_embedding(timesteps, _freqs(dim // 2, max_period, timesteps.device))
    if dim % 2:
        emb = th.cat([emb, th.zeros_like(emb[:, :1])], dim=-1)
    return emb


def checkpoint(func, inputs, params, flag):
    """Inference-only build: gradient checkpointing degenerates to a plain call (reference :124-139)."""
    return func(*inputs)
