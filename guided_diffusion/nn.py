"""Wrapper modules that contribute the `wrapped_module` path segment of the reference's state-dict
names (reference guided_diffusion/nn.py:341-367) and the flash-attn shim name.  They own nothing
but their child; layout changes ((b n) folding, b t c h w <-> b c t h w) do not exist here because
activations are channels-last [B,T,H,W,C] from the first kernel to the last."""
from __future__ import annotations

import torch.nn as nn


class PlaceHolder(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.wrapped_module = module


class LazyReshaper2D(PlaceHolder):
    """Reference: applies a 2-D op frame by frame."""


class LazyReshaper3D(PlaceHolder):
    """Reference: applies an op over (T,H,W) of each batch element."""


class FalshAttn(nn.Module):
    """Name kept for module-tree parity; the temporal window attention is flair_attn_temporal."""
