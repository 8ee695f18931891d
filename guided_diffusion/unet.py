"""Cross-frame modules of the bicubic (SR3) denoiser — names of the reference's guided_diffusion/unet.py
that `sr3.py` imports (:113-254 ResBlock with a (3,1,1) temporal kernel, :313-595 BasicVSRPP that owns the
shared SPyNet and resizes `lqs`, :598-661 SecondOrderDeformableAlignment, :664-758 TemporalAttention).
They are the same kernels as the blur UNet's modules (guided_diffusion/unet_new.py); only construction
details differ.  The dead `CrossFrameUNetModel` / `AttentionBlock` of the reference file are not provided."""
from __future__ import annotations

import torch.nn as nn

from . import unet_new as _u
from .unet_new import (SecondOrderDeformableAlignment, SPyNet, TemporalAttention, TemporalWrapper, TimestepBlock,  # noqa: F401
                       convert_module_to_f16, convert_module_to_f32)


class ResBlock(_u.ResBlock):
    """unet.ResBlock (reference :113-254): same block, constructor takes the conv kernel / padding; sr3 uses
    dims=3 with kernel (3,1,1), plain `h + emb` conditioning and 32-group norms."""

    def __init__(self, channels, emb_channels, dropout, kernel_size=3, padding=1, stride=1, padding_mode="zeros",
                 out_channels=None, use_conv=False, use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False,
                 down=False, norm_type="group_norm", win_size=5):
        if stride != 1 or padding_mode != "zeros" or norm_type != "group_norm":
            raise NotImplementedError("only the FLAIR configuration (stride 1, zero padding, group norm) is supported")
        super().__init__(channels, emb_channels, dropout, out_channels=out_channels, use_conv=use_conv,
                         use_scale_shift_norm=use_scale_shift_norm, dims=dims, use_checkpoint=use_checkpoint, up=up,
                         down=down, kernel_size=kernel_size, padding=padding)


class BasicVSRPP(_u.BasicVSRPP):
    """unet.BasicVSRPP (reference :313-595): registers the shared SPyNet under every instance (so the state
    dict repeats it, like the reference) — flows are computed once per window by sr3.UNet and cached."""

    def __init__(self, mid_channels=64, max_residue_magnitude=10, shared_spynet=None, use_checkpoint=False):
        super().__init__(mid_channels=mid_channels, max_residue_magnitude=max_residue_magnitude,
                         use_checkpoint=use_checkpoint)
        self.spynet = shared_spynet if shared_spynet is not None else SPyNet(pretrained=None)
