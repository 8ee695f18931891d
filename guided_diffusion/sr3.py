"""SR3-style UNet of the x8 / x16 bicubic tasks (reference guided_diffusion/sr3.py) — boundary class.

Round-1 status: the class exists so that `respace._WrappedModel` can dispatch on it; the forward is
not implemented yet ("next")."""
import torch.nn as nn


class UNet(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("sr3.UNet (x8/x16 bicubic denoiser) is scheduled after the blur/JPEG UNet")
