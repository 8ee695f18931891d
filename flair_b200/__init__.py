"""flair_b200 — B200-native (sm_100a) kernels behind the FLAIR hot path.

`guided_diffusion/` (repo root) is the drop-in Python boundary that mirrors the
reference's API; this package holds the CUDA kernels, their C ABI binding and
the executor that strings them into a UNet forward / sampler step.
"""
__all__ = ["build", "ops"]
