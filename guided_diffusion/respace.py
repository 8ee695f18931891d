"""Timestep respacing (reference guided_diffusion/respace.py:7-167): which of the original
1000/2000 steps are kept, betas re-derived for the kept ones, and the model wrapper that maps the
respaced index back to what each UNet family expects."""
from __future__ import annotations

import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts, mode="uniform"):
    """Set of original steps to keep (reference :7-66).  "100" -> 100 evenly spread steps;
    "ddimN" -> integer stride; comma list -> per-section counts; mode "quad" -> quadratic list."""
    if mode == "quad":
        seq = np.linspace(0, np.sqrt(num_timesteps * 0.8), int(section_counts)) ** 2
        return [int(s) for s in seq]
    if mode != "uniform":
        return None
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for sec, count in enumerate(section_counts):
        size = base + (1 if sec < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):  # accumulate like the reference: float additions, python round()
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """Diffusion restricted to `use_timesteps` of a base process (reference :78-135)."""

    def __init__(self, use_timesteps, noise_schedule="linear", **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.noise_schedule = noise_schedule
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        self._wrapped = {}

    def _run_model(self, model, *args, **kwargs):
        return super()._run_model(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        wrapped = self._wrapped.get(id(model))
        if wrapped is None or wrapped.model is not model:
            wrapped = _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps,
                                    noise_schedule=self.noise_schedule,
                                    sqrt_alphas_cumprod_prev=self.sqrt_alphas_cumprod_prev)
            self._wrapped = {id(model): wrapped}
        return wrapped

    def _scale_timesteps(self, t):
        return t  # done by the wrapper


class _WrappedModel:
    """Respaced index -> model conditioning (reference :138-167): the blur/JPEG UNet receives the
    original timestep, the SR3 UNet the continuous noise level sqrt(alpha_bar_{t}) (table shifted by
    one).  Lookup tables live on the device (the reference rebuilds them from python lists per call)."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, noise_schedule="linear",
                 sqrt_alphas_cumprod_prev=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self.noise_schedule = noise_schedule
        self.sqrt_alphas_cumprod_prev = sqrt_alphas_cumprod_prev
        self._dev = {}

    def _tables(self, device):
        tabs = self._dev.get(str(device))
        if tabs is None:
            tabs = (th.tensor(self.timestep_map, device=device, dtype=th.long),
                    th.from_numpy(np.asarray(self.sqrt_alphas_cumprod_prev)).to(device, th.float32))
            self._dev = {str(device): tabs}
        return tabs

    def __call__(self, x, ts, **kwargs):
        from .sr3 import UNet as SR3_UNet
        tmap, levels = self._tables(ts.device)
        kwargs["old_ts"] = ts
        if isinstance(self.model, SR3_UNet):
            return self.model(x, levels[ts + 1], **kwargs)
        new_ts = tmap[ts].to(ts.dtype)
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
