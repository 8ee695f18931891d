"""Diffusion schedule tables (float64 numpy), restating
guided_diffusion/gaussian_diffusion.py:15-36,112-173 (named betas, coefficient tables),
guided_diffusion/respace.py:7-61,87-102 (uniform respacing, re-derived betas) and
gaussian_diffusion.py:648-657 (data-consistency weights gamma_t)."""
from __future__ import annotations

import numpy as np


def named_betas(name: str, n: int) -> np.ndarray:
    # gaussian_diffusion.py:24-34
    if name == "face_blur":
        s = 1000 / n
        return np.linspace(s * 1e-4, s * 2e-2, n, dtype=np.float64)
    if name == "face_bicubic":
        return np.linspace(1e-6, 1e-2, 2000, dtype=np.float64)
    raise NotImplementedError(name)


def kept_steps(n: int, count: int) -> list:
    # respace.py:40-61 with a single section: stride (n-1)/(count-1), python round()
    if count <= 1:
        return [0]
    stride = (n - 1) / (count - 1)
    cur, out = 0.0, []
    for _ in range(count):
        out.append(round(cur))
        cur += stride
    return sorted(set(out))


class Tables:
    """Respaced coefficient tables for `count` sampling steps of an n-step process."""

    def __init__(self, name: str, n: int, count: int = 100):
        base_ac = np.cumprod(1.0 - named_betas(name, n))
        self.timestep_map = kept_steps(n, count)
        betas, last = [], 1.0
        for i in self.timestep_map:  # respace.py:94-101
            betas.append(1 - base_ac[i] / last)
            last = base_ac[i]
        self.betas = np.array(betas, dtype=np.float64)
        ac = np.cumprod(1.0 - self.betas)  # gaussian_diffusion.py:134-149
        self.num_timesteps = len(betas)
        self.alphas_cumprod = ac
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.sqrt_alphas_cumprod_prev = np.sqrt(np.append(1.0, ac))  # length T+1
        self.sqrt_one_minus_alphas_cumprod_prev = np.append(0.0, np.sqrt(1.0 - ac[:-1]))
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)

    def gammas(self, zeta: float, noise_level) -> np.ndarray:
        # gaussian_diffusion.py:648-657 (noise_level is used raw, not /255)
        if zeta == -1:
            return np.ones_like(self.betas)
        g = zeta * (noise_level ** 2 / (self.sqrt_one_minus_alphas_cumprod / self.sqrt_alphas_cumprod) ** 2)
        g[g >= 1] = 0.991
        g[g <= 1e-1] = 1e-6
        return 1 - g
