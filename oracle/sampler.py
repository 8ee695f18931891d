"""FLAIR sampling step and loop on the CPU (torch fp32), restating
guided_diffusion/gaussian_diffusion.py:250-342 (p_mean_variance, EPSILON model,
6->3 channel split), :344-365 (x0<->eps), :423-517 (p_sample: data consistency,
prev_recon overwrite, rho-mixed update), :589-689 (loop) with the aux face prior
disabled (tau = num_timesteps) and respace.py:155-167 (timestep
mapping / SR3 noise level)."""
from __future__ import annotations

import numpy as np
import torch

from .schedule import Tables


def _c(arr, t):  # gaussian_diffusion.py:692-705: float64 table -> fp32 scalar
    return torch.tensor(arr[t], dtype=torch.float64).float()


def p_sample_step(tab: Tables, x_t, model_out, t: int, noise, restore_fn=None, gamma=1.0,
                  rho=0.35, prev_recon=None, num_frames=None):
    """One x_t -> x_{t-1} update.  model_out may carry 6 channels (eps, var): only eps is used."""
    eps = model_out[:, :3]
    a, b = _c(tab.sqrt_recip_alphas_cumprod, t), _c(tab.sqrt_recipm1_alphas_cumprod, t)
    x0 = (a * x_t - b * eps).clamp(-1, 1)  # :344-349, :311-314
    if restore_fn is not None:  # :465-470
        g = torch.tensor(float(gamma), dtype=torch.float64).float()
        x0 = (x0 - g * restore_fn(x0)).clamp(-1, 1)
    if prev_recon is not None:  # :497-506  (prev_recon: (b, k, 3, H, W))
        x0 = x0.reshape(-1, num_frames, *x0.shape[1:]).clone()
        x0[:, : prev_recon.shape[1]] = prev_recon
        x0 = x0.reshape(-1, *x0.shape[2:])
    eps_hat = (a * x_t - x0) / b  # :361-365
    c = _c(tab.sqrt_alphas_cumprod_prev, t)
    d = _c(tab.sqrt_one_minus_alphas_cumprod_prev, t)
    nz = 0.0 if t == 0 else 1.0
    s1, s2 = float(np.sqrt(1 - rho)), float(np.sqrt(rho))  # python floats times fp32 tensors (:514)
    sample = c * x0 + nz * (s1 * d * eps_hat + s2 * d * noise)  # :508-515
    return sample, x0


def model_time_input(tab: Tables, t: int, sr3: bool):
    """respace.py:155-167: blur UNet gets the original timestep index, SR3 a noise level."""
    if sr3:
        return float(np.float32(tab.sqrt_alphas_cumprod_prev[t + 1]))
    return tab.timestep_map[t]


def sample_loop(tab: Tables, model_fn, x_T, tape, restore_fn=None, rho=0.35, zeta=-1, noise_level=None,
                prev_recon=None, num_frames=None, t_start=-1, sr3=False):
    """model_fn(x, t_input) -> (N, 3|6, H, W).  tape[i] is the noise of the i-th executed step."""
    gam = tab.gammas(zeta, noise_level)
    steps = list(range(tab.num_timesteps))
    if t_start != -1:
        steps = steps[: t_start + 1]
    x = x_T
    for i, t in enumerate(steps[::-1]):
        out = model_fn(x, model_time_input(tab, t, sr3))
        x, _ = p_sample_step(tab, x, out, t, tape[i], restore_fn, gam[t], rho, prev_recon, num_frames)
    return x


def q_sample(tab: Tables, x0, t: int, noise):
    # gaussian_diffusion.py:206-224
    return _c(tab.sqrt_alphas_cumprod, t) * x0 + _c(tab.sqrt_one_minus_alphas_cumprod, t) * noise
