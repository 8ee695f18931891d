"""FLAIR diffusion sampler behind the reference API, executed by fused sm_100a kernels.

Mirrors (names, keyword lists, return conventions) guided_diffusion/gaussian_diffusion.py of the
reference: `get_named_beta_schedule` (:15-36), the enums (:59-92), `GaussianDiffusion`
(:95-689: q_sample, q_posterior_mean_variance, p_mean_variance, p_sample, p_sample_loop,
p_sample_loop_progressive, sample) and `_extract_into_tensor` (:692-705).

What differs, by design:
  * coefficient tables are uploaded once per device as one fp32 table (the float64 -> fp32 cast is
    the one the reference performs after indexing, :702) and indexed *inside* the kernels by the
    device-resident timestep tensor: no host->device traffic and no host sync per step;
  * `p_sample` is model call + ONE fused update kernel (x0 from eps, data-consistency projection,
    prev_recon overwrite, eps re-derivation, rho-mixed x_{t-1});  the learned-range variance and
    the posterior mean that the reference computes and then ignores (:278-292,328-330) are only
    computed by the public `p_mean_variance`;
  * `prev_recon` stays on the device.
The auxiliary face prior branch (:471-496) calls the user-supplied reference PyTorch modules
unchanged (out of scope per BASELINE.json north_star).
"""
from __future__ import annotations

import enum
import math

import numpy as np
import torch as th

from flair_b200 import ops


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """Named beta schedules of the two FLAIR model families (reference :15-36)."""
    if schedule_name == "face_blur":
        k = 1000 / num_diffusion_timesteps
        return np.linspace(k * 1e-4, k * 2e-2, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "face_bicubic":
        return np.linspace(1e-6, 1e-2, 2000, dtype=np.float64)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _gamma_table(diffusion, zeta, noise_level):
    """Data-consistency weights gamma_t (reference :648-657; noise_level is used raw)."""
    if zeta == -1:
        return np.ones_like(diffusion.betas)
    sigma_t = diffusion.sqrt_one_minus_alphas_cumprod / diffusion.sqrt_alphas_cumprod
    g = zeta * (noise_level ** 2 / sigma_t ** 2)
    g[g >= 1] = 0.991
    g[g <= 1e-1] = 1e-6
    return 1 - g


def _aux_weight_table(num_timesteps, start_timestep, tau, w):
    """Blend weights of the aux face prior (reference :632-646)."""
    if start_timestep - tau > 0:
        ws = np.exp(-np.linspace(0, 1, start_timestep - tau + 1))
        ws = 1 - (ws - ws.min()) / (ws.max() - ws.min()) * (1 - w)
        ws = np.append(ws, np.ones(num_timesteps - start_timestep - 1))
        return np.concatenate([np.ones(tau), ws])
    return np.ones(num_timesteps) * w


class GaussianDiffusion:
    """Sampling utilities with the reference's constructor and method signatures (:112-120)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps

        betas = np.array(betas, dtype=np.float64)
        if betas.ndim != 1:
            raise AssertionError("betas must be 1-D")
        if not ((betas > 0).all() and (betas <= 1).all()):
            raise AssertionError("betas must lie in (0, 1]")
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])

        ac = np.cumprod(1.0 - betas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = ac_prev
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod_prev = np.sqrt(np.append(1.0, ac))  # NB: length T+1 (reference :141)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.sqrt_one_minus_alphas_cumprod_prev = np.append(0.0, np.sqrt(1.0 - ac[:-1]))
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - ac_prev) * np.sqrt(1.0 - betas) / (1.0 - ac)
        self.posterior_mean_coef3 = self.posterior_mean_coef1 + self.posterior_mean_coef2 * np.sqrt(ac)
        self.posterior_mean_coef4 = self.posterior_mean_coef2 * np.sqrt(1 - ac)
        self._dev_tables = {}

    # ------------------------------------------------------------------ device tables
    def _table(self, device, kind, gammas=None):
        """fp32 device tables [T, 8] consumed by the kernels (see include/flair_b200.h)."""
        key = (str(device), kind, None if gammas is None else gammas.tobytes())
        tab = self._dev_tables.get(key)
        if tab is None:
            T = self.num_timesteps
            cols = np.zeros((T, 8), dtype=np.float64)
            if kind == "update":
                cols[:, 0] = self.sqrt_recip_alphas_cumprod
                cols[:, 1] = self.sqrt_recipm1_alphas_cumprod
                cols[:, 2] = self.sqrt_alphas_cumprod_prev[:T]
                cols[:, 3] = self.sqrt_one_minus_alphas_cumprod_prev
                cols[:, 4] = 1.0 if gammas is None else gammas
            else:  # "posterior"
                cols[:, 0] = self.posterior_mean_coef1
                cols[:, 1] = self.posterior_mean_coef2
                cols[:, 2] = self.posterior_log_variance_clipped
                cols[:, 3] = np.log(self.betas)
                if self.model_var_type == ModelVarType.FIXED_LARGE:
                    v = np.append(self.posterior_variance[1], self.betas[1:])
                    cols[:, 4], cols[:, 5] = v, np.log(v)
                else:
                    cols[:, 4], cols[:, 5] = self.posterior_variance, self.posterior_log_variance_clipped
            tab = th.from_numpy(cols).float().to(device)
            if len(self._dev_tables) > 64:
                self._dev_tables.clear()
            self._dev_tables[key] = tab
        return tab

    # ------------------------------------------------------------------ forward process
    def q_mean_variance(self, x_start, t):
        mean = _extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = _extract_into_tensor(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = _extract_into_tensor(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def q_sample(self, x_start, t, noise=None):
        """x_t ~ q(x_t | x_0) (reference :206-224).  `t` must be uniform over the batch (it is in
        every FLAIR call site, scripts/video_sample.py:379-397)."""
        if noise is None:
            noise = th.randn_like(x_start)
        if noise.shape != x_start.shape:
            raise AssertionError("noise / x_start shape mismatch")
        ti = int(t.reshape(-1)[0]) if th.is_tensor(t) else int(t)
        a = float(np.float32(self.sqrt_alphas_cumprod[ti]))
        b = float(np.float32(self.sqrt_one_minus_alphas_cumprod[ti]))
        return ops.axpby(x_start, noise, a, b)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        if x_start.shape != x_t.shape:
            raise AssertionError("x_start / x_t shape mismatch")
        mean, var, logvar = ops.mean_variance(x_t, x_start, None, self._table(x_t.device, "posterior"),
                                              learned_range=False, t_arr=t.long())
        if self.model_var_type == ModelVarType.FIXED_LARGE:  # this method always reports the true posterior
            var = _extract_into_tensor(self.posterior_variance, t, x_t.shape)
            logvar = _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    # ------------------------------------------------------------------ reverse process
    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _run_model(self, model, x, t, model_kwargs):
        if model_kwargs is None:
            model_kwargs = {}
        # the reference injects these two tables into the kwargs (:273-274); models swallow them
        model_kwargs["sqrt_recip_alphas_cumprod"] = self.sqrt_recip_alphas_cumprod
        model_kwargs["sqrt_recipm1_alphas_cumprod"] = self.sqrt_recipm1_alphas_cumprod
        B, C = x.shape[:2]
        if tuple(t.shape) != (B,):
            raise AssertionError(f"t must have shape ({B},)")
        out = model(x, self._scale_timesteps(t), **model_kwargs)
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            if tuple(out.shape) != (B, C * 2, *x.shape[2:]):
                raise AssertionError(f"model output shape {tuple(out.shape)} != {(B, C * 2, *x.shape[2:])}")
        return out

    def p_mean_variance(self, model, x, t, clip_denoised=True, model_kwargs=None):
        """Full reference dict {mean, variance, log_variance, pred_xstart} (:250-342), EPSILON models."""
        if self.model_mean_type != ModelMeanType.EPSILON:
            raise NotImplementedError("FLAIR models predict epsilon (scripts/video_sample.py:35-68)")
        if self.model_var_type == ModelVarType.LEARNED:
            raise NotImplementedError("ModelVarType.LEARNED is not used by FLAIR")
        out = self._run_model(model, x, t, model_kwargs)
        t64 = t.long()
        x0 = ops.pred_xstart(x, out, self._table(x.device, "update"), t_arr=t64, clip_denoised=clip_denoised)
        learned = self.model_var_type == ModelVarType.LEARNED_RANGE
        mean, var, logvar = ops.mean_variance(x, x0, out if learned else None,
                                              self._table(x.device, "posterior"), learned_range=learned, t_arr=t64)
        return {"mean": mean, "variance": var, "log_variance": logvar, "pred_xstart": x0}

    def _predict_xstart_from_eps(self, x_t, t, eps):
        if x_t.shape != eps.shape:
            raise AssertionError("x_t / eps shape mismatch")
        return ops.pred_xstart(x_t, eps, self._table(x_t.device, "update"), t_arr=t.long(), clip_denoised=False)

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - pred_xstart) / \
            _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def sample(self, model, noise, model_kwargs, restore_fn, face_restore_helper, aux_model, post_fn,
               clip_denoised=True, sample_mode="ddpm", device=None, progress=False, w=0.5, tau=None,
               aligned=False, affine_matrices=None, rho=0.35, noise_level=None, prev_recon=None, zeta=-1,
               t_start=-1):
        if tau is None:
            tau = 0
        if sample_mode != "ddpm":
            raise NotImplementedError(sample_mode)  # the reference returns an unbound name here (:397-421)
        return self.p_sample_loop(model=model, shape=noise.shape, noise=noise, clip_denoised=clip_denoised,
                                  model_kwargs=model_kwargs, progress=progress, device=device,
                                  restore_fn=restore_fn, face_restore_helper=face_restore_helper,
                                  aux_model=aux_model, post_fn=post_fn, w=w, tau=tau, aligned=aligned,
                                  affine_matrices=affine_matrices, rho=rho, noise_level=noise_level,
                                  prev_recon=prev_recon, zeta=zeta, t_start=t_start)

    def p_sample(self, model, x, t, clip_denoised=True, model_kwargs=None, restore_fn=None,
                 affine_matrices=None, face_restore_helper=None, aux_model=None, w=0.5, start_timestep=None,
                 tau=None, aligned=False, rho=0.35, prev_recon=None, gamma=None, _t_host=None, _noise=None,
                 _generator=None, _coef=None):
        """x_t -> {sample: x_{t-1}, pred_xstart} (reference :423-517).

        restore_fn may be any callable Tensor->Tensor (protocol of :465-468); callables exposing
        `.fused_lr(x0)` (this repo's blur operator) get the projection evaluated inside the update
        kernel instead of through a full-resolution intermediate."""
        model_out = self._run_model(model, x, t, model_kwargs)
        if self.model_var_type not in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE) and model_out.shape[1] == 6:
            model_out = model_out[:, :3, ...]
        t64 = t.long()
        # _coef: update table with the gamma_t column already filled in (the graphed step indexes it by the
        # device-resident t); otherwise gamma arrives per call like in the reference (:466)
        coef = self._table(x.device, "update") if _coef is None else _coef
        gamma_arr = None
        if restore_fn is not None and _coef is None:
            if gamma is None:
                raise AssertionError("restore_fn needs gamma (reference p_sample, :466)")
            gamma_arr = gamma.reshape(gamma.shape[0], -1)[:, 0].contiguous() if th.is_tensor(gamma) else \
                th.full((x.shape[0],), float(gamma), device=x.device)
        noise = _noise if _noise is not None else (
            th.randn_like(x) if _generator is None else th.randn(x.shape, device=x.device, generator=_generator))
        frames = None if model_kwargs is None else model_kwargs.get("num_frames")
        prev = None
        if prev_recon is not None:
            prev = prev_recon.to(device=x.device, dtype=th.float32)
            if frames is None:
                raise AssertionError("prev_recon needs model_kwargs['num_frames']")

        aux_active = False
        if aux_model is not None and start_timestep is not None and tau is not None:
            th_ = _t_host if _t_host is not None else int(t64.max())  # uniform t in every FLAIR call site
            tl_ = _t_host if _t_host is not None else int(t64.min())
            aux_active = th_ <= start_timestep and tl_ >= tau

        if not aux_active:
            kw = dict(model_out=model_out, noise=noise, t_arr=t64, gamma_arr=gamma_arr, prev=prev,
                      frames_per_window=frames or 1, rho=rho, clip_denoised=clip_denoised)
            fused = getattr(restore_fn, "fused_lr", None)
            if restore_fn is None:
                sample, x0 = ops.sampler_update(x, coef, **kw)
            elif fused is not None:
                x0_pre = ops.pred_xstart(x, model_out, coef, t_arr=t64, clip_denoised=clip_denoised)
                q_lr, taps, sf, pre = fused(x0_pre)
                sample, x0 = ops.sampler_update(x, coef, q_lr=q_lr, up_taps=taps, sf=sf, pre_stride=pre, **kw)
            else:
                x0_pre = ops.pred_xstart(x, model_out, coef, t_arr=t64, clip_denoised=clip_denoised)
                sample, x0 = ops.sampler_update(x, coef, R=restore_fn(x0_pre), **kw)
            return {"sample": sample, "pred_xstart": x0}

        # ---- aux face prior active: stage x0 explicitly, blend with the reference modules, then update
        x0 = ops.pred_xstart(x, model_out, coef, t_arr=t64, clip_denoised=clip_denoised)
        if restore_fn is not None:
            x0 = ops.dc_apply(x0, restore_fn(x0), gamma_arr=gamma_arr, clip_denoised=clip_denoised)
        if not aligned:
            aux_face = face_restore_helper.get_crop_face_from_affine_matrices(x0, affine_matrices)
            aux_xt = face_restore_helper.get_crop_face_from_affine_matrices(x, affine_matrices)
        else:
            aux_face, aux_xt = x0, x
        aux_face = aux_model(aux_face, t, aux_xt)
        if not aligned:
            inv_face, inv_mask = face_restore_helper.inverse_faces(aux_face, affine_matrices)
        else:  # x_with_face = aux_face: the blend with a mask of ones (x0 * 0 + face * 1 is exact)
            inv_face = aux_face
            inv_mask = self._ones_mask(x0)
        # x_with_face = x0 (1 - m) + face m; clamp; x0 <- w x0 + (1 - w) x_with_face  (:488-496), one launch
        x0 = ops.aux_blend(x0, inv_face, inv_mask, float(w), clip_denoised=clip_denoised)
        if prev is not None:
            x0 = x0.reshape(-1, frames, *x0.shape[1:]).clone()
            x0[:, : prev.shape[1]].copy_(prev)
            x0 = x0.reshape(-1, *x0.shape[2:])
        sample, x0 = ops.sampler_update(x, coef, x0_in=x0, noise=noise, t_arr=t64, rho=rho)
        return {"sample": sample, "pred_xstart": x0}

    def _ones_mask(self, x0):
        key = (x0.device, x0.shape[0], x0.shape[2], x0.shape[3])
        cache = self.__dict__.setdefault("_ones_masks", {})
        if key not in cache:
            cache.clear()
            cache[key] = th.ones(x0.shape[0], 1, x0.shape[2], x0.shape[3], device=x0.device)
        return cache[key]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, model_kwargs=None, device=None,
                      progress=False, affine_matrices=None, restore_fn=None, face_restore_helper=None,
                      aux_model=None, post_fn=None, w=0.5, tau=None, aligned=False, rho=0.35, noise_level=None,
                      prev_recon=None, zeta=-1, t_start=-1):
        final = None
        for out in self.p_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, model_kwargs=model_kwargs, device=device,
                progress=progress, restore_fn=restore_fn, affine_matrices=affine_matrices,
                face_restore_helper=face_restore_helper, aux_model=aux_model, w=w, tau=tau, aligned=aligned,
                rho=rho, noise_level=noise_level, prev_recon=prev_recon, zeta=zeta, t_start=t_start,
                _views=post_fn is None):
            if post_fn is not None:
                post_fn(out)
            final = out
        return final["sample"].clone() if post_fn is None else final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, model_kwargs=None,
                                  device=None, progress=False, affine_matrices=None, face_restore_helper=None,
                                  aux_model=None, restore_fn=None, w=0.5, tau=None, aligned=False, rho=0.35,
                                  noise_level=None, prev_recon=None, zeta=-1, t_start=-1, noise_tape=None,
                                  generator=None, _views=False):
        """Generator over per-step dicts {sample, pred_xstart, t} (reference :589-689).

        `noise_tape` / `generator` (optional, not in the reference) supply the per-step normals /
        the torch generator they are drawn from: entry i is used by the i-th executed step (parity runs).
        `_views=True` (internal): yield the graphed step's static output buffers without cloning them."""
        if device is None:
            device = next(model.parameters()).device
        if not isinstance(shape, (tuple, list, th.Size)):
            raise AssertionError("shape must be a tuple or list")
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))
        if t_start != -1:
            if t_start < 0 or t_start >= self.num_timesteps:
                raise ValueError("t_start must be in [0, num_timesteps)")
            indices = indices[: t_start + 1]
        indices = indices[::-1]
        start_timestep = indices[0]
        if tau is None:
            tau = 0
        ws = _aux_weight_table(self.num_timesteps, start_timestep, tau, w) if aux_model is not None \
            else np.ones(self.num_timesteps)
        gammas = _gamma_table(self, zeta, noise_level)
        gam_dev = th.from_numpy(gammas).float().to(device)
        if prev_recon is not None:
            prev_recon = prev_recon.to(device=device, dtype=th.float32)  # resident, not re-uploaded per step
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        n = shape[0]
        sg = None
        if aux_model is None and _StepGraph.eligible(model, restore_fn, img):
            sg = _StepGraph.get(self, model, img, model_kwargs or {}, restore_fn, prev_recon, gammas, rho,
                                clip_denoised)
        if sg is not None:
            # One CUDA-graph replay per step: UNet forward, x0, data consistency and the fused update are a single
            # graph (no host work between the ~3000 kernels of a step).  Yielded tensors are the graph's static
            # buffers unless `_views` is False (they are overwritten by the next step).
            with th.no_grad():
                sg.load_window(img, model_kwargs or {}, restore_fn, prev_recon)
                for step_no, i in enumerate(indices):
                    out = sg.step(i, None if noise_tape is None else noise_tape[step_no], generator)
                    if not _views:
                        out = {k: v.clone() for k, v in out.items()}
                    yield out
            return
        with th.no_grad():
            for step_no, i in enumerate(indices):
                t = th.full((n,), i, device=device, dtype=th.long)
                out = self.p_sample(
                    model, img, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs, restore_fn=restore_fn,
                    affine_matrices=affine_matrices, face_restore_helper=face_restore_helper, aux_model=aux_model,
                    w=float(ws[i]), start_timestep=start_timestep, tau=tau, aligned=aligned, rho=rho,
                    prev_recon=prev_recon, gamma=gam_dev[i].expand(n), _t_host=i,
                    _noise=None if noise_tape is None else noise_tape[step_no], _generator=generator)
                img = out["sample"]
                out["t"] = t
                yield out


class _StepGraph:
    """One whole sampling step — `_WrappedModel` + UNet forward, x0 from eps, the task's data-consistency operator,
    `prev_recon` overwrite and the rho-mixed update (reference p_sample, :423-517) — captured ONCE per signature into
    a CUDA graph over static buffers and replayed per step.  The step index lives in a device tensor (the kernels index
    their coefficient tables with it), the per-window inputs (conditioning frames, flows, degraded frames,
    `prev_recon`) are copied into the static buffers when a window starts, and the graph's last node copies x_{t-1}
    back into the x_t buffer, so a step costs the host one `fill_`, one `normal_` and one replay.
    (FLAIR_STEP_GRAPH=0 falls back to one graph per UNet forward + eager update kernels.)"""

    @staticmethod
    def _unwrap(model):
        return getattr(model, "model", model) if type(model).__name__ == "_WrappedModel" else model

    @staticmethod
    def eligible(model, restore_fn, img):
        import os
        m = _StepGraph._unwrap(model)
        if os.environ.get("FLAIR_STEP_GRAPH", "1") == "0" or not img.is_cuda:
            return False
        if not (getattr(m, "use_cuda_graph", False) and hasattr(m, "_forward_impl") and hasattr(m, "flows_for")):
            return False
        return restore_fn is None or all(hasattr(restore_fn, a) for a in ("signature", "static_clone", "load_from"))

    @staticmethod
    def get(diffusion, model, img, kwargs, restore_fn, prev_recon, gammas, rho, clip_denoised):
        m = _StepGraph._unwrap(model)
        ksig = tuple(sorted((k, (tuple(v.shape), str(v.dtype)) if th.is_tensor(v) else
                             (v if isinstance(v, (int, float, bool, str, type(None))) else id(v)))
                            for k, v in kwargs.items() if not isinstance(v, np.ndarray)))
        key = (id(m), m._param_stamp(), getattr(m, "compute_dtype", None), getattr(m, "stream_dtype", None),
               tuple(img.shape), str(img.device), ksig, None if restore_fn is None else restore_fn.signature(),
               None if prev_recon is None else tuple(prev_recon.shape), gammas.tobytes(), float(rho),
               bool(clip_denoised))
        cache = diffusion.__dict__.setdefault("_step_graphs", {})
        sg = cache.get(key)
        if sg is None:
            if len(cache) >= 6:
                cache.clear()
            sg = cache[key] = _StepGraph(diffusion, model, img, kwargs, restore_fn, prev_recon, gammas, rho,
                                         clip_denoised)
        return sg

    def __init__(self, diffusion, model, img, kwargs, restore_fn, prev_recon, gammas, rho, clip_denoised):
        from flair_b200 import _lib as L
        self.diffusion, self.model = diffusion, model
        m = self._unwrap(model)
        dev = img.device
        self.x = img.float().clone()
        self.noise = th.empty_like(self.x)
        self.t = th.zeros(img.shape[0], dtype=th.long, device=dev)
        self.kw = {k: (v.clone() if th.is_tensor(v) else v) for k, v in kwargs.items() if not isinstance(v, np.ndarray)}
        self.frames = self.kw.get("num_frames")
        fl = m.flows_for(self.x.shape, self.kw.get("low_res_input"), self.kw.get("rnn_input"), self.frames,
                         self.kw.get("enable_cross_frames", True))
        self.flows = {r: tuple(f.clone() for f in ff) for r, ff in fl.items()}
        self.restore = None if restore_fn is None else restore_fn.static_clone()
        self.prev = None if prev_recon is None else prev_recon.to(device=dev, dtype=th.float32).clone()
        self.coef = diffusion._table(dev, "update", gammas)  # gamma_t in column 4; the reference keeps it
        run_kw = dict(self.kw, _static_flows=self.flows)

        def run():
            out = diffusion.p_sample(model, self.x, self.t, clip_denoised=clip_denoised, model_kwargs=dict(run_kw),
                                     restore_fn=self.restore, rho=rho, prev_recon=self.prev, _noise=self.noise,
                                     _coef=self.coef)
            return out

        self.t.fill_(diffusion.num_timesteps - 1)
        self.noise.normal_()
        side = th.cuda.Stream()
        side.wait_stream(th.cuda.current_stream())
        with th.cuda.stream(side):  # warm-up: packs weights, sets kernel attributes, sizes the pool
            run()
        th.cuda.current_stream().wait_stream(side)
        self.graph = th.cuda.CUDAGraph()
        n0 = L.LAUNCHES[0]
        with th.cuda.graph(self.graph):
            self.out = run()
            self.x.copy_(self.out["sample"])  # x_{t-1} becomes the next step's x_t inside the graph
        self.launches = L.LAUNCHES[0] - n0
        self.out["t"] = self.t

    def load_window(self, img, kwargs, restore_fn, prev_recon):
        m = self._unwrap(self.model)
        self.x.copy_(img)
        for k, v in kwargs.items():
            if th.is_tensor(v):
                self.kw[k].copy_(v)
        fl = m.flows_for(self.x.shape, self.kw.get("low_res_input"), self.kw.get("rnn_input"), self.frames,
                         self.kw.get("enable_cross_frames", True))
        for r, ff in fl.items():
            for dst, src in zip(self.flows[r], ff):
                dst.copy_(src)
        if restore_fn is not None:
            self.restore.load_from(restore_fn)
        if prev_recon is not None:
            self.prev.copy_(prev_recon)

    def step(self, i, noise, generator):
        from flair_b200 import _lib as L
        self.t.fill_(i)
        if noise is not None:
            self.noise.copy_(noise)
        else:
            self.noise.normal_(generator=generator)
        self.graph.replay()
        L.LAUNCHES[0] += self.launches
        return self.out


def _extract_into_tensor(arr, timesteps, broadcast_shape, dtype=th.float32):
    """Gather per-sample coefficients from a 1-D numpy table and broadcast (reference :692-705).
    Kept for API parity; the kernels index device-resident tables instead."""
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].to(dtype)
    while res.dim() < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)
