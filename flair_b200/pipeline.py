"""Window-level driver of the FLAIR sampler: what scripts/video_sample.py:361-485 does per clip
(10-frame windows with 3-frame overlap, `init` / `rnn_input` / `noise` preparation, the 100-step
sampling loop with the task's data-consistency operator, `prev_recon` chaining and stitching), as an
in-memory API so the CLI, the bench and the multi-GPU sharding all call the same code.

Pre-/post-processing that runs once per window (area / bicubic resize of the LR frames, stitching) is
plain torch on the device; everything that runs once per *sampling step* is flair_b200 kernels."""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial

import numpy as np
import torch
import torch.nn.functional as F

FRAME_SLICE_LEN = 10  # scripts/video_sample.py:202
OVERLAP = 3           # scripts/video_sample.py:203


def windows(n_frames: int, size: int = FRAME_SLICE_LEN, overlap: int = OVERLAP):
    """Frame index ranges produced by more_itertools.windowed(frames, size, step=size-overlap) with the
    None padding filtered (scripts/video_sample.py:361-368)."""
    step = size - overlap
    if not 0 <= overlap < size:
        raise ValueError(f"windows: need 0 <= overlap < size, got size={size} overlap={overlap}")
    out, start = [], 0
    if n_frames <= 0:       # windowed() of an empty sequence yields nothing
        return out
    while True:
        end = min(start + size, n_frames)
        out.append((start, end))
        if end >= n_frames:
            return out
        start += step


@dataclass
class TaskKnobs:
    """Per-demo sampler settings (scripts/video_sample.py:499-556)."""
    rho: float
    noise_level: float
    zeta: float
    jpeg_qf: int = -1
    factor: int = 4


KNOBS = {
    "gaussian": TaskKnobs(rho=0.25, noise_level=2.55, zeta=1.0, factor=4),
    "jpeg": TaskKnobs(rho=0.5, noise_level=12.75, zeta=1.0, jpeg_qf=60, factor=4),
    "x8_bicubic": TaskKnobs(rho=0.85, noise_level=0.0, zeta=-1, factor=8),
    "x16_bicubic": TaskKnobs(rho=0.85, noise_level=0.0, zeta=-1, factor=16),
}


class BlurRestore:
    """restore_fn of the gaussian / jpeg tasks (gaussian_restore, scripts/video_sample.py:183-193) bound to
    one window's degraded frames.  Callable like the reference closure; `fused_lr` lets the sampler
    evaluate Up(.) inside its update kernel.  InvhTh(y) is constant over the steps of a window and is
    computed once here (the reference recomputes it at every step, pseudoSR.py:262-270)."""

    def __init__(self, A, degraded, jpeg_qf=-1, _inv=None):
        self.A, self.y, self.qf = A, degraded.reshape(-1, *degraded.shape[-3:]).contiguous(), jpeg_qf
        self.inv_y = A.Conv_LR_with_Inv_hTh_OP(self.y) if _inv is None else _inv

    def _codec(self):
        if self.qf == -1:
            return None, None
        from guided_diffusion.jpeg import jpeg_decode, jpeg_encode
        return (lambda im: jpeg_encode(im, self.qf)), (lambda pl: jpeg_decode(pl, self.qf))

    def __call__(self, x):
        enc, dec = self._codec()
        return self.A.A_pinv(self.y, x, jpeg_encode=enc, jpeg_decode=dec)

    def fused_lr(self, x):
        enc, dec = self._codec()
        return (self.A.lr_correction(None, x, jpeg_decode=dec, jpeg_encode=enc, inv_LR=self.inv_y),
                self.A.Upscale_OP.Filter_OP.taps, int(self.A.ds_factor), int(self.A.pre_stride[0]))

    # ---- static-buffer protocol of the graphed sampling step (guided_diffusion.gaussian_diffusion._StepGraph)
    def signature(self):
        return ("blur", id(self.A), self.qf, tuple(self.y.shape))

    def static_clone(self):
        """A copy bound to buffers the step graph owns; `load_from` refreshes them for the next window."""
        return BlurRestore(self.A, self.y.clone(), self.qf, _inv=self.inv_y.clone())

    def load_from(self, other):
        self.y.copy_(other.y)
        self.inv_y.copy_(other.inv_y)


class BicubicRestore:
    """bicubic_restore (scripts/video_sample.py:177-181): A^+(A x - y)."""

    def __init__(self, A, degraded):
        self.A, self.y = A, degraded.reshape(-1, *degraded.shape[-3:]).contiguous()

    def __call__(self, x):
        return self.A.restore(x, self.y)

    def signature(self):
        return ("bicubic", id(self.A), tuple(self.y.shape))

    def static_clone(self):
        return BicubicRestore(self.A, self.y.clone())

    def load_from(self, other):
        self.y.copy_(other.y)


def init_frames(task, lr01, size):
    """INIT_FUNC (scripts/video_sample.py:158-163): LR frames in [0,1] -> HR-size conditioning in [0,1]."""
    if task in ("gaussian", "jpeg"):
        return F.interpolate(lr01, (size, size), mode="area").clamp(0, 1)
    from torchvision.transforms import functional as VF
    return VF.resize(lr01, (size, size), VF.InterpolationMode.BICUBIC).clamp(0, 1)


def rnn_frames(task, lr_pm1, size):
    """model_kwargs["rnn_input"] of the gaussian / jpeg tasks (scripts/video_sample.py:406-425)."""
    if task not in ("gaussian", "jpeg"):
        return None
    from torchvision.transforms import functional as VF
    x = VF.normalize(lr_pm1, 0.5, 0.5)
    x = VF.resize(x, (size, size), VF.InterpolationMode.BICUBIC)
    return VF.normalize(x, -1, 2).clamp(-1, 1)


BACKGROUND_WEIGHT = {"x8_bicubic": 0.93, "x16_bicubic": 0.98}   # scripts/video_sample.py:427-444


def background_weights(task, init_pm1, face_parse):
    """model_kwargs["vsrpp_weights"] of the bicubic tasks (scripts/video_sample.py:427-444): the BasicVSR++ branch is
    damped on pixels the parsing network labels background (class 0).  init_pm1: (T,3,S,S) in [-1,1]; face_parse: the
    reference's parsing module (output[0] = class logits), caller-supplied PyTorch.  Returns (1,T,1,S,S), or the
    script's DEFAULT_WEIGHT (1.0) for the other tasks / without a parsing module.  Runs once per window."""
    bg = BACKGROUND_WEIGHT.get(task)
    if bg is None or face_parse is None:
        return 1.0
    mask = (face_parse(init_pm1)[0].argmax(1, keepdim=True) == 0).float()
    weight = mask * bg + (1 - mask) * 1.0
    return weight[None].contiguous()


@torch.no_grad()
def restore_window(model, diffusion, A, task, lr01, *, image_size, prev_recon=None, knobs=None, t_start=-1,
                   noise_tape=None, generator=None, vsrpp_weights=1.0, aux=None, face_parse=None):
    """One window: lr01 (T,3,h,w) in [0,1] on the device -> restored (T,3,S,S) in [-1,1] (fp32, device).

    aux: None, or the aux face prior of the reference script (scripts/video_sample.py:446-475) as a dict
    {aux_model, face_restore_helper, affine_matrices, w, tau, aligned}: `aux_model` is the reference's PyTorch prior
    (CodeFormer), `face_restore_helper` a guided_diffusion.facelib FaceRestoreHelper (device-side crops / inverse
    warps), `affine_matrices` one 2x3 matrix per frame of the window."""
    knobs = knobs or KNOBS[task]
    T = lr01.shape[0]
    dev = lr01.device
    init = (init_frames(task, lr01, image_size) - 0.5) / 0.5
    degraded = (lr01 - 0.5) / 0.5
    t0 = diffusion.num_timesteps - 1 if t_start == -1 else t_start
    if noise_tape is not None:
        q_noise, tape = noise_tape[0], noise_tape[1:]
    else:
        q_noise, tape = torch.randn(init.shape, device=dev, generator=generator), None
    noise = diffusion.q_sample(init, torch.full((T,), t0, device=dev, dtype=torch.long), noise=q_noise)
    if face_parse is not None:
        vsrpp_weights = background_weights(task, init, face_parse)
    model_kwargs = {"low_res_input": init[None], "num_frames": T, "enable_cross_frames": True,
                    "vsrpp_weights": vsrpp_weights}
    rnn = rnn_frames(task, degraded, image_size)
    if rnn is not None:
        model_kwargs["rnn_input"] = rnn[None].contiguous()
    restore = BlurRestore(A, degraded, knobs.jpeg_qf) if task in ("gaussian", "jpeg") else BicubicRestore(A, degraded)
    final = None
    for out in diffusion.p_sample_loop_progressive(
            model, noise.shape, noise=noise, model_kwargs=model_kwargs, device=dev, restore_fn=restore,
            rho=knobs.rho, noise_level=knobs.noise_level, zeta=knobs.zeta, prev_recon=prev_recon,
            t_start=t_start, noise_tape=tape, generator=generator, _views=True,
            **({"aux_model": None} if aux is None else aux)):
        final = out
    return final["sample"].clone()  # the graphed step yields its static buffers: detach the result from them


@torch.no_grad()
def restore_clip(model, diffusion, A, task, lr01, *, image_size, chained=True, knobs=None, noise_tapes=None,
                 generator=None, overlap=OVERLAP, t_start=-1, face_parse=None):
    """Whole clip: lr01 (N,3,h,w) in [0,1] (device) -> (N,3,S,S) in [0,1].

    chained=True reproduces the reference script: windows run in order, every later window is
    hard-conditioned on the previous window's last `overlap` restored frames at every step and
    contributes only its non-overlapping frames (scripts/video_sample.py:369,476-485)."""
    outs, prev = [], None
    if lr01.shape[0] == 0:  # the reference script dies in torch.cat([]) here; say why
        raise ValueError("restore_clip: empty clip (no frames to restore)")
    for k, (a, b) in enumerate(windows(lr01.shape[0], FRAME_SLICE_LEN, overlap)):
        sample = restore_window(model, diffusion, A, task, lr01[a:b], image_size=image_size,
                                prev_recon=prev if chained else None, knobs=knobs, t_start=t_start,
                                noise_tape=None if noise_tapes is None else noise_tapes[k], generator=generator,
                                face_parse=face_parse)
        keep = sample if k == 0 else sample[overlap:]
        prev = sample[None, -overlap:].clone() if chained else None
        outs.append(((keep.clamp(-1, 1) + 1) / 2))
    return torch.cat(outs, 0)


def make_diffusion(task):
    """SpacedDiffusion exactly as scripts/video_sample.py:311-325 builds it."""
    from guided_diffusion import gaussian_diffusion as gd
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    blur = task in ("gaussian", "jpeg")
    steps = 1000 if blur else 2000
    sched = "face_blur" if blur else "face_bicubic"
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, "100", "uniform"), betas=gd.get_named_beta_schedule(sched, steps),
        noise_schedule=sched, model_mean_type=gd.ModelMeanType.EPSILON,
        model_var_type=gd.ModelVarType.LEARNED_RANGE if blur else gd.ModelVarType.FIXED_SMALL,
        loss_type=gd.LossType.RESCALED_MSE if blur else gd.LossType.MSE, rescale_timesteps=False)


def make_operator(task, device, image_size, kernels_mat=None):
    """get_A_func (scripts/video_sample.py:205-263) with the HR size as a parameter (the reference
    hard-wires 512)."""
    if task in ("x8_bicubic", "x16_bicubic"):
        from guided_diffusion.restore_util import SRConv
        factor = KNOBS[task].factor
        k = np.zeros(factor * 4)
        for i in range(factor * 4):
            x = abs((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
            a = -0.5
            k[i] = ((a + 2) * x ** 3 - (a + 3) * x ** 2 + 1) if x <= 1 else (
                (a * x ** 3 - 5 * a * x ** 2 + 8 * a * x - 4 * a) if x < 2 else 0)
        k = torch.from_numpy(k / np.sum(k)).float().to(device)
        return SRConv(k / k.sum(), 3, image_size, device, stride=factor)
    import guided_diffusion.pseudoSR as psr
    if kernels_mat is None:
        raise ValueError("the gaussian / jpeg operator needs the 25x25 blur kernel (miscs/kernels_12.mat [0,3])")
    conf = psr.Get_pseudoSR_Conf(4)
    conf.sigmoid_range_limit = False
    conf.input_range = np.array(None)
    return psr.pseudoSR(conf, upscale_kernel=kernels_mat, kernel_indx=10).WrapArchitecture_PyTorch().to(device)
