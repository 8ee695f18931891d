"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container:  python tools/gen_golden.py [ops|unet|all]
The fixtures pin the oracle (oracle/) and, through it, the CUDA path.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ref_env  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
ref_env.setup()

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.io import loadmat  # noqa: E402

import guided_diffusion.gaussian_diffusion as gd  # noqa: E402  (reference)
from guided_diffusion import jpeg as rjpeg  # noqa: E402
from guided_diffusion import pseudoSR as rpsr  # noqa: E402
from guided_diffusion.respace import SpacedDiffusion, space_timesteps  # noqa: E402
from guided_diffusion.restore_util import SRConv  # noqa: E402
from guided_diffusion.dct import LinearDCT  # noqa: E402
from flair_b200 import synth  # noqa: E402

assert "/root/reference" in gd.__file__, gd.__file__


def make_diffusion(task):
    steps, sched, var = {
        "gaussian": (1000, "face_blur", gd.ModelVarType.LEARNED_RANGE),
        "bicubic": (2000, "face_bicubic", gd.ModelVarType.FIXED_SMALL),
    }[task]
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, "100", "uniform"),
        betas=gd.get_named_beta_schedule(sched, steps),
        noise_schedule=sched, model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=var,
        loss_type=gd.LossType.MSE, rescale_timesteps=False)


def blur_A():
    kernel = loadmat("/root/reference/miscs/kernels_12.mat")["kernels"]
    conf = rpsr.Get_pseudoSR_Conf(4)
    conf.sigmoid_range_limit = False
    conf.input_range = np.array(None)
    host = rpsr.pseudoSR(conf, upscale_kernel=kernel[0, 3], kernel_indx=10)
    return host, host.WrapArchitecture_PyTorch(), kernel[0, 3]


def bicubic_kernel(factor):
    def cub(x, a=-0.5):
        if abs(x) <= 1:
            return (a + 2) * abs(x) ** 3 - (a + 3) * abs(x) ** 2 + 1
        elif 1 < abs(x) and abs(x) < 2:
            return a * abs(x) ** 3 - 5 * a * abs(x) ** 2 + 8 * a * abs(x) - 4 * a
        return 0
    k = np.zeros((factor * 4))
    for i in range(factor * 4):
        k[i] = cub((1 / factor) * (i - np.floor(factor * 4 / 2) + 0.5))
    k = k / np.sum(k)
    kernel = torch.from_numpy(k).float()
    return kernel / kernel.sum()


def gen_ops():
    OUT.mkdir(parents=True, exist_ok=True)
    g = torch.Generator().manual_seed(7)
    # ---- schedules --------------------------------------------------------------------
    sched = {}
    for task in ("gaussian", "bicubic"):
        d = make_diffusion(task)
        sched[task] = {k: torch.from_numpy(np.asarray(getattr(d, k), dtype=np.float64)) for k in (
            "betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
            "sqrt_alphas_cumprod_prev", "sqrt_one_minus_alphas_cumprod_prev",
            "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod")}
        sched[task]["timestep_map"] = torch.tensor(d.timestep_map)
    torch.save(sched, OUT / "schedule.pt")

    # ---- blur x4 operator ---------------------------------------------------------------
    host, A, raw_kernel = blur_A()
    taps = {"raw_kernel": torch.from_numpy(np.asarray(raw_kernel, dtype=np.float64)),
            "ds_kernel": torch.from_numpy(np.asarray(host.ds_kernel, dtype=np.float64)),
            "inv_hTh": torch.from_numpy(np.asarray(host.inv_hTh, dtype=np.float64)),
            "pre_stride": torch.tensor(np.asarray(A.pre_stride)),
            "post_stride": torch.tensor(np.asarray(A.post_stride)),
            "w_down": A.DownscaleOP.Filter_OP.weight[0, 0].clone(),
            "w_inv": A.Conv_LR_with_Inv_hTh_OP.Filter_OP.weight[0, 0].clone(),
            "w_up": A.Upscale_OP.Filter_OP.weight[0, 0].clone()}
    torch.save(taps, OUT / "pseudosr_taps.pt")
    hr = synth.synthetic_clip(2, 64, seed=3) * 2 - 1
    x = (hr + 0.3 * torch.randn(hr.shape, generator=g)).clamp(-1, 1)
    y = A.DownscaleOP(hr)
    fx = {"x": x, "y": y, "down_x": A.DownscaleOP(x), "inv_y": A.Conv_LR_with_Inv_hTh_OP(y),
          "up_inv_y": A.A_pinv(y), "R": A.A_pinv(y, x)}
    torch.save(fx, OUT / "dc_gaussian.pt")

    # ---- JPEG -----------------------------------------------------------------------------
    qf = 60
    img = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1)
    enc = rjpeg.jpeg_encode(img.clone(), qf)
    dec = rjpeg.jpeg_decode([e.clone() for e in enc], qf)
    yj = rjpeg.jpeg_decode(rjpeg.jpeg_encode(A.DownscaleOP(hr), qf), qf)
    Rj = A.A_pinv(yj, x, jpeg_encode=lambda im: rjpeg.jpeg_encode(im, qf),
                  jpeg_decode=lambda im: rjpeg.jpeg_decode(im, qf))
    q1, q2 = rjpeg.quantization_matrix(qf)
    fx = {"qf": qf, "img": img, "enc_luma": enc[0], "enc_chroma": enc[1], "dec": dec,
          "dct": LinearDCT(8, "dct", norm="ortho").weight.clone(),
          "idct": LinearDCT(8, "idct", norm="ortho").weight.clone(),
          "q_luma": q1.reshape(8, 8), "q_chroma": q2.reshape(8, 8), "x": x, "y": yj, "R": Rj}
    torch.save(fx, OUT / "dc_jpeg.pt")

    # ---- SRConv x8 / x16 (img_dim 64) -----------------------------------------------------
    for factor in (8, 16):
        A_sr = SRConv(bicubic_kernel(factor), 3, 64, torch.device("cpu"), stride=factor)
        ysr = A_sr.A(hr.reshape(2, -1)).reshape(2, 3, 64 // factor, 64 // factor)
        R = A_sr.A_pinv(A_sr.A(x.reshape(2, -1)) - ysr.reshape(2, -1)).reshape(*x.shape)
        fx = {"factor": factor, "x": x, "y": ysr, "R": R, "U": A_sr.U_small, "S": A_sr.singulars_small,
              "V": A_sr.V_small, "taps": bicubic_kernel(factor)}
        torch.save(fx, OUT / f"dc_srconv_x{factor}.pt")

    # ---- single sampler steps + a short loop (dummy eps model) ----------------------------
    d = make_diffusion("gaussian")
    tape = synth.noise_tape((2, 3, 64, 64), 8, seed=5)
    it = iter(tape)
    gd.th.randn_like = lambda t: next(it).to(t)  # RNG source only
    restore = lambda v: A.A_pinv(y, v)
    steps = {}
    x_t = d.q_sample(hr, torch.full((2,), 99), noise=next(it))
    mout = torch.randn(2, 6, 64, 64, generator=g)
    model = lambda xx, ts, **kw: mout
    gammas = 1 - np.clip(1.0 * (2.55 ** 2 / (d.sqrt_one_minus_alphas_cumprod / d.sqrt_alphas_cumprod) ** 2), None, None)
    for t in (99, 50, 1, 0):
        # gamma table exactly as p_sample_loop_progressive builds it (gaussian demo knobs)
        gm = 1.0 * (2.55 ** 2 / (d.sqrt_one_minus_alphas_cumprod / d.sqrt_alphas_cumprod) ** 2)
        gm[gm >= 1] = 0.991
        gm[gm <= 1e-1] = 1e-6
        gm = 1 - gm
        tt = torch.full((2,), t)
        noise_used = tape[1 + len(steps)]
        out = d.p_sample(model, x_t, tt, model_kwargs={}, restore_fn=restore, aux_model=None, rho=0.25,
                         gamma=gd._extract_into_tensor(gm, tt, x_t.shape))
        steps[t] = {"sample": out["sample"], "pred_xstart": out["pred_xstart"], "noise": noise_used,
                    "gamma": float(gm[t])}
    # prev_recon variant
    prev = (torch.rand(1, 1, 3, 64, 64, generator=g) * 2 - 1)
    out = d.p_sample(model, x_t, torch.full((2,), 50), model_kwargs={"num_frames": 2}, restore_fn=restore,
                     aux_model=None, rho=0.25, prev_recon=prev,
                     gamma=gd._extract_into_tensor(gm, torch.full((2,), 50), x_t.shape))
    steps["prev"] = {"sample": out["sample"], "pred_xstart": out["pred_xstart"], "noise": tape[5], "prev": prev}
    torch.save({"x_t": x_t, "model_out": mout, "y": y, "hr": hr, "steps": steps, "gammas": torch.from_numpy(gm)},
               OUT / "sampler_step.pt")

    # short loop, t_start = 5 (6 steps), toy deterministic eps model
    tape2 = synth.noise_tape((2, 3, 64, 64), 6, seed=6)
    it2 = iter(tape2[1:])
    gd.th.randn_like = lambda t: next(it2).to(t)
    def toy(xx, ts, **kw):
        e = 0.3 * torch.roll(xx, 1, -1) - 0.1 * xx + 0.001 * ts.float().view(-1, 1, 1, 1)
        return torch.cat([e, torch.zeros_like(e)], 1)
    x5 = d.q_sample(hr, torch.full((2,), 5), noise=tape2[0])
    final = d.sample(toy, x5, model_kwargs={}, restore_fn=restore, face_restore_helper=None,
                     aux_model=lambda *a, **k: None, post_fn=None, tau=d.num_timesteps, affine_matrices=None,
                     aligned=False, sample_mode="ddpm", prev_recon=None, rho=0.25, noise_level=2.55, zeta=1.0,
                     t_start=5, device=torch.device("cpu"))
    torch.save({"x_start": x5, "final": final, "hr": hr, "y": y}, OUT / "sampler_loop.pt")
    print("ops fixtures written to", OUT)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("ops", "all"):
        gen_ops()
    if what in ("unet", "all"):
        import gen_golden_unet
        gen_golden_unet.main(OUT)
