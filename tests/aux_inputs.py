"""Seeded inputs of the aux-warp fixtures (the same functions tools/gen_golden_aux.py used)."""
import numpy as np

from flair_b200 import synth


def logits_from_seed(seed, n, size):
    """(n, 19, size, size) fp32: piecewise-constant class regions (16 x 16 blocks) plus small noise."""
    rng = np.random.default_rng(seed)
    cells = size // 16
    coarse = rng.standard_normal((n, 19, cells, cells)).astype(np.float32) * 3
    yy, xx = np.mgrid[0:cells, 0:cells]
    r = np.hypot(yy - cells / 2 + 0.5, xx - cells / 2 + 0.5) / (cells / 2)
    coarse[:, 0] += (6 * (r - 0.75)).astype(np.float32)
    fine = np.kron(coarse, np.ones((16, 16), np.float32))
    return (fine + 0.05 * rng.standard_normal(fine.shape).astype(np.float32)).astype(np.float32)


def case_inputs(c):
    n, img, face = c["n"], c["img"], c["face"]
    frames = synth.synthetic_clip(n, img, seed=c["frames_seed"][0]) * 2 - 1
    frames = (frames + 0.3 * synth.noise_tape((n, 3, img, img), 1, seed=c["frames_seed"][1])[0]).float()
    faces = (synth.synthetic_clip(n, face, seed=c["faces_seed"]) * 2 - 1).float()
    return frames, faces, logits_from_seed(c["logits_seed"], n, face)


# ---------------------------------------------------------------------------------------------------------------
# The two aux-prior sampling steps (reference gaussian_diffusion.py:423-517 with aux_model active): seeded inputs shared by
# tools/gen_golden_aux.py psample (which runs the UNMODIFIED reference on them -> tests/golden/aux_psample.pt), the CPU
# check of the oracle composition (tests/test_oracle_aux.py) and the GPU tests (test_gpu_round2.py / test_gpu_aux.py).
def _aux_aligned(x0, tt, xt):
    return (0.8 * x0 + 0.1).clamp(-1, 1)


def _aux_unaligned(face, tt, xt):
    return (0.7 * face + 0.2 * xt.clamp(-1, 1) * 0.1 + 0.05).clamp(-1, 1)


def aligned_case():
    """aligned=True (no face helper), blur data consistency, t = 97 of the 100-step face_blur schedule."""
    import torch
    from pathlib import Path
    from oracle import degrade
    from oracle.schedule import Tables
    g = torch.load(Path(__file__).resolve().parent / "golden" / "pseudosr_taps.pt", weights_only=True)
    ds, inv = g["ds_kernel"].float(), g["inv_hTh"].float()
    N, S, t, w, rho = 5, 64, 97, 0.6, 0.25
    gen = torch.Generator().manual_seed(5)
    x_t = torch.randn(N, 3, S, S, generator=gen)
    mout = 0.5 * torch.randn(N, 6, S, S, generator=gen)
    noise = torch.randn(N, 3, S, S, generator=gen)
    y = degrade.blur_down(torch.rand(N, 3, S, S, generator=gen) * 2 - 1, ds)
    gam = Tables("face_blur", 1000).gammas(1.0, 2.55)
    return dict(N=N, S=S, t=t, w=w, rho=rho, x_t=x_t, mout=mout, noise=noise, y=y, ds=ds, inv=inv, gammas=gam,
                gamma=gam[t], aux=_aux_aligned)


def aligned_oracle(c):
    """Oracle composition of that step: (pred_xstart, sample)."""
    import torch
    from oracle import degrade, sampler
    from oracle.schedule import Tables
    tab, t, x_t, mout, w, rho = Tables("face_blur", 1000), c["t"], c["x_t"], c["mout"], c["w"], c["rho"]
    a, b = sampler._c(tab.sqrt_recip_alphas_cumprod, t), sampler._c(tab.sqrt_recipm1_alphas_cumprod, t)
    x0 = (a * x_t - b * mout[:, :3]).clamp(-1, 1)
    x0 = (x0 - torch.tensor(float(c["gamma"]), dtype=torch.float64).float()
          * degrade.blur_restore(x0, c["y"], c["ds"], c["inv"])).clamp(-1, 1)
    x0 = w * x0 + (1 - w) * c["aux"](x0, None, None).clamp(-1, 1)
    eps_hat = (a * x_t - x0) / b
    cc, dd = sampler._c(tab.sqrt_alphas_cumprod_prev, t), sampler._c(tab.sqrt_one_minus_alphas_cumprod_prev, t)
    return x0, cc * x0 + (float(np.sqrt(1 - rho)) * dd * eps_hat + float(np.sqrt(rho)) * dd * c["noise"])


def unaligned_case():
    """aligned=False: image 128x128 -> face 128x128 -> back, no data consistency, t = 60."""
    import torch
    N, S, t, w, rho = 3, 128, 60, 0.4, 0.25
    gen = torch.Generator().manual_seed(11)
    x_t = torch.randn(N, 3, S, S, generator=gen)
    mout = 0.5 * torch.randn(N, 6, S, S, generator=gen)
    noise = torch.randn(N, 3, S, S, generator=gen)
    logits = logits_from_seed(3, N, S)
    rng = np.random.default_rng(5)
    Ms = []
    for _ in range(N):
        sc, a = rng.uniform(1.3, 1.7), rng.uniform(-0.2, 0.2)
        R = np.array([[sc * np.cos(a), -sc * np.sin(a)], [sc * np.sin(a), sc * np.cos(a)]])
        tr = np.array([S / 2, S / 2]) - R @ np.array([S * rng.uniform(0.45, 0.55), S * rng.uniform(0.45, 0.55)])
        Ms.append(np.concatenate([R, tr[:, None]], 1))
    return dict(N=N, S=S, t=t, w=w, rho=rho, x_t=x_t, mout=mout, noise=noise, logits=logits, Ms=Ms, aux=_aux_unaligned)


def unaligned_oracle(u):
    import torch
    from oracle import face_warp as fw
    from oracle import sampler
    from oracle.schedule import Tables
    tab, t, x_t, mout, w, rho, S, Ms = Tables("face_blur", 1000), u["t"], u["x_t"], u["mout"], u["w"], u["rho"], u["S"], u["Ms"]
    a_, b_ = sampler._c(tab.sqrt_recip_alphas_cumprod, t), sampler._c(tab.sqrt_recipm1_alphas_cumprod, t)
    x0 = (a_ * x_t - b_ * mout[:, :3]).clamp(-1, 1)
    face = torch.from_numpy(fw.crop_faces(x0.numpy(), Ms, (S, S)))
    face_xt = torch.from_numpy(fw.crop_faces(x_t.numpy(), Ms, (S, S)))
    restored = u["aux"](face, None, face_xt)
    inv_f, inv_m = fw.inverse_faces(restored.numpy(), u["logits"], Ms)
    x0 = torch.from_numpy(fw.blend(x0.numpy(), inv_f, inv_m, w))
    eps_hat = (a_ * x_t - x0) / b_
    c_, d_ = sampler._c(tab.sqrt_alphas_cumprod_prev, t), sampler._c(tab.sqrt_one_minus_alphas_cumprod_prev, t)
    return x0, c_ * x0 + (float(np.sqrt(1 - rho)) * d_ * eps_hat + float(np.sqrt(rho)) * d_ * u["noise"])
