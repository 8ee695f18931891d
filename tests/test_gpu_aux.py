"""GPU parity of the aux face-prior warps through the C ABI:

  * `FaceRestoreHelper.get_crop_face_from_affine_matrices` / `.inverse_faces` (flair_warp_affine_cubic_f32,
    flair_parse_mask_f32, flair_gaussian_blur_f32) against the outputs of the UNMODIFIED reference methods + OpenCV
    (tests/golden/aux_warp.pt, tools/gen_golden_aux.py), 1e-5 absolute in the [-1, 1] / [0, 1] domains;
  * the blend (flair_aux_blend_f32) against the reference expression;
  * `p_sample` with the prior active and `aligned=False` (crop -> aux model -> inverse warp -> masked blend -> update)
    against the same step composed from the CPU oracle (oracle/sampler.py + oracle/face_warp.py)."""
import numpy as np
import pytest
import torch

from aux_inputs import case_inputs, logits_from_seed
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from flair_b200 import _lib as L
    L.check(L.lib().flair_check_device(0))
    return torch.device("cuda:0")


def _helper(dev, face, logits):
    from guided_diffusion.facelib.utils.face_restoration_helper import FaceRestoreHelper
    lg = torch.from_numpy(logits).to(dev)
    return FaceRestoreHelper(face_size=face, device=dev, face_parse=lambda x: (lg,))


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_warps_vs_reference(dev, golden, idx):
    c = golden("aux_warp.pt")["cases"][idx]
    frames, faces, logits = case_inputs(c)
    s = c["stride"]
    h = _helper(dev, c["face"], logits)
    Ms = list(c["matrices"].numpy())
    crops = h.get_crop_face_from_affine_matrices(frames.to(dev), Ms).cpu()[:, :, ::s, ::s]
    inv_faces, inv_masks = h.inverse_faces(faces.to(dev), Ms)
    e = [float((crops - c["crops"]).abs().max()),
         float((inv_faces.cpu()[:, :, ::s, ::s] - c["inv_faces"]).abs().max()),
         float((inv_masks.cpu()[:, :, ::s, ::s] - c["inv_masks"]).abs().max())]
    print(f"case {idx} (image {c['img']}, face {c['face']}): max abs err crop {e[0]:.2e} inverse face {e[1]:.2e} mask {e[2]:.2e}")
    assert max(e) < 1e-5
    # the warp is a pure function of its inputs: a second call gives the same bits
    again = h.get_crop_face_from_affine_matrices(frames.to(dev), Ms).cpu()[:, :, ::s, ::s]
    assert torch.equal(again, crops)


def test_blend_vs_reference(dev, golden):
    from flair_b200 import ops, synth
    fx = golden("aux_warp.pt")
    c, b = fx["cases"][0], fx["blend"]
    x0 = (synth.synthetic_clip(c["n"], c["img"], seed=b["x0_seed"]) * 2 - 1).float()
    out = ops.aux_blend(x0.to(dev), c["inv_faces"].to(dev), c["inv_masks"].to(dev), b["w"])
    assert float((out.cpu() - b["out"]).abs().max()) < 1e-6


def test_p_sample_aux_unaligned_vs_oracle(dev, golden):
    """One sampling step with the prior active through the device-side helper (image 64x64 -> face 128x128 -> back)."""
    from oracle import face_warp as fw
    from oracle import sampler
    from oracle.schedule import Tables
    import guided_diffusion.gaussian_diffusion as gd
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "100", "uniform"),
                        betas=gd.get_named_beta_schedule("face_blur", 1000), noise_schedule="face_blur",
                        model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.LEARNED_RANGE,
                        loss_type=gd.LossType.MSE, rescale_timesteps=False)
    tab = Tables("face_blur", 1000)
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
    aux = lambda face, tt, xt: (0.7 * face + 0.2 * xt.clamp(-1, 1) * 0.1 + 0.05).clamp(-1, 1)
    h = _helper(dev, S, logits)
    out = d.p_sample(lambda xx, ts, **kw: mout.to(dev), x_t.to(dev), torch.full((N,), t, device=dev), model_kwargs={},
                     aux_model=aux, face_restore_helper=h, affine_matrices=Ms, w=w, start_timestep=99, tau=5,
                     aligned=False, rho=rho, _noise=noise.to(dev))
    # oracle composition (reference gaussian_diffusion.py:465-515)
    a_, b_ = sampler._c(tab.sqrt_recip_alphas_cumprod, t), sampler._c(tab.sqrt_recipm1_alphas_cumprod, t)
    x0 = (a_ * x_t - b_ * mout[:, :3]).clamp(-1, 1)
    face = torch.from_numpy(fw.crop_faces(x0.numpy(), Ms, (S, S)))
    face_xt = torch.from_numpy(fw.crop_faces(x_t.numpy(), Ms, (S, S)))
    restored = aux(face, None, face_xt)
    inv_f, inv_m = fw.inverse_faces(restored.numpy(), logits, Ms)
    x0 = torch.from_numpy(fw.blend(x0.numpy(), inv_f, inv_m, w))
    eps_hat = (a_ * x_t - x0) / b_
    c_, d_ = sampler._c(tab.sqrt_alphas_cumprod_prev, t), sampler._c(tab.sqrt_one_minus_alphas_cumprod_prev, t)
    ref = c_ * x0 + (float(np.sqrt(1 - rho)) * d_ * eps_hat + float(np.sqrt(rho)) * d_ * noise)
    assert rel_err(out["pred_xstart"].cpu(), x0) < 1e-5
    assert rel_err(out["sample"].cpu(), ref) < 1e-5
    # the same step run by the UNMODIFIED reference p_sample through its own FaceRestoreHelper + cv2
    # (tests/golden/aux_psample.pt; the composition above is within 1.5e-8 of it, tests/test_oracle_aux.py)
    fx = golden("aux_psample.pt")["unaligned"]
    assert rel_err(out["pred_xstart"].cpu(), fx["pred_xstart"]) < 1.5e-5
    assert rel_err(out["sample"].cpu(), fx["sample"]) < 1.5e-5
