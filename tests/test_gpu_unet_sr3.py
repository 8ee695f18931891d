"""GPU parity of the SR3 (x8 / x16 bicubic) UNet forward against outputs of the reference sr3.UNet
(tests/golden/unet_sr3.pt), plus one bicubic sampling step through the pipeline against the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
MODES = [(torch.float16, torch.float16, 4e-3), (torch.bfloat16, torch.float32, 1e-2)]


@pytest.fixture(scope="module")
def model_and_fx(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from flair_b200 import synth
    from guided_diffusion.sr3 import UNet
    fx = golden("unet_sr3.pt")
    model = UNet(**fx["cfg"], dtype=torch.float16, use_checkpoint=True)
    model.load_state_dict({k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()})
    model.convert_to_fp16()
    model.eval().cuda()
    return model, fx


@pytest.mark.parametrize("cdt,sdt,tol", MODES)
def test_image_mode(model_and_fx, cdt, sdt, tol):
    model, fx = model_and_fx
    model.compute_dtype, model.stream_dtype = cdt, sdt
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["image_level"].to(dev), low_res_input=fx["low_res"][:, None].to(dev),
                num_frames=1, enable_cross_frames=False)
    assert out.shape == fx["image_out"].shape and out.dtype == torch.float32
    err = rel_err(out.cpu(), fx["image_out"])
    print("sr3 image-mode rel L2", cdt, err)
    assert err < tol


@pytest.mark.parametrize("cdt,sdt,tol", MODES)
def test_video_mode(model_and_fx, cdt, sdt, tol):
    model, fx = model_and_fx
    model.compute_dtype, model.stream_dtype = cdt, sdt
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["video_level"].to(dev), low_res_input=fx["low_res"][None].to(dev), num_frames=4,
                enable_cross_frames=True, vsrpp_weights=fx["vsrpp_weights"].to(dev))
    err = rel_err(out.cpu(), fx["video_out"])
    print("sr3 video-mode rel L2", cdt, err)
    assert err < tol


@pytest.mark.parametrize("task", ["x8_bicubic", "x16_bicubic"])
def test_bicubic_sampling_step_vs_oracle(model_and_fx, golden, task):
    """One p_sample step of the bicubic tasks (SR3 UNet + SRConv data consistency, zeta=-1 -> gamma 1,
    rho 0.85) through SpacedDiffusion on the GPU against the oracle sampler + oracle UNet on the CPU."""
    from flair_b200 import pipeline, synth
    from oracle import degrade, sampler
    from oracle.schedule import Tables
    from oracle.unet_sr3 import SR3UNetOracle
    model, fx = model_and_fx
    model.compute_dtype = model.stream_dtype = torch.float16
    dev = torch.device("cuda:0")
    S, T, t = 64, 4, 61
    factor = pipeline.KNOBS[task].factor
    A = pipeline.make_operator(task, dev, S)
    diffusion = pipeline.make_diffusion(task)
    tab = Tables("face_bicubic", 2000)
    hr = synth.synthetic_clip(T, S, seed=12) * 2 - 1
    U, Sv, V = A.U_small.cpu(), A.singulars_small.cpu(), A.V_small.cpu()
    M = U @ torch.diag(Sv) @ V[:, : Sv.shape[0]].t()
    y = M @ hr @ M.t()
    init = torch.nn.functional.interpolate((y + 1) / 2, (S, S), mode="bicubic").clamp(0, 1) * 2 - 1
    tape = synth.noise_tape((T, 3, S, S), 1, seed=13)
    x_t = sampler.q_sample(tab, init, t, tape[0])
    restore = pipeline.BicubicRestore(A, y.to(dev))
    out = diffusion.p_sample(model, x_t.to(dev), torch.full((T,), t, device=dev), model_kwargs={
        "low_res_input": init[None].to(dev), "num_frames": T, "enable_cross_frames": True, "vsrpp_weights": 1.0},
        restore_fn=restore, rho=0.85, gamma=torch.ones(T, device=dev), _noise=tape[1].to(dev))
    sd = {k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()}
    om = SR3UNetOracle(fx["cfg"], sd)
    level = torch.full((T,), sampler.model_time_input(tab, t, sr3=True))
    eps = om.forward(x_t, level, init[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
    ref, ref_x0 = sampler.p_sample_step(tab, x_t, eps, t, tape[1], lambda v: degrade.srconv_restore(v, y, U, Sv, V),
                                        gamma=1.0, rho=0.85)
    assert rel_err(out["pred_xstart"].cpu(), ref_x0) < 1e-2
    assert rel_err(out["sample"].cpu(), ref) < 1e-2
