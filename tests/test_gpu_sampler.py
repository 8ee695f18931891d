"""End-to-end parity of the sampler: reference `SpacedDiffusion.sample` + reference UNetModel (video
mode) + reference blur operator on CPU (tests/golden/sampler_*.pt, made by tools/gen_golden_unet.py)
against this repo's pipeline on the GPU, on identical inputs, weights and noise tape.
Target (BASELINE.json north_star): final restored frames >= 40 dB PSNR against the reference output."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def psnr(a, b, peak=2.0):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10 * math.log10(peak * peak / max(mse, 1e-20))


@pytest.fixture(scope="module")
def setup(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pathlib import Path
    from flair_b200 import pipeline, synth
    from guided_diffusion.script_util import blur_unet_config
    from guided_diffusion.unet_new import UNetModel
    dev = torch.device("cuda:0")
    model = UNetModel(**blur_unet_config(64))
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    model.convert_to_fp16()
    model.eval().to(dev)
    kern = np.load(Path(pipeline.__file__).parent / "data" / "blur_kernel_k03.npy")
    A = pipeline.make_operator("gaussian", dev, 64, kernels_mat=kern)
    return model, pipeline.make_diffusion("gaussian"), A, dev


def _run(setup, fx):
    from flair_b200 import pipeline, synth
    model, diffusion, A, dev = setup
    T, S, ts = fx["frames"], fx["size"], fx["t_start"]
    n_steps = 100 if ts == -1 else ts + 1
    tape = synth.noise_tape((T, 3, S, S), n_steps, seed=fx["noise_seed"]).to(dev)
    out = pipeline.restore_window(model, diffusion, A, "gaussian", fx["lr01"].to(dev), image_size=S, t_start=ts,
                                  noise_tape=tape)
    return out.cpu()


def test_ten_steps(setup, golden):
    fx = golden("sampler_t9.pt")
    out = _run(setup, fx)
    p = psnr(out, fx["sample"])
    print("10-step PSNR vs reference:", p)
    assert p >= 40.0


def test_full_sampler_psnr(setup, golden):
    try:
        fx = golden("sampler_full.pt")
    except FileNotFoundError:
        pytest.skip("tests/golden/sampler_full.pt not generated")
    out = _run(setup, fx)
    p = psnr(out, fx["sample"])
    print("full 100-step PSNR vs reference:", p)
    assert p >= 40.0


def test_chained_windows_match_single_window_prefix(setup):
    """Window chaining: with prev_recon the first OVERLAP frames of window 2 are overwritten by window 1's
    tail at every step, so the stitched clip has exactly N frames and window 1's frames are untouched."""
    from flair_b200 import pipeline, synth
    model, diffusion, A, dev = setup
    hr = synth.synthetic_clip(12, 64, seed=5).to(dev)
    lr01 = ((A.DownscaleOP(hr * 2 - 1) + 1) / 2).clamp(0, 1)
    g = torch.Generator(device=dev).manual_seed(0)
    full = pipeline.restore_clip(model, diffusion, A, "gaussian", lr01, image_size=64, chained=True, generator=g,
                                 t_start=4)
    assert full.shape == (12, 3, 64, 64) and bool(torch.isfinite(full).all())
    g = torch.Generator(device=dev).manual_seed(0)
    first = pipeline.restore_window(model, diffusion, A, "gaussian", lr01[:10], image_size=64, t_start=4, generator=g)
    assert torch.allclose(full[:10], (first.clamp(-1, 1) + 1) / 2, atol=1e-6)


@pytest.mark.parametrize("task", ["gaussian", "jpeg", "x8_bicubic", "x16_bicubic"])
def test_every_task_end_to_end(golden, task):
    """All four demo tasks (BASELINE.json configs 2-5) through the public pipeline on a 12-frame 64x64 clip, two
    chained windows, last 3 sampler steps: shape, range, finiteness, and the same seed gives the same bits."""
    from pathlib import Path
    from flair_b200 import pipeline, synth
    dev = torch.device("cuda:0")
    S, N = 64, 12
    if task in ("gaussian", "jpeg"):
        from guided_diffusion.unet_new import UNetModel
        model = UNetModel(**golden("unet_blur.pt")["cfg"], use_fp16=True)
        kern = np.load(Path(pipeline.__file__).parent / "data" / "blur_kernel_k03.npy")
        A = pipeline.make_operator(task, dev, S, kernels_mat=kern)
    else:
        from guided_diffusion.sr3 import UNet
        model = UNet(**golden("unet_sr3.pt")["cfg"], dtype=torch.float16)
        A = pipeline.make_operator(task, dev, S)
    model.load_state_dict(synth.synthetic_state_dict(model, seed=21))
    model.convert_to_fp16()
    model.eval().to(dev)
    diffusion = pipeline.make_diffusion(task)
    hr = synth.synthetic_clip(N, S, seed=9).to(dev) * 2 - 1
    if task in ("gaussian", "jpeg"):
        lr = A.DownscaleOP(hr)
        if task == "jpeg":
            from guided_diffusion.jpeg import jpeg_decode, jpeg_encode
            lr = jpeg_decode(jpeg_encode(lr, 60), 60)
    else:
        f = pipeline.KNOBS[task].factor
        lr = A.A(hr.reshape(N, -1)).reshape(N, 3, S // f, S // f)
    lr01 = ((lr + 1) / 2).clamp(0, 1)
    outs = []
    for _ in range(2):
        g = torch.Generator(device=dev).manual_seed(3)
        outs.append(pipeline.restore_clip(model, diffusion, A, task, lr01, image_size=S, chained=True, generator=g,
                                          t_start=2))
    out = outs[0]
    assert out.shape == (N, 3, S, S) and bool(torch.isfinite(out).all())
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert torch.equal(outs[0], outs[1])
