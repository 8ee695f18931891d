"""GPU parity: the CUDA sampler / data-consistency kernels, called through the drop-in
`guided_diffusion` boundary (which goes through the C ABI), against the CPU oracle and the
reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from flair_b200 import _lib as L
    L.check(L.lib().flair_check_device(0))
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def blur_op(dev, golden):
    from guided_diffusion import pseudoSR as P
    g = golden("pseudosr_taps.pt")
    host = P.pseudoSR(P.Get_pseudoSR_Conf(4), upscale_kernel=g["raw_kernel"].numpy().astype(np.float32),
                      kernel_indx=10)
    return host.WrapArchitecture_PyTorch().to(dev)


def test_blur_pieces(dev, golden, blur_op):
    fx = golden("dc_gaussian.pt")
    x, y = fx["x"].to(dev), fx["y"].to(dev)
    assert rel_err(blur_op.DownscaleOP(x).cpu(), fx["down_x"]) < 1e-5
    assert rel_err(blur_op.Conv_LR_with_Inv_hTh_OP(y).cpu(), fx["inv_y"]) < 1e-5
    assert rel_err(blur_op.A_pinv(y).cpu(), fx["up_inv_y"]) < 1e-5
    assert rel_err(blur_op.A_pinv(y, x).cpu(), fx["R"]) < 1e-5


def test_blur_restore_vs_oracle_256(dev, golden, blur_op):
    """BASELINE size (256^2 HR, 10 frames) against the oracle + linearity property."""
    from flair_b200 import synth
    from oracle import degrade
    g = golden("pseudosr_taps.pt")
    ds, inv = g["ds_kernel"].float(), g["inv_hTh"].float()
    hr = synth.synthetic_clip(3, 256, seed=11) * 2 - 1
    gen = torch.Generator().manual_seed(3)
    x = (hr + 0.2 * torch.randn(hr.shape, generator=gen)).clamp(-1, 1)
    y = degrade.blur_down(hr, ds)
    R = blur_op.A_pinv(y.to(dev), x.to(dev)).cpu()
    assert rel_err(R, degrade.blur_restore(x, y, ds, inv)) < 1e-5
    # and against the UNMODIFIED reference operator on the same inputs (tests/golden/dc_gaussian_256.pt; the oracle is
    # bit-equal to it on the fixture's host, tests/test_oracle_ops.py::test_blur_restore_at_the_benchmarked_size)
    assert rel_err(R, golden("dc_gaussian_256.pt")["R"]) < 1e-5
    # linearity: R(x; y) - R(x'; y) == R(x - x'; 0)
    x2 = x.flip(0)
    lhs = blur_op.A_pinv(y.to(dev), x.to(dev)) - blur_op.A_pinv(y.to(dev), x2.to(dev))
    rhs = blur_op.A_pinv(torch.zeros_like(y).to(dev), (x - x2).to(dev))
    assert rel_err(lhs.cpu(), rhs.cpu()) < 1e-4


def test_jpeg_codec(dev, golden):
    from guided_diffusion import jpeg
    fx = golden("dc_jpeg.pt")
    enc = jpeg.jpeg_encode(fx["img"].to(dev), fx["qf"])
    assert enc[0].shape == fx["enc_luma"].shape and enc[1].shape == fx["enc_chroma"].shape
    assert int((enc[0].cpu() != fx["enc_luma"]).sum()) <= 2, "luma coefficients differ"
    assert int((enc[1].cpu() != fx["enc_chroma"]).sum()) <= 2, "chroma coefficients differ"
    dec = jpeg.jpeg_decode([fx["enc_luma"].to(dev), fx["enc_chroma"].to(dev)], fx["qf"])
    assert rel_err(dec.cpu(), fx["dec"]) < 1e-5
    rt = jpeg.jpeg_roundtrip(fx["img"].to(dev), fx["qf"])
    assert rel_err(rt.cpu(), jpeg.jpeg_decode(enc, fx["qf"]).cpu()) < 1e-6


def test_jpeg_restore(dev, golden, blur_op):
    from guided_diffusion import jpeg
    fx = golden("dc_jpeg.pt")
    qf = fx["qf"]
    R = blur_op.A_pinv(fx["y"].to(dev), fx["x"].to(dev), jpeg_encode=lambda im: jpeg.jpeg_encode(im, qf),
                       jpeg_decode=lambda im: jpeg.jpeg_decode(im, qf))
    assert rel_err(R.cpu(), fx["R"]) < 1e-4  # one flipped round() moves a block by a quant step


def test_jpeg_idempotent_on_decoded(dev):
    """Property: re-encoding a decoded image at the same QF is (nearly) a fixed point."""
    from guided_diffusion import jpeg
    gen = torch.Generator().manual_seed(9)
    img = (torch.rand(4, 3, 64, 64, generator=gen) * 2 - 1).to(dev)
    once = jpeg.jpeg_roundtrip(img, 60)
    twice = jpeg.jpeg_roundtrip(once, 60)
    assert rel_err(twice.cpu(), once.cpu()) < 5e-2


@pytest.mark.parametrize("factor", [8, 16])
def test_srconv(dev, golden, factor):
    from guided_diffusion.restore_util import SRConv
    fx = golden(f"dc_srconv_x{factor}.pt")
    A = SRConv(fx["taps"].to(dev), 3, 64, dev, stride=factor)
    x, y = fx["x"].to(dev), fx["y"].to(dev)
    n = x.shape[0]
    Ax = A.A(x.reshape(n, -1))
    R = A.A_pinv(Ax - y.reshape(n, -1)).reshape(x.shape)
    assert rel_err(R.cpu(), fx["R"]) < 1e-5
    assert rel_err(A.restore(x, y).cpu(), fx["R"]) < 1e-5
    # A(A^+ y) == y: all singular values are kept
    assert rel_err(A.A(A.A_pinv(y.reshape(n, -1))).cpu(), y.reshape(n, -1).cpu()) < 1e-4


@pytest.mark.parametrize("factor", [8, 16])
def test_srconv_256_vs_oracle(dev, factor):
    from flair_b200 import synth
    from guided_diffusion.restore_util import SRConv
    from oracle import degrade
    taps = degrade.bicubic_taps(factor)
    A = SRConv(taps.to(dev), 3, 256, dev, stride=factor)
    hr = synth.synthetic_clip(2, 256, seed=5) * 2 - 1
    x = hr.roll(3, -1) * 0.9
    U, S, V = A.U_small.cpu(), A.singulars_small.cpu(), A.V_small.cpu()
    y = (U @ torch.diag(S) @ V[:, : S.shape[0]].t()) @ hr @ (U @ torch.diag(S) @ V[:, : S.shape[0]].t()).t()
    assert rel_err(A.restore(x.to(dev), y.to(dev)).cpu(), degrade.srconv_restore(x, y, U, S, V)) < 1e-5


def _diffusion():
    from guided_diffusion import gaussian_diffusion as gd
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    return SpacedDiffusion(use_timesteps=space_timesteps(1000, "100", "uniform"),
                           betas=gd.get_named_beta_schedule("face_blur", 1000), noise_schedule="face_blur",
                           model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.LEARNED_RANGE,
                           loss_type=gd.LossType.MSE, rescale_timesteps=False)


def test_p_sample_steps(dev, golden, blur_op):
    fx = golden("sampler_step.pt")
    d = _diffusion()
    y = fx["y"].to(dev)
    mout = fx["model_out"].to(dev)
    model = lambda xx, ts, **kw: mout
    x_t = fx["x_t"].to(dev)
    for generic in (False, True):
        if generic:
            restore = lambda v: blur_op.A_pinv(y, v)
        else:
            class Fused:
                def __call__(self, v):
                    return blur_op.A_pinv(y, v)
                def fused_lr(self, v):
                    return (blur_op.lr_correction(y, v), blur_op.Upscale_OP.Filter_OP.taps, 4, 1)
            restore = Fused()
        for t in (99, 50, 1, 0):
            st = fx["steps"][t]
            out = d.p_sample(model, x_t, torch.full((2,), t, device=dev), model_kwargs={}, restore_fn=restore,
                             rho=0.25, gamma=torch.full((2,), st["gamma"], device=dev), _noise=st["noise"].to(dev))
            assert rel_err(out["pred_xstart"].cpu(), st["pred_xstart"]) < 1e-5, (generic, t)
            assert rel_err(out["sample"].cpu(), st["sample"]) < 1e-5, (generic, t)
    st = fx["steps"]["prev"]
    out = d.p_sample(model, x_t, torch.full((2,), 50, device=dev), model_kwargs={"num_frames": 2},
                     restore_fn=lambda v: blur_op.A_pinv(y, v), rho=0.25, prev_recon=st["prev"].to(dev),
                     gamma=torch.full((2,), fx["steps"][50]["gamma"], device=dev), _noise=st["noise"].to(dev))
    assert rel_err(out["pred_xstart"].cpu(), st["pred_xstart"]) < 1e-5
    assert rel_err(out["sample"].cpu(), st["sample"]) < 1e-5


def test_p_mean_variance_keys(dev, golden):
    fx = golden("sampler_step.pt")
    d = _diffusion()
    mout = fx["model_out"].to(dev)
    out = d.p_mean_variance(lambda xx, ts, **kw: mout, fx["x_t"].to(dev), torch.full((2,), 50, device=dev))
    assert set(out) == {"mean", "variance", "log_variance", "pred_xstart"}
    # oracle for the unused-by-p_sample pieces: plain formulas of gaussian_diffusion.py:278-292,226-248
    x0 = out["pred_xstart"].cpu()
    frac = (fx["model_out"][:, 3:] + 1) / 2
    lv = frac * float(np.log(d.betas[50])) + (1 - frac) * float(d.posterior_log_variance_clipped[50])
    assert rel_err(out["log_variance"].cpu(), lv) < 1e-5
    mean = float(d.posterior_mean_coef1[50]) * x0 + float(d.posterior_mean_coef2[50]) * fx["x_t"]
    assert rel_err(out["mean"].cpu(), mean) < 1e-5


def test_sample_loop(dev, golden, blur_op):
    from flair_b200 import synth
    fx = golden("sampler_loop.pt")
    d = _diffusion()
    y = fx["y"].to(dev)
    tape = synth.noise_tape((2, 3, 64, 64), 6, seed=6).to(dev)

    def toy(xx, ts, **kw):
        e = 0.3 * torch.roll(xx, 1, -1) - 0.1 * xx + 0.001 * ts.float().view(-1, 1, 1, 1)
        return torch.cat([e, torch.zeros_like(e)], 1)

    x5 = d.q_sample(fx["hr"].to(dev), torch.full((2,), 5, device=dev), noise=tape[0])
    assert rel_err(x5.cpu(), fx["x_start"]) < 1e-6
    final = None
    for out in d.p_sample_loop_progressive(toy, x5.shape, noise=x5, model_kwargs={}, device=dev,
                                           restore_fn=lambda v: blur_op.A_pinv(y, v), aux_model=None, rho=0.25,
                                           noise_level=2.55, zeta=1.0, t_start=5, noise_tape=tape[1:]):
        final = out
    assert set(final) == {"sample", "pred_xstart", "t"}
    assert rel_err(final["sample"].cpu(), fx["final"]) < 1e-5
