"""GPU parity of the blur/JPEG UNet forward (through the drop-in UNetModel -> C ABI) against
outputs of the reference UNetModel (tests/golden/unet_blur.pt).  Tolerance: BASELINE.json north_star,
1e-2 relative L2 for a bf16 forward against the fp32 reference."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
# (operand dtype, stream dtype, tolerance): fp16 is the default (the reference's torso dtype) and is held to
# a tighter bound than the north_star's 1e-2; bf16 operands with an fp32 stream must meet the 1e-2.
MODES = [(torch.float16, torch.float16, 4e-3), (torch.bfloat16, torch.float32, 1e-2)]


@pytest.fixture(scope="module")
def model_and_fx(golden):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from flair_b200 import synth
    from guided_diffusion.unet_new import UNetModel
    fx = golden("unet_blur.pt")
    model = UNetModel(**fx["cfg"], use_fp16=True, use_checkpoint=True)
    model.load_state_dict({k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()})
    model.convert_to_fp16()
    model.eval().cuda()
    return model, fx


@pytest.mark.parametrize("cdt,sdt,tol", MODES)
def test_image_mode(model_and_fx, cdt, sdt, tol):
    model, fx = model_and_fx
    model.compute_dtype, model.stream_dtype = cdt, sdt
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["image_t"].to(dev), low_res_input=fx["low_res"][:, None].to(dev), num_frames=1,
                enable_cross_frames=False, vsrpp_weights=1.0)
    assert out.shape == fx["image_out"].shape and out.dtype == torch.float32
    err = rel_err(out.cpu(), fx["image_out"])
    print("image-mode rel L2", cdt, err)
    assert err < tol


def test_image_mode_is_per_frame(model_and_fx):
    """Frames are independent in image mode: a batch equals the frames run one by one."""
    model, fx = model_and_fx
    model.compute_dtype = model.stream_dtype = torch.float16
    dev = "cuda"
    full = model(fx["x"].to(dev), fx["image_t"].to(dev), low_res_input=fx["low_res"][:, None].to(dev), num_frames=1,
                 enable_cross_frames=False)
    one = model(fx["x"][2:3].to(dev), fx["image_t"][2:3].to(dev), low_res_input=fx["low_res"][2:3, None].to(dev),
                num_frames=1, enable_cross_frames=False)
    assert rel_err(one.cpu(), full[2:3].cpu()) < 1e-6


@pytest.mark.parametrize("cdt,sdt,tol", MODES)
def test_video_mode(model_and_fx, cdt, sdt, tol):
    model, fx = model_and_fx
    model.compute_dtype, model.stream_dtype = cdt, sdt
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["video_t"].to(dev), low_res_input=fx["low_res"][None].to(dev), num_frames=4,
                rnn_input=fx["rnn_input"][None].to(dev), enable_cross_frames=True, vsrpp_weights=1.0)
    err = rel_err(out.cpu(), fx["video_out"])
    print("video-mode rel L2", cdt, err)
    assert err < tol


def test_video_mode_weight_map(model_and_fx):
    model, fx = model_and_fx
    model.compute_dtype = model.stream_dtype = torch.float16
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["video_t"].to(dev), low_res_input=fx["low_res"][None].to(dev), num_frames=4,
                rnn_input=None, enable_cross_frames=True, vsrpp_weights=fx["vsrpp_weights"].to(dev))
    err = rel_err(out.cpu(), fx["video_out_weighted"])
    print("video-mode (weight map) rel L2", err)
    assert err < 4e-3


def test_video_mode_fused_group_norm_statistics(model_and_fx, monkeypatch):
    """The optional path where every conv that feeds a GroupNorm leaves the statistics from its own epilogue
    (FLAIR_FUSED_GN=1; off by default, see flair_b200/ops.py): same forward, eager launches (the cached graph of the
    fixture was captured without it)."""
    from flair_b200 import ops
    model, fx = model_and_fx
    model.compute_dtype = model.stream_dtype = torch.float16
    monkeypatch.setattr(ops, "FUSED_GN", True)
    monkeypatch.setattr(model, "use_cuda_graph", False)
    dev = "cuda"
    out = model(fx["x"].to(dev), fx["video_t"].to(dev), low_res_input=fx["low_res"][None].to(dev), num_frames=4,
                rnn_input=fx["rnn_input"][None].to(dev), enable_cross_frames=True, vsrpp_weights=1.0)
    err = rel_err(out.cpu(), fx["video_out"])
    print("video-mode rel L2 with fused GroupNorm statistics", err)
    assert err < 4e-3
