"""Pin the CPU oracle of the blur/JPEG UNet against outputs of the reference UNetModel itself
(tests/golden/unet_blur.pt, made by tools/gen_golden.py unet).  Weights are the deterministic
synthetic state dict (flair_b200.synth), regenerated here from the recorded key/shape list."""
import pytest
import torch

from conftest import rel_err
from flair_b200 import synth
from oracle.unet_blur import BlurUNetOracle


@pytest.fixture(scope="module")
def oracle_model(golden):
    fx = golden("unet_blur.pt")
    sd = {k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()}
    return BlurUNetOracle(fx["cfg"], sd), fx


def test_image_mode(oracle_model):
    m, fx = oracle_model
    out = m.forward(fx["x"], fx["image_t"], fx["low_res"][:, None], num_frames=1, enable_cross_frames=False)
    assert rel_err(out, fx["image_out"]) < 2e-5


def test_video_mode(oracle_model):
    m, fx = oracle_model
    out = m.forward(fx["x"], fx["video_t"], fx["low_res"][None], num_frames=4, rnn_input=fx["rnn_input"][None],
                    enable_cross_frames=True, vsrpp_weights=1.0)
    assert rel_err(out, fx["video_out"]) < 2e-5


def test_video_mode_weight_map(oracle_model):
    m, fx = oracle_model
    out = m.forward(fx["x"], fx["video_t"], fx["low_res"][None], num_frames=4, rnn_input=None,
                    enable_cross_frames=True, vsrpp_weights=fx["vsrpp_weights"])
    assert rel_err(out, fx["video_out_weighted"]) < 2e-5


# ------------------------------------------------------------------------------------------------ SR3
@pytest.fixture(scope="module")
def sr3_oracle(golden):
    from oracle.unet_sr3 import SR3UNetOracle
    fx = golden("unet_sr3.pt")
    sd = {k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()}
    return SR3UNetOracle(fx["cfg"], sd), fx


def test_sr3_image_mode(sr3_oracle):
    m, fx = sr3_oracle
    out = m.forward(fx["x"], fx["image_level"], fx["low_res"][:, None], num_frames=1, enable_cross_frames=False)
    assert rel_err(out, fx["image_out"]) < 2e-5


def test_sr3_video_mode(sr3_oracle):
    m, fx = sr3_oracle
    out = m.forward(fx["x"], fx["video_level"], fx["low_res"][None], num_frames=4, enable_cross_frames=True,
                    vsrpp_weights=fx["vsrpp_weights"])
    assert rel_err(out, fx["video_out"]) < 2e-5


# ------------------------------------------------------------------------------------------------ TemporalAttention
@pytest.mark.parametrize("name", ["f5", "f7"])
def test_temporal_attention_isolated(golden, name):
    """oracle.temporal_attention against the reference modules in isolation (unet_new.TemporalAttention F=5,
    unet.TemporalAttention F=7) on T=9 frames: interior frames with F-1 distinct neighbours + both padded ends
    (tests/golden/temporal_attention.pt, tools/gen_golden_big.py tattn)."""
    fx = golden("temporal_attention.pt")
    sd = {"ta." + k: v for k, v in fx[name]["sd"].items()}
    m = BlurUNetOracle.__new__(BlurUNetOracle)
    m.sd, m.heads_dim = sd, 64
    m.p = lambda n: sd[n].float()
    out = m.temporal_attention(fx["x"], "ta", frames=fx[name]["frames"])
    assert rel_err(out, fx[name]["out"]) < 2e-5


# ------------------------------------------------------------------------------------------------ ragged windows
@pytest.fixture(scope="module")
def ragged(golden):
    from oracle.unet_sr3 import SR3UNetOracle
    fx = golden("unet_ragged.pt")
    from guided_diffusion.unet_new import UNetModel
    blur_sd = synth.synthetic_state_dict(UNetModel(**fx["blur_cfg"], use_fp16=False), seed=fx["blur_weights_seed"])
    keys = golden("unet_sr3.pt")["keys"]
    sr3_sd = {k: synth.synthetic_tensor(k, shp, fx["sr3_weights_seed"]) for k, shp in keys.items()}
    return fx, BlurUNetOracle(fx["blur_cfg"], blur_sd), SR3UNetOracle(fx["sr3_cfg"], sr3_sd)


@pytest.mark.parametrize("T", [1, 2, 3, 5, 7, 9])
def test_video_mode_every_window_length(ragged, T):
    """tests/golden/unet_ragged.pt (tools/gen_golden_big.py ragged: the unmodified reference at every window length the
    script can produce) pins the oracle at T != 4, 10; tests/test_gpu_ragged.py runs the GPU path on the same seeds."""
    fx, blur, _ = ragged
    S = fx["size"]
    x = torch.randn(T, 3, S, S, generator=torch.Generator().manual_seed(100 + T))
    low = synth.synthetic_clip(T, S, seed=40 + T) * 2 - 1
    out = blur.forward(x, torch.full((T,), fx["t"]), low[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
    assert rel_err(out, fx["blur"][T]) < 2e-5


@pytest.mark.parametrize("T", [2, 5, 9])
def test_sr3_video_mode_every_window_length(ragged, T):
    fx, _, sr3 = ragged
    S = fx["size"]
    x = torch.randn(T, 3, S, S, generator=torch.Generator().manual_seed(200 + T))
    low = synth.synthetic_clip(T, S, seed=60 + T) * 2 - 1
    out = sr3.forward(x, torch.full((T,), fx["level"]), low[None], num_frames=T, enable_cross_frames=True,
                      vsrpp_weights=1.0)
    assert rel_err(out, fx["sr3"][T]) < 2e-5
