"""Golden vectors at the BENCHMARKED shapes and for every demo task, produced by the UNMODIFIED reference
(/root/reference) on CPU (round 2; VERDICT r01 "parity at the shapes you benchmark, for every task").

  python tools/gen_golden_big.py blur256      # unet_new.UNetModel video mode, T=10, 256x256  (~8 min on 8 cores)
  python tools/gen_golden_big.py sr3_256      # sr3.UNet video mode, T=10, 256x256, image_size=256 config
  python tools/gen_golden_big.py sampler jpeg|x8_bicubic|x16_bicubic   # 100-step reference sampler, T=4, 64x64
  python tools/gen_golden_big.py tattn        # isolated TemporalAttention modules (unet_new F=5, unet F=7), T=9
  python tools/gen_golden_big.py ops256       # blur data-consistency operator at 256x256 (3 frames)
  python tools/gen_golden_big.py ragged       # video-mode forwards at T = 1..9 (blur) / 2, 5, 9 (SR3), 64x64

Inputs are functions of seeds (flair_b200.synth) so the fixtures hold only the seeds and the reference OUTPUT
(fp16 for the 256x256 forwards: 2^-11 relative rounding against a 1e-2 tolerance)."""
import math
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ref_env  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
ref_env.setup()

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import guided_diffusion.nn as rnn  # noqa: E402  (reference)
from flair_b200 import synth  # noqa: E402

assert "/root/reference" in rnn.__file__, rnn.__file__


def cpu_attn(self, q, k, v):
    """CPU stand-in for flash_attn_func (CUDA only): softmax(q k^T / sqrt(d)) v in fp32."""
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(q.shape[-1])
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v.float()).to(q.dtype)


rnn.FalshAttn.forward = cpu_attn


def _with_cuda_flag(ctor):
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True  # deform_align is only constructed "on CUDA" (unet_new.py:650)
    try:
        return ctor()
    finally:
        torch.cuda.is_available = real


BLUR_CFG = dict(in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
                attention_resolutions=(16, 32, 64), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4),
                num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True, temporal_block=True)
SR3_CFG_64 = dict(image_size=64, in_channel=6, out_channel=3, inner_channel=64, norm_groups=16,
                  channel_mults=(1, 2, 4, 8, 16), attn_res=(16, 8), vsrpp_res=(64,), temporal_attn=True, res_blocks=1,
                  cross_frame_module=True, num_frames=7, head_dim=64)
SR3_CFG_256 = dict(image_size=256, in_channel=6, out_channel=3, inner_channel=64, norm_groups=16,
                   channel_mults=(1, 2, 4, 8, 16), attn_res=(64, 32), vsrpp_res=(512, 256), temporal_attn=True,
                   res_blocks=1, cross_frame_module=True, num_frames=7, head_dim=64)


def blur256():
    import guided_diffusion.unet_new as runet
    S, T = 256, 10
    cfg = dict(image_size=S, **BLUR_CFG)
    model = _with_cuda_flag(lambda: runet.UNetModel(**cfg, use_fp16=False, use_checkpoint=False))
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    g = torch.Generator().manual_seed(121)
    x = torch.randn(T, 3, S, S, generator=g)
    clip = synth.synthetic_clip(T, S, seed=14) * 2 - 1
    tv = torch.full((T,), 500)
    t0 = time.time()
    out = model(x, tv, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
    print("blur256 reference forward:", time.time() - t0, "s", tuple(out.shape), float(out.std()))
    torch.save({"cfg": cfg, "size": S, "frames": T, "x_seed": 121, "clip_seed": 14, "t": 500, "weights_seed": 1234,
                "out_f16": out.half(), "out_norm": float(out.double().norm()),
                "cpu_seconds": time.time() - t0, "cores": torch.get_num_threads()}, OUT / "unet_blur_256.pt")


def sr3_256():
    import guided_diffusion.sr3 as rsr3
    S, T = 256, 10
    cfg = dict(SR3_CFG_256)
    model = _with_cuda_flag(lambda: rsr3.UNet(**cfg, spatial_attn=False, dropout=0.0, dtype=torch.float32,
                                              use_checkpoint=False))
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    g = torch.Generator().manual_seed(131)
    x = torch.randn(T, 3, S, S, generator=g)
    clip = synth.synthetic_clip(T, S, seed=16) * 2 - 1
    lv = torch.full((T,), 0.37)
    wmap = ((clip.mean(1, keepdim=True) > 0).float() * 0.07 + 0.93)[None]   # (1,T,1,S,S) like the face-parse weights
    t0 = time.time()
    out = model(x, lv, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=wmap)
    print("sr3_256 reference forward:", time.time() - t0, "s", tuple(out.shape), float(out.std()))
    torch.save({"cfg": cfg, "size": S, "frames": T, "x_seed": 131, "clip_seed": 16, "level": 0.37,
                "weights_seed": 1234, "out_f16": out.half(), "out_norm": float(out.double().norm()),
                "cpu_seconds": time.time() - t0, "cores": torch.get_num_threads()}, OUT / "unet_sr3_256.pt")


RAGGED_BLUR_CFG = dict(image_size=64, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
                       attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
                       resblock_updown=True, use_scale_shift_norm=True, temporal_block=True)


def ragged():
    """Video-mode forward of the UNMODIFIED reference at every window length the script can produce (a short clip is
    one window of 1..9 frames, the last window of a clip holds 4..10): pins the oracle — and through it the GPU path,
    tests/test_gpu_ragged.py uses the same seeds — at T != 4, 10.  Outputs only (inputs are functions of seeds)."""
    import guided_diffusion.sr3 as rsr3
    import guided_diffusion.unet_new as runet
    S = 64
    fx = {"blur": {}, "sr3": {}, "blur_cfg": RAGGED_BLUR_CFG, "sr3_cfg": dict(SR3_CFG_64), "blur_weights_seed": 99,
          "sr3_weights_seed": 1234, "t": 370, "level": 0.37, "size": S}
    model = _with_cuda_flag(lambda: runet.UNetModel(**RAGGED_BLUR_CFG, use_fp16=False, use_checkpoint=False))
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=99))
    for T in (1, 2, 3, 5, 7, 9):
        x = torch.randn(T, 3, S, S, generator=torch.Generator().manual_seed(100 + T))
        low = synth.synthetic_clip(T, S, seed=40 + T) * 2 - 1
        t0 = time.time()
        out = model(x, torch.full((T,), 370), low_res_input=low[None], num_frames=T, enable_cross_frames=True,
                    vsrpp_weights=1.0)
        fx["blur"][T] = out.clone()
        print("ragged blur T =", T, tuple(out.shape), float(out.std()), f"{time.time() - t0:.1f} s")
    model = _with_cuda_flag(lambda: rsr3.UNet(**SR3_CFG_64, spatial_attn=False, dropout=0.0, dtype=torch.float32,
                                              use_checkpoint=False))
    model.eval()
    keys = torch.load(OUT / "unet_sr3.pt", weights_only=True)["keys"]
    model.load_state_dict({k: synth.synthetic_tensor(k, shp, 1234) for k, shp in keys.items()})
    for T in (2, 5, 9):
        x = torch.randn(T, 3, S, S, generator=torch.Generator().manual_seed(200 + T))
        low = synth.synthetic_clip(T, S, seed=60 + T) * 2 - 1
        out = model(x, torch.full((T,), 0.37), low_res_input=low[None], num_frames=T, enable_cross_frames=True,
                    vsrpp_weights=1.0)
        fx["sr3"][T] = out.clone()
        print("ragged sr3 T =", T, tuple(out.shape), float(out.std()))
    torch.save(fx, OUT / "unet_ragged.pt")


def ops256():
    """The blur data-consistency operator of the UNMODIFIED reference (pseudoSR.py:248-281) at the BENCHMARKED size:
    A_pinv(y, x) on 3 frames of 256x256, inputs exactly those of tests/test_gpu_ops.py::test_blur_restore_vs_oracle_256
    (functions of seeds) -> tests/golden/dc_gaussian_256.pt (outputs only)."""
    from scipy.io import loadmat
    import guided_diffusion.pseudoSR as rpsr
    from oracle import degrade
    kernel = loadmat("/root/reference/miscs/kernels_12.mat")["kernels"]
    conf = rpsr.Get_pseudoSR_Conf(4); conf.sigmoid_range_limit = False; conf.input_range = np.array(None)
    A = rpsr.pseudoSR(conf, upscale_kernel=kernel[0, 3], kernel_indx=10).WrapArchitecture_PyTorch()
    taps = torch.load(OUT / "pseudosr_taps.pt", weights_only=True)
    ds = taps["ds_kernel"].float()
    hr = synth.synthetic_clip(3, 256, seed=11) * 2 - 1
    x = (hr + 0.2 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(3))).clamp(-1, 1)
    y = degrade.blur_down(hr, ds)
    y_ref = A.DownscaleOP(hr)
    print("oracle blur_down vs reference DownscaleOP at 256:", float((y - y_ref).abs().max()))
    R = A.A_pinv(y, x)
    torch.save({"R": R.clone(), "down_hr": y_ref.clone(), "hr_seed": 11, "noise_seed": 3, "frames": 3, "size": 256},
               OUT / "dc_gaussian_256.pt")
    print("wrote dc_gaussian_256.pt", tuple(R.shape), float(R.abs().mean()))


def _bicubic_kernel(factor):
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
    k = torch.from_numpy(k).float()
    return k / k.sum()


def sampler(task):
    """Full 100-step respaced sampling of one 4-frame 64x64 window with the reference sampler, the reference UNet
    (video mode) and the reference data-consistency operator of `task`, on a recorded noise tape
    (scripts/video_sample.py:372-485 with the HR size as a parameter and the aux prior disabled as in SURVEY 8c)."""
    import guided_diffusion.gaussian_diffusion as gd
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    from guided_diffusion.jpeg import jpeg_decode, jpeg_encode
    from torchvision.transforms import functional as VF
    S, T = 64, 4
    blur = task in ("gaussian", "jpeg")
    knobs = {"gaussian": (0.25, 2.55, 1.0, -1), "jpeg": (0.5, 12.75, 1.0, 60),
             "x8_bicubic": (0.85, 0.0, -1, -1), "x16_bicubic": (0.85, 0.0, -1, -1)}[task]
    rho, noise_level, zeta, qf = knobs
    if blur:
        import guided_diffusion.unet_new as runet
        import guided_diffusion.pseudoSR as rpsr
        from scipy.io import loadmat
        model = _with_cuda_flag(lambda: runet.UNetModel(image_size=S, **BLUR_CFG, use_fp16=False, use_checkpoint=False))
        d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "100", "uniform"),
                            betas=gd.get_named_beta_schedule("face_blur", 1000), noise_schedule="face_blur",
                            model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.LEARNED_RANGE,
                            loss_type=gd.LossType.RESCALED_MSE, rescale_timesteps=False)
        kernel = loadmat("/root/reference/miscs/kernels_12.mat")["kernels"]
        conf = rpsr.Get_pseudoSR_Conf(4); conf.sigmoid_range_limit = False; conf.input_range = np.array(None)
        A = rpsr.pseudoSR(conf, upscale_kernel=kernel[0, 3], kernel_indx=10).WrapArchitecture_PyTorch()
        hr = synth.synthetic_clip(T, S, seed=28)
        lr = A.DownscaleOP(hr * 2 - 1)
        if qf != -1:
            lr = jpeg_decode(jpeg_encode(lr, qf), qf)
        lr01 = ((lr + 1) / 2).clamp(0, 1)
        init = (F.interpolate(lr01, (S, S), mode="area").clamp(0, 1) - 0.5) / 0.5
        degraded = (lr01 - 0.5) / 0.5
        if qf != -1:
            restore = lambda x: A.A_pinv(degraded, x, jpeg_encode=lambda im: jpeg_encode(im, qf),
                                         jpeg_decode=lambda pl: jpeg_decode(pl, qf))
        else:
            restore = lambda x: A.A_pinv(degraded, x)
        rnn_in = VF.normalize(VF.resize(VF.normalize(degraded, 0.5, 0.5), (S, S), VF.InterpolationMode.BICUBIC),
                              -1, 2).clamp(-1, 1)
        kwargs = {"low_res_input": init[None], "num_frames": T, "enable_cross_frames": True, "vsrpp_weights": 1.0,
                  "rnn_input": rnn_in[None]}
        wmap = None
    else:
        import guided_diffusion.sr3 as rsr3
        from guided_diffusion.restore_util import SRConv
        factor = 8 if task == "x8_bicubic" else 16
        model = _with_cuda_flag(lambda: rsr3.UNet(**SR3_CFG_64, spatial_attn=False, dropout=0.0, dtype=torch.float32,
                                                  use_checkpoint=False))
        d = SpacedDiffusion(use_timesteps=space_timesteps(2000, "100", "uniform"),
                            betas=gd.get_named_beta_schedule("face_bicubic", 2000), noise_schedule="face_bicubic",
                            model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.FIXED_SMALL,
                            loss_type=gd.LossType.MSE, rescale_timesteps=False)
        A = SRConv(_bicubic_kernel(factor), 3, S, torch.device("cpu"), stride=factor)
        hr = synth.synthetic_clip(T, S, seed=28)
        lr = A.A((hr * 2 - 1).reshape(T, -1)).reshape(T, 3, S // factor, S // factor)
        lr01 = ((lr + 1) / 2).clamp(0, 1)
        init = (VF.resize(lr01, (S, S), VF.InterpolationMode.BICUBIC).clamp(0, 1) - 0.5) / 0.5
        degraded = (lr01 - 0.5) / 0.5
        restore = lambda x: A.A_pinv(A.A(x.reshape(x.shape[0], -1)) - degraded.reshape(x.shape[0], -1)).reshape(*x.size())
        # stands in for the face-parse background mask of scripts/video_sample.py:427-444
        mask = (init.mean(1, keepdim=True) > 0).float()
        wmap = (mask * (0.93 if factor == 8 else 0.98) + (1 - mask) * 1.0)[None]
        kwargs = {"low_res_input": init[None], "num_frames": T, "enable_cross_frames": True, "vsrpp_weights": wmap}
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    tape = synth.noise_tape((T, 3, S, S), 100, seed=2)
    it = iter(tape[1:])
    gd.th.randn_like = lambda t: next(it).to(t)
    noise = d.q_sample(init, torch.full((T,), 99), noise=tape[0])
    t0 = time.time()
    sample = d.sample(model, noise, model_kwargs=kwargs, device=torch.device("cpu"), progress=False,
                      clip_denoised=True, restore_fn=restore, post_fn=None, face_restore_helper=None,
                      aux_model=lambda *a, **k: None, w=0.75, tau=d.num_timesteps, affine_matrices=None, aligned=False,
                      sample_mode="ddpm", rho=rho, noise_level=noise_level, prev_recon=None, zeta=zeta, t_start=-1)
    print(task, "reference sampling took", time.time() - t0, "s; range", float(sample.min()), float(sample.max()))
    torch.save({"task": task, "lr01": lr01, "hr": hr, "sample": sample, "size": S, "frames": T, "t_start": -1,
                "noise_seed": 2, "weights_seed": 1234, "vsrpp_weights": wmap,
                "cfg": (dict(image_size=S, **BLUR_CFG) if blur else dict(SR3_CFG_64))}, OUT / f"sampler_full_{task}.pt")


def tattn():
    """Isolated TemporalAttention: unet_new (F=5, blur UNet) and unet (F=7, SR3) modules on a T=9 feature map, so
    interior frames see 4 / 6 distinct neighbours and the window ends exercise the replicate padding."""
    import guided_diffusion.unet_new as runet
    import guided_diffusion.unet as runet_old
    fx = {}
    T, H, W, C = 9, 8, 8, 128
    g = torch.Generator().manual_seed(77)
    x = torch.randn(1, T, C, H, W, generator=g)
    for name, mod, frames in (("f5", runet, 5), ("f7", runet_old, 7)):
        m = mod.TemporalAttention(C, frames, num_heads=C // 64, num_head_channels=64, use_checkpoint=False)
        m.eval()
        sd = {k: synth.synthetic_tensor("tattn." + name + "." + k, v.shape, 55) for k, v in m.state_dict().items()}
        m.load_state_dict(sd)
        y = m(x)  # (B,T,C,H,W) in and out (unet_new.py:473, unet.py:712)
        fx[name] = {"sd": sd, "out": y.reshape(1, T, C, H, W), "frames": frames}
        print(name, tuple(y.shape), float(y.std()))
    fx["x"] = x
    torch.save(fx, OUT / "temporal_attention.pt")


def script_meta():
    """The reference demo script's constants / config dictionaries and, for both model families as the script builds
    them (MODEL_TYPE[task](**MODEL_CONFIG[task]) followed by convert_to_fp16(), scripts/video_sample.py:327-330),
    the checkpoint layout a `torch.load`-ed state dict must have: key -> (shape, dtype)."""
    import importlib
    import json
    sys.path.insert(0, "/root/reference/scripts")
    m = importlib.import_module("scripts.video_sample")

    def plain(v):
        if isinstance(v, dict):
            return {k: plain(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [plain(x) for x in v]
        if isinstance(v, (int, float, str, bool, type(None))):
            return v
        return str(v)

    meta = {"FRAME_SLICE_LEN": m.FRAME_SLICE_LEN, "OVERLAP": m.OVERLAP, "DEFAULT_WEIGHT": m.DEFAULT_WEIGHT,
            "CKPT_PATH": plain(m.CKPT_PATH), "DIFFUSION_CONFIG": plain(m.DIFFUSION_CONFIG),
            "MODEL_CONFIG": plain(m.MODEL_CONFIG), "MODEL_TYPE": {k: v.__module__ + "." + v.__name__ for k, v in m.MODEL_TYPE.items()},
            "RESTORE_FUNC": {k: v.__name__ for k, v in m.RESTORE_FUNC.items()},
            "commands": sorted(n for n in dir(m) if n.endswith("_demo")), "checkpoint_layout": {}}
    import inspect
    sig = inspect.signature(m.main)
    meta["main_signature"] = {k: (None if p.default is inspect._empty else plain(p.default)) for k, p in sig.parameters.items()}
    for task in ("gaussian", "x8_bicubic"):
        model = _with_cuda_flag(lambda: m.MODEL_TYPE[task](**m.MODEL_CONFIG[task]))
        model.convert_to_fp16()
        meta["checkpoint_layout"][task] = {k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()}
        print(task, len(meta["checkpoint_layout"][task]), "state-dict entries")
        del model
    (OUT / "script_meta.json").write_text(json.dumps(meta, indent=0))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "script":
        script_meta()
        raise SystemExit(0)
    if what == "blur256":
        blur256()
    elif what == "sr3_256":
        sr3_256()
    elif what == "sampler":
        sampler(sys.argv[2])
    elif what == "tattn":
        tattn()
    elif what == "ragged":
        ragged()
    elif what == "ops256":
        ops256()
    else:
        raise SystemExit(__doc__)
