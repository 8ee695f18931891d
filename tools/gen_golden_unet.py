"""UNet golden vectors: run the reference blur/JPEG UNetModel (unmodified file) on CPU with the
deterministic synthetic state dict and dump input/output pairs (imported by gen_golden.py)."""
import math

import torch


def main(OUT):
    import guided_diffusion.nn as rnn
    import guided_diffusion.unet_new as runet
    from flair_b200 import synth
    assert "/root/reference" in runet.__file__

    # CPU stand-in for flash_attn_func (CUDA only): softmax(q k^T / sqrt(d)) v, fp32  
    def cpu_attn(self, q, k, v):
        s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(q.shape[-1])
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v.float()).to(q.dtype)
    rnn.FalshAttn.forward = cpu_attn

    cfg = dict(image_size=64, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
               attention_resolutions=(16, 32, 64), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4),
               use_fp16=False, num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True,
               temporal_block=True, use_checkpoint=False)
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True  # deform_align is only constructed "on CUDA" (unet_new.py:650)
    model = runet.UNetModel(**cfg)
    torch.cuda.is_available = real
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}

    g = torch.Generator().manual_seed(21)
    clip = synth.synthetic_clip(4, 64, seed=4) * 2 - 1
    x = torch.randn(4, 3, 64, 64, generator=g)
    fx = {"cfg": {k: v for k, v in cfg.items() if k not in ("use_fp16", "use_checkpoint")}, "keys": keys,
          "x": x, "low_res": clip}
    # image mode: every frame is its own batch element (num_frames=1, cross-frame modules skipped)
    ts = torch.tensor([0, 10, 500, 999])
    fx["image_t"] = ts
    fx["image_out"] = model(x, ts, low_res_input=clip[:, None], num_frames=1, enable_cross_frames=False,
                            vsrpp_weights=1.0)
    # video mode, one window of 4 frames
    tv = torch.full((4,), 500)
    fx["video_t"] = tv
    rnn_in = (clip * 0.9).roll(1, -1)
    fx["rnn_input"] = rnn_in
    fx["video_out"] = model(x, tv, low_res_input=clip[None], num_frames=4, rnn_input=rnn_in[None],
                            enable_cross_frames=True, vsrpp_weights=1.0)
    wmap = (torch.rand(1, 4, 1, 64, 64, generator=g) > 0.5).float() * 0.07 + 0.93
    fx["vsrpp_weights"] = wmap
    fx["video_out_weighted"] = model(x, tv, low_res_input=clip[None], num_frames=4, rnn_input=None,
                                     enable_cross_frames=True, vsrpp_weights=wmap)
    for k in ("image_out", "video_out", "video_out_weighted"):
        print(k, tuple(fx[k].shape), float(fx[k].abs().mean()), float(fx[k].std()))
    torch.save(fx, OUT / "unet_blur.pt")
    print("unet fixture written")


def sampler_full(OUT, steps_t_start=-1, name="sampler_full.pt"):
    """Full respaced sampling of one 4-frame 64x64 window with the reference sampler + reference UNet
    (video mode) + reference blur operator, on a recorded noise tape.  Golden for the 40 dB PSNR target."""
    import time
    import numpy as np
    from scipy.io import loadmat
    import guided_diffusion.gaussian_diffusion as gd
    import guided_diffusion.nn as rnn
    import guided_diffusion.unet_new as runet
    import guided_diffusion.pseudoSR as rpsr
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    from torchvision.transforms import functional as VF
    import torch.nn.functional as F
    from flair_b200 import synth

    def cpu_attn(self, q, k, v):
        s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(q.shape[-1])
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v.float()).to(q.dtype)
    rnn.FalshAttn.forward = cpu_attn
    S, T = 64, 4
    cfg = dict(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
               attention_resolutions=(16, 32, 64), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4),
               use_fp16=False, num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True,
               temporal_block=True, use_checkpoint=False)
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    model = runet.UNetModel(**cfg)
    torch.cuda.is_available = real
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "100", "uniform"),
                        betas=gd.get_named_beta_schedule("face_blur", 1000), noise_schedule="face_blur",
                        model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.LEARNED_RANGE,
                        loss_type=gd.LossType.RESCALED_MSE, rescale_timesteps=False)
    kernel = loadmat("/root/reference/miscs/kernels_12.mat")["kernels"]
    conf = rpsr.Get_pseudoSR_Conf(4); conf.sigmoid_range_limit = False; conf.input_range = np.array(None)
    A = rpsr.pseudoSR(conf, upscale_kernel=kernel[0, 3], kernel_indx=10).WrapArchitecture_PyTorch()
    hr = synth.synthetic_clip(T, S, seed=8)
    lr01 = ((A.DownscaleOP(hr * 2 - 1) + 1) / 2).clamp(0, 1)            # degraded frames in [0,1]
    # --- exactly the per-window preparation of scripts/video_sample.py:372-425 (HR size S instead of 512)
    init = (F.interpolate(lr01, (S, S), mode="area").clamp(0, 1) - 0.5) / 0.5
    degraded = (lr01 - 0.5) / 0.5
    n_steps = 100 if steps_t_start == -1 else steps_t_start + 1
    tape = synth.noise_tape((T, 3, S, S), n_steps, seed=2)
    it = iter(tape[1:])
    gd.th.randn_like = lambda t: next(it).to(t)
    t0 = 99 if steps_t_start == -1 else steps_t_start
    noise = d.q_sample(init, torch.full((T,), t0), noise=tape[0])
    rnn_in = VF.normalize(VF.resize(VF.normalize(degraded, 0.5, 0.5), (S, S), VF.InterpolationMode.BICUBIC), -1, 2).clamp(-1, 1)
    kwargs = {"low_res_input": init[None], "num_frames": T, "enable_cross_frames": True, "vsrpp_weights": 1.0,
              "rnn_input": rnn_in[None]}
    restore = lambda x: A.A_pinv(degraded, x)
    t_begin = time.time()
    sample = d.sample(model, noise, model_kwargs=kwargs, device=torch.device("cpu"), progress=False,
                      clip_denoised=True, restore_fn=restore, post_fn=None, face_restore_helper=None,
                      aux_model=lambda *a, **k: None, w=0.75, tau=d.num_timesteps, affine_matrices=None, aligned=False,
                      sample_mode="ddpm", rho=0.25, noise_level=2.55, prev_recon=None, zeta=1.0, t_start=steps_t_start)
    print("reference sampling took", time.time() - t_begin, "s; out range", float(sample.min()), float(sample.max()))
    torch.save({"lr01": lr01, "hr": hr, "sample": sample, "size": S, "frames": T, "t_start": steps_t_start,
                "noise_seed": 2}, OUT / name)


def sr3_golden(OUT):
    """SR3 (bicubic) UNet: reference sr3.UNet on CPU, synthetic weights, image + video mode."""
    import guided_diffusion.nn as rnn
    import guided_diffusion.sr3 as rsr3
    from flair_b200 import synth
    assert "/root/reference" in rsr3.__file__

    def cpu_attn(self, q, k, v):
        s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(q.shape[-1])
        return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), v.float()).to(q.dtype)
    rnn.FalshAttn.forward = cpu_attn
    cfg = dict(image_size=64, in_channel=6, out_channel=3, inner_channel=64, norm_groups=16, channel_mults=(1, 2, 4, 8, 16),
               attn_res=(16, 8), vsrpp_res=(64,), spatial_attn=False, temporal_attn=True, res_blocks=1, dropout=0.0,
               dtype=torch.float32, cross_frame_module=True, use_checkpoint=False, num_frames=7, head_dim=64)
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    model = rsr3.UNet(**cfg)
    torch.cuda.is_available = real
    model.eval()
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(31)
    clip = synth.synthetic_clip(4, 64, seed=6) * 2 - 1
    x = torch.randn(4, 3, 64, 64, generator=g)
    fx = {"cfg": {k: v for k, v in cfg.items() if k not in ("dtype", "use_checkpoint", "spatial_attn", "dropout")},
          "keys": keys, "x": x, "low_res": clip}
    lv = torch.tensor([0.9995, 0.93, 0.41, 0.012])
    fx["image_level"] = lv
    fx["image_out"] = model(x, lv, low_res_input=clip[:, None], num_frames=1, enable_cross_frames=False)
    lvv = torch.full((4,), 0.37)
    fx["video_level"] = lvv
    wmap = (torch.rand(1, 4, 1, 64, 64, generator=g) > 0.5).float() * 0.07 + 0.93
    fx["vsrpp_weights"] = wmap
    fx["video_out"] = model(x, lvv, low_res_input=clip[None], num_frames=4, enable_cross_frames=True, vsrpp_weights=wmap)
    for k in ("image_out", "video_out"):
        print(k, tuple(fx[k].shape), float(fx[k].abs().mean()), float(fx[k].std()))
    torch.save(fx, OUT / "unet_sr3.pt")
    print("sr3 fixture written")

if __name__ == "__main__":
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import ref_env
    OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
    ref_env.setup()
    if sys.argv[1] == "sr3":
        sr3_golden(OUT)
    elif sys.argv[1] == "full":
        sampler_full(OUT, -1, "sampler_full.pt")
    else:
        sampler_full(OUT, int(sys.argv[1]), f"sampler_t{sys.argv[1]}.pt")
