"""Model / diffusion factories with the reference's names (guided_diffusion/script_util.py:31-310):
`create_model_and_diffusion`, `model_and_diffusion_defaults`, `create_model`,
`create_gaussian_diffusion`, `args_to_dict`, `add_dict_to_argparser`, `str2bool`.

The reference's `create_model_and_diffusion(task, ...)` picks the UNet family and the schedule from
`task` ("face_blur" -> unet_new.UNetModel + 1000-step face_blur schedule with learned sigma;
"face_bicubic" -> sr3.UNet + 2000-step face_bicubic schedule) and ignores most of its other
arguments; any other task leaves `t_schedule` unbound there (:100-108) — here it is a ValueError."""
from __future__ import annotations

import argparse

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps

NUM_CLASSES = 1000


def diffusion_defaults():
    return dict(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing="",
                use_kl=False, predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False)


def model_and_diffusion_defaults():
    res = dict(task="face_blur", image_size=64, num_channels=128, num_res_blocks=2, num_heads=4,
               num_heads_upsample=-1, num_head_channels=-1, attention_resolutions="16,8", rnn_resolutions="1,2",
               channel_mult="", dropout=0.0, class_cond=False, use_checkpoint=False, use_scale_shift_norm=True,
               resblock_updown=False, use_fp16=False, use_new_attention_order=False, temporal_block=True,
               in_channels=6, out_channels=3, mean_type="epsilon", var_type="fixedlarge")
    res.update(diffusion_defaults())
    return res


def blur_unet_config(image_size=512, use_fp16=True, use_checkpoint=True):
    """MODEL_CONFIG["gaussian"] / ["jpeg"] of scripts/video_sample.py:116-155."""
    return dict(image_size=image_size, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
                attention_resolutions=(image_size // 32, image_size // 16, image_size // 8) if image_size == 512
                else (16, 32, 64), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4), use_fp16=use_fp16,
                num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True, temporal_block=True,
                use_checkpoint=use_checkpoint)


def create_model(task, image_size, use_fp16=True, use_checkpoint=True, **unused):
    if task == "face_blur":
        from .unet_new import UNetModel
        return UNetModel(**blur_unet_config(image_size, use_fp16, use_checkpoint))
    if task == "face_bicubic":
        from .sr3 import UNet
        return UNet(image_size=image_size, in_channel=6, out_channel=3, inner_channel=64, norm_groups=16,
                    channel_mults=(1, 2, 4, 8, 16), attn_res=(64, 32), vsrpp_res=(512, 256), spatial_attn=False,
                    temporal_attn=True, res_blocks=1, dropout=0.0, cross_frame_module=True,
                    use_checkpoint=use_checkpoint, num_frames=7, head_dim=64)
    raise ValueError(f"unknown task {task!r}: expected 'face_blur' or 'face_bicubic'")


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing="", mean_type="epsilon",
                              var_type="fixedlarge"):
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    loss_type = gd.LossType.RESCALED_KL if use_kl else (
        gd.LossType.RESCALED_MSE if rescale_learned_sigmas else gd.LossType.MSE)
    if not timestep_respacing:
        timestep_respacing = [steps]
    mean = {"epsilon": gd.ModelMeanType.EPSILON, "xstart": gd.ModelMeanType.START_X,
            "xprev": gd.ModelMeanType.PREVIOUS_X}[mean_type]
    if learn_sigma:
        var = gd.ModelVarType.LEARNED_RANGE
    else:
        var = {"fixedlarge": gd.ModelVarType.FIXED_LARGE, "fixedsmall": gd.ModelVarType.FIXED_SMALL}[var_type]
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing), noise_schedule=noise_schedule,
                           betas=betas, model_mean_type=mean, model_var_type=var, loss_type=loss_type,
                           rescale_timesteps=rescale_timesteps)


def create_model_and_diffusion(task, image_size, *args, use_fp16=True, use_checkpoint=True,
                               timestep_respacing="100", **kwargs):
    """(model, diffusion) for task in {"face_blur", "face_bicubic"} (reference :65-148)."""
    model = create_model(task, image_size, use_fp16=use_fp16, use_checkpoint=use_checkpoint)
    if task == "face_blur":
        diffusion = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="face_blur",
                                              rescale_learned_sigmas=True, timestep_respacing=timestep_respacing)
    else:
        diffusion = create_gaussian_diffusion(steps=2000, learn_sigma=False, noise_schedule="face_bicubic",
                                              timestep_respacing=timestep_respacing, var_type="fixedsmall")
    return model, diffusion


def add_dict_to_argparser(parser, default_dict):
    for k, v in default_dict.items():
        v_type = type(v)
        if v is None:
            v_type = str
        elif isinstance(v, bool):
            v_type = str2bool
        parser.add_argument(f"--{k}", default=v, type=v_type)


def args_to_dict(args, keys):
    return {k: getattr(args, k) for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")
