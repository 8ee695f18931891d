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
