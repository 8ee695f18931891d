"""GPU probe: is a video-mode forward bit-reproducible across graph replays / eager runs?"""
import os, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flair_b200 import ops, synth
from guided_diffusion.unet_new import UNetModel
S, T = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
           attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
           resblock_updown=True, use_scale_shift_norm=True, temporal_block=True, use_fp16=True)
model = UNetModel(**cfg)
model.load_state_dict(synth.synthetic_state_dict(model, seed=99)); model.convert_to_fp16(); model.eval().cuda()
dev = "cuda"
clip = (synth.synthetic_clip(T, S) * 2 - 1).to(dev)
ts = torch.full((T,), 500, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn(T, 3, S, S, device=dev, generator=g) for _ in range(3)]
fwd = lambda x: model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
for graph in (False, True):
    model.use_cuda_graph = graph
    ref = [fwd(x).clone() for x in xs]
    bad = 0
    for rep in range(6):
        for i, x in enumerate(xs):
            y = fwd(x)
            if not torch.equal(y, ref[i]):
                bad += 1
                d = (y - ref[i]).abs()
                print(f"graph={graph} rep {rep} input {i}: max diff {float(d.max()):.3e}, differing elements {int((d > 0).sum())}")
    print(f"graph={graph}: {bad} of 18 forwards differ")
