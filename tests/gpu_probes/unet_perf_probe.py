"""GPU probe: wall / device time of one UNet forward at the BASELINE size (256^2, T=10)."""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flair_b200 import synth
from guided_diffusion.unet_new import UNetModel

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
mode = sys.argv[2] if len(sys.argv) > 2 else "video"
SZ = int(sys.argv[3]) if len(sys.argv) > 3 else 256
import os
NOVSR = bool(int(os.environ.get('NOVSR', '0')))
cfg = dict(image_size=SZ, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=2,
           attention_resolutions=(16, 32, 64), rnn_resolutions=() if NOVSR else (1, 2), channel_mult=(0.5, 1, 1, 2, 2, 4, 4),
           use_fp16=True, num_head_channels=64, resblock_updown=True, use_scale_shift_norm=True,
           temporal_block=True, use_checkpoint=True)
t0 = time.time()
model = UNetModel(**cfg)
model.load_state_dict(synth.synthetic_state_dict(model))
model.convert_to_fp16(); model.eval().cuda()
model.use_cuda_graph = not bool(int(os.environ.get('NOGRAPH', '0')))
print("model ready", time.time() - t0, flush=True)
dev = "cuda"
clip = (synth.synthetic_clip(T, SZ) * 2 - 1).to(dev)
x = torch.randn(T, 3, SZ, SZ, device=dev)
ts = torch.full((T,), 500, device=dev)
def fwd():
    if mode == "video":
        return model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
    return model(x, ts, low_res_input=clip[:, None], num_frames=1, enable_cross_frames=False)
for _ in range(int(os.environ.get('NWARM', '2'))): fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n = int(os.environ.get('NTIMED', '3'))
t0 = time.time(); e0.record()
for _ in range(n): o = fwd()
e1.record(); torch.cuda.synchronize()
wall = (time.time() - t0) / n * 1e3
gpu = e0.elapsed_time(e1) / n
gf = {"video": 2370.2, "image": 248.0}[mode] * T * (SZ / 256) ** 2
print(f"{mode} T={T}: wall {wall:.1f} ms  device {gpu:.1f} ms  -> {gf/gpu:.1f} TFLOP/s algorithmic; finite={bool(torch.isfinite(o).all())}")
print("peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
