"""GPU probe: device time of one SR3 (x8 bicubic) video-mode forward at the BASELINE size (T=10, 256^2)."""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flair_b200 import synth
from guided_diffusion.script_util import create_model
T, SZ = 10, 256
model = create_model("face_bicubic", SZ)
model.load_state_dict(synth.synthetic_state_dict(model))
model.convert_to_fp16(); model.eval().cuda()
dev = "cuda"
clip = (synth.synthetic_clip(T, SZ) * 2 - 1).to(dev)
x = torch.randn(T, 3, SZ, SZ, device=dev)
lvl = torch.full((T,), 0.7, device=dev)
fwd = lambda: model(x, lvl, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
for _ in range(2): o = fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3): o = fwd()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"sr3 video T={T} {SZ}x{SZ}: {ms:.1f} ms per forward -> {988.1 * T / ms:.0f} TFLOP/s algorithmic (988.1 GF/frame, SPyNet cached); finite={bool(torch.isfinite(o).all())}")
