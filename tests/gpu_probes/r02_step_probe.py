"""GPU probe (round 2): device time per sampling step of one 10-frame 256x256 window (blur UNet, gaussian task) with
the whole step in one CUDA graph (FLAIR_STEP_GRAPH=1, default) against one graph per UNet forward + eager tail."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pathlib import Path
from flair_b200 import pipeline, synth
from guided_diffusion.script_util import blur_unet_config
from guided_diffusion.unet_new import UNetModel

S, T, STEPS = 256, 10, int(os.environ.get("STEPS", "12"))
task = sys.argv[1] if len(sys.argv) > 1 else "gaussian"
dev = torch.device("cuda:0")
model = UNetModel(**blur_unet_config(S))
model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
model.convert_to_fp16(); model.eval().to(dev)
diffusion = pipeline.make_diffusion(task)
A = pipeline.make_operator(task, dev, S, kernels_mat=np.load(Path(pipeline.__file__).parent / "data" / "blur_kernel_k03.npy"))
hr = synth.synthetic_clip(T, S, seed=1).to(dev)
lr = A.DownscaleOP(hr * 2 - 1)
if task == "jpeg":
    from guided_diffusion.jpeg import jpeg_roundtrip
    lr = jpeg_roundtrip(lr, 60)
lr01 = ((lr + 1) / 2).clamp(0, 1)
res = {}
for mode in ("1", "0", "1"):
    os.environ["FLAIR_STEP_GRAPH"] = mode
    for rep in range(2):  # first call captures
        gen = torch.Generator(device=dev).manual_seed(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.time(); e0.record()
        out = pipeline.restore_window(model, diffusion, A, task, lr01, image_size=S, t_start=STEPS - 1, generator=gen)
        e1.record(); torch.cuda.synchronize()
        wall = (time.time() - t0) / STEPS * 1e3
        gpu = e0.elapsed_time(e1) / STEPS
    res[mode] = out
    print(f"{task} step_graph={mode}: {gpu:.2f} ms/step device, {wall:.2f} ms/step wall (T={T}, {S}x{S}, {STEPS} steps)", flush=True)
print("graphed == eager:", bool(torch.equal(res["1"], res["0"])), float((res["1"] - res["0"]).abs().max()))
print("peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
