"""GPU probe: achieved HBM bandwidth of the memory-bound kernels (algorithmic bytes / CUDA-event time).
Buffers are rotated so that every launch reads data that is not in the 126 MB L2."""
import json, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from flair_b200 import ops, pipeline, synth
dev = "cuda"
PEAK = 6559.4
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

def timeit(fns, reps=5):
    """fns: list of closures over DIFFERENT buffers (rotated); returns device us per call.  The calls are captured
    into a CUDA graph so that the host-side cost of the Python/ctypes wrappers is not charged to the kernels."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns: f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))

def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print(f"{name:58s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f} % of {PEAK:.0f}", flush=True)

R = 4  # rotation depth (4 x >= 84 MB > L2)
# ---- GroupNorm on the three largest map shapes of the forward
for (T, H, C) in ((10, 256, 64), (10, 128, 128), (10, 64, 256)):
    xs = [torch.randn(1, T, H, H, C, device=dev).half() for _ in range(R)]
    outs = [torch.empty_like(x) for x in xs]
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    film = torch.randn(T, 2 * C, device=dev) * 0.1
    n = xs[0].numel()
    stats = [ops.gn_stats(x) for x in xs]
    report(f"gn_stats   [1,{T},{H},{H},{C}] fp16", timeit([lambda x=x: ops.gn_stats(x) for x in xs]), n * 2)
    report(f"gn_apply+FiLM+SiLU [1,{T},{H},{H},{C}] fp16", timeit([lambda x=x, s=s, o=o: ops.gn_apply(x, s, gamma, beta, scale=film[:, :C], shift=film[:, C:], silu=True, out=o) for x, s, o in zip(xs, stats, outs)]), n * 4)
# ---- fused sampler update + blur data consistency, 64 frames per launch
N, S = 64, 256
diffusion = pipeline.make_diffusion("gaussian")
kern = np.load("flair_b200/data/blur_kernel_k03.npy")
A = pipeline.make_operator("gaussian", torch.device(dev), S, kernels_mat=kern)
hr = synth.synthetic_clip(N, S, seed=1).to(dev) * 2 - 1
y = A.DownscaleOP(hr)
restore = pipeline.BlurRestore(A, y)
sets = []
for _ in range(R):
    sets.append(dict(x=torch.randn(N, 3, S, S, device=dev), mo=torch.randn(N, 6, S, S, device=dev), z=torch.randn(N, 3, S, S, device=dev)))
t = torch.full((N,), 50, device=dev, dtype=torch.long)
gam = torch.full((N,), 0.5, device=dev)
def step(d):
    return diffusion.p_sample(lambda *a, **k: d["mo"], d["x"], t, model_kwargs={}, restore_fn=restore, rho=0.25, gamma=gam, _noise=d["z"])
try:
    us = timeit([lambda d=d: step(d) for d in sets], reps=3)
    report(f"p_sample tail (x0, blur DC, update) {N} frames 256^2", us, N * 3.93e6)
except Exception as e:
    print("p_sample probe failed:", repr(e)[:200])
# ---- stand-alone blur restore (G2) and jpeg codec
xs = [torch.randn(N, 3, S, S, device=dev) for _ in range(R)]
report(f"A_pinv blur restore R(x) {N} frames 256^2", timeit([lambda x=x: restore(x) for x in xs], reps=3), N * 1.62e6)
