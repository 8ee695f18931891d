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
# ---- sampler tail at 64 frames per launch: pred_xstart -> Down -> InvhTh -> fused update (blur DC in-register)
N, S = 64, 256
diffusion = pipeline.make_diffusion("gaussian")
kern = np.load("flair_b200/data/blur_kernel_k03.npy")
A = pipeline.make_operator("gaussian", torch.device(dev), S, kernels_mat=kern)
coef = diffusion._table(torch.device(dev), "update")
t = torch.full((N,), 50, device=dev, dtype=torch.long)
gam = torch.full((N,), 0.5, device=dev)
taps_ds = A.DownscaleOP.Filter_OP.taps
taps_inv = A.Conv_LR_with_Inv_hTh_OP.Filter_OP.taps
taps_up = A.Upscale_OP.Filter_OP.taps
sf, pre = int(A.ds_factor), int(A.pre_stride[0])
sets = [dict(x=torch.randn(N, 3, S, S, device=dev), mo=torch.randn(N, 6, S, S, device=dev), z=torch.randn(N, 3, S, S, device=dev),
             x0=torch.randn(N, 3, S, S, device=dev).clamp(-1, 1), lr=torch.randn(N, 3, S // 4, S // 4, device=dev),
             sub=torch.randn(N, 3, S // 4, S // 4, device=dev)) for _ in range(R)]
plane = N * 3 * S * S * 4
report(f"pred_xstart (read x_t, eps; write x0) {N} frames", timeit([lambda d=d: ops.pred_xstart(d["x"], d["mo"], coef, t_arr=t) for d in sets]), 3 * plane)
report(f"blur_down 9x9 /4 (read x0; write LR) {N} frames", timeit([lambda d=d: ops.blur_down(d["x0"], taps_ds, sf, pre) for d in sets]), plane + plane // 16)
report(f"filter_same 39x39 on LR (read, sub; write) {N} frames", timeit([lambda d=d: ops.filter_same(d["lr"], taps_inv, sub=d["sub"]) for d in sets]), 3 * plane // 16)
report(f"sampler_update + in-register Up(q) (3.93 MB/frame) {N} fr", timeit([lambda d=d: ops.sampler_update(d["x"], coef, model_out=d["mo"], noise=d["z"], t_arr=t, gamma_arr=gam, q_lr=d["lr"], up_taps=taps_up, sf=sf, pre_stride=pre, rho=0.25) for d in sets]), N * 3.93e6)
report(f"sampler_update, no DC (3.15 MB/frame + x0) {N} frames", timeit([lambda d=d: ops.sampler_update(d["x"], coef, model_out=d["mo"], noise=d["z"], t_arr=t, rho=0.25) for d in sets]), 5 * plane)
