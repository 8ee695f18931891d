#!/usr/bin/env python
"""bench.py — FLAIR hot-path benchmark on B200 (driver contract: see the task prompt / DESIGN.md §measurement).

Workload (BASELINE.json configs[1], "gaussian-demo"): Gaussian-blur restoration of a synthetic
16-frame 256x256 clip with the full 100-step respaced sampler, blur/JPEG UNet in video mode, two
chained 10-frame windows (overlap 3), random-init (synthetic) weights.  One *step* = one full
restoration of the clip (200 UNet forwards + 200 data-consistency/update steps).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path (oracle/)

Under torchrun (N > 1) every rank restores its own clip (weak scaling, no data-path collective);
time = max over ranks of the CUDA-event time of K steps between barriers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch

METRIC = "restored_frames_per_s_full_sampler"
UNIT = "frames/s"
FRAMES, SIZE, SAMPLER_STEPS = 16, 256, 100
# algorithmic dense-contraction FLOPs of one video-mode forward per frame (SPyNet cached)
GF_PER_FRAME_FWD = 2370.2


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops=d.get("bf16_tflops_sustained", 1400.0), hbm=d.get("hbm_gbs", 6650.0), src="measured")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/) — the `--impl reference` arm and the cpu_baseline leg
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(size=64, frames=3):
    """One sampler step (video-mode UNet forward + blur data consistency + update) of the CPU oracle on a
    bounded sample: `frames` frames at size x size.  Returns (step_fn, description, scale_to_256)."""
    from flair_b200 import synth
    from oracle import degrade, sampler
    from oracle.schedule import Tables
    from oracle.unet_blur import BlurUNetOracle, default_config
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    cfg = default_config(size)
    keys = _blur_unet_keys(cfg)
    sd = {k: synth.synthetic_tensor(k, shp, 1234) for k, shp in keys.items()}
    model = BlurUNetOracle(cfg, sd)
    kern = np.load(ROOT / "flair_b200" / "data" / "blur_kernel_k03.npy")
    taps = torch.load(ROOT / "tests" / "golden" / "pseudosr_taps.pt", map_location="cpu", weights_only=False)
    ds, inv = taps["ds_kernel"].float(), taps["inv_hTh"].float()
    del kern
    tab = Tables("face_blur", 1000)
    hr = synth.synthetic_clip(frames, size, seed=1) * 2 - 1
    y = degrade.blur_down(hr, ds)
    init = torch.nn.functional.interpolate((y + 1) / 2, (size, size), mode="area").clamp(0, 1) * 2 - 1
    tape = synth.noise_tape((frames, 3, size, size), 1, seed=2)
    t = 50
    x_t = sampler.q_sample(tab, init, t, tape[0])
    gam = tab.gammas(1.0, 2.55)[t]

    def step():
        eps = model.forward(x_t, torch.full((frames,), tab.timestep_map[t]), init[None], num_frames=frames,
                            enable_cross_frames=True, vsrpp_weights=1.0)
        return sampler.p_sample_step(tab, x_t, eps, t, tape[1], lambda v: degrade.blur_restore(v, y, ds, inv),
                                     gamma=gam, rho=0.25)[0]

    desc = (f"1 sampler step (video-mode blur UNet fwd + blur DC + update) on {frames} frames at {size}x{size}, "
            f"oracle port (torch fp32 CPU); frames/s scaled by ({size}/{SIZE})^2 area and /{SAMPLER_STEPS} steps")
    return step, desc, (size / SIZE) ** 2, frames


def _blur_unet_keys(cfg):
    """State-dict key -> shape of the blur UNet for `cfg` (from this repo's module tree; CPU, no kernels)."""
    from guided_diffusion.unet_new import UNetModel
    m = UNetModel(**cfg, use_fp16=False)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def time_cpu(steps, warmup):
    step, desc, area, frames = cpu_step_factory()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    fps = frames / (SAMPLER_STEPS * dt) * area
    return fps, dt, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, dt, desc = time_cpu(args.steps, max(args.warmup, 1))
    cores = os.cpu_count() or 1
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "note": "CPU port of the reference path; each step is a bounded sample"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_name():
    return (f"gaussian-demo: blur restoration of a synthetic {FRAMES}-frame {SIZE}x{SIZE} clip, full {SAMPLER_STEPS}-step "
            "respaced sampler, blur UNet video mode (2 chained 10-frame windows, overlap 3), synthetic weights")


# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    from flair_b200 import _lib as L
    from flair_b200 import ops, pipeline, synth
    from guided_diffusion.script_util import blur_unet_config
    from guided_diffusion.unet_new import UNetModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.check(L.lib().flair_check_device(local))
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG >= VERSION (WARN included)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    model = UNetModel(**blur_unet_config(SIZE))
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    model.convert_to_fp16()
    model.eval().to(dev)
    diffusion = pipeline.make_diffusion("gaussian")
    A = pipeline.make_operator("gaussian", dev, SIZE,
                               kernels_mat=np.load(ROOT / "flair_b200" / "data" / "blur_kernel_k03.npy"))
    # every rank restores its own clip (weak scaling): HR clip -> degraded LR frames with the same operator
    hr = synth.synthetic_clip(FRAMES, SIZE, seed=1 + rank).to(dev)
    lr01 = ((A.DownscaleOP(hr * 2 - 1) + 1) / 2).clamp(0, 1)
    lr_host = lr01.cpu().pin_memory()
    out_host = torch.empty(FRAMES, 3, SIZE, SIZE, dtype=torch.float32).pin_memory()
    gen = torch.Generator(device=dev).manual_seed(2 + rank)

    def restore(lr_dev):
        return pipeline.restore_clip(model, diffusion, A, "gaussian", lr_dev, image_size=SIZE, chained=True,
                                     generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        restore(lr_host.to(dev, non_blocking=True))
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    n0 = L.LAUNCHES[0]
    with ClockSampler(local) as clocks:
        barrier()
        for k in range(args.steps):
            ev[k][0].record()
            lr_dev = lr_host.to(dev, non_blocking=True)      # H2D of this step's inputs (pinned)
            ev[k][1].record()
            out = restore(lr_dev)                             # inputs resident: the `value` region
            ev[k][2].record()
            out_host.copy_(out, non_blocking=True)            # D2H of the restored frames
            ev[k][3].record()
        barrier()
    launches = L.LAUNCHES[0] - n0
    inner = sum(e[1].elapsed_time(e[2]) for e in ev) / 1e3
    outer = sum(e[0].elapsed_time(e[3]) for e in ev) / 1e3
    if world > 1:
        t = torch.tensor([inner, outer], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        inner, outer = float(t[0]), float(t[1])
    total_frames = FRAMES * world * args.steps
    value, e2e = total_frames / inner, total_frames / outer

    if rank == 0:
        pk = peaks()
        # ---- roofline of the dominant kernel (flair_conv_igemm): CUDA-event time of every conv launch of one
        # eager forward (T=10) on the launch stream, against the algorithmic FLOPs of those launches.
        roof = conv_roofline(model, ops, dev, pk)
        fwd_tflops = GF_PER_FRAME_FWD * 1e9 * (10 + 9) * SAMPLER_STEPS * world * args.steps / inner / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, desc = time_cpu(1, 1)
            cpu = {"value": fps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc,
                   "sample_s_per_step": dt}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": inner / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(), "frames_per_gpu": FRAMES, "size": SIZE,
                       "sampler_steps": SAMPLER_STEPS, "l2": "activations per forward (2.9 GB) exceed the 126 MB L2",
                       "operands": "fp16 tcgen05 (fp32 accumulate), same tensor rate as bf16",
                       "unet_fwd_tflops_algorithmic": fwd_tflops, "unet_fwd_frac_of_peak": fwd_tflops / pk["tflops"] / world},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": lr_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(model, ops, dev, pk):
    """Roofline of the dominant kernel (conv_igemm_kernel, ~60 % of the device time of a forward).

    Every flair_conv_igemm launch of one video-mode forward (T=10, 256x256) is recorded with its arguments during an
    eager forward, then the same launches are captured back to back into a CUDA graph and the graph is timed with
    CUDA events on its stream: achieved = sum of the algorithmic FLOPs of those launches / device time of one
    replay.  (Timing them inside the eager forward would charge the host-side launch cost of this CPU-bound eager
    mode to the kernel.)  Inputs are the real activations of that forward; between two launches of the same
    buffer ~3 GB of other maps stream through the 126 MB L2, so operands are not artificially cache-hot."""
    from flair_b200 import synth
    T = 10
    rec = []
    real = ops.conv

    def recording_conv(x, wpk, cout, ksize=(1, 3, 3), **kw):
        y = real(x, wpk, cout, ksize, **kw)
        B, F_, H, W, cin = x.shape
        s = kw.get("stride", 1)
        kw2 = dict(kw)
        kw2["out"] = y if not kw.get("nchw_out") else None
        if kw2["out"] is None:
            kw2.pop("out")
        taps = ksize[0] * ksize[1] * ksize[2]
        nbytes = x.numel() * x.element_size() + y.numel() * y.element_size() + taps * cout * cin * 2
        for r in ("residual", "residual2"):
            if kw.get(r) is not None:
                nbytes += y.numel() * kw[r].element_size()
        rec.append(((x, wpk, cout, ksize), kw2, 2.0 * B * F_ * (H // s) * (W // s) * cin * cout * taps, nbytes))
        return y

    clip = (synth.synthetic_clip(T, SIZE) * 2 - 1).to(dev)
    x = torch.randn(T, 3, SIZE, SIZE, device=dev)
    ts = torch.full((T,), 500, device=dev)
    graph_flag = model.use_cuda_graph
    model.use_cuda_graph = False
    import guided_diffusion.unet_new as U
    try:
        model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)  # warm
        ops.conv = recording_conv
        U.ops.conv = recording_conv
        model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
        torch.cuda.synchronize()
    finally:
        ops.conv = real
        U.ops.conv = real
        model.use_cuda_graph = graph_flag

    def replay_all():
        for a, kw, _, _ in rec:
            real(*a, **kw)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        replay_all()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        replay_all()
    g.replay()
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) / 1e3 / reps
    flops = sum(f for _, _, f, _ in rec)
    alg_bytes = sum(nb for _, _, _, nb in rec)
    achieved = flops / secs / 1e12
    # DRAM bytes per launch of the same 1622 launches from an ncu capture (profiles/r01_conv_traffic.json; ncu cannot
    # run inside the bench).  Next to it: the algorithmic bytes (operands + result + residuals, each once).
    traffic = traffic_note = None
    tj = ROOT / "profiles" / "r01_conv_traffic.json"
    if tj.exists():
        t = json.loads(tj.read_text())
        traffic = t["dram_bytes_per_launch"]
        traffic_note = {"source": "profiles/r01_conv_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                  "average over the conv launches of one forward, cold-cache)",
                        "l2_to_sm_bytes_per_launch": t["l2_to_sm_bytes_per_launch"]}
    return {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05/TMA implicit GEMM)", "achieved": achieved,
            "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
            "traffic_note": traffic_note, "alg_bytes_per_launch": alg_bytes / len(rec),
            "peak_source": pk["src"] + " (bf16_tflops_sustained; fp16 and bf16 share the kind::f16 pipe)",
            "launches": len(rec), "avg_launch_us": secs / len(rec) * 1e6, "alg_gflop_per_forward": flops / 1e9,
            "how": "all conv launches of one video-mode forward (T=10, 256x256) replayed back to back from a CUDA "
                   "graph, CUDA events on the replay stream; algorithmic FLOPs = 2*M*Cout*Cin*taps per launch"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
