#!/usr/bin/env python
"""bench.py — FLAIR hot-path benchmark on B200 (driver contract: task prompt / DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W                 # this repo's sm_100a path, BASELINE.json configs[1]
  python bench.py --impl reference --steps K --warmup W         # the reference path's CPU port (oracle/), config 1
  python bench.py --workload {gaussian,jpeg,x8,x16} --mode {replicas,sharded}
  python bench.py --workload fwd-sweep                          # config 5b: image-mode UNet forward, 1..64 frames

Workloads (BASELINE.json `configs`):
  gaussian  [1] Gaussian-blur restoration of a synthetic 16-frame 256x256 clip, full 100-step respaced sampler, blur
                UNet in video mode, two chained 10-frame windows (overlap 3, scripts/video_sample.py:202-203,361-368).
                One *step* = one full restoration of the clip (200 UNet forwards + 200 data-consistency/update steps).
  x8        [2] x8 bicubic SR of a 32-frame clip (SR3 UNet);  x16 [3] x16 bicubic SR of a 64-frame clip;
  jpeg      [4] DCT-JPEG artifact removal of a 64-frame clip;  fwd-sweep [4b] UNet-forward-only batch sweep.

Scaling modes (N > 1, torchrun, one rank per GPU):
  replicas  WEAK scaling: every rank restores its own clip of the workload, no data-path message; value = all frames
            of all ranks / max-over-ranks device time.  This is the line's `value` (`"scaling": "weak"`).
  sharded   STRONG scaling: ONE clip lives on rank 0; its 10-frame windows (overlap 2 -> 32 frames = 4 windows,
            64 frames = 8 windows) are scattered over the ranks with NCCL send/recv (each rank receives its
            windows' frames including the 2-frame halo it shares with the neighbour), restored independently, and the
            restored frames are gathered and stitched on rank 0 — scatter, compute and gather all inside the timed
            region.  Under `--mode auto` (default) with N > 1 the line additionally carries `"strong": [...]`: the
            BASELINE config quoted for that N (2/4 GPUs: x8 32 frames; 8 GPUs: x16 and jpeg 64 frames), one warm-up +
            one timed step each.
Timing: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks.
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
SIZE, SAMPLER_STEPS = 256, 100
# algorithmic dense-contraction FLOPs per frame of one forward (hooks on the reference, SPyNet cached)
GF_VIDEO = {"blur": 2370.2, "sr3": 988.1}
GF_IMAGE = {"blur": 248.0, "sr3": 308.3}
WORKLOADS = {
    "gaussian": dict(task="gaussian", frames=16, model="blur", cfg=1),
    "x8": dict(task="x8_bicubic", frames=32, model="sr3", cfg=2),
    "x16": dict(task="x16_bicubic", frames=64, model="sr3", cfg=3),
    "jpeg": dict(task="jpeg", frames=64, model="blur", cfg=4),
}
STRONG_BY_N = {2: ["x8"], 4: ["x8"], 8: ["x16", "jpeg"]}   # BASELINE.json configs 3-5
SHARD_OVERLAP = 2                                        # 32 / 64 frames -> exactly 4 / 8 windows


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops=d.get("bf16_tflops_sustained", 1400.0), tflops_burst=d.get("bf16_tflops", 1590.0),
                    hbm=d.get("hbm_gbs", 6650.0), src="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


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


def workload_name(wl="gaussian"):
    w = WORKLOADS[wl]
    desc = {"gaussian": "gaussian-demo: blur restoration", "jpeg": "jpeg-demo: DCT-JPEG (qf 60) artifact removal",
            "x8": "x8-bicubic-demo: x8 bicubic super-resolution", "x16": "x16-bicubic-demo: x16 bicubic super-resolution"}[wl]
    unet = "blur UNet" if w["model"] == "blur" else "SR3 UNet (image_size=256 layout)"
    return (f"{desc} of a synthetic {w['frames']}-frame {SIZE}x{SIZE} clip, full {SAMPLER_STEPS}-step respaced sampler, "
            f"{unet} in video mode, 10-frame windows, synthetic weights (BASELINE.json configs[{w['cfg']}])")


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference path's CPU port (oracle/) on SURVEY 8(d)'s config 1 — exactly, no extrapolation
# ------------------------------------------------------------------------------------------------
def cpu_config1_factory():
    """BASELINE.json configs[0]: ONE UNetModel denoising forward (image mode: num_frames=1, cross-frame modules off)
    + ONE gaussian-demo data-consistency step + the sampler update on ONE 256x256 frame, random-init weights, fp32,
    all host cores.  Returns (step_fn, description)."""
    from flair_b200 import synth
    from oracle import degrade, sampler
    from oracle.schedule import Tables
    from oracle.unet_blur import BlurUNetOracle, default_config
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_grad_enabled(False)
    cfg = default_config(SIZE)
    keys = _blur_unet_keys(cfg)
    sd = {k: synth.synthetic_tensor(k, shp, 1234) for k, shp in keys.items()}
    model = BlurUNetOracle(cfg, sd)
    taps = torch.load(ROOT / "tests" / "golden" / "pseudosr_taps.pt", map_location="cpu", weights_only=True)
    ds, inv = taps["ds_kernel"].float(), taps["inv_hTh"].float()
    tab = Tables("face_blur", 1000)
    hr = synth.synthetic_clip(1, SIZE, seed=1) * 2 - 1
    y = degrade.blur_down(hr, ds)
    init = torch.nn.functional.interpolate((y + 1) / 2, (SIZE, SIZE), mode="area").clamp(0, 1) * 2 - 1
    tape = synth.noise_tape((1, 3, SIZE, SIZE), 1, seed=2)
    t = 50
    x_t = sampler.q_sample(tab, init, t, tape[0])
    gam = tab.gammas(1.0, 2.55)[t]

    def step():
        eps = model.forward(x_t, torch.full((1,), tab.timestep_map[t]), init[:, None], num_frames=1,
                            enable_cross_frames=False)
        return sampler.p_sample_step(tab, x_t, eps, t, tape[1], lambda v: degrade.blur_restore(v, y, ds, inv),
                                     gamma=gam, rho=0.25)[0]

    desc = (f"BASELINE.json configs[0] exactly: 1 UNetModel forward (image mode) + 1 gaussian data-consistency step + "
            f"sampler update on ONE {SIZE}x{SIZE} frame, fp32, torch {torch.__version__} on {torch.get_num_threads()} "
            f"host threads, oracle port of the reference path; frames/s = 1 / ({SAMPLER_STEPS} steps x step time); "
            "image mode is 248 GF/frame-step against 2370 GF/frame-step of the video-mode path the GPU arm runs")
    return step, desc


def _blur_unet_keys(cfg):
    """State-dict key -> shape of the blur UNet for `cfg` (from this repo's module tree; CPU, no kernels)."""
    from guided_diffusion.unet_new import UNetModel
    m = UNetModel(**cfg, use_fp16=False)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def time_cpu(steps, warmup, best_of=False):
    step, desc = cpu_config1_factory()
    for _ in range(max(warmup, 1)):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    dt = min(times) if best_of else sum(times) / len(times)
    return 1.0 / (SAMPLER_STEPS * dt), dt, desc, times


def reference_video_note():
    """A real measurement of the UNMODIFIED reference's video-mode forward (T=10, 256x256), taken while the parity
    fixture was generated in the build container (tools/gen_golden_big.py blur256): context only."""
    p = ROOT / "tests" / "golden" / "unet_blur_256.pt"
    try:
        fx = torch.load(p, map_location="cpu", weights_only=True)
        return {"reference_video_forward_T10_256_cpu_s": round(float(fx["cpu_seconds"]), 1), "cores": int(fx["cores"]),
                "where": "build container, tools/gen_golden_big.py blur256 (unmodified /root/reference)"}
    except Exception:
        return None


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    fps, dt, desc, times = time_cpu(args.steps, max(args.warmup, 1))
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("gaussian"),
                   "note": "CPU arm: each step is the bounded sample described in cpu_baseline.sample",
                   "reference_video_mode": reference_video_note()},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "step_seconds": times},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class Job:
    """Model + sampler + operator + synthetic degraded clip of one workload on one device."""

    def __init__(self, wl, dev, seed_offset=0):
        from flair_b200 import pipeline, synth
        from guided_diffusion.script_util import create_model
        w = WORKLOADS[wl]
        self.wl, self.task, self.frames, self.kind, self.dev = wl, w["task"], w["frames"], w["model"], dev
        self.model = create_model("face_blur" if self.kind == "blur" else "face_bicubic", SIZE)
        self.model.load_state_dict(synth.synthetic_state_dict(self.model, seed=1234))
        self.model.convert_to_fp16()
        self.model.eval().to(dev)
        self.diffusion = pipeline.make_diffusion(self.task)
        kern = np.load(ROOT / "flair_b200" / "data" / "blur_kernel_k03.npy") if self.kind == "blur" else None
        self.A = pipeline.make_operator(self.task, dev, SIZE, kernels_mat=kern)
        hr = synth.synthetic_clip(self.frames, SIZE, seed=1 + seed_offset).to(dev) * 2 - 1
        if self.kind == "blur":
            lr = self.A.DownscaleOP(hr)
            if self.task == "jpeg":
                from guided_diffusion.jpeg import jpeg_roundtrip
                lr = jpeg_roundtrip(lr, pipeline.KNOBS["jpeg"].jpeg_qf)
        else:
            f = pipeline.KNOBS[self.task].factor
            lr = self.A.A(hr.reshape(self.frames, -1)).reshape(self.frames, 3, SIZE // f, SIZE // f)
        self.lr_host = ((lr + 1) / 2).clamp(0, 1).cpu().pin_memory()
        self.out_host = torch.empty(self.frames, 3, SIZE, SIZE, dtype=torch.float32).pin_memory()
        self.gen = torch.Generator(device=dev).manual_seed(2 + seed_offset)

    def restore_chained(self, lr_dev):
        """The reference script's semantics: chained windows, overlap 3."""
        from flair_b200 import pipeline
        return pipeline.restore_clip(self.model, self.diffusion, self.A, self.task, lr_dev, image_size=SIZE,
                                     chained=True, generator=self.gen)

    def restore_window(self, lr_win, widx):
        """One independent window (sharded mode): noise keyed by the window index, so the stitched clip does not
        depend on the number of ranks."""
        from flair_b200 import pipeline
        g = torch.Generator(device=self.dev).manual_seed(1000 + widx)
        return pipeline.restore_window(self.model, self.diffusion, self.A, self.task, lr_win, image_size=SIZE,
                                       generator=g)


def _launch_count():
    """C-ABI launch calls issued so far by this process (a replayed CUDA graph adds the launches it captured)."""
    from flair_b200 import _lib as L
    return L.LAUNCHES[0]


def _timed_steps(n_steps, barrier, body):
    """body(k, ev) records ev[0..3] around H2D / compute / D2H; returns (inner_s, outer_s) summed over steps."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_steps)]
    barrier()
    for k in range(n_steps):
        body(k, ev[k])
    barrier()
    inner = sum(e[1].elapsed_time(e[2]) for e in ev) / 1e3
    outer = sum(e[0].elapsed_time(e[3]) for e in ev) / 1e3
    return inner, outer


def run_replicas(job, steps, warmup, world, barrier, reduce_max):
    dev = job.dev
    for _ in range(warmup):
        job.restore_chained(job.lr_host.to(dev, non_blocking=True))

    def body(k, e):
        e[0].record()
        lr_dev = job.lr_host.to(dev, non_blocking=True)      # H2D of this step's inputs (pinned)
        e[1].record()
        out = job.restore_chained(lr_dev)                     # inputs resident: the `value` region
        e[2].record()
        job.out_host.copy_(out, non_blocking=True)            # D2H of the restored frames
        e[3].record()

    n0 = _launch_count()
    inner, outer = reduce_max(*_timed_steps(steps, barrier, body))
    launches = _launch_count() - n0                          # C-ABI launches of THIS rank inside the timed region
    total = job.frames * world * steps
    return dict(value=total / inner, e2e=total / outer, ms_per_step=inner / steps * 1e3,
                h2d=job.lr_host.numel() * 4, d2h=job.out_host.numel() * 4, launches=launches)


def run_sharded(job, steps, warmup, world, rank, barrier, reduce_max):
    """One clip on rank 0, windows (overlap 2) scattered / restored / gathered; everything inside the timed region."""
    from flair_b200 import parallel
    dev = job.dev
    tail = tuple(job.lr_host.shape[1:])
    stats = {}

    def one(lr_dev):
        return parallel.restore_clip_windows(job.restore_window, lr_dev, job.frames, dev, tail,
                                             overlap=SHARD_OVERLAP, stats=stats)

    for _ in range(warmup):
        one(job.lr_host.to(dev, non_blocking=True) if rank == 0 else None)

    def body(k, e):
        e[0].record()
        lr_dev = job.lr_host.to(dev, non_blocking=True) if rank == 0 else None
        e[1].record()
        out = one(lr_dev)                                     # scatter (NCCL p2p) + windows + gather (NCCL p2p)
        e[2].record()
        if rank == 0:
            job.out_host.copy_(out, non_blocking=True)
        e[3].record()

    n0 = _launch_count()
    inner, outer = reduce_max(*_timed_steps(steps, barrier, body))
    launches = _launch_count() - n0
    total = job.frames * steps
    return dict(value=total / inner, e2e=total / outer, ms_per_step=inner / steps * 1e3,
                h2d=job.lr_host.numel() * 4, d2h=job.out_host.numel() * 4, launches=launches,
                p2p_bytes_per_step=stats.get("p2p_bytes", 0), windows=stats.get("windows"),
                windows_per_rank_max=stats.get("windows_per_rank_max"),
                compute_s_max_rank=reduce_max(stats.get("compute_s", 0.0), 0.0)[0])


def run_native(args):
    import torch.distributed as dist
    from flair_b200 import _lib as L
    from flair_b200 import ops as _ops

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(a, b):
        if world > 1:
            t = torch.tensor([a, b], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0]), float(t[1])
        return a, b

    if args.workload == "fwd-sweep":
        if rank == 0:
            print(json.dumps(fwd_sweep_line(args, dev)))
        if world > 1:
            dist.destroy_process_group()
        return

    mode = args.mode if args.mode != "auto" else "replicas"
    job = Job(args.workload, dev, seed_offset=rank if mode == "replicas" else 0)
    with ClockSampler(local) as clocks:
        if mode == "replicas":
            res = run_replicas(job, args.steps, args.warmup, world, barrier, reduce_max)
        else:
            res = run_sharded(job, args.steps, args.warmup, world, rank, barrier, reduce_max)
    launches = res["launches"]      # rank 0's launches inside the timed region (warm-up steps excluded)

    strong = []
    if args.mode == "auto" and world > 1:
        jobs = {args.workload: job}
        for wl in STRONG_BY_N.get(world, []):
            j = jobs.get(wl) or Job(wl, dev)
            r = run_sharded(j, 1, 1, world, rank, barrier, reduce_max)
            strong.append({"workload": workload_name(wl), "scaling": "strong", "n_gpus": world,
                           "window_overlap": SHARD_OVERLAP, "value": r["value"], "unit": UNIT,
                           "e2e": r["e2e"], "s_per_clip": r["ms_per_step"] / 1e3, "windows": r["windows"],
                           "windows_per_rank_max": r["windows_per_rank_max"],
                           "nccl_p2p_bytes_per_clip": r["p2p_bytes_per_step"],
                           "compute_s_max_rank": r["compute_s_max_rank"],
                           "parallel_efficiency_vs_compute": (r["compute_s_max_rank"] / (r["ms_per_step"] / 1e3))
                           if r["ms_per_step"] else None})
            if j is not job:
                del j
                torch.cuda.empty_cache()

    if rank == 0:
        pk = peaks()
        w = WORKLOADS[args.workload]
        # frame evaluations per clip: chained windows of <= 10 frames with overlap 3 (replicas) / 2 (sharded)
        from flair_b200.pipeline import windows as _windows
        ov = 3 if mode == "replicas" else SHARD_OVERLAP
        fevals = sum(b - a for a, b in _windows(w["frames"], 10, ov))
        clips = (world if mode == "replicas" else 1) * args.steps
        secs = res["ms_per_step"] / 1e3 * args.steps
        fwd_tflops = GF_VIDEO[w["model"]] * 1e9 * fevals * SAMPLER_STEPS * clips / secs / 1e12
        roof = conv_roofline(job, dev, pk)
        roof["forward"] = {"achieved": fwd_tflops / world, "unit": "TFLOP/s per GPU", "frac": fwd_tflops / pk["tflops"] / world,
                           "what": "algorithmic dense-contraction FLOPs of every UNet forward of the timed region "
                                   f"({GF_VIDEO[w['model']]} GF per frame per step, SURVEY 8d) / device time of the "
                                   "WHOLE region (all kernels, sampler tail and host gaps included)"}
        hbm = hbm_rooflines(job, dev, pk)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, desc, times = time_cpu(3, 1, best_of=True)
            cpu = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": desc,
                   "step_seconds_best_of_3": dt, "reference_video_mode": reference_video_note()}
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if mode == "replicas" else "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "mode": mode,
                       "frames_per_gpu": w["frames"] if mode == "replicas" else w["frames"] / world, "size": SIZE,
                       "sampler_steps": SAMPLER_STEPS, "frame_evaluations_per_clip": fevals,
                       "l2": "inputs larger than L2: activations of one forward (2.9 GB) exceed the 126 MB L2",
                       "operands": "fp16 tcgen05 (fp32 accumulate), same tensor rate as bf16",
                       "step_graph": os.environ.get("FLAIR_STEP_GRAPH", "1") != "0",
                       "fused_gn_statistics": bool(_ops.FUSED_GN),
                       "unet_fwd_tflops_algorithmic": fwd_tflops,
                       "unet_fwd_frac_of_peak": fwd_tflops / pk["tflops"] / world},
            "clocks": clocks.summary(),
            "e2e": {"value": res["e2e"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"]},
            "gpu_launches": launches, "roofline": roof, "roofline_hbm": hbm, "cpu_baseline": cpu,
        }
        if mode == "sharded":
            line["config"].update(window_overlap=SHARD_OVERLAP, windows=res["windows"],
                                  nccl_p2p_bytes_per_step=res["p2p_bytes_per_step"],
                                  compute_s_max_rank=res["compute_s_max_rank"])
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_sweep and w["model"] == "blur":
            line["config"]["fwd_sweep_image_mode"] = fwd_sweep(job.model, dev, pk, (1, 8, 64))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# rooflines
# ------------------------------------------------------------------------------------------------
def _graph_time(fns, reps=3):
    """Device seconds per pass over `fns`, captured into one CUDA graph (no host launch cost), CUDA events."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps


def conv_roofline(job, dev, pk):
    """Roofline of the dominant kernel (conv_igemm_kernel, ~60 % of the device time of a forward).

    Every flair_conv_igemm launch of one video-mode forward (T=10, 256x256) is recorded with its arguments during an
    eager forward, then the same launches are captured back to back into a CUDA graph and the graph is timed with
    CUDA events on its stream: achieved = sum of the algorithmic FLOPs of those launches / device time of one
    replay.  Inputs are the real activations of that forward; between two launches of the same buffer ~3 GB of other
    maps stream through the 126 MB L2, so operands are not artificially cache-hot."""
    from flair_b200 import ops, synth
    model = job.model
    T = 10
    rec = []
    real = ops.conv

    def recording_conv(x, wpk, cout, ksize=(1, 3, 3), **kw):
        y = real(x, wpk, cout, ksize, **kw)
        B, F_, H, W, cin = x.shape
        s = kw.get("stride", 1)
        kw2 = dict(kw)
        if not kw.get("nchw_out"):
            kw2["out"] = y
        taps = ksize[0] * ksize[1] * ksize[2]
        nbytes = x.numel() * x.element_size() + y.numel() * y.element_size() + taps * cout * cin * 2
        for r in ("residual", "residual2", "preadd"):
            if kw.get(r) is not None:
                nbytes += y.numel() * kw[r].element_size()
        rec.append(((x, wpk, cout, ksize), kw2, 2.0 * B * F_ * (H // s) * (W // s) * cin * cout * taps, nbytes))
        return y

    clip = (synth.synthetic_clip(T, SIZE) * 2 - 1).to(dev)
    x = torch.randn(T, 3, SIZE, SIZE, device=dev)
    ts = torch.full((T,), 500, device=dev) if job.kind == "blur" else torch.full((T,), 0.5, device=dev)
    graph_flag = model.use_cuda_graph
    model.use_cuda_graph = False
    import guided_diffusion.sr3 as S3
    import guided_diffusion.unet_new as U
    try:
        model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)  # warm
        ops.conv = U.ops.conv = S3.ops.conv = recording_conv
        model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
        torch.cuda.synchronize()
    finally:
        ops.conv = U.ops.conv = S3.ops.conv = real
        model.use_cuda_graph = graph_flag

    secs = _graph_time([lambda a=a, kw=kw: real(*a, **kw) for a, kw, _, _ in rec])
    flops = sum(f for _, _, f, _ in rec)
    alg_bytes = sum(nb for _, _, _, nb in rec)
    achieved = flops / secs / 1e12
    # DRAM bytes per launch from the committed ncu capture of the conv launches of one forward (ncu cannot run inside
    # the bench): newest profiles/r*_conv_traffic.json.  Next to it: algorithmic bytes (operands + result + addends once).
    traffic = traffic_note = None
    cands = sorted((ROOT / "profiles").glob("r*_conv_traffic.json"))
    if cands:
        t = json.loads(cands[-1].read_text())
        traffic = t["dram_bytes_per_launch"]
        traffic_note = {"source": f"profiles/{cands[-1].name} (ncu dram__bytes_read.sum + dram__bytes_write.sum, average "
                                  "over the conv launches of one forward, cold-cache)",
                        "l2_to_sm_bytes_per_launch": t.get("l2_to_sm_bytes_per_launch")}
    return {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05/TMA implicit GEMM)", "achieved": achieved,
            "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
            "traffic_note": traffic_note, "alg_bytes_per_launch": alg_bytes / len(rec),
            "peak_source": pk["src"] + ": bf16_tflops_sustained (kernel timed inside a long replay); fp16 and bf16 "
                                       "share the kind::f16 pipe",
            "launches": len(rec), "avg_launch_us": secs / len(rec) * 1e6, "alg_gflop_per_forward": flops / 1e9,
            "how": "all conv launches of one video-mode forward (T=10, 256x256) replayed back to back from a CUDA "
                   "graph, CUDA events on the replay stream; algorithmic FLOPs = 2*M*Cout*Cin*taps per launch"}


def hbm_rooflines(job, dev, pk):
    """Achieved HBM GB/s of the memory-bound kernels, measured in this run: algorithmic bytes / CUDA-event time of
    graph replays over 4 rotated buffer sets (each > L2, so no launch re-reads cached data).  Reported against the
    measured copy peak (MEASURED_PEAKS.json) and against the 8 TB/s datasheet figure BASELINE.md quotes."""
    from flair_b200 import ops
    out = []
    R = 4

    def add(name, secs_per_call, nbytes, note):
        gbs = nbytes / secs_per_call / 1e9
        out.append({"kernel": name, "bound": "hbm", "achieved": gbs, "unit": "GB/s", "peak": pk["hbm"],
                    "frac": gbs / pk["hbm"], "frac_of_8000": gbs / 8000.0, "alg_bytes_per_launch": nbytes,
                    "launch_us": secs_per_call * 1e6, "what": note})

    T, H, C = 10, SIZE, 64
    xs = [torch.randn(1, T, H, H, C, device=dev).half() for _ in range(R)]
    outs = [torch.empty_like(x) for x in xs]
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    film = torch.randn(T, 2 * C, device=dev) * 0.1
    n = xs[0].numel()
    stats = [ops.gn_stats(x) for x in xs]
    add("gn_stats_kernel", _graph_time([lambda x=x: ops.gn_stats(x) for x in xs]) / R, n * 2,
        f"GroupNorm32 statistics of a [1,{T},{H},{H},{C}] fp16 map (2 B/element read)")
    add("gn_apply_fast_kernel", _graph_time([lambda x=x, s=s, o=o: ops.gn_apply(
        x, s, gamma, beta, scale=film[:, :C], shift=film[:, C:], silu=True, out=o) for x, s, o in zip(xs, stats, outs)]) / R,
        n * 4, "normalise + FiLM + SiLU of the same map (2 B read + 2 B written per element)")
    del xs, outs, stats
    N = 64
    coef = job.diffusion._table(dev, "update")
    t = torch.full((N,), 50, device=dev, dtype=torch.long)
    sets = [dict(x=torch.randn(N, 3, SIZE, SIZE, device=dev), mo=torch.randn(N, 6, SIZE, SIZE, device=dev),
                 z=torch.randn(N, 3, SIZE, SIZE, device=dev)) for _ in range(R)]
    plane = N * 3 * SIZE * SIZE * 4
    add("pred_xstart_kernel", _graph_time([lambda d=d: ops.pred_xstart(d["x"], d["mo"], coef, t_arr=t) for d in sets]) / R,
        3 * plane, f"x0 from eps at {N} frames per launch (read x_t, eps; write x0)")
    add("sampler_update_kernel", _graph_time([lambda d=d: ops.sampler_update(
        d["x"], coef, model_out=d["mo"], noise=d["z"], t_arr=t, rho=0.25) for d in sets]) / R, 5 * plane,
        f"fused p_sample update at {N} frames per launch (read x_t, eps, z; write x_(t-1), x0): 3.15 MB/frame + x0")
    if job.kind == "blur":
        A = job.A
        taps_up = A.Upscale_OP.Filter_OP.taps
        sf, pre = int(A.ds_factor), int(A.pre_stride[0])
        gam = torch.full((N,), 0.5, device=dev)
        lrs = [torch.randn(N, 3, SIZE // 4, SIZE // 4, device=dev) for _ in range(R)]
        add("sampler_update_kernel (+ in-register Up)", _graph_time([lambda d=d, q=q: ops.sampler_update(
            d["x"], coef, model_out=d["mo"], noise=d["z"], t_arr=t, gamma_arr=gam, q_lr=q, up_taps=taps_up, sf=sf,
            pre_stride=pre, rho=0.25) for d, q in zip(sets, lrs)]) / R, int(N * 3.93e6),
            f"fused update + blur data consistency (R = Up(q) in-register) at {N} frames: 3.93 MB/frame (SURVEY 8d)")
        taps_ds = A.DownscaleOP.Filter_OP.taps
        x0s = [d["x"].clamp(-1, 1) for d in sets]
        add("blur_down_tiled_kernel", _graph_time([lambda v=v: ops.blur_down(v, taps_ds, sf, pre) for v in x0s]) / R,
            plane + plane // 16, f"9x9 blur + /4 decimation at {N} frames (read x0; write the 64x64 LR map)")
    return out


def fwd_sweep(model, dev, pk, batches):
    """BASELINE.json configs[4b]: UNet-forward-only batch sweep in IMAGE mode (num_frames=1, cross-frame modules off,
    reference unet_new.py:1311-1362): frames per launch chain -> algorithmic TFLOP/s (248.0 GF per frame)."""
    from flair_b200 import synth
    rows = []
    for n in batches:
        clip = (synth.synthetic_clip(n, SIZE) * 2 - 1).to(dev)
        x = torch.randn(n, 3, SIZE, SIZE, device=dev)
        ts = torch.full((n,), 500, device=dev)
        f = lambda: model(x, ts, low_res_input=clip[:, None], num_frames=1, enable_cross_frames=False)
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        reps = 3 if n >= 16 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = GF_IMAGE["blur"] * n / ms   # GF per ms == TFLOP/s
        rows.append({"frames": n, "ms": ms, "tflops": tf, "frac_of_peak": tf / pk["tflops"]})
    return rows


def fwd_sweep_line(args, dev):
    from flair_b200 import synth
    from guided_diffusion.script_util import create_model
    pk = peaks()
    model = create_model("face_blur", SIZE)
    model.load_state_dict(synth.synthetic_state_dict(model, seed=1234))
    model.convert_to_fp16()
    model.eval().to(dev)
    with ClockSampler(dev.index or 0) as clocks:
        rows = fwd_sweep(model, dev, pk, (1, 2, 4, 8, 16, 32, 64))
    best = max(rows, key=lambda r: r["tflops"])
    return {"metric": "unet_fwd_tflops_image_mode", "value": best["tflops"], "unit": "TFLOP/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "UNet-forward-only batch sweep, image mode (num_frames=1), 1-64 frames of 256x256, blur "
                                   "UNet, synthetic weights (BASELINE.json configs[4], second half); 248.0 GF per frame",
                       "sweep": rows},
            "roofline": {"bound": "tensor", "achieved": best["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": best["tflops"] / pk["tflops"], "traffic": None,
                         "what": f"whole image-mode forward at {best['frames']} frames (all kernels)"},
            "clocks": clocks.summary()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["fwd-sweep"], default="gaussian")
    ap.add_argument("--mode", choices=["auto", "replicas", "sharded"], default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
