"""2-GPU probe (torchrun): window-granularity sharding (flair_b200.parallel.restore_clip_windows, what bench.py's
strong-scaling mode runs): a 18-frame clip = 2 windows of 10 frames with overlap 2, one per rank, NCCL scatter of the
degraded frames, independent sampling (noise keyed by the window index), NCCL gather + stitching on rank 0 — compared
bit for bit with the same clip restored by ONE process (world-size independence of the stitched result).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_probes/sharded_windows_probe.py
"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from flair_b200 import parallel, pipeline, synth
from guided_diffusion.unet_new import UNetModel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
torch.set_grad_enabled(False)
S, N, T_START, OV = 64, 18, 3, 2
cfg = dict(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
           attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
           resblock_updown=True, use_scale_shift_norm=True, temporal_block=True, use_fp16=True)
model = UNetModel(**cfg)
model.load_state_dict(synth.synthetic_state_dict(model, seed=99)); model.convert_to_fp16(); model.eval().to(dev)
diffusion = pipeline.make_diffusion("gaussian")
A = pipeline.make_operator("gaussian", dev, S, kernels_mat=np.load("flair_b200/data/blur_kernel_k03.npy"))


def restore_window(lr_win, widx):
    g = torch.Generator(device=dev).manual_seed(1000 + widx)   # noise keyed by the window, not by the rank
    return pipeline.restore_window(model, diffusion, A, "gaussian", lr_win, image_size=S, t_start=T_START, generator=g)


lr01 = None
if rank == 0:
    hr = synth.synthetic_clip(N, S, seed=5).to(dev)
    lr01 = ((A.DownscaleOP(hr * 2 - 1) + 1) / 2).clamp(0, 1)
stats = {}
out = parallel.restore_clip_windows(restore_window, lr01, N, dev, (3, S // 4, S // 4), overlap=OV, stats=stats)
if rank == 0:
    assert out.shape == (N, 3, S, S) and bool(torch.isfinite(out).all())
    # the same clip, every window on this process
    parts = []
    for widx, (a, b) in enumerate(pipeline.windows(N, 10, OV)):
        o = restore_window(lr01[a:b], widx)
        parts.append(((o if widx == 0 else o[OV:]).clamp(-1, 1) + 1) / 2)
    ref = torch.cat(parts, 0)
    d = float((out - ref).abs().max())
    print(f"windows {stats['windows']}, p2p bytes {stats['p2p_bytes']}, max |sharded - single process| = {d:.3e}")
    print("WINDOWS OK" if d == 0.0 else "WINDOWS MISMATCH")
dist.barrier()
dist.destroy_process_group()
