"""2-GPU probe (torchrun): one 16-frame clip sharded over two ranks with flair_b200.parallel — NCCL point-to-point
scatter of the degraded segments (incl. the 3 overlap frames), per-rank windowed sampling, NCCL gather + stitching —
checked on rank 0 against the same segments restored locally with the same seeds.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_probes/sharded_probe.py
"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from flair_b200 import parallel, pipeline, synth
from guided_diffusion.script_util import blur_unet_config
from guided_diffusion.unet_new import UNetModel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
torch.set_grad_enabled(False)
S, N, T_START = 64, 16, 3
cfg = dict(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
           attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
           resblock_updown=True, use_scale_shift_norm=True, temporal_block=True, use_fp16=True)
model = UNetModel(**cfg)
model.load_state_dict(synth.synthetic_state_dict(model, seed=99)); model.convert_to_fp16(); model.eval().to(dev)
diffusion = pipeline.make_diffusion("gaussian")
A = pipeline.make_operator("gaussian", dev, S, kernels_mat=np.load("flair_b200/data/blur_kernel_k03.npy"))


def restore_segment(seed):
    def fn(lr_seg):
        g = torch.Generator(device=dev).manual_seed(seed)
        return pipeline.restore_clip(model, diffusion, A, "gaussian", lr_seg, image_size=S, chained=True, generator=g,
                                     t_start=T_START)
    return fn


lr01 = None
if rank == 0:
    hr = synth.synthetic_clip(N, S, seed=5).to(dev)
    lr01 = ((A.DownscaleOP(hr * 2 - 1) + 1) / 2).clamp(0, 1)
plan = parallel.segment_plan(N, world)
out = parallel.restore_clip_sharded(restore_segment(100 + rank), lr01, N, dev, (3, S // 4, S // 4))
if rank == 0:
    assert out.shape == (N, 3, S, S) and bool(torch.isfinite(out).all())
    ok = True
    for r in range(world):
        a, b, drop = plan[r]
        if b <= a:
            continue
        ref = restore_segment(100 + r)(lr01[a:b])[drop:]
        got = out[a + drop:b]
        d = float((got - ref).abs().max())
        print(f"rank {r}: frames [{a},{b}) drop {drop}: max |sharded - local| = {d:.3e}")
        ok &= d == 0.0
    print("SHARDED OK" if ok else "SHARDED MISMATCH", "plan", plan)
dist.barrier()
dist.destroy_process_group()
