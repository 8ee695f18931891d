"""N > 1 host logic on CPU: world_size-2 gloo run of the segment split / P2P halo scatter / stitch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flair_b200 import parallel
from flair_b200.pipeline import windows


def test_segment_plan_covers_clip():
    for n, world in [(16, 2), (32, 4), (64, 8), (64, 2), (10, 4), (25, 3)]:
        plan = parallel.segment_plan(n, world)
        kept = sum(max(0, b - a - d) for a, b, d in plan)
        assert kept == n, (n, world, plan)
        assert plan[0][0] == 0 and max(b for _, b, _ in plan) == n
    # OVERLAP=2 gives one window per GPU for 64 frames on 8 GPUs
    assert len(windows(64, 10, 2)) == 8
    assert all(b - a == 10 for a, b, _ in parallel.segment_plan(64, 8, 10, 2)[:-1])


def _fake_restore(seg):
    # deterministic per-frame function: stitching must reproduce the single-process result
    return seg * 2.0 + 1.0


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip = torch.arange(n_frames * 3 * 4 * 4, dtype=torch.float32).reshape(n_frames, 3, 4, 4) if rank == 0 else None
        out = parallel.restore_clip_sharded(_fake_restore, clip, n_frames, torch.device("cpu"), (3, 4, 4))
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [16, 31])
def test_two_rank_gloo_roundtrip(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_frames
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    clip = torch.arange(n_frames * 3 * 4 * 4, dtype=torch.float32).reshape(n_frames, 3, 4, 4)
    assert torch.equal(out, _fake_restore(clip))


# ------------------------------------------------------------------------------------------------ window granularity
def test_window_plan_is_world_size_independent():
    """Strong-scaling mode: independent windows (overlap 2).  32 frames = 4 windows, 64 frames = 8; every
    frame is produced exactly once whatever the number of ranks."""
    assert len(windows(32, 10, 2)) == 4 and len(windows(64, 10, 2)) == 8
    for n, world in [(64, 8), (64, 1), (32, 2), (32, 4), (16, 3), (10, 4)]:
        plan = parallel.window_plan(n, world, 10, 2)
        flat = [w for ws in plan for w in ws]
        assert [w[0] for w in flat] == list(range(len(flat)))          # contiguous, in order
        kept = sum((b - a) - (0 if i == 0 else 2) for i, a, b in flat)
        assert kept == n, (n, world, plan)
    assert all(len(ws) == 1 for ws in parallel.window_plan(64, 8, 10, 2))


def _fake_window(seg, widx):
    # depends on the window index like the per-window noise seed does; maps [0,1] input to [-1,1]
    return seg * 2.0 - 1.0 + 0.001 * widx


def _worker_windows(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip = torch.linspace(0, 1, n_frames * 3 * 4 * 4).reshape(n_frames, 3, 4, 4) if rank == 0 else None
        stats = {}
        out = parallel.restore_clip_windows(_fake_window, clip, n_frames, torch.device("cpu"), (3, 4, 4), overlap=2,
                                            stats=stats)
        if rank == 0:
            q.put((out, stats))
        dist.barrier()
    finally:
        dist.destroy_process_group()


# (32, 4) and (64, 8) are the BASELINE strong-scaling layouts bench.py --gpus 4 / 8 runs: one window per rank
@pytest.mark.parametrize("n_frames,world", [(32, 2), (26, 3), (32, 4), (64, 8)])
def test_window_sharding_matches_single_process(n_frames, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + n_frames + world
    procs = [ctx.Process(target=_worker_windows, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, stats = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    clip = torch.linspace(0, 1, n_frames * 3 * 4 * 4).reshape(n_frames, 3, 4, 4)
    single = parallel.restore_clip_windows(_fake_window, clip, n_frames, torch.device("cpu"), (3, 4, 4), overlap=2)
    assert torch.equal(out, single)          # same bits for any number of ranks
    assert stats["windows"] == len(windows(n_frames, 10, 2)) and stats["p2p_bytes"] > 0
    if (n_frames, world) in ((32, 4), (64, 8)):
        assert stats["windows_per_rank_max"] == 1
        # rank 0 sends every other rank its window (halo included) and receives it back minus the 2 shared frames
        wins = windows(n_frames, 10, 2)[1:]
        assert stats["p2p_bytes"] == sum((b - a) + (b - a - 2) for a, b in wins) * 3 * 4 * 4 * 4
