"""Multi-GPU sharding of one long clip: one process per GPU (torchrun), contiguous groups of the
reference's 10-frame windows per rank, NCCL point-to-point for the input halo (the overlap frames a
segment shares with its neighbour) and for collecting the restored frames.  No collective sits on the
data path: GroupNorm-over-T, Conv3d, TemporalAttention and the BasicVSR++ recurrence all
stay inside one window, hence inside one GPU.

Semantics: inside a segment the windows are chained exactly like the reference script
(`prev_recon`, scripts/video_sample.py:476-483); the first window of every segment is sampled
unconditioned (the reference's first-window behaviour) and, for segments after the first, its first
`overlap` frames are dropped at stitching time like any later window's."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist

from .pipeline import FRAME_SLICE_LEN, OVERLAP, windows


def segment_plan(n_frames: int, world: int, size: int = FRAME_SLICE_LEN, overlap: int = OVERLAP) -> List[Tuple[int, int, int]]:
    """Per rank (first_frame, end_frame, frames_to_drop_at_stitching).  Ranks beyond the number of
    windows get an empty segment (first == end)."""
    wins = windows(n_frames, size, overlap)
    per, extra = divmod(len(wins), world)
    plan, k = [], 0
    for r in range(world):
        cnt = per + (1 if r < extra else 0)
        if cnt == 0:
            plan.append((n_frames, n_frames, 0))
            continue
        a, b = wins[k][0], wins[k + cnt - 1][1]
        plan.append((a, b, overlap if k > 0 else 0))
        k += cnt
    return plan


def scatter_frames(lr01, plan, rank: int, world: int, device, shape_tail=None):
    """Rank 0 holds the (N,3,h,w) degraded clip; every rank returns its segment (with halo frames).
    Point-to-point only: rank 0 -> r for each r (NCCL send/recv on GPU, gloo on CPU)."""
    a, b, _ = plan[rank]
    if world == 1:
        return lr01[a:b]
    if rank == 0:
        reqs = []
        for r in range(1, world):
            ra, rb, _ = plan[r]
            if rb > ra:
                reqs.append(dist.isend(lr01[ra:rb].contiguous(), dst=r))
        for q in reqs:
            q.wait()
        return lr01[a:b]
    if b <= a:
        return torch.empty(0, *shape_tail, device=device)
    buf = torch.empty(b - a, *shape_tail, device=device)
    dist.recv(buf, src=0)
    return buf


def gather_frames(out_seg, plan, rank: int, world: int, n_frames: int):
    """Stitch on rank 0: segment r contributes its frames minus the ones it shares with segment r-1."""
    a, b, drop = plan[rank]
    mine = out_seg[drop:] if b > a else out_seg
    if world == 1:
        return mine
    if rank != 0:
        if b > a:
            dist.send(mine.contiguous(), dst=0)
        return None
    parts = [mine]
    for r in range(1, world):
        ra, rb, rdrop = plan[r]
        if rb <= ra:
            continue
        buf = torch.empty(rb - ra - rdrop, *mine.shape[1:], device=mine.device, dtype=mine.dtype)
        dist.recv(buf, src=r)
        parts.append(buf)
    out = torch.cat(parts, 0)
    assert out.shape[0] == n_frames, (out.shape, n_frames)
    return out


def restore_clip_sharded(restore_segment: Callable[[torch.Tensor], torch.Tensor], lr01, n_frames: int, device,
                         shape_tail, size: int = FRAME_SLICE_LEN, overlap: int = OVERLAP):
    """restore_segment(lr_segment) -> restored segment (same frame count).  Returns the stitched clip on
    rank 0 and None elsewhere.  With world == 1 this is restore_segment(lr01)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    plan = segment_plan(n_frames, world, size, overlap)
    seg = scatter_frames(lr01, plan, rank, world, device, shape_tail)
    out = restore_segment(seg) if seg.shape[0] > 0 else seg.new_empty(0)
    return gather_frames(out, plan, rank, world, n_frames)
