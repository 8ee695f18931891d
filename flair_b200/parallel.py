"""Multi-GPU sharding of one long clip: one process per GPU (torchrun), NCCL point-to-point for the input frames (each
rank receives the frames of its windows, i.e. including the `overlap`-frame halo it shares with its neighbours) and
for collecting the restored frames.  No collective sits on the data path: GroupNorm-over-T, Conv3d,
TemporalAttention and the BasicVSR++ recurrence all stay inside one 10-frame window, hence inside one GPU.

Two granularities:

* `restore_clip_windows` (what bench.py's strong-scaling mode uses, SURVEY 8e "recommended partitioning"): the clip is
  cut with the reference's own `windowed(frames, 10, step=10-overlap)`; EVERY window is sampled independently
  (`prev_recon=None`, the reference's first-window semantics) and contributes its frames minus the `overlap` it shares
  with the previous window (scripts/video_sample.py:481-485).  Contiguous windows go to each rank.  The noise of a
  window is keyed by its index by the caller, so the stitched clip is the same bits for ANY number of ranks.
  With overlap 2 a 32-frame clip is exactly 4 windows and a 64-frame clip exactly 8 (one per GPU of an 8-GPU box).

* `restore_clip_sharded` (segment granularity): contiguous groups of windows per rank, chained inside a segment like
  the reference script (`prev_recon`, scripts/video_sample.py:476-483); the first window of every segment is
  sampled unconditioned.  NOTE: its output depends on the world size (the windows at segment starts lose the hard
  conditioning on the previous window) — windows of one clip are strictly sequential in the reference, so exact
  chained semantics cannot be sharded; use one GPU (`pipeline.restore_clip(chained=True)`) for script parity."""
from __future__ import annotations

import time
from typing import Callable, List, Tuple

import torch
import torch.distributed as dist

from .pipeline import FRAME_SLICE_LEN, OVERLAP, windows


def _world():
    return (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)


# ------------------------------------------------------------------------------------------------ window granularity
def window_plan(n_frames: int, world: int, size: int = FRAME_SLICE_LEN, overlap: int = 2):
    """Per rank: list of (window_index, first_frame, end_frame).  Contiguous, as even as possible."""
    wins = windows(n_frames, size, overlap)
    per, extra = divmod(len(wins), world)
    plan, k = [], 0
    for r in range(world):
        cnt = per + (1 if r < extra else 0)
        plan.append([(k + j, *wins[k + j]) for j in range(cnt)])
        k += cnt
    return plan


def restore_clip_windows(restore_window: Callable[[torch.Tensor, int], torch.Tensor], lr01, n_frames: int, device,
                         shape_tail, size: int = FRAME_SLICE_LEN, overlap: int = 2, stats: dict | None = None):
    """restore_window(lr_window, window_index) -> restored window (same frame count, [-1,1]).  `lr01` (N,3,h,w) lives
    on rank 0 (None elsewhere).  Returns the stitched clip in [0,1] on rank 0 and None elsewhere."""
    world, rank = _world()
    plan = window_plan(n_frames, world, size, overlap)
    p2p = 0
    # ---- scatter: rank 0 -> r, the frame range spanned by r's windows (halo frames included)
    span = lambda ws: (ws[0][1], ws[-1][2]) if ws else (0, 0)
    a0, b0 = span(plan[rank])
    if world == 1:
        seg = lr01
    elif rank == 0:
        reqs = []
        for r in range(1, world):
            ra, rb = span(plan[r])
            if rb > ra:
                chunk = lr01[ra:rb].contiguous()
                p2p += chunk.numel() * chunk.element_size()
                reqs.append(dist.isend(chunk, dst=r))
        for q in reqs:
            q.wait()
        seg = lr01[a0:b0]
    elif b0 > a0:
        seg = torch.empty(b0 - a0, *shape_tail, device=device)
        dist.recv(seg, src=0)
    else:
        seg = None
    # ---- independent windows of this rank
    t0 = time.perf_counter()
    parts = []
    for (widx, a, b) in plan[rank]:
        out = restore_window(seg[a - a0:b - a0], widx)
        keep = out if widx == 0 else out[overlap:]
        parts.append((keep.clamp(-1, 1) + 1) / 2)
    mine = torch.cat(parts, 0) if parts else None
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    compute_s = time.perf_counter() - t0
    # ---- gather on rank 0 (rank order == frame order)
    result = None
    if world == 1:
        result = mine
    elif rank != 0:
        if mine is not None:
            dist.send(mine.contiguous(), dst=0)
    else:
        outs = [mine]
        for r in range(1, world):
            ws = plan[r]
            if not ws:
                continue
            cnt = sum((b - a) - (0 if widx == 0 else overlap) for widx, a, b in ws)
            buf = torch.empty(cnt, *mine.shape[1:], device=mine.device, dtype=mine.dtype)
            dist.recv(buf, src=r)
            p2p += buf.numel() * buf.element_size()
            outs.append(buf)
        result = torch.cat(outs, 0)
        assert result.shape[0] == n_frames, (result.shape, n_frames)
    if stats is not None:
        stats.update(p2p_bytes=p2p, windows=sum(len(w) for w in plan), windows_per_rank_max=max(len(w) for w in plan),
                     compute_s=compute_s)
    return result


# ------------------------------------------------------------------------------------------------ segment granularity
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
    world, rank = _world()
    plan = segment_plan(n_frames, world, size, overlap)
    seg = scatter_frames(lr01, plan, rank, world, device, shape_tail)
    out = restore_segment(seg) if seg.shape[0] > 0 else seg.new_empty(0)
    return gather_frames(out, plan, rank, world, n_frames)
