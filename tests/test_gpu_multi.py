"""Two-GPU tests of the sharded restoration (NCCL point-to-point scatter / gather around independent per-rank sampling).
Skipped on boxes with fewer than two devices; the host-side plumbing is covered on CPU by tests/test_parallel_gloo.py."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, marker):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "gpu_probes" / script)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert marker in r.stdout, r.stdout[-2000:]


def test_segment_sharding_two_ranks_matches_local():
    """restore_clip_sharded: every rank's chained segment equals the same segment restored locally (same seeds)."""
    _torchrun("sharded_probe.py", "SHARDED OK")


def test_window_sharding_two_ranks_is_world_size_independent():
    """restore_clip_windows (bench.py strong scaling): the stitched clip from 2 ranks == the one-process result."""
    _torchrun("sharded_windows_probe.py", "WINDOWS OK")
