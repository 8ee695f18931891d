"""bench.py contract on CPU: the reference arm prints exactly one JSON line with the agreed keys, and the native arm
fails loudly (no CPU fallback) when there is no GPU."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "restored_frames_per_s_full_sampler" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_native_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_timing_drivers_host_logic(monkeypatch):
    """run_replicas / run_sharded host logic with a stub job and stub CUDA events (no device work): warm-up steps are
    not timed and not counted, `value` = frames of all ranks / inner time, `e2e` uses the outer (copy-inclusive)
    interval, `launches` counts only the timed region."""
    sys.path.insert(0, str(ROOT))
    import bench
    from flair_b200 import _lib as L

    clock = [0.0]

    class FakeEvent:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self):
            self.t = clock[0]

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)

    class StubJob:
        frames, dev = 16, torch.device("cpu")

        def __init__(self):
            self.lr_host = torch.zeros(16, 3, 4, 4)
            self.out_host = torch.zeros(16, 3, 16, 16)
            self.calls = 0

        def restore_chained(self, lr_dev):
            self.calls += 1
            clock[0] += 2.0                 # two "seconds" of device time per clip
            L.LAUNCHES[0] += 100
            return torch.zeros(16, 3, 16, 16)

        def restore_window(self, lr_win, widx):
            clock[0] += 0.5
            L.LAUNCHES[0] += 10
            return torch.zeros(lr_win.shape[0], 3, 16, 16)

    # .to(dev, non_blocking=True) of a CPU tensor is a no-op; make the "copies" advance the fake clock
    orig_to, orig_copy = torch.Tensor.to, torch.Tensor.copy_

    def slow_to(self, *a, **k):
        clock[0] += 0.25
        return orig_to(self, *a, **k)

    def slow_copy(self, *a, **k):
        clock[0] += 0.25
        return orig_copy(self, *a, **k)

    job = StubJob()
    monkeypatch.setattr(torch.Tensor, "to", slow_to)
    monkeypatch.setattr(torch.Tensor, "copy_", slow_copy)
    barrier = lambda: None
    reduce_max = lambda a, b: (a, b)
    n0 = L.LAUNCHES[0]
    res = bench.run_replicas(job, 2, 3, 4, barrier, reduce_max)          # 2 timed steps, 3 warm-up, "4 ranks"
    assert job.calls == 5 and L.LAUNCHES[0] - n0 == 500 and res["launches"] == 200
    assert abs(res["ms_per_step"] - 2000.0) < 1e-6
    assert abs(res["value"] - 16 * 4 * 2 / 4.0) < 1e-9                   # all ranks' frames / inner seconds
    assert abs(res["e2e"] - 16 * 4 * 2 / 5.0) < 1e-9                     # + H2D and D2H inside the outer interval
    assert res["h2d"] == 16 * 3 * 4 * 4 * 4 and res["d2h"] == 16 * 3 * 16 * 16 * 4

    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    res = bench.run_sharded(job, 1, 1, 1, 0, barrier, reduce_max)        # one rank: 2 windows (overlap 2) of 16 frames
    assert res["windows"] == 2 and res["windows_per_rank_max"] == 2 and res["p2p_bytes_per_step"] == 0
    assert res["launches"] == 20 and abs(res["ms_per_step"] - 1000.0) < 1e-6
    assert abs(res["value"] - 16.0) < 1e-9 and res["e2e"] < res["value"]
