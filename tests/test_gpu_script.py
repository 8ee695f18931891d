"""GPU: the demo script end to end — `scripts/video_sample.py::main` reads a folder of degraded PNG
frames, builds the demo model with synthetic weights (the checkpoints are not available offline), runs the windowed
sampler through flair_b200.pipeline and writes restored PNGs, like the reference script (:265-492)."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _script():
    spec = importlib.util.spec_from_file_location("flair_video_sample", ROOT / "scripts" / "video_sample.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_gaussian_demo_main_reads_and_writes_pngs(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cv2
    from flair_b200 import synth
    m = _script()
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    lr = (synth.synthetic_clip(4, 64, seed=8) * 255).round().byte().permute(0, 2, 3, 1).numpy()   # 4 LR frames 64x64
    for i, f in enumerate(lr):
        cv2.imwrite(str(src / f"{i}.png"), cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    out = m.main("gaussian", str(src), str(dst), device="cuda", t_start=1, image_size=256, weights="synthetic", seed=3)
    assert out.shape == (4, 3, 256, 256) and bool(torch.isfinite(out).all())
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    pngs = sorted(dst.glob("*.png"))
    assert [p.name for p in pngs] == ["0000.png", "0001.png", "0002.png", "0003.png"]
    back = np.stack([cv2.cvtColor(cv2.imread(str(p)), cv2.COLOR_BGR2RGB) for p in pngs])
    assert np.array_equal(back, (out * 255).byte().permute(0, 2, 3, 1).cpu().numpy())
    # two sampler steps from t = 1 keep the restored frames close to the blurred input they start from
    up = torch.nn.functional.interpolate(torch.from_numpy(lr).permute(0, 3, 1, 2).float() / 255, (256, 256), mode="area")
    assert float((out.cpu() - up).abs().mean()) < 0.3
