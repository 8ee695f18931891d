"""Probe (not a test): a ONE-frame clip in video mode (enable_cross_frames=True, T = 1) — the degenerate window of a
single-frame input.  Prints the GPU result against the oracle, or the exception either side raises."""
import sys
import traceback
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from test_gpu_ragged import CFG, S  # noqa: E402


def main():
    from flair_b200 import synth
    from guided_diffusion.unet_new import UNetModel
    from oracle.unet_blur import BlurUNetOracle
    model = UNetModel(**CFG, use_fp16=True)
    sd = synth.synthetic_state_dict(model, seed=99)
    model.load_state_dict(sd)
    x = torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(1))
    low = synth.synthetic_clip(1, S, seed=41) * 2 - 1
    t = torch.full((1,), 370)
    ref = None
    try:
        ref = BlurUNetOracle(CFG, sd).forward(x, t, low[None], num_frames=1, enable_cross_frames=True, vsrpp_weights=1.0)
        print("[t1] oracle ok", tuple(ref.shape))
    except Exception:
        print("[t1] oracle raised:", traceback.format_exc(limit=3).splitlines()[-1])
    if not torch.cuda.is_available():
        return
    model.convert_to_fp16()
    model.eval().cuda()
    try:
        out = model(x.cuda(), t.cuda(), low_res_input=low[None].cuda(), num_frames=1, enable_cross_frames=True,
                    vsrpp_weights=1.0)
        torch.cuda.synchronize()
        print("[t1] gpu ok", tuple(out.shape), "finite", bool(torch.isfinite(out).all()))
        if ref is not None:
            print("[t1] rel L2 vs oracle", float((out.cpu() - ref).norm() / ref.norm()))
    except Exception:
        print("[t1] gpu raised:", traceback.format_exc(limit=4).splitlines()[-1])


if __name__ == "__main__":
    main()
