"""GPU probe: UNet forward error vs the reference golden for different compute dtypes + timing."""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flair_b200 import synth
from guided_diffusion.unet_new import UNetModel

fx = torch.load("tests/golden/unet_blur.pt", weights_only=True)
model = UNetModel(**fx["cfg"], use_fp16=True, use_checkpoint=True)
model.load_state_dict({k: synth.synthetic_tensor(k, shp, 1234) for k, shp in fx["keys"].items()})
model.convert_to_fp16(); model.eval().cuda()
dev = "cuda"
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for dt, sdt in ((torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float16), (torch.float16, torch.float16)):
    model.compute_dtype = dt; model.stream_dtype = sdt
    print("operand", dt, "stream", sdt)
    o = model(fx["x"].to(dev), fx["image_t"].to(dev), low_res_input=fx["low_res"][:, None].to(dev), num_frames=1, enable_cross_frames=False)
    print(dt, "image", rel(o.cpu(), fx["image_out"]), flush=True)
    o = model(fx["x"].to(dev), fx["video_t"].to(dev), low_res_input=fx["low_res"][None].to(dev), num_frames=4,
              rnn_input=fx["rnn_input"][None].to(dev), enable_cross_frames=True, vsrpp_weights=1.0)
    print(dt, "video", rel(o.cpu(), fx["video_out"]), flush=True)
