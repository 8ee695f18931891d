"""GPU probe: spatial attention (tensor-core kernel vs fp32 oracle) error and time per launch.
env FLAIR_ATTN_TC=0/1, FLAIR_ATTN_VMODE=0/1/2.  argv: heads H W frames"""
import os, sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
from oracle import kernels as K
heads, H, W, N = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (4, 16, 16, 10)
g = torch.Generator().manual_seed(1)
qkv = (torch.randn(1, N, H, W, heads * 192, generator=g) * 1.5).half()
x = qkv.cuda()
y = ops.attn_spatial(x, heads)
torch.cuda.synchronize()
ref = K.qkv_attention_legacy(qkv, heads)
err = float((y.cpu().double() - ref.double()).norm() / ref.double().norm())
run = lambda: [ops.attn_spatial(x, heads) for _ in range(20)]
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): run()
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): run()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): gr.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 100
fl = 4.0 * N * heads * (H * W) ** 2 * 64
print(f"attn heads={heads} L={H*W} frames={N} TC={os.environ.get('FLAIR_ATTN_TC','1')} VMODE={os.environ.get('FLAIR_ATTN_VMODE','0')}: "
      f"rel err {err:.3e}, {us:.2f} us per launch ({fl / us / 1e6:.1f} TFLOP/s)")
