"""GPU probe: device time per launch of a conv when a chain of identical launches is replayed from a CUDA graph
(no host launch cost).  argv: T H Cin Cout [n]"""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
T, H, cin, cout = [int(v) for v in sys.argv[1:5]]
n = int(sys.argv[5]) if len(sys.argv) > 5 else 40
dt = torch.float16
x = torch.randn(1, T, H, H, cin, device="cuda").to(dt)
w = torch.randn(cout, cin, 1, 3, 3, device="cuda") / (cin * 9) ** 0.5
wpk = ops.pack_conv_weight(w, dt)
b = torch.zeros(cout, device="cuda")
ys = [torch.empty(1, T, H, H, cout, dtype=dt, device="cuda") for _ in range(2)]
run = lambda: [ops.conv(x, wpk, cout, (1, 3, 3), bias=b, out=ys[i & 1]) for i in range(n)]
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): run()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): run()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * n)
gf = 2.0 * T * H * H * cin * cout * 9 / 1e9
print(f"conv T={T} {H}x{H} {cin}->{cout}: {us:.2f} us per launch in a graph chain  ({gf / us * 1e3:.0f} TFLOP/s)")
