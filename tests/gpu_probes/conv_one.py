"""GPU probe: run one conv shape a few times (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
T, H, cin, cout, kt = [int(v) for v in sys.argv[1:6]]
n = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dt = torch.float16
x = torch.randn(1, T, H, H, cin, device="cuda").to(dt)
w = torch.randn(cout, cin, kt, 3, 3, device="cuda") / (cin * 9 * kt) ** 0.5
wpk = ops.pack_conv_weight(w, dt)
b = torch.zeros(cout, device="cuda")
y = ops.conv(x, wpk, cout, (kt, 3, 3), bias=b)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
torch.cuda.synchronize(); e0.record()
for _ in range(n): ops.conv(x, wpk, cout, (kt, 3, 3), bias=b, out=y)
e1.record(); torch.cuda.synchronize()
print("us per launch", e0.elapsed_time(e1) / n * 1e3)
