"""GPU probe: intra-kernel timeline of CTA 0 of a conv launch.  Needs a library built with the trace marks:
    touch flair_b200/csrc/conv_igemm.cu; FLAIR_BUILD_TRACE=1 python -m flair_b200.build
(rebuild without the variable afterwards: the marks sit in the issue loops)."""
import ctypes, os, sys
os.environ["FLAIR_CONV_TRACE"] = "1"
import torch
sys.path.insert(0, ".")
from flair_b200 import ops, _lib as L
T, H, cin, cout = [int(v) for v in sys.argv[1:5]]
dt = torch.float16
x = torch.randn(1, T, H, H, cin, device="cuda").to(dt)
w = torch.randn(cout, cin, 1, 3, 3, device="cuda") / (cin * 9) ** 0.5
wpk = ops.pack_conv_weight(w, dt)
b = torch.zeros(cout, device="cuda")
y = ops.conv(x, wpk, cout, (1, 3, 3), bias=b)
for _ in range(5): ops.conv(x, wpk, cout, (1, 3, 3), bias=b, out=y)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16)()
L.lib().flair_debug_conv_trace.argtypes = [ctypes.c_void_p]
L.check(L.lib().flair_debug_conv_trace(buf))
t = list(buf)
names = ["entry", "prologue done", "TMA loop start", "TMA tile 1 start", "weights resident (MMA)", "MMA tile0 go", "MMA tile1 go",
         "MMA tile2 go", "MMA tile3 go", "acc tile0 ready", "acc tile1 ready", "acc tile2 ready", "acc tile3 ready",
         "epilogue thread done", "all warps joined"]
print(f"conv T={T} {H}x{H} {cin}->{cout}: cycles since entry (1.965 GHz: 1000 cycles = 0.51 us)")
for n, v in zip(names, t):
    if v: print(f"  {n:26s} {v - t[0]:8d} cycles  {(v - t[0]) / 1965:7.2f} us")
