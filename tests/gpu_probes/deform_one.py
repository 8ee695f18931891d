"""GPU probe: a few launches of flair_deform_conv at the model shapes (for ncu captures).  argv: C H W [smooth]"""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
C, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
smooth = len(sys.argv) > 4
dt, mrm, N = torch.float16, 10.0, 1
g = torch.Generator().manual_seed(1)
dev = "cuda"
xa = torch.randn(N, H, W, C, generator=g).to(dt); xb = torch.randn(N, H, W, C, generator=g).to(dt)
o = torch.randn(N, H, W, 432, generator=g) * 0.5
f1 = torch.randn(N, 2, H, W, generator=g) * 2; f2 = torch.randn(N, 2, H, W, generator=g) * 3
if smooth:
    sm = lambda t, k: torch.nn.functional.avg_pool2d(t, k, 1, k // 2) * k
    o = sm(o.permute(0, 3, 1, 2), 9).permute(0, 2, 3, 1).contiguous() * 0.3
    f1, f2 = sm(f1, 15), sm(f2, 15)
w = (torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5); b = torch.randn(C, generator=g) * 0.1
om_p = o.half()[..., ops.deform_offset_perm()].contiguous().to(dev)
wpk = ops.pack_deform_weight(w, dt).to(dev)
gm = ops.pair_planes
xa_g, xb_g, f1d, f2d, bd = gm(xa.to(dev)), gm(xb.to(dev)), f1.to(dev), f2.to(dev), b.to(dev)
out = torch.empty(N, H, W, C, dtype=dt, device=dev)
run = lambda: ops.deform_conv(xa_g, xb_g, om_p, f1d, f2d, wpk, bd, mrm, out=out)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"C={C} {H}x{W} smooth={smooth}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
