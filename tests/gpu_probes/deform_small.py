import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
C, H, W, N, dt, mrm = int(sys.argv[3]) if len(sys.argv) > 3 else 64, int(sys.argv[1]), int(sys.argv[2]), 1, torch.float16, 10.0
g = torch.Generator().manual_seed(1)
dev = "cuda"
xa = torch.randn(N, H, W, C, generator=g).to(dt); xb = torch.randn(N, H, W, C, generator=g).to(dt)
o = (torch.randn(N, H, W, 432, generator=g) * 0.5).to(dt)
f1 = torch.randn(N, 2, H, W, generator=g) * 2; f2 = torch.randn(N, 2, H, W, generator=g) * 3
w = (torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5); b = torch.randn(C, generator=g) * 0.1
perm = ops.deform_offset_perm()
om_p = o[..., perm].contiguous().to(dev)
wpk = ops.pack_conv_weight(w.permute(0, 2, 3, 1).reshape(C, -1), dt).to(dev)
xa_d, xb_d, f1d, f2d, bd = xa.to(dev), xb.to(dev), f1.to(dev), f2.to(dev), b.to(dev)
outs = []
for rep in range(4):
    outs.append(ops.deform_conv(xa_d, xb_d, om_p, f1d, f2d, wpk, bd, mrm).float())
torch.cuda.synchronize()
for rep in range(1, 4):
    d = (outs[rep] - outs[0]).abs().reshape(-1, C)
    print("rep", rep, "max diff vs rep0", float(d.max()), "bad pixels", int((d.amax(1) > 0.01).sum()))
