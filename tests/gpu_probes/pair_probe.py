"""GPU probe: (1) conv epilogue pair-plane output == ops.pair_planes(out); (2) deform_conv determinism at small shapes."""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops, _lib as L
dev, dt = "cuda", torch.float16
g = torch.Generator().manual_seed(0)
for C, H in ((64, 64), (128, 32), (64, 256)):
    x = torch.randn(1, 1, H, H, 2 * C, generator=g).to(dt).to(dev)
    w = torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5
    wpk = ops.pack_conv_weight(w, dt).to(dev)
    res = torch.randn(1, 1, H, H, C, generator=g).to(dt).to(dev)
    wide = torch.zeros(1, 1, H, H, 3 * C, dtype=dt, device=dev)
    out = wide[..., C:2 * C]
    P = torch.full((8, H * H, 2, C // 8), 7.0, dtype=dt, device=dev)
    ops.conv(x, wpk, C, (1, 3, 3), residual=res, out=out, out2=P)
    ref = ops.pair_planes(out[0].contiguous())
    ref[:, -1, 1] = 7.0
    print(f"conv out2 C={C} {H}x{H}: pair planes equal {bool(torch.equal(P, ref))}  max diff {float((P.float() - ref.float()).abs().max()):.3e}")
    # deform determinism
    xa, xb = torch.randn(1, H, H, C, generator=g).to(dt).to(dev), torch.randn(1, H, H, C, generator=g).to(dt).to(dev)
    om = (torch.randn(1, H, H, 432, generator=g) * 0.5).half().to(dev)
    f1, f2 = (torch.randn(1, 2, H, H, generator=g) * 2).to(dev), (torch.randn(1, 2, H, H, generator=g) * 3).to(dev)
    wd = ops.pack_conv_weight(torch.randn(C, 18 * C, generator=g) / (18 * C) ** 0.5, dt).to(dev)
    xa_g, xb_g = ops.pair_planes(xa), ops.pair_planes(xb)
    outs = [ops.deform_conv(xa_g, xb_g, om, f1, f2, wd, None, 10.0, out=torch.zeros(1, H, H, 3 * C, dtype=dt, device=dev)[..., C:2 * C]).clone()
            for _ in range(20)]
    torch.cuda.synchronize()
    nd = sum(0 if torch.equal(o, outs[0]) else 1 for o in outs)
    print(f"deform C={C} {H}x{H}: {nd} of 20 launches differ from the first")
