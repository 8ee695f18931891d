"""GPU probe: fused deformable conv (flair_deform_conv) vs torchvision.ops.deform_conv2d (CPU fp32) and
vs the two-kernel path (im2col + GEMM), with timings."""
import sys
import torch, torchvision
sys.path.insert(0, ".")
from flair_b200 import ops

def run(C, H, W, N=1, dt=torch.float16, mrm=10.0, time_it=False, smooth=False):
    g = torch.Generator().manual_seed(C + H)
    dev = "cuda"
    xa = torch.randn(N, H, W, C, generator=g).to(dt)
    xb = torch.randn(N, H, W, C, generator=g).to(dt)
    o = torch.randn(N, H, W, 432, generator=g) * 0.5
    if smooth:
        o = torch.nn.functional.avg_pool2d(o.permute(0, 3, 1, 2), 9, 1, 4).permute(0, 2, 3, 1).contiguous() * 3
    o = o.half()  # reference channel order; the offset map is always fp16
    f1 = torch.randn(N, 2, H, W, generator=g) * 2
    f2 = torch.randn(N, 2, H, W, generator=g) * 3
    w = (torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5)
    b = torch.randn(C, generator=g) * 0.1
    # ---- reference (CPU fp32 on the 16-bit-rounded inputs)
    of = o.float().permute(0, 3, 1, 2)
    o1, o2, mask = torch.chunk(of, 3, dim=1)
    offset = mrm * torch.tanh(torch.cat((o1, o2), 1))
    off1, off2 = torch.chunk(offset, 2, 1)
    off1 = off1 + f1.flip(1).repeat(1, off1.size(1) // 2, 1, 1)
    off2 = off2 + f2.flip(1).repeat(1, off2.size(1) // 2, 1, 1)
    x = torch.cat([xa, xb], -1).float().permute(0, 3, 1, 2)
    ref = torchvision.ops.deform_conv2d(x, torch.cat([off1, off2], 1), w.to(dt).float(), b, 1, 1, 1, torch.sigmoid(mask))
    ref = ref.permute(0, 2, 3, 1)
    # ---- fused
    perm = ops.deform_offset_perm()
    om_p = o[..., perm].contiguous().to(dev)
    wd = w.permute(0, 2, 3, 1).reshape(C, -1)
    wpk = ops.pack_deform_weight(wd, dt).to(dev)
    xa_d, xb_d, f1d, f2d, bd = xa.to(dev), xb.to(dev), f1.to(dev), f2.to(dev), b.to(dev)
    xa_g, xb_g = ops.pair_planes(xa_d), ops.pair_planes(xb_d)
    out = ops.deform_conv(xa_g, xb_g, om_p, f1d, f2d, wpk, bd, mrm)
    out_g = ops.deform_conv(xa_g, xb_g, om_p, f1d, f2d, wpk, bd, mrm)  # second launch: must be bit-identical
    # ---- two-kernel path
    o_d = o.to(dev)
    cols = ops.deform_im2col(xa_d, xb_d, o_d, f1d, f2d, 16, mrm)
    old = ops.conv(cols, wpk, C, (1, 1, 1), bias=bd)[0]
    torch.cuda.synchronize()
    e = lambda y: float((y.float().cpu() - ref).norm() / ref.norm())
    same = bool((out == out_g).all())
    print(f"C={C} {H}x{W} N={N} {dt}: fused rel-L2 {e(out):.3e}  rerun identical {same}  im2col+gemm {e(old):.3e}  "
          f"fused-vs-old max {float((out.float() - old.float()).abs().max()):.3e}", flush=True)
    ok = e(out) < 4e-3 and same
    if time_it:
        def t(fn, n=20):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        print(f"   fused {t(lambda: ops.deform_conv(xa_g, xb_g, om_p, f1d, f2d, wpk, bd, mrm, out=out)):.1f} us   "
              f"im2col {t(lambda: ops.deform_im2col(xa_d, xb_d, o_d, f1d, f2d, 16, mrm)):.1f} us + gemm "
              f"{t(lambda: ops.conv(cols, wpk, C, (1, 1, 1), bias=bd)):.1f} us", flush=True)
    return ok

ok = True
ok &= run(64, 16, 24)
ok &= run(64, 40, 40, N=2)
ok &= run(128, 24, 16)
ok &= run(64, 32, 32, dt=torch.bfloat16)
ok &= run(64, 256, 256, time_it=True)
ok &= run(64, 256, 256, time_it=True, smooth=True)
ok &= run(128, 128, 128, time_it=True)
ok &= run(128, 128, 128, time_it=True, smooth=True)
print("OK" if ok else "BAD")
sys.exit(0 if ok else 1)
