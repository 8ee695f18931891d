"""GPU probe: per-shape throughput of flair_conv_igemm on the conv shapes of the blur UNet
(T frames of 256x256), CUDA-event timed, inputs re-used (L2-warm like the model)."""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops, _lib as L

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda"; dt = torch.float16
shapes = []  # (tag, B, T, H, W, cin, cout, ks, count per forward)
def add(tag, res, cin, cout, ks, count, frames=T, B=1):
    shapes.append((tag, B, frames, res, res, cin, cout, ks, count))
# 2-D ResBlock convs (count = number of such convs per forward, approx from SURVEY U1)
for res, cin, cout, cnt in [(256, 64, 64, 14), (256, 128, 64, 2), (256, 192, 64, 1), (128, 128, 128, 18), (128, 64, 128, 1),
                            (128, 256, 128, 2), (64, 128, 128, 20), (64, 256, 256, 4), (32, 256, 256, 18), (32, 512, 256, 2),
                            (16, 256, 256, 18), (16, 512, 512, 4), (8, 512, 512, 18), (8, 1024, 512, 2), (4, 512, 512, 26),
                            (4, 1024, 512, 3)]:
    add("2d", res, cin, cout, (1, 3, 3), cnt)
for res, c, cnt in [(256, 64, 10), (128, 128, 22), (64, 128, 10), (64, 256, 2), (32, 256, 12), (16, 256, 6), (16, 512, 6), (8, 512, 12), (4, 512, 16)]:
    add("3d", res, c, c, (3, 3, 3), cnt)
# BasicVSR++ per-frame convs (frames=1): offset net, deform GEMM, backbone, reconstruction
for res, c in [(256, 64), (128, 128)]:
    add("vsr off0", res, 3 * c + 4, c, (1, 3, 3), 90, frames=1)
    add("vsr off1", res, c, c, (1, 3, 3), 180 + 300, frames=1)
    add("vsr off3", res, c, 432, (1, 3, 3), 90, frames=1)
    add("vsr dgemm", res, 18 * c, c, (1, 1, 1), 90, frames=1)
    add("vsr bb3c", res, 3 * c, c, (1, 3, 3), 100, frames=1)
add("qkv", 16, 256, 768, (1, 1, 1), 5)
add("1x1 skip", 256, 192, 64, (1, 1, 1), 1)

tot = 0.0
print(f"{'tag':10s} {'shape':34s} {'us':>9s} {'TF/s':>8s} {'cnt':>4s} {'ms/fwd':>8s}")
for tag, B, F, H, W, cin, cout, ks, cnt in shapes:
    cs = (cin + 7) // 8 * 8
    xb = torch.randn(B, F, H, W, cs, device=dev).to(dt)
    x = xb[..., :cin]
    w = torch.randn(cout, cin, *ks, device=dev) / (cin * ks[0] * ks[1] * ks[2]) ** 0.5
    wpk = ops.pack_conv_weight(w, dt)
    b = torch.zeros(cout, device=dev)
    y = ops.conv(x, wpk, cout, ks, bias=b)
    for _ in range(3): ops.conv(x, wpk, cout, ks, bias=b, out=y)
    n = 10
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): ops.conv(x, wpk, cout, ks, bias=b, out=y)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * F * H * W * cin * cout * ks[0] * ks[1] * ks[2]
    tot += us * cnt / 1e3
    print(f"{tag:10s} {f'{F}x{H}x{W} {cin}->{cout} k{ks}':34s} {us:9.1f} {fl/us/1e6:8.1f} {cnt:4d} {us*cnt/1e3:8.2f}", flush=True)
print("estimated conv ms per forward:", tot)
