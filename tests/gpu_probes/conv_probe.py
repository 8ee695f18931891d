"""GPU probe: tcgen05 implicit-GEMM conv vs torch (run under gpurun, not pytest)."""
import sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from flair_b200 import ops, _lib as L

torch.manual_seed(0)
dev = "cuda"
L.check(L.lib().flair_check_device(0))


def run(B, T, H, W, cin, cout, ks, dtype=torch.bfloat16, residual=False, act=L.ACT_NONE, bias=True,
        nchw=False, stride=1, time_it=False, tag=""):
    kt, kh, kw = ks
    cs = (cin + 7) // 8 * 8
    x = torch.randn(B, T, H, W, cs, device=dev).to(dtype)[..., :cin]
    w = (torch.randn(cout, cin, kt, kh, kw, device=dev) / (cin * kt * kh * kw) ** 0.5)
    b = torch.randn(cout, device=dev) if bias else None
    wpk = ops.pack_conv_weight(w, dtype)
    res = torch.randn(B, T, (H + stride - 1) // stride, (W + stride - 1) // stride, cout, device=dev).to(dtype) if residual else None
    y = ops.conv(x, wpk, cout, ks, bias=b, residual=res, act=act, nchw_out=nchw, stride=stride)
    torch.cuda.synchronize()
    # reference: fp32 math on the rounded operands
    xr = x.float().permute(0, 4, 1, 2, 3)  # B C T H W
    wr = w.to(dtype).float()
    ref = F.conv3d(xr, wr, b, stride=(1, stride, stride), padding=(kt // 2, kh // 2, kw // 2))
    if act == L.ACT_SILU: ref = F.silu(ref)
    elif act == L.ACT_RELU: ref = F.relu(ref)
    elif act == L.ACT_LRELU01: ref = F.leaky_relu(ref, 0.1)
    ref = ref.permute(0, 2, 3, 4, 1)  # B T H W C
    if residual: ref = ref + res.float()
    if nchw:
        got = y.reshape(B, T, cout, *y.shape[-2:]).permute(0, 1, 3, 4, 2)
    else:
        got = y.float()
    err = (got - ref).norm() / ref.norm()
    mx = (got - ref).abs().max()
    msg = f"{tag} B{B} T{T} {H}x{W} {cin}->{cout} k{ks} s{stride} {str(dtype)[6:]}: rel_l2={err:.3e} max={mx:.3e}"
    if time_it:
        for _ in range(3): ops.conv(x, wpk, cout, ks, bias=b, out=y)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        n = 20
        for _ in range(n): ops.conv(x, wpk, cout, ks, bias=b, out=y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * T * (H // stride) * (W // stride) * cin * cout * kt * kh * kw
        msg += f"  {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s"
    print(msg, flush=True)
    return float(err)


bad = 0
cases = [
    dict(B=1, T=1, H=16, W=16, cin=64, cout=64, ks=(1, 1, 1), bias=False, tag="gemm-min"),
    dict(B=1, T=1, H=16, W=16, cin=64, cout=64, ks=(1, 3, 3), tag="conv-min"),
    dict(B=1, T=2, H=32, W=32, cin=128, cout=128, ks=(1, 3, 3), residual=True, tag="res"),
    dict(B=1, T=3, H=8, W=8, cin=256, cout=512, ks=(1, 3, 3), act=L.ACT_SILU, tag="ntiles2"),
    dict(B=1, T=10, H=4, W=4, cin=512, cout=512, ks=(1, 3, 3), tag="lowres"),
    dict(B=1, T=10, H=4, W=4, cin=512, cout=512, ks=(3, 3, 3), tag="conv3d-low"),
    dict(B=2, T=5, H=16, W=16, cin=64, cout=64, ks=(3, 3, 3), tag="conv3d-b2"),
    dict(B=1, T=4, H=64, W=64, cin=196, cout=64, ks=(1, 3, 3), act=L.ACT_LRELU01, tag="cin196"),
    dict(B=1, T=2, H=64, W=64, cin=64, cout=432, ks=(1, 3, 3), tag="cout432"),
    dict(B=1, T=2, H=32, W=32, cin=64, cout=6, ks=(1, 3, 3), nchw=True, tag="cout6-nchw"),
    dict(B=1, T=1, H=1, W=40, cin=512, cout=1536, ks=(1, 1, 1), tag="linear"),
    dict(B=1, T=3, H=16, W=16, cin=64, cout=64, ks=(3, 1, 1), tag="t311"),
    dict(B=1, T=2, H=32, W=32, cin=64, cout=128, ks=(1, 3, 3), dtype=torch.float16, tag="fp16"),
    dict(B=1, T=2, H=32, W=32, cin=64, cout=128, ks=(1, 3, 3), stride=2, tag="stride2"),
    dict(B=1, T=3, H=24, W=40, cin=64, cout=64, ks=(1, 3, 3), residual=True, tag="ragged-24x40"),
    dict(B=2, T=3, H=32, W=16, cin=128, cout=256, ks=(3, 3, 3), tag="3d-halo-b2"),
    dict(B=1, T=10, H=16, W=16, cin=192, cout=64, ks=(1, 3, 3), act=L.ACT_RELU, tag="192->64"),
    dict(B=1, T=4, H=64, W=64, cin=128, cout=128, ks=(1, 3, 3), tag="128-resident"),
]
for c in cases:
    try:
        e = run(**c)
        if not (e < 5e-3): bad += 1
    except Exception as ex:  # noqa
        print("FAIL", c.get("tag"), repr(ex), flush=True); bad += 1
print("---- timing ----", flush=True)
for c in [
    dict(B=1, T=10, H=256, W=256, cin=64, cout=64, ks=(1, 3, 3), tag="64@256"),
    dict(B=1, T=10, H=128, W=128, cin=128, cout=128, ks=(1, 3, 3), tag="128@128"),
    dict(B=1, T=10, H=32, W=32, cin=256, cout=256, ks=(1, 3, 3), tag="256@32"),
    dict(B=1, T=10, H=16, W=16, cin=512, cout=512, ks=(1, 3, 3), tag="512@16"),
    dict(B=1, T=10, H=128, W=128, cin=128, cout=128, ks=(3, 3, 3), tag="3d 128@128"),
    dict(B=1, T=1, H=64, W=128, cin=4096, cout=4096, ks=(1, 1, 1), tag="gemm 8192x4096x4096"),
]:
    try:
        run(time_it=True, **c)
    except Exception as ex:
        print("FAIL", c.get("tag"), repr(ex), flush=True); bad += 1
print("BAD", bad)
sys.exit(1 if bad else 0)
