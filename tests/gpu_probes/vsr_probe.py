"""GPU probe: device time of ONE BasicVSR++ module forward (T frames, CxHxW), replayed from a CUDA graph, and the
in-situ marginal cost of each per-frame launch kind (the launch is skipped: results are then meaningless, only the
time difference is read).
argv: C H [T]   (C=64 H=256 and C=128 H=128 are the two shapes of the blur UNet; 5 modules each per forward)"""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
from flair_b200 import _lib as L
import guided_diffusion.unet_new as U

C, H = int(sys.argv[1]), int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
torch.manual_seed(0)
m = U.BasicVSRPP(C, 10).to(dev).half()
with torch.no_grad():
    for p in m.parameters():
        if float(p.detach().abs().max()) == 0:
            p.data.normal_(0, 0.02)
ff = torch.randn(1, T - 1, 2, H, H, device=dev) * 2
fb = torch.randn(1, T - 1, 2, H, H, device=dev) * 2
flows = {H: (ff, fb, U.UNetModel._flow_pack(ff, True), U.UNetModel._flow_pack(fb, False))}
ctx = U._Ctx(None, flows, 1.0, True, torch.float16, T, torch.float16)
x = torch.randn(1, T, H, H, C, device=dev).half()

SKIP = set()
tag = {"name": None, "i": 0}
_scratch = {}
def _wrap_method(cls, meth, name):
    orig = getattr(cls, meth)
    def w(self, *a, **k):
        tag.update(name=name, i=0)
        try:
            return orig(self, *a, **k)
        finally:
            tag["name"] = None
    setattr(cls, meth, w)
_wrap_method(U.SecondOrderDeformableAlignment, "run", "da")
_wrap_method(U.ResidualBlocksWithInputConv, "run_split", "bb")
_conv, _deform, _warp2, _warp = ops.conv, ops.deform_conv, ops.flow_warp2, ops.flow_warp
def conv(xx, wpk, cout, *a, **k):
    if tag["name"]:
        kind = f'{tag["name"]}{tag["i"]}'; tag["i"] += 1
        if kind in SKIP:
            if k.get("out") is not None: return k["out"]
            key = (tuple(xx.shape[:4]), cout, k.get("out_dtype"))
            if key not in _scratch: _scratch[key] = torch.zeros(*xx.shape[:4], cout, dtype=k.get("out_dtype") or xx.dtype, device=xx.device)
            return _scratch[key]
    return _conv(xx, wpk, cout, *a, **k)
ops.conv = conv
ops.deform_conv = lambda *a, **k: (k["out"] if "deform" in SKIP else _deform(*a, **k))
ops.flow_warp2 = lambda *a, **k: (None if "warp" in SKIP else _warp2(*a, **k))
ops.flow_warp = lambda *a, **k: (k.get("out") if ("warp" in SKIP and k.get("out") is not None) else _warp(*a, **k))

def measure(skip):
    SKIP.clear(); SKIP.update(skip)
    run = lambda: m(x, ctx)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): run(); run()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    n0 = L.LAUNCHES[0]
    with torch.cuda.graph(g): y = run()
    n = L.LAUNCHES[0] - n0
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, n, y

base, n, y = measure([])
print(f"BasicVSRPP C={C} {H}x{H} T={T}: {base:.3f} ms per module forward ({base * 1e3 / (2 * T):.1f} us per frame-step), {n} C-ABI launches; finite={bool(torch.isfinite(y).all())}")
per = 2 * (T - 1)   # per-frame steps that run the alignment
for name, skip in [("flow_warp(2)", ["warp"]), ("offset conv0 (K=2C, preadd)", ["da0"]), ("offset conv1", ["da1"]), ("offset conv2", ["da2"]),
                   ("offset conv3 (->432)", ["da3"]), ("deform_conv", ["deform"]), ("backbone conv0 (preadd)", ["bb0"]),
                   ("backbone res conv1", ["bb1"]), ("backbone res conv2 (+2 residuals, pair planes)", ["bb2"]),
                   ("ALL per-frame launches", ["warp", "da0", "da1", "da2", "da3", "deform", "bb0", "bb1", "bb2"])]:
    ms, n2, _ = measure(skip)
    div = 2 * T if skip[0].startswith("bb") else per
    print(f"  without {name:48s}: {ms:.3f} ms  -> {(base - ms) * 1e3 / div:6.1f} us per launch in situ ({n - n2} launches)")
