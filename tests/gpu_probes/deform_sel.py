"""Which k-block of the fused deformable conv is corrupted?  Selector weights: out[:, n] = sum of k-block n of the A row."""
import sys
import torch
sys.path.insert(0, ".")
from flair_b200 import ops
C, H, W, N, dt, mrm = 64, int(sys.argv[1]), int(sys.argv[2]), 1, torch.float16, 10.0
g = torch.Generator().manual_seed(1)
dev = "cuda"
xa = torch.randn(N, H, W, C, generator=g).to(dt); xb = torch.randn(N, H, W, C, generator=g).to(dt)
o = (torch.randn(N, H, W, 432, generator=g) * 0.5).to(dt)
f1 = torch.randn(N, 2, H, W, generator=g) * 2; f2 = torch.randn(N, 2, H, W, generator=g) * 3
perm = ops.deform_offset_perm()
om_p = o[..., perm].contiguous().to(dev)
K = 18 * C
wsel = torch.zeros(C, K)
for n in range(18):
    wsel[n, n * 64:(n + 1) * 64] = 1.0
for n in range(18):  # columns 18..35: first-half chunks only (quad even), 36..53: second half
    wsel[18 + n, n * 64:n * 64 + 32] = 1.0
    wsel[36 + n, n * 64 + 32:(n + 1) * 64] = 1.0
wpk = ops.pack_conv_weight(wsel, dt).to(dev)
xa_d, xb_d, f1d, f2d = xa.to(dev), xb.to(dev), f1.to(dev), f2.to(dev)
cols = ops.deform_im2col(xa_d, xb_d, o.to(dev), f1d, f2d, 16, mrm)[0, 0].reshape(-1, K).float()  # [pix][K]
E = cols.reshape(-1, 18, 64).sum(-1)                     # [pix][18]
Elo = cols.reshape(-1, 18, 2, 32).sum(-1)                # [pix][18][2]
for rep in range(3):
    out = ops.deform_conv(xa_d, xb_d, om_p, f1d, f2d, wpk, None, mrm).float().reshape(-1, C)
    torch.cuda.synchronize()
    d = (out[:, :18] - E).abs()
    bad = (d > 0.08).nonzero()
    print(f"rep {rep}: bad (pixel,kblock) entries: {bad.shape[0]}")
    import collections
    cnt = collections.Counter(int(b[1]) for b in bad)
    print("   by k-block:", sorted(cnt.items()))
    for p, n in bad[:12].tolist():
        obs = float(out[p, n]); olo, ohi = float(out[p, 18 + n]), float(out[p, 36 + n])
        e = float(E[p, n]); elo, ehi = float(Elo[p, n, 0]), float(Elo[p, n, 1])
        stale = [float(E[p, m]) for m in range(18)]
        # which other k-block of the same pixel (or previous tile pixel p-128*k) matches?
        match = [m for m in range(18) if abs(stale[m] - obs) < 0.03]
        mlo = [m for m in range(18) if abs(float(Elo[p, m, 0]) - olo) < 0.03]
        mhi = [m for m in range(18) if abs(float(Elo[p, m, 1]) - ohi) < 0.03]
        print(f"   pix {p} (tile {p // 128} row {p % 128}) kb {n}: obs {obs:.3f} exp {e:.3f} | lo obs {olo:.3f} exp {elo:.3f} hi obs {ohi:.3f} exp {ehi:.3f} | full-match kb {match} lo-match {mlo} hi-match {mhi}")

# ---- hypothesis: the warp used the offsets of another tap t' at tap t's base position
o_d = o.to(dev)
def swapped_cols(shift):
    """offsets/mask of tap (t + shift) % 9 used at tap t"""
    oo = o_d.clone().float()
    ref = o_d.float()
    for gq in range(16):
        for t in range(9):
            t2 = (t + shift) % 9
            oo[..., (gq * 9 + t) * 2] = ref[..., (gq * 9 + t2) * 2]
            oo[..., (gq * 9 + t) * 2 + 1] = ref[..., (gq * 9 + t2) * 2 + 1]
            oo[..., 288 + gq * 9 + t] = ref[..., 288 + gq * 9 + t2]
    c = ops.deform_im2col(xa_d, xb_d, oo.to(dt), f1d, f2d, 16, mrm)[0, 0].reshape(-1, K).float()
    return c.reshape(-1, 18, 2, 32).sum(-1)
alts = {s: swapped_cols(s) for s in range(1, 9)}
out = ops.deform_conv(xa_d, xb_d, om_p, f1d, f2d, wpk, None, mrm).float().reshape(-1, C)
torch.cuda.synchronize()
d = (out[:, :18] - E).abs()
bad = (d > 0.08).nonzero()
print("final run bad entries", bad.shape[0])
hist = collections.Counter()
for p, n in bad.tolist():
    olo = float(out[p, 18 + n]); ohi = float(out[p, 36 + n])
    ms = [s for s in alts if abs(float(alts[s][p, n, 0]) - olo) < 0.03]
    ms_hi = [s for s in alts if abs(float(alts[s][p, n, 1]) - ohi) < 0.03]
    hist[(n, tuple(ms))] += 1
print("matches (kblock, tap shifts whose offsets reproduce the wrong lo value):", sorted(hist.items()))
