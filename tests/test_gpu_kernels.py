"""Per-kernel parity: every UNet-side kernel family of libflair_b200.so (through the C ABI / ctypes wrappers in
flair_b200.ops) against the CPU oracle (oracle/kernels.py) on the same 16-bit-rounded operands.

Tolerances (relative L2): fp16 operands 1e-3, bf16 operands 5e-3 — the storage rounding of the 16-bit result
(2^-11 / 2^-8 per element) plus fp32 accumulation-order noise; fp32 outputs 1e-5."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from flair_b200 import _lib as L
    L.check(L.lib().flair_check_device(0))
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _tol(dtype):
    return 1e-3 if dtype == torch.float16 else 5e-3


CONV_CASES = [
    # B, T, H, W, cin, cout, ks, stride, act, residual, dtype, nchw
    (1, 1, 16, 16, 64, 64, (1, 1, 1), 1, None, False, torch.bfloat16, False),      # minimal GEMM
    (1, 2, 32, 32, 128, 128, (1, 3, 3), 1, None, True, torch.float16, False),      # ResBlock conv + skip
    (1, 3, 8, 8, 256, 512, (1, 3, 3), 1, "silu", False, torch.bfloat16, False),    # two N tiles, low res
    (1, 10, 4, 4, 512, 512, (3, 3, 3), 1, None, False, torch.float16, False),      # conv3d at 4x4 (tiles span frames)
    (2, 5, 16, 16, 64, 64, (3, 3, 3), 1, None, False, torch.float16, False),       # conv3d, B = 2 (no frame bleed)
    (1, 2, 64, 64, 196, 64, (1, 3, 3), 1, "lrelu", False, torch.float16, False),   # offset net: Cin % 64 != 0
    (1, 2, 64, 64, 64, 432, (1, 3, 3), 1, None, False, torch.float16, False),      # offset net: Cout = 27 * 16
    (1, 2, 32, 32, 64, 6, (1, 3, 3), 1, None, False, torch.float16, True),         # last conv: fp32 NCHW, Cout = 6
    (1, 1, 1, 40, 512, 1536, (1, 1, 1), 1, None, False, torch.float16, False),     # Linear as 1x1
    (1, 3, 16, 16, 64, 64, (3, 1, 1), 1, None, False, torch.float16, False),       # sr3 temporal (3,1,1)
    (1, 2, 32, 32, 64, 128, (1, 3, 3), 2, None, False, torch.float16, False),      # sr3 Downsample: stride 2
    (1, 3, 24, 40, 64, 64, (1, 3, 3), 1, "relu", True, torch.float16, False),      # ragged map (partial tiles)
    (1, 4, 64, 64, 128, 128, (1, 3, 3), 1, None, False, torch.bfloat16, False),    # halo tiles, streamed weight ring
    (1, 1, 256, 256, 64, 64, (1, 3, 3), 1, "lrelu", True, torch.float16, False),   # halo, resident, 4 tiles per CTA
    (1, 1, 128, 128, 128, 128, (1, 3, 3), 1, "relu", False, torch.float16, False), # halo, streamed, one tile per CTA
    (1, 2, 48, 24, 68, 64, (1, 3, 3), 1, None, False, torch.float16, False),       # [cur | flows]: Cin = C + 4, ragged
    (1, 4, 32, 16, 128, 64, (3, 3, 3), 1, None, True, torch.float16, False),       # conv3d halo: 27 taps streamed
    (1, 3, 16, 8, 256, 256, (1, 3, 3), 1, "silu", False, torch.float16, False),    # halo at the minimum map, N = 256 / split
    (1, 10, 16, 16, 512, 512, (1, 3, 3), 1, None, False, torch.float16, False),    # low-res: N tile split for parallelism
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: f"{c[4]}to{c[5]}_{c[2]}x{c[3]}_k{''.join(map(str, c[6]))}_s{c[7]}")
def test_conv_igemm(dev, case):
    from flair_b200 import _lib as L
    from flair_b200 import ops
    from oracle import kernels as K
    B, T, H, W, cin, cout, ks, stride, act, residual, dt, nchw = case
    g = torch.Generator().manual_seed(cin * 7 + cout)
    cs = (cin + 7) // 8 * 8
    xfull = torch.randn(B, T, H, W, cs, generator=g).to(dt)
    x = xfull[..., :cin]   # channel stride > Cin when Cin % 8 != 0
    w = torch.randn(cout, cin, *ks, generator=g) / (cin * ks[0] * ks[1] * ks[2]) ** 0.5
    b = torch.randn(cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    res = torch.randn(B, T, Ho, Wo, cout, generator=g).to(dt) if residual else None
    acts = {None: L.ACT_NONE, "silu": L.ACT_SILU, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU01}
    xd = xfull.to(dev)[..., :cin]
    y = ops.conv(xd, ops.pack_conv_weight(w, dt).to(dev), cout, ks, bias=b.to(dev), stride=stride, act=acts[act],
                 residual=None if res is None else res.to(dev), nchw_out=nchw)
    ref = K.conv_cl(x, w.to(dt), b, stride=stride, act=act, residual=res)
    if nchw:
        y = y.reshape(B, T, cout, Ho, Wo).permute(0, 1, 3, 4, 2)
        assert _rel(y, ref) < 1e-5
    else:
        assert _rel(y, ref) < _tol(dt)


@pytest.mark.parametrize("C,H", [(64, 32), (128, 16), (64, 40)])
def test_conv_pair_plane_output(dev, C, H):
    """flair_conv_params.out2: the pair-plane copy equals the NHWC output rearranged, bit for bit, also when the
    NHWC output is a channel slice of a wider buffer."""
    from flair_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(1, 1, H, H, 2 * C, generator=g).half().to(dev)
    wpk = ops.pack_conv_weight(torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5, torch.float16).to(dev)
    wide = torch.zeros(1, 1, H, H, 3 * C, dtype=torch.float16, device=dev)
    out = wide[..., C:2 * C]
    P = torch.full((8, H * H, 2, C // 8), 7.0, dtype=torch.float16, device=dev)
    nb = H if H == 40 else 1   # one case with vertical pairs (neighbor = W); flair_deform_conv uses horizontal ones
    ops.conv(x, wpk, C, (1, 3, 3), out=out, out2=P, out2_neighbor=nb)
    ref = ops.pair_planes(out[0].contiguous(), vertical=nb != 1)
    ref[:, -nb:, 1] = 7.0  # slot 1 of the last `neighbor` entries is never written
    assert torch.equal(P, ref)
    assert float(wide[..., :C].abs().max()) == 0.0 and float(wide[..., 2 * C:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,H,Cin,C,ks,groups,res", [(1, 3, 32, 64, 64, (1, 3, 3), 32, False), (1, 2, 32, 64, 128, (1, 3, 3), 32, True),
                                                       (4, 1, 16, 128, 256, (1, 3, 3), 32, True), (1, 4, 8, 256, 512, (1, 1, 1), 32, True),
                                                       (2, 3, 16, 64, 64, (3, 3, 3), 32, False), (1, 2, 16, 128, 1024, (1, 3, 3), 16, False),
                                                       (3, 1, 4, 64, 64, (1, 3, 3), 32, False)])
def test_conv_fused_group_norm_statistics(dev, monkeypatch, B, T, H, Cin, C, ks, groups, res):
    """flair_conv_params.gn_partial + flair_gn_finalize: the (mean, rstd) left by the conv epilogue equal the
    statistics kernel's on the stored output (group sizes 2..64, 3x3 / 1x1 / 3x3x3, residual, several batch elements;
    the last case has M tiles spanning frames of different batch elements -> ops.conv falls back, no attribute)."""
    from flair_b200 import ops
    monkeypatch.setattr(ops, "FUSED_GN", True)   # optional path, off by default (flair_b200/ops.py)
    g = torch.Generator().manual_seed(C + H + B)
    x = torch.randn(B, T, H, H, Cin, generator=g).half().to(dev)
    w = torch.randn(C, Cin, *ks, generator=g) / (Cin * ks[0] * ks[1] * ks[2]) ** 0.5
    wpk = ops.pack_conv_weight(w, torch.float16).to(dev)
    r = torch.randn(B, T, H, H, C, generator=g).half().to(dev) if res else None
    y = ops.conv(x, wpk, C, ks, bias=torch.randn(C, generator=g).to(dev), residual=r, gn_groups=groups)
    y_plain = ops.conv(x, wpk, C, ks, bias=torch.randn(C, generator=torch.Generator().manual_seed(C + H + B)).to(dev) * 0, residual=r)
    fused = getattr(y, "_flair_gn", None)
    if B == 3:
        assert fused is None   # 4x4 maps: one 128-pixel tile holds 8 frames = several batch elements
        return
    assert fused is not None and fused[0] == groups
    fin = fused[2][0]
    delattr(y, "_flair_gn")
    ref = ops.gn_stats(y, groups)[0]   # the statistics kernel on the same stored map
    assert torch.allclose(fin[..., 0], ref[..., 0], rtol=0, atol=2e-5 * float(ref[..., 0].abs().max() + 1))
    assert torch.allclose(fin[..., 1], ref[..., 1], rtol=2e-5, atol=0)
    # deterministic: a second launch leaves the same bits
    y2 = ops.conv(x, wpk, C, ks, bias=None, residual=r, gn_groups=groups)
    y3 = ops.conv(x, wpk, C, ks, bias=None, residual=r, gn_groups=groups)
    assert torch.equal(y2._flair_gn[2][0], y3._flair_gn[2][0]) and y_plain.shape == y.shape


@pytest.mark.parametrize("C,groups,T,H,dt", [(64, 32, 3, 32, torch.float16), (256, 32, 2, 8, torch.bfloat16),
                                              (128, 16, 4, 16, torch.float16), (192, 32, 2, 16, torch.float16)])
def test_group_norm_film_silu(dev, C, groups, T, H, dt):
    from flair_b200 import ops
    from oracle import kernels as K
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(2, T, H, H, C, generator=g) * 1.7 + 0.3).to(dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    film = torch.randn(2 * T, 2 * C, generator=g) * 0.2
    xd = x.to(dev)
    fd = film.to(dev)
    y = ops.gn_apply(xd, ops.gn_stats(xd, groups), gamma.to(dev), beta.to(dev), scale=fd[:, :C], shift=fd[:, C:],
                     silu=True, groups=groups)
    ref = K.group_norm_cl(x, gamma, beta, groups, scale=film[:, :C], shift=film[:, C:], silu=True)
    assert _rel(y, ref) < _tol(dt)
    y2 = ops.gn_apply(xd, ops.gn_stats(xd, groups), gamma.to(dev), beta.to(dev), groups=groups, out_dtype=torch.float32)
    assert _rel(y2, K.group_norm_cl(x, gamma, beta, groups)) < 1e-5


@pytest.mark.parametrize("C,groups,T,H,dt,rs", [(64, 32, 3, 32, torch.float16, 1), (64, 32, 3, 32, torch.float16, 2),
                                                 (256, 32, 2, 8, torch.bfloat16, 1), (128, 32, 2, 16, torch.float16, 2),
                                                 (192, 32, 2, 16, torch.float16, 2)])
def test_group_norm_resample(dev, C, groups, T, H, dt, rs):
    """GroupNorm + FiLM + SiLU followed by the ResBlock's nearest x2 upsample / 2x2 average pool (unet_new.py:249-254),
    and the plain resample of the skip path (no norm), in one launch each."""
    import torch.nn.functional as F
    from flair_b200 import ops
    from oracle import kernels as K
    g = torch.Generator().manual_seed(C + rs)
    x = (torch.randn(2, T, H, H, C, generator=g) * 1.3 - 0.2).to(dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    film = torch.randn(2 * T, 2 * C, generator=g) * 0.2
    xd, fd = x.to(dev), film.to(dev)

    def resample(t):   # (B, T, H, W, C) fp32
        n = t.reshape(-1, H, H, C).permute(0, 3, 1, 2)
        n = F.interpolate(n, scale_factor=2, mode="nearest") if rs == 1 else F.avg_pool2d(n, 2)
        return n.permute(0, 2, 3, 1).reshape(2, T, n.shape[2], n.shape[3], C)

    y = ops.gn_apply(xd, ops.gn_stats(xd, groups), gamma.to(dev), beta.to(dev), scale=fd[:, :C], shift=fd[:, C:],
                     silu=True, groups=groups, resample=rs)
    ref = resample(K.group_norm_cl(x, gamma, beta, groups, scale=film[:, :C], shift=film[:, C:], silu=True).float())
    assert y.shape == ref.shape and _rel(y, ref) < _tol(dt)
    y2 = ops.gn_apply(xd, None, resample=rs)
    assert _rel(y2, resample(x.float())) < (1e-6 if rs == 1 else _tol(dt))


@pytest.mark.parametrize("heads,L", [(4, 16), (8, 8), (8, 4)])
def test_spatial_attention(dev, heads, L):
    from flair_b200 import ops
    from oracle import kernels as K
    g = torch.Generator().manual_seed(heads * L)
    qkv = torch.randn(1, 3, L, L, heads * 3 * 64, generator=g).half()
    y = ops.attn_spatial(qkv.to(dev), heads)
    assert _rel(y, K.qkv_attention_legacy(qkv, heads)) < 1e-3


@pytest.mark.parametrize("heads,hw,dtype,bias", [(4, (16, 16), torch.float16, False), (8, (16, 8), torch.float16, True),
                                                 (2, (32, 32), torch.float16, False), (4, (16, 16), torch.bfloat16, True)])
def test_spatial_attention_tensor_core(dev, heads, hw, dtype, bias):
    """L = H*W a multiple of 128 takes the tcgen05 kernel (QK^T and PV on the tensor cores, fp32 online softmax from
    TMEM; L = 1024 walks 8 key blocks).  Checked against the fp32 oracle and against the SIMT kernel (FLAIR_ATTN_TC=0
    is read once per process, so the SIMT result comes from a map whose token count is not a multiple of 128)."""
    from flair_b200 import ops
    from oracle import kernels as K
    H, W = hw
    g = torch.Generator().manual_seed(heads * H + W)
    qkv = (torch.randn(1, 2, H, W, heads * 3 * 64, generator=g) * 1.5).to(dtype)
    rb = torch.randn(2, heads * 64, generator=g) if bias else None
    y = ops.attn_spatial(qkv.to(dev), heads, rowbias=None if rb is None else rb.to(dev))
    ref = K.qkv_attention_legacy(qkv.float(), heads)
    if rb is not None:
        ref = ref + rb[None, :, None, None, :]
    tol = 2e-3 if dtype == torch.float16 else 1e-2
    assert _rel(y, ref) < tol
    # a strided qkv view (channel stride > 3C) and a wider output go through the same tensor map
    wide = torch.zeros(1, 2, H, W, heads * 3 * 64 + 64, dtype=dtype, device=dev)
    wide[..., :heads * 3 * 64] = qkv.to(dev)
    y2 = ops.attn_spatial(wide[..., :heads * 3 * 64], heads, rowbias=None if rb is None else rb.to(dev))
    assert torch.equal(y, y2)


def test_flow_warp(dev):
    from flair_b200 import ops
    from oracle import kernels as K
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 24, 40, 64, generator=g).half()
    flow = torch.randn(2, 2, 24, 40, generator=g) * 6  # includes samples outside the map (zeros padding)
    y = ops.flow_warp(x.to(dev), flow.to(dev))
    assert _rel(y, K.flow_warp_cl(x, flow)) < 1e-3


def test_flow_warp2_equals_two_warps(dev):
    """The fused first-/second-order warp launch writes the same bits as two flow_warp launches, into channel slices."""
    from flair_b200 import ops
    g = torch.Generator().manual_seed(5)
    C, H, W = 64, 40, 24
    xa, xb = torch.randn(1, H, W, C, generator=g).half().to(dev), torch.randn(1, H, W, C, generator=g).half().to(dev)
    f1, f2 = (torch.randn(1, 2, H, W, generator=g) * 3).to(dev), (torch.randn(1, 2, H, W, generator=g) * 7).to(dev)
    cond = torch.zeros(1, H, W, 3 * C + 8, dtype=torch.float16, device=dev)
    ops.flow_warp2(xa, f1, cond[..., :C], xb, f2, cond[..., 2 * C:3 * C])
    assert torch.equal(cond[..., :C], ops.flow_warp(xa, f1))
    assert torch.equal(cond[..., 2 * C:3 * C], ops.flow_warp(xb, f2))
    assert float(cond[..., C:2 * C].abs().max()) == 0.0 and float(cond[..., 3 * C:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,H,C,groups", [(1, 10, 64, 64, 32), (2, 3, 16, 192, 32), (1, 4, 8, 512, 32), (1, 2, 32, 128, 16)])
def test_group_norm_statistics(dev, B, T, H, C, groups):
    """flair_gn_stats: (mean, rstd) finalised by the last CTA of the launch == torch statistics; repeated launches
    agree bit for bit (the ticket counter is self-cleaning, the reduction order is fixed)."""
    from flair_b200 import ops
    g = torch.Generator().manual_seed(C + T)
    x = (torch.randn(B, T, H, H, C, generator=g) * 2.3 - 0.7).half()
    xd = x.to(dev)
    fin, n = ops.gn_stats(xd, groups)
    assert n == 0 and fin.shape == (B, groups, 2)
    xf = x.float().reshape(B, T * H * H, groups, C // groups).permute(0, 2, 1, 3).reshape(B, groups, -1).double()
    mean, var = xf.mean(-1), xf.var(-1, unbiased=False)
    assert _rel(fin[..., 0], mean) < 1e-5
    assert _rel(fin[..., 1], 1.0 / torch.sqrt(var + 1e-5)) < 1e-5
    for _ in range(3):
        assert torch.equal(ops.gn_stats(xd, groups)[0], fin)


DEFORM_CASES = [(64, 16, 24, 1, torch.float16), (64, 40, 40, 2, torch.float16), (128, 24, 16, 1, torch.float16),
                (64, 32, 32, 1, torch.bfloat16), (128, 32, 32, 1, torch.float16), (64, 160, 160, 1, torch.float16)]


@pytest.mark.parametrize("C,H,W,N,dt", DEFORM_CASES, ids=lambda v: str(v).replace("torch.", ""))
def test_deform_conv_fused(dev, C, H, W, N, dt):
    """flair_deform_conv (offset post-processing + torchvision.ops.deform_conv2d in one launch) vs the oracle;
    ragged maps (partial 16x8 tiles), N > 1, more tiles than SMs (160x160), flows that leave the map, and two
    launches must agree bit for bit (the kernel had two ring races while it was written)."""
    from flair_b200 import ops
    from oracle import kernels as K
    g = torch.Generator().manual_seed(C + H)
    xa, xb = torch.randn(N, H, W, C, generator=g).to(dt), torch.randn(N, H, W, C, generator=g).to(dt)
    o = (torch.randn(N, H, W, 432, generator=g) * 0.5).half()           # reference channel order
    f1, f2 = torch.randn(N, 2, H, W, generator=g) * 2, torch.randn(N, 2, H, W, generator=g) * 5
    w = (torch.randn(C, 2 * C, 3, 3, generator=g) / (18 * C) ** 0.5).to(dt)
    b = torch.randn(C, generator=g) * 0.1
    nchw = lambda t: t.float().permute(0, 3, 1, 2)
    ref = K.deform_align_core(nchw(xa), nchw(xb), nchw(o), f1, f2, w, b, 10.0).permute(0, 2, 3, 1)
    om = o[..., ops.deform_offset_perm()].contiguous().to(dev)
    wpk = ops.pack_deform_weight(w.float(), dt).to(dev)
    xa_p, xb_p = ops.pair_planes(xa.to(dev)), ops.pair_planes(xb.to(dev))
    wide = torch.zeros(N, H, W, 2 * C, dtype=dt, device=dev)
    y1 = ops.deform_conv(xa_p, xb_p, om, f1.to(dev), f2.to(dev), wpk, b.to(dev), 10.0, out=wide[..., C:]).clone()
    y2 = ops.deform_conv(xa_p, xb_p, om, f1.to(dev), f2.to(dev), wpk, b.to(dev), 10.0)
    assert torch.equal(y1, y2)
    assert float(wide[..., :C].abs().max()) == 0.0
    assert _rel(y1, ref) < _tol(dt)


def test_deform_conv_matches_two_kernel_path(dev):
    """The generic path (flair_deform_im2col + 1x1 GEMM, used for shapes the fused kernel does not cover) and the
    fused kernel implement the same operator."""
    from flair_b200 import ops
    g = torch.Generator().manual_seed(11)
    C, H, W = 64, 48, 32
    xa, xb = torch.randn(1, H, W, C, generator=g).half().to(dev), torch.randn(1, H, W, C, generator=g).half().to(dev)
    o = (torch.randn(1, H, W, 432, generator=g) * 0.5).half().to(dev)
    f1, f2 = (torch.randn(1, 2, H, W, generator=g) * 2).to(dev), (torch.randn(1, 2, H, W, generator=g) * 3).to(dev)
    w2d = torch.randn(C, 18 * C, generator=g) / (18 * C) ** 0.5
    wpk = ops.pack_conv_weight(w2d, torch.float16).to(dev)          # tap-major K (im2col + GEMM path)
    wpk_f = ops.pack_deform_weight(w2d, torch.float16).to(dev)      # channel-block-major K (fused kernel)
    cols = ops.deform_im2col(xa, xb, o, f1, f2, 16, 10.0)
    old = ops.conv(cols, wpk, C, (1, 1, 1))[0]
    new = ops.deform_conv(ops.pair_planes(xa), ops.pair_planes(xb), o[..., ops.deform_offset_perm().to(dev)].contiguous(),
                          f1, f2, wpk_f, None, 10.0)
    assert _rel(new, old) < 1e-3


def test_video_forward_is_reproducible(dev):
    """Two eager forwards and two graph replays of a video-mode UNet (BasicVSR++ at C = 64 and C = 128) give
    bit-identical results — the whole launch chain is deterministic (no atomics, no races)."""
    from flair_b200 import synth
    from guided_diffusion.unet_new import UNetModel
    S, T = 64, 4
    model = UNetModel(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
                      attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
                      resblock_updown=True, use_scale_shift_norm=True, temporal_block=True, use_fp16=True)
    model.load_state_dict(synth.synthetic_state_dict(model, seed=7))
    model.convert_to_fp16()
    model.eval().to(dev)
    clip = (synth.synthetic_clip(T, S) * 2 - 1).to(dev)
    ts = torch.full((T,), 321, device=dev)
    x = torch.randn(T, 3, S, S, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    outs = []
    for graph in (False, False, True, True, True):
        model.use_cuda_graph = graph
        outs.append(model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0).clone())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
