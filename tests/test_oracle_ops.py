"""Pin the CPU oracle (oracle/) against vectors dumped from the reference itself
(tests/golden, made by tools/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import degrade, sampler
from oracle.schedule import Tables


@pytest.mark.parametrize("task,name,n", [("gaussian", "face_blur", 1000), ("bicubic", "face_bicubic", 2000)])
def test_schedule_tables(golden, task, name, n):
    ref = golden("schedule.pt")[task]
    tab = Tables(name, n, 100)
    assert tab.timestep_map == ref["timestep_map"].tolist()
    for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_alphas_cumprod_prev",
              "sqrt_one_minus_alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
        np.testing.assert_array_equal(getattr(tab, k), ref[k].numpy(), err_msg=k)


def test_timestep_maps_match_survey():
    assert Tables("face_blur", 1000).timestep_map[:3] == [0, 10, 20]
    assert Tables("face_blur", 1000).timestep_map[-3:] == [979, 989, 999]
    assert Tables("face_bicubic", 2000).timestep_map[:4] == [0, 20, 40, 61]


def _taps(golden):
    t = golden("pseudosr_taps.pt")
    return t["ds_kernel"].float(), t["inv_hTh"].float()


def test_blur_operator_pieces(golden):
    fx = golden("dc_gaussian.pt")
    ds, inv = _taps(golden)
    assert rel_err(degrade.blur_down(fx["x"], ds), fx["down_x"]) < 1e-6
    assert rel_err(degrade.inv_hth(fx["y"], inv), fx["inv_y"]) < 1e-6
    assert rel_err(degrade.up_blur(degrade.inv_hth(fx["y"], inv), ds), fx["up_inv_y"]) < 1e-6


def test_blur_restore(golden):
    fx = golden("dc_gaussian.pt")
    ds, inv = _taps(golden)
    assert rel_err(degrade.blur_restore(fx["x"], fx["y"], ds, inv), fx["R"]) < 1e-5


def test_blur_restore_at_the_benchmarked_size(golden):
    """256 x 256 (BASELINE size): the oracle against the UNMODIFIED reference operator on the seeded inputs of
    tests/test_gpu_ops.py::test_blur_restore_vs_oracle_256 (tests/golden/dc_gaussian_256.pt, gen_golden_big.py ops256)."""
    from flair_b200 import synth
    fx = golden("dc_gaussian_256.pt")
    ds, inv = _taps(golden)
    hr = synth.synthetic_clip(fx["frames"], fx["size"], seed=fx["hr_seed"]) * 2 - 1
    x = (hr + 0.2 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(fx["noise_seed"]))).clamp(-1, 1)
    y = degrade.blur_down(hr, ds)
    assert rel_err(y, fx["down_hr"]) < 1e-6          # (bit-equal on the host that made the fixture)
    err = rel_err(degrade.blur_restore(x, y, ds, inv), fx["R"])
    print("oracle vs reference, blur restore at 256:", err)
    assert err < 1e-6


def test_filter_weights_are_the_taps(golden):
    t = golden("pseudosr_taps.pt")
    ds, inv = _taps(golden)
    assert torch.equal(t["w_down"], torch.flip(ds, (0, 1)))
    assert torch.equal(t["w_inv"], inv)
    assert torch.equal(t["w_up"], (t["ds_kernel"] * 16).float())
    assert t["pre_stride"].tolist() == [1, 1] and t["post_stride"].tolist() == [2, 2]


def test_jpeg_tables_and_dct(golden):
    fx = golden("dc_jpeg.pt")
    q1, q2 = degrade.quant_tables(fx["qf"])
    assert torch.equal(q1, fx["q_luma"]) and torch.equal(q2, fx["q_chroma"])
    # FFT-route matrices (reference dct.py:31-60): the last bit follows the host's FFT code path, so <= 1 ulp
    # against the fixture (made on another CPU); tests/test_boundary_cpu.py::test_jpeg_tables checks bit equality
    # with the reference module run on this host.
    assert (degrade.dct8_matrix() - fx["dct"]).abs().max().item() <= 6e-8
    assert (degrade.idct8_matrix() - fx["idct"]).abs().max().item() <= 6e-8


def test_jpeg_codec(golden):
    fx = golden("dc_jpeg.pt")
    enc = degrade.jpeg_encode(fx["img"], fx["qf"])
    # quantised coefficients are integers: allow no more than a handful of round() ties to flip
    assert (enc[0] != fx["enc_luma"]).sum() <= 2 and (enc[1] != fx["enc_chroma"]).sum() <= 2
    dec = degrade.jpeg_decode([fx["enc_luma"], fx["enc_chroma"]], fx["qf"])
    assert rel_err(dec, fx["dec"]) < 1e-6


def test_jpeg_restore(golden):
    fx = golden("dc_jpeg.pt")
    ds, inv = _taps(golden)
    R = degrade.blur_restore(fx["x"], fx["y"], ds, inv, jpeg_qf=fx["qf"])
    assert rel_err(R, fx["R"]) < 1e-4  # a flipped round() moves one 8x8 block by one quant step


@pytest.mark.parametrize("factor", [8, 16])
def test_srconv(golden, factor):
    fx = golden(f"dc_srconv_x{factor}.pt")
    assert torch.equal(degrade.bicubic_taps(factor), fx["taps"])
    A = degrade.srconv_matrix(fx["taps"], 64, factor)
    U, S, V = fx["U"], fx["S"], fx["V"]
    assert rel_err(U @ torch.diag(S) @ V[:, : S.shape[0]].t(), A) < 1e-5
    assert rel_err(degrade.srconv_restore(fx["x"], fx["y"], U, S, V), fx["R"]) < 1e-5
    # size-independent property: A(A^+ y) == y (all singular values non-zero)
    assert float(S.min()) > 3e-2


def test_sampler_steps(golden):
    fx = golden("sampler_step.pt")
    ds, inv = _taps(golden)
    tab = Tables("face_blur", 1000)
    np.testing.assert_allclose(tab.gammas(1.0, 2.55), fx["gammas"].numpy(), rtol=0, atol=0)
    restore = lambda v: degrade.blur_restore(v, fx["y"], ds, inv)
    for t in (99, 50, 1, 0):
        st = fx["steps"][t]
        s, x0 = sampler.p_sample_step(tab, fx["x_t"], fx["model_out"], t, st["noise"], restore,
                                      gamma=st["gamma"], rho=0.25)
        assert rel_err(x0, st["pred_xstart"]) < 1e-5, t
        assert rel_err(s, st["sample"]) < 1e-5, t
    st = fx["steps"]["prev"]
    s, x0 = sampler.p_sample_step(tab, fx["x_t"], fx["model_out"], 50, st["noise"], restore,
                                  gamma=fx["steps"][50]["gamma"], rho=0.25, prev_recon=st["prev"], num_frames=2)
    assert rel_err(x0, st["pred_xstart"]) < 1e-5 and rel_err(s, st["sample"]) < 1e-5


def test_sampler_loop(golden):
    from flair_b200 import synth
    fx = golden("sampler_loop.pt")
    ds, inv = _taps(golden)
    tab = Tables("face_blur", 1000)
    tape = synth.noise_tape((2, 3, 64, 64), 6, seed=6)
    restore = lambda v: degrade.blur_restore(v, fx["y"], ds, inv)

    def toy(x, t_in):
        e = 0.3 * torch.roll(x, 1, -1) - 0.1 * x + 0.001 * float(t_in)
        return torch.cat([e, torch.zeros_like(e)], 1)

    x5 = sampler.q_sample(tab, fx["hr"], 5, tape[0])
    assert rel_err(x5, fx["x_start"]) < 1e-6
    out = sampler.sample_loop(tab, toy, fx["x_start"], tape[1:], restore, rho=0.25, zeta=1.0, noise_level=2.55,
                              t_start=5)
    assert rel_err(out, fx["final"]) < 1e-5


# ------------------------------------------------------------------------------------------------ windowing
def test_windowed_restatement_known_answers():
    """oracle/windowing.windowed against the documented examples of more_itertools.windowed."""
    from oracle.windowing import windowed
    assert list(windowed([1, 2, 3, 4, 5], 3)) == [(1, 2, 3), (2, 3, 4), (3, 4, 5)]
    assert list(windowed([1, 2, 3], 4)) == [(1, 2, 3, None)]
    assert list(windowed([1, 2, 3, 4, 5, 6], 3, fillvalue="!", step=2)) == [(1, 2, 3), (3, 4, 5), (5, 6, "!")]
    assert list(windowed([1, 2, 3, 4, 5, 6, 7, 8], 3, step=4)) == [(1, 2, 3), (5, 6, 7)]   # step > n skips items
    assert list(windowed([], 3)) == []
    assert list(windowed([1, 2], 0)) == [()]


@pytest.mark.parametrize("size,overlap", [(10, 3), (10, 2), (4, 1), (5, 0), (3, 2)])
def test_pipeline_windows_match_the_script(size, overlap):
    """flair_b200.pipeline.windows == the reference script's windowing for every clip length, ragged tails included;
    stitching (drop `overlap` frames of every later window) yields every frame exactly once, in order."""
    from flair_b200.pipeline import windows
    from oracle.windowing import script_stitch, script_windows
    for n in range(0, 81):
        want = script_windows(n, size, overlap)
        got = windows(n, size, overlap)
        assert [list(range(a, b)) for a, b in got] == want, (n, got, want)
        assert script_stitch(want, overlap) == list(range(n)), n
        assert all(len(w) > overlap for w in want[1:]), n     # a later window always contributes >= 1 frame
    with pytest.raises(ValueError):
        windows(5, size, size)


# ------------------------------------------------------------------------------------------------ operator properties
def test_blur_data_consistency_properties(golden):
    """Size-independent properties of the blur operator's correction R = A_pinv(y, x) (pseudoSR.py:248-281): x that
    already explains y is a fixed point (R == 0 exactly), R is linear in (x, y), and ONE full correction makes the
    observation hold: Down(x - R) == y away from the replicate-padded border ring."""
    ds, inv = _taps(golden)
    g = torch.Generator().manual_seed(0)
    R = lambda x, y: degrade.blur_restore(x, y, ds, inv)
    S = 128
    x, x2, z = (torch.rand(1, 3, S, S, generator=g) * 2 - 1 for _ in range(3))
    y, y2 = degrade.blur_down(z, ds), torch.rand(1, 3, S // 4, S // 4, generator=g)
    assert float(R(z, y).abs().max()) == 0.0
    assert rel_err(R(0.3 * x - 1.7 * x2, 0.3 * y - 1.7 * y2), 0.3 * R(x, y) - 1.7 * R(x2, y2)) < 5e-6
    before = degrade.blur_down(x, ds) - y
    after = degrade.blur_down(x - R(x, y), ds) - y
    assert float(before.abs().max()) > 0.5
    assert float(after[..., 1:-1, 1:-1].abs().max()) < 5e-6      # interior: the truncated inverse filter is exact to fp32
    assert float(after.abs().max()) < 0.2                        # border ring: replicate padding, not a projection


def test_jpeg_codec_is_idempotent(golden):
    """encode(decode(encode(x))) == encode(x): a decoded image re-quantises to the same integer coefficients
    (jpeg.py:72-167), so the JPEG data-consistency step leaves an already consistent estimate alone."""
    fx = golden("dc_jpeg.pt")
    e1 = degrade.jpeg_encode(fx["img"], fx["qf"])
    d1 = degrade.jpeg_decode(e1, fx["qf"])
    e2 = degrade.jpeg_encode(d1, fx["qf"])
    assert torch.equal(e1[0], e2[0]) and torch.equal(e1[1], e2[1])
    assert torch.equal(e1[0], e1[0].round()) and torch.equal(e1[1], e1[1].round())     # integer planes
    assert torch.equal(degrade.jpeg_decode(e2, fx["qf"]), d1)


@pytest.mark.parametrize("factor", [8, 16])
def test_srconv_projection_properties(golden, factor):
    """bicubic_restore (scripts/video_sample.py:177-181): after x <- x - A^+(A x - y) the observation holds exactly
    (A A^+ = I: all singular values non-zero) and a second correction is zero (idempotent projection)."""
    fx = golden(f"dc_srconv_x{factor}.pt")
    U, S, V = fx["U"], fx["S"], fx["V"]
    small = U.shape[0]
    V1, ss = V[:, :small], S[:, None] * S[None, :]
    A = lambda X: U @ (ss * (V1.t() @ X @ V1)) @ U.t()
    x, y = fx["x"], fx["y"]
    x1 = x - degrade.srconv_restore(x, y, U, S, V)
    assert rel_err(A(x), y) > 1e-2 and rel_err(A(x1), y) < 2e-6
    assert float(degrade.srconv_restore(x1, y, U, S, V).abs().max()) < 5e-6
