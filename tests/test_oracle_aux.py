"""CPU: the oracle of the aux face-prior warps (oracle/face_warp.py) against the outputs of the UNMODIFIED reference
methods + OpenCV stored by tools/gen_golden_aux.py (tests/golden/aux_warp.pt), and the host-side matrix helpers of
the drop-in FaceRestoreHelper."""
import numpy as np
import pytest
import torch

from flair_b200 import synth
from oracle import face_warp as fw

from aux_inputs import case_inputs


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_oracle_matches_reference_warps(golden, idx):
    fx = golden("aux_warp.pt")
    c = fx["cases"][idx]
    frames, faces, logits = case_inputs(c)
    s = c["stride"]
    crops = fw.crop_faces(frames.numpy(), c["matrices"].numpy(), (c["face"], c["face"]))[:, :, ::s, ::s]
    assert np.abs(crops - c["crops"].numpy()).max() < 2e-6          # [-1, 1] domain (255-domain float rounding / 127.5)
    inv_faces, inv_masks = fw.inverse_faces(faces.numpy(), logits, c["matrices"].numpy())
    assert np.abs(inv_faces[:, :, ::s, ::s] - c["inv_faces"].numpy()).max() < 2e-6
    assert np.abs(inv_masks[:, :, ::s, ::s] - c["inv_masks"].numpy()).max() < 2e-6
    for M, Mi in zip(c["matrices"].numpy(), c["inverse_matrices"].numpy()):
        assert np.array_equal(fw.invert_affine(M), Mi)


def test_oracle_blend_matches_reference(golden):
    fx = golden("aux_warp.pt")
    c, b = fx["cases"][0], fx["blend"]
    x0 = (synth.synthetic_clip(c["n"], c["img"], seed=b["x0_seed"]) * 2 - 1).float().numpy()
    out = fw.blend(x0, c["inv_faces"].numpy(), c["inv_masks"].numpy(), b["w"])
    assert np.abs(out - b["out"].numpy()).max() < 1e-6


def test_helper_host_side(golden):
    """Matrix helpers of the drop-in class (no GPU needed): same inverse as cv2.invertAffineTransform, same taps."""
    from guided_diffusion.facelib.utils import face_restoration_helper as frh
    fx = golden("aux_warp.pt")
    for c in fx["cases"]:
        for M, Mi in zip(c["matrices"].numpy(), c["inverse_matrices"].numpy()):
            assert np.array_equal(frh.invert_affine(M), Mi)
    assert np.array_equal(frh.gaussian_taps(), fw.gaussian_kernel())
    assert frh.MASK_COLORMAP == fw.MASK_COLORMAP and frh.CROP_BORDER == fw.CROP_BORDER
    with pytest.raises(RuntimeError):
        frh.FaceRestoreHelper(device="cpu")


def test_aux_prior_sampling_step_matches_reference(golden):
    """The aux-prior branch of p_sample (reference gaussian_diffusion.py:465-515): the oracle compositions the GPU tests
    check against (tests/aux_inputs.py) versus the UNMODIFIED reference p_sample on the same seeded inputs
    (tests/golden/aux_psample.pt, tools/gen_golden_aux.py psample) — aligned=True with blur data consistency, and
    aligned=False through the reference FaceRestoreHelper + cv2."""
    import aux_inputs as ai
    from conftest import rel_err
    fx = golden("aux_psample.pt")
    x0, sample = ai.aligned_oracle(ai.aligned_case())
    ea = rel_err(x0, fx["aligned"]["pred_xstart"]), rel_err(sample, fx["aligned"]["sample"])
    x0, sample = ai.unaligned_oracle(ai.unaligned_case())
    eu = rel_err(x0, fx["unaligned"]["pred_xstart"]), rel_err(sample, fx["unaligned"]["sample"])
    print("aux p_sample, oracle vs reference: aligned", ea, "unaligned", eu)
    assert max(ea) < 2e-6 and max(eu) < 2e-6
