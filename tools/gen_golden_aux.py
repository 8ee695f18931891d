"""Golden vectors for the aux face-prior warps, produced by the UNMODIFIED reference methods
`FaceRestoreHelper.get_crop_face_from_affine_matrices` / `.inverse_faces`
(/root/reference/guided_diffusion/facelib/utils/face_restoration_helper.py:225-345) with the build container's
OpenCV (cv2 4.13.0) on CPU.

  python tools/gen_golden_aux.py            ->  tests/golden/aux_warp.pt
  python tools/gen_golden_aux.py psample    ->  tests/golden/aux_psample.pt  (reference p_sample with the prior active)

The helper object is created without running its constructor (which downloads the RetinaFace / ParseNet checkpoints:
no network here); `face_parse` is a stand-in returning fixed, seeded logits (the real parsing network is a reference
PyTorch module outside this path).  Inputs are functions of seeds (flair_b200.synth + `logits_from_seed` below, which
tests re-create) so the fixture holds the matrices and the reference OUTPUTS only; the 512 x 512 case stores every
4th pixel of each output."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ref_env  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
ref_env.setup()

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from guided_diffusion.facelib.utils import face_restoration_helper as ref_frh  # noqa: E402  (reference)
from flair_b200 import synth  # noqa: E402

assert "/root/reference" in ref_frh.__file__, ref_frh.__file__


def logits_from_seed(seed, n, size):
    """(n, 19, size, size) fp32: piecewise-constant class regions (16 x 16 blocks) plus small noise."""
    rng = np.random.default_rng(seed)
    cells = size // 16
    coarse = rng.standard_normal((n, 19, cells, cells)).astype(np.float32) * 3
    # a face-like prior: classes 1..13 win in the middle, background (0) at the rim
    yy, xx = np.mgrid[0:cells, 0:cells]
    r = np.hypot(yy - cells / 2 + 0.5, xx - cells / 2 + 0.5) / (cells / 2)
    coarse[:, 0] += (6 * (r - 0.75)).astype(np.float32)
    fine = np.kron(coarse, np.ones((16, 16), np.float32))
    return (fine + 0.05 * rng.standard_normal(fine.shape).astype(np.float32)).astype(np.float32)


def matrices(n, img, face):
    """similarity maps of a face of ~45 % of the image into the `face` frame (what estimateAffinePartial2D returns)."""
    rng = np.random.default_rng(100 + img + face)
    out = []
    for k in range(n):
        s = face / (0.45 * img) * rng.uniform(0.9, 1.1)
        a = rng.uniform(-0.25, 0.25)
        cx, cy = img * rng.uniform(0.42, 0.58), img * rng.uniform(0.42, 0.58)
        R = np.array([[s * np.cos(a), -s * np.sin(a)], [s * np.sin(a), s * np.cos(a)]])
        t = np.array([face / 2, face / 2]) - R @ np.array([cx, cy])
        out.append(np.concatenate([R, t[:, None]], 1).astype(np.float64))
    return out


def run_case(img, face, n, stride):
    helper = object.__new__(ref_frh.FaceRestoreHelper)   # no constructor: it needs the network
    helper.face_size = (face, face)
    helper.device = torch.device("cpu")
    logits = torch.from_numpy(logits_from_seed(7 + face, n, face))
    helper.face_parse = lambda x: (logits,)
    frames = synth.synthetic_clip(n, img, seed=21) * 2 - 1
    frames = (frames + 0.3 * synth.noise_tape((n, 3, img, img), 1, seed=22)[0]).float()   # also outside [-1, 1]
    faces = (synth.synthetic_clip(n, face, seed=23) * 2 - 1).float()
    Ms = matrices(n, img, face)
    crops = helper.get_crop_face_from_affine_matrices(frames, Ms)
    inv_faces, inv_masks = helper.inverse_faces(faces, Ms)
    inv_M = helper.get_inverse_affine(Ms)
    sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
    return dict(img=img, face=face, n=n, stride=stride, matrices=torch.from_numpy(np.stack(Ms)), inverse_matrices=torch.from_numpy(np.stack(inv_M)),
                crops=crops[sub].contiguous(), inv_faces=inv_faces[sub].contiguous(),
                inv_masks=inv_masks[sub].float().contiguous(),
                frames_seed=(21, 22), faces_seed=23, logits_seed=7 + face)


def psample():
    """The aux-prior branch of the UNMODIFIED reference `GaussianDiffusion.p_sample` (gaussian_diffusion.py:423-517) on
    the exact inputs of the two GPU tests of that branch (tests/aux_inputs.py: aligned_case / unaligned_case), so the
    oracle compositions those tests use are pinned to the reference -> tests/golden/aux_psample.pt (outputs only)."""
    import numpy as np
    from scipy.io import loadmat
    import guided_diffusion.gaussian_diffusion as gd
    import guided_diffusion.pseudoSR as rpsr
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
    import aux_inputs as ai
    assert "/root/reference" in gd.__file__
    d = SpacedDiffusion(use_timesteps=space_timesteps(1000, "100", "uniform"),
                        betas=gd.get_named_beta_schedule("face_blur", 1000), noise_schedule="face_blur",
                        model_mean_type=gd.ModelMeanType.EPSILON, model_var_type=gd.ModelVarType.LEARNED_RANGE,
                        loss_type=gd.LossType.MSE, rescale_timesteps=False)
    out = {}
    # ---- aligned=True, blur data consistency with a per-frame gamma (tests/test_gpu_round2.py)
    c = ai.aligned_case()
    kernel = loadmat("/root/reference/miscs/kernels_12.mat")["kernels"]
    conf = rpsr.Get_pseudoSR_Conf(4); conf.sigmoid_range_limit = False; conf.input_range = np.array(None)
    A = rpsr.pseudoSR(conf, upscale_kernel=kernel[0, 3], kernel_indx=10).WrapArchitecture_PyTorch()
    gd.th.randn_like = lambda t: c["noise"].clone()
    N = c["x_t"].shape[0]
    r = d.p_sample(lambda xx, ts, **kw: c["mout"].clone(), c["x_t"].clone(), torch.full((N,), c["t"]), model_kwargs={},
                   restore_fn=lambda v: A.A_pinv(c["y"], v), aux_model=c["aux"], w=c["w"], start_timestep=99, tau=5,
                   aligned=True, rho=c["rho"], gamma=torch.full((N, 1, 1, 1), float(c["gamma"]), dtype=torch.float32),
                   face_restore_helper=None, affine_matrices=None)
    out["aligned"] = {"sample": r["sample"].clone(), "pred_xstart": r["pred_xstart"].clone()}
    # ---- aligned=False: crops -> aux -> inverse warps + parsing mask -> blend, through the reference helper + cv2
    u = ai.unaligned_case()
    helper = object.__new__(ref_frh.FaceRestoreHelper)
    helper.face_size = (u["S"], u["S"])
    helper.device = torch.device("cpu")
    logits = torch.from_numpy(u["logits"])
    helper.face_parse = lambda x: (logits,)
    gd.th.randn_like = lambda t: u["noise"].clone()
    N = u["x_t"].shape[0]
    r = d.p_sample(lambda xx, ts, **kw: u["mout"].clone(), u["x_t"].clone(), torch.full((N,), u["t"]), model_kwargs={},
                   aux_model=u["aux"], face_restore_helper=helper, affine_matrices=u["Ms"], w=u["w"], start_timestep=99,
                   tau=5, aligned=False, rho=u["rho"])
    out["unaligned"] = {"sample": r["sample"].clone(), "pred_xstart": r["pred_xstart"].clone()}
    out["cv2_version"] = cv2.__version__
    torch.save(out, OUT / "aux_psample.pt")
    print("wrote", OUT / "aux_psample.pt", (OUT / "aux_psample.pt").stat().st_size / 1e6, "MB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "psample":
        psample()
        raise SystemExit(0)
    fx = dict(cv2_version=cv2.__version__, cases=[run_case(256, 256, 1, 1), run_case(256, 512, 2, 4),
                                                  run_case(512, 512, 1, 4)])
    # the blend of gaussian_diffusion.py:488-496 on the first case, in torch like the reference writes it
    c = fx["cases"][0]
    x0 = (synth.synthetic_clip(1, 256, seed=31) * 2 - 1).float()
    w = 0.35
    xw = (x0 * (1 - c["inv_masks"]) + c["inv_faces"] * c["inv_masks"]).clamp(-1, 1)
    fx["blend"] = dict(w=w, x0_seed=31, out=w * x0 + (1 - w) * xw)
    torch.save(fx, OUT / "aux_warp.pt")
    print("wrote", OUT / "aux_warp.pt", (OUT / "aux_warp.pt").stat().st_size / 1e6, "MB")
