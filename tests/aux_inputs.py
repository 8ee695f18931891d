"""Seeded inputs of the aux-warp fixtures (the same functions tools/gen_golden_aux.py used)."""
import numpy as np

from flair_b200 import synth


def logits_from_seed(seed, n, size):
    """(n, 19, size, size) fp32: piecewise-constant class regions (16 x 16 blocks) plus small noise."""
    rng = np.random.default_rng(seed)
    cells = size // 16
    coarse = rng.standard_normal((n, 19, cells, cells)).astype(np.float32) * 3
    yy, xx = np.mgrid[0:cells, 0:cells]
    r = np.hypot(yy - cells / 2 + 0.5, xx - cells / 2 + 0.5) / (cells / 2)
    coarse[:, 0] += (6 * (r - 0.75)).astype(np.float32)
    fine = np.kron(coarse, np.ones((16, 16), np.float32))
    return (fine + 0.05 * rng.standard_normal(fine.shape).astype(np.float32)).astype(np.float32)


def case_inputs(c):
    n, img, face = c["n"], c["img"], c["face"]
    frames = synth.synthetic_clip(n, img, seed=c["frames_seed"][0]) * 2 - 1
    frames = (frames + 0.3 * synth.noise_tape((n, 3, img, img), 1, seed=c["frames_seed"][1])[0]).float()
    faces = (synth.synthetic_clip(n, face, seed=c["faces_seed"]) * 2 - 1).float()
    return frames, faces, logits_from_seed(c["logits_seed"], n, face)
