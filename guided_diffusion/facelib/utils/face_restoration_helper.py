"""`FaceRestoreHelper` — the three methods `GaussianDiffusion.p_sample` calls every step while the aux face prior is
active, on the device (drop-in for guided_diffusion/facelib/utils/face_restoration_helper.py of the reference):

  get_crop_face_from_affine_matrices (:225-253), get_inverse_affine (:255-262), inverse_faces (:264-345).

The reference moves every frame to the host, calls cv2.warpAffine / cv2.GaussianBlur per frame and moves the result
back (two crops + one inverse per step = 4 host round trips); here the same arithmetic (OpenCV's float bicubic path,
see flair_b200/csrc/face_warp.cu) runs as a few launches on the sampler's stream and nothing leaves the GPU.  Face
detection / landmark alignment (`get_crop_face`, :127-223: RetinaFace + cv2.estimateAffinePartial2D, once per window,
outside the sampling loop) is not part of the hot path and is not reimplemented: pass the affine matrices it produced.
"""
from __future__ import annotations

import numpy as np
import torch

from flair_b200 import ops

MASK_COLORMAP = (0, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 0, 0, 0, 0, 0)  # reference :283-303
CROP_BORDER = (135.0, 133.0, 132.0)                                                             # reference :243


def invert_affine(M) -> np.ndarray:
    """cv2.invertAffineTransform (and the inversion inside cv2.warpAffine), double precision."""
    M = np.asarray(M, np.float64).reshape(2, 3)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    A12, A21 = M[0, 1] * -D, M[1, 0] * -D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], np.float64)


def gaussian_taps(ksize=101, sigma=26.0) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma) (double)."""
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return k / k.sum()


class FaceRestoreHelper:
    """Same constructor keywords as the reference (:60-69) for the fields this path uses.  `face_parse` is the parsing
    network (reference: `init_parsing_model("parsenet")`, a PyTorch module whose output[0] holds 19 class logits); the
    detection model is not constructed here."""

    def __init__(self, face_size=512, crop_ratio=(1, 1), det_model="retinaface_resnet50", save_ext="png",
                 template_3points=False, device=None, face_parse=None):
        assert crop_ratio[0] >= 1 and crop_ratio[1] >= 1, "crop ration only supports >=1"
        self.crop_ratio = crop_ratio
        self.face_size = (int(face_size * crop_ratio[1]), int(face_size * crop_ratio[0]))   # (w, h)
        self.device = torch.device("cuda") if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("flair_b200 FaceRestoreHelper runs on CUDA only (no CPU fallback)")
        self.face_parse = face_parse
        self.face_det = None
        self._lut_bits = sum(1 << i for i, v in enumerate(MASK_COLORMAP) if v)
        self._taps = None
        self._mcache = {}

    # --- matrices: inverted on the host in double (6 numbers per frame), cached on the device per window
    def _device_maps(self, affine_matrices, inverse: bool):
        key = (inverse, tuple(np.asarray(m, np.float64).tobytes() for m in affine_matrices))
        hit = self._mcache.get(key)
        if hit is None:
            if len(self._mcache) > 64:
                self._mcache.clear()
            mats = [np.asarray(m, np.float64).reshape(2, 3) for m in affine_matrices]
            # cv2.warpAffine(M) samples the source at M^-1; inverse_faces passes M^-1, so it samples at (M^-1)^-1
            maps = [invert_affine(invert_affine(m)) if inverse else invert_affine(m) for m in mats]
            hit = torch.from_numpy(np.stack(maps).reshape(-1, 6)).to(self.device)
            self._mcache[key] = hit
        return hit

    def get_inverse_affine(self, affine_matrices):
        return [invert_affine(m) for m in affine_matrices]

    def get_crop_face_from_affine_matrices(self, bathed_imgs: torch.Tensor, affine_matrices):
        if len(affine_matrices) == 0:
            return None
        assert bathed_imgs.shape[0] == len(affine_matrices) and bathed_imgs.shape[1] == 3
        fw, fh = self.face_size
        return ops.warp_affine_cubic(bathed_imgs, self._device_maps(affine_matrices, False), (fh, fw),
                                     border=CROP_BORDER, in_mode=1, out_mode=1)

    def parse_masks(self, parse_logits: torch.Tensor) -> torch.Tensor:
        """(B, 19, h, w) logits -> (B, 1, h, w) blurred mask in [0, 1] (reference :281-318)."""
        if self._taps is None:
            self._taps = torch.from_numpy(gaussian_taps().astype(np.float32)).to(self.device)
        mask = ops.parse_mask(parse_logits, self._lut_bits)
        tmp = torch.empty_like(mask)
        ops.gaussian_blur101(mask, self._taps, out=mask, tmp=tmp)
        return ops.gaussian_blur101(mask, self._taps, finish=True, thres=10, scale=255.0, out=mask, tmp=tmp)

    def inverse_faces(self, restored_face_imgs: torch.Tensor, affine_matrices):
        if self.face_parse is None:
            raise RuntimeError("inverse_faces needs the parsing network: FaceRestoreHelper(face_parse=<ParseNet module>)")
        logits = self.face_parse(restored_face_imgs)[0]
        mask = self.parse_masks(logits.float())
        h, w = restored_face_imgs.shape[-2:]
        maps = self._device_maps(affine_matrices, True)
        inv_faces = ops.warp_affine_cubic(restored_face_imgs, maps, (h, w), in_mode=1, out_mode=1)
        inv_masks = ops.warp_affine_cubic(mask, maps, (h, w))
        return inv_faces, inv_masks
