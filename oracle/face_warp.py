"""CPU oracle (test infrastructure only) of the aux-prior warps of the FLAIR sampler: the crop of every frame into
the aligned-face frame of the prior and the way back with the parsing mask.

Restates guided_diffusion/facelib/utils/face_restoration_helper.py:225-253 (get_crop_face_from_affine_matrices),
:255-262 (get_inverse_affine) and :264-345 (inverse_faces) of the reference, which delegate the arithmetic to
OpenCV (third-party dependency, `opencv-python`; the build container has 4.13.0): `cv2.warpAffine(..., INTER_CUBIC)`,
`cv2.invertAffineTransform`, `cv2.GaussianBlur((101, 101), 26)`.  OpenCV's published algorithm for those calls
(modules/imgproc/src/imgwarp.cpp: WarpAffineInvoker + remapBicubic; smooth.dispatch.cpp / filter.simd.hpp) is restated
here in numpy:

* warpAffine inverts the 2x3 matrix in double, evaluates the source position of a destination pixel in 1/1024 pixel
  fixed point (`saturate_cast<int>` = round half to even) with a rounding offset of 16, keeps 5 fractional bits
  (1/32 pixel) and interpolates with the 32-entry float bicubic table (A = -0.75); taps outside the source take the
  constant border value, written as cval + sum((S - cval) * w) on the border and sum(S * w) in the interior.
* GaussianBlur is a separable 101-tap filter (double kernel exp(-(i-50)^2 / (2 sigma^2)) normalised to sum 1) with
  BORDER_REFLECT_101.

Pinned: tools/gen_golden_aux.py runs the reference's own (unmodified) FaceRestoreHelper methods with cv2 and stores
inputs / outputs in tests/golden/aux_warp.pt; tests/test_oracle_aux.py checks this file against them."""
from __future__ import annotations

import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, TAB = 1 << AB_BITS, 1 << INTER_BITS
CROP_BORDER = (135.0, 133.0, 132.0)   # face_restoration_helper.py:243
MASK_COLORMAP = (0, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 0, 0, 0, 0, 0)   # :283-303


def cubic_table() -> np.ndarray:
    """[32][4] float32 bicubic weights (imgwarp.cpp interpolateCubic, A = -0.75, fraction i/32)."""
    A = np.float32(-0.75)
    tab = np.zeros((TAB, 4), np.float32)
    one = np.float32(1)
    for i in range(TAB):
        x = np.float32(i) * np.float32(1.0 / TAB)
        c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
        c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
        c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
        tab[i] = (c0, c1, c2, one - c0 - c1 - c2)
    return tab


def invert_affine(M) -> np.ndarray:
    """cv2.invertAffineTransform / the inversion inside cv2.warpAffine (double)."""
    M = np.asarray(M, np.float64).reshape(2, 3)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    A12, A21 = M[0, 1] * -D, M[1, 0] * -D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], np.float64)


def fixed_point_coords(Minv: np.ndarray, out_h: int, out_w: int):
    """Integer source position (sx, sy) and 1/32 fractions (fx, fy) of every destination pixel."""
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(Minv[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(Minv[1, 0] * xs * AB_SCALE).astype(np.int64)
    rd = AB_SCALE // TAB // 2
    X0 = np.rint((Minv[0, 1] * ys + Minv[0, 2]) * AB_SCALE).astype(np.int64) + rd
    Y0 = np.rint((Minv[1, 1] * ys + Minv[1, 2]) * AB_SCALE).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    return sx, sy, X & (TAB - 1), Y & (TAB - 1)


def warp_affine_cubic(src: np.ndarray, M, out_hw, border=0.0) -> np.ndarray:
    """cv2.warpAffine(src, M, (w, h), flags=INTER_CUBIC, borderMode=BORDER_CONSTANT, borderValue=border) for a
    float (H, W, C) or (H, W) image; float32 arithmetic like OpenCV's float path."""
    squeeze = src.ndim == 2
    s = np.asarray(src, np.float32)
    if squeeze:
        s = s[:, :, None]
    H, W, C = s.shape
    oh, ow = out_hw
    cval = np.broadcast_to(np.asarray(border, np.float32), (C,)).astype(np.float32)
    sx, sy, fx, fy = fixed_point_coords(invert_affine(M), oh, ow)
    sx = sx - 1
    sy = sy - 1
    tab = cubic_table()
    wx, wy = tab[fx], tab[fy]                                  # (oh, ow, 4)
    w2 = (wy[:, :, :, None] * wx[:, :, None, :]).astype(np.float32)   # [ky][kx], float product like the 2-D table
    out = np.empty((oh, ow, C), np.float32)
    pad = np.empty((H + 8, W + 8, C), np.float32)              # taps outside the source read the border value
    pad[:] = cval
    pad[4:-4, 4:-4] = s
    outside = (sx >= W) | (sx + 4 <= 0) | (sy >= H) | (sy + 4 <= 0)
    interior = (sx >= 0) & (sx < max(W - 3, 0)) & (sy >= 0) & (sy < max(H - 3, 0))
    cx = np.clip(sx, -4, W) + 4
    cy = np.clip(sy, -4, H) + 4
    acc_i = np.zeros((oh, ow, C), np.float32)                  # interior form: sum(S * w), row by row
    acc_b = np.zeros((oh, ow, C), np.float32)                  # border form: sum((S - cval) * w), tap by tap
    for ky in range(4):
        yy = np.minimum(cy + ky, H + 7)
        row = np.zeros((oh, ow, C), np.float32)
        for kx in range(4):
            xx = np.minimum(cx + kx, W + 7)
            v = pad[yy, xx]                                    # (oh, ow, C)
            wk = w2[:, :, ky, kx][:, :, None]
            row = row + v * wk if kx else v * wk
            acc_b = acc_b + (v - cval) * wk
        acc_i = acc_i + row
    out = np.where(interior[:, :, None], acc_i, acc_b + cval)
    out = np.where(outside[:, :, None], cval, out).astype(np.float32)
    return out[:, :, 0] if squeeze else out


def gaussian_kernel(ksize=101, sigma=26.0) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma, CV_64F)."""
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return k / k.sum()


def gaussian_blur(img: np.ndarray, ksize=101, sigma=26.0) -> np.ndarray:
    """cv2.GaussianBlur(img, (ksize, ksize), sigma) of a float64 (H, W) image (BORDER_REFLECT_101, separable)."""
    k = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    a = np.pad(np.asarray(img, np.float64), r, mode="reflect")
    H, W = img.shape
    rows = np.zeros((H + 2 * r, W), np.float64)
    for t in range(ksize):
        rows += k[t] * a[:, t:t + W]
    out = np.zeros((H, W), np.float64)
    for t in range(ksize):
        out += k[t] * rows[t:t + H]
    return out


def to_u8_range(x):
    """VF.normalize(x, [-1]*3, [2]*3).clamp(0, 1) * 255 on a (..., ) float32 array."""
    x = np.asarray(x, np.float32)
    return np.clip((x - np.float32(-1)) / np.float32(2), 0, 1).astype(np.float32) * np.float32(255)


def from_u8_range(x):
    """VF.normalize(x / 255, [0.5]*3, [0.5]*3).clamp(-1, 1)."""
    x = np.asarray(x, np.float32) / np.float32(255)
    return np.clip((x - np.float32(0.5)) / np.float32(0.5), -1, 1).astype(np.float32)


def crop_faces(imgs: np.ndarray, affine_matrices, face_size=(512, 512)) -> np.ndarray:
    """get_crop_face_from_affine_matrices (:225-253): imgs (B, 3, H, W) in [-1, 1] -> (B, 3, fh, fw) in [-1, 1]."""
    fw, fh = face_size
    hwc = np.transpose(to_u8_range(imgs), (0, 2, 3, 1))
    faces = [warp_affine_cubic(im, M, (fh, fw), border=CROP_BORDER) for im, M in zip(hwc, affine_matrices)]
    return from_u8_range(np.transpose(np.stack(faces, 0), (0, 3, 1, 2)))


def parse_mask(parse_idx: np.ndarray) -> np.ndarray:
    """inverse_faces :281-318: class index map (H, W) -> blurred float mask in [0, 1] (float64)."""
    lut = np.asarray(MASK_COLORMAP, np.float64)
    mask = lut[parse_idx]
    mask = gaussian_blur(gaussian_blur(mask))
    thres = 10
    mask[:thres, :] = 0
    mask[-thres:, :] = 0
    mask[:, :thres] = 0
    mask[:, -thres:] = 0
    return mask / 255.0


def inverse_faces(faces: np.ndarray, parse_logits: np.ndarray, affine_matrices):
    """inverse_faces (:264-345): faces (B, 3, h, w) in [-1, 1], parse_logits (B, 19, h, w) (= face_parse(faces)[0])
    -> (inverse-warped faces (B, 3, h, w) in [-1, 1], inverse-warped masks (B, 1, h, w))."""
    parse = parse_logits.argmax(axis=1)
    hwc = np.transpose(to_u8_range(faces), (0, 2, 3, 1))
    inv_faces, inv_masks = [], []
    for face, M, p in zip(hwc, affine_matrices, parse):
        Minv = invert_affine(M)                     # get_inverse_affine (:255-262)
        h, w, _ = face.shape
        inv_faces.append(warp_affine_cubic(face, Minv, (h, w), border=0.0))
        inv_masks.append(warp_affine_cubic(parse_mask(p).astype(np.float32), Minv, (h, w), border=0.0))
    inv = from_u8_range(np.transpose(np.stack(inv_faces, 0), (0, 3, 1, 2)))
    return inv, np.stack(inv_masks, 0)[:, None].astype(np.float32)


def blend(x0: np.ndarray, inv_face: np.ndarray, inv_mask: np.ndarray, w: float, clip=True) -> np.ndarray:
    """gaussian_diffusion.py:488-496: x_with_face = x0 (1 - m) + face m; x0 <- w x0 + (1 - w) x_with_face."""
    x0 = np.asarray(x0, np.float32)
    xw = x0 * (np.float32(1) - inv_mask) + inv_face * inv_mask
    if clip:
        xw = np.clip(xw, -1, 1)
    return (np.float32(w) * x0 + np.float32(1 - w) * xw).astype(np.float32)
