"""ctypes binding of libflair_b200.so (the C ABI declared in include/flair_b200.h).

The library is built in-tree by `flair_b200.build`; there is deliberately no
fallback: if the shared object is missing the import of any op fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

# FLAIR_B200_LIB: A/B measurements against a library built from another commit (tools/build_ref_lib.sh); never set in
# production — the default is the in-tree library built by `python -m flair_b200.build`
import os as _os
_LIB_PATH = Path(_os.environ.get("FLAIR_B200_LIB") or Path(__file__).resolve().parent / "libflair_b200.so")

BF16, F32, F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU01, ACT_SILU = 0, 1, 2, 3
OUT_NHWC, OUT_NCHW = 0, 1


class ConvParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("x_cstride", C.c_int), ("wgt", C.c_void_p), ("Cout", C.c_int),
        ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("stride_hw", C.c_int),
        ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rowbias_stride", C.c_int),
        ("rowscale", C.c_void_p), ("rowscale_stride", C.c_int),
        ("residual", C.c_void_p), ("residual_dtype", C.c_int), ("residual_cstride", C.c_int),
        ("residual2", C.c_void_p), ("residual2_dtype", C.c_int), ("residual2_cstride", C.c_int),
        ("out", C.c_void_p), ("out_dtype", C.c_int), ("out_layout", C.c_int),
        ("out_cstride", C.c_int), ("act", C.c_int), ("in_dtype", C.c_int),
        ("out_scale", C.c_float), ("gn_partial", C.c_void_p), ("gn_groups", C.c_int),
        ("out2", C.c_void_p), ("out2_group_channels", C.c_int), ("out2_group_stride", C.c_longlong),
        ("preadd", C.c_void_p), ("preadd_dtype", C.c_int), ("preadd_cstride", C.c_int),
        ("out2_neighbor", C.c_int),
    ]


class UpdateParams(C.Structure):
    _fields_ = [
        ("x_t", C.c_void_p), ("model_out", C.c_void_p), ("model_ch", C.c_int), ("noise", C.c_void_p),
        ("R", C.c_void_p), ("q_lr", C.c_void_p), ("up_taps", C.c_void_p),
        ("up_k", C.c_int), ("sf", C.c_int), ("pre_stride", C.c_int),
        ("prev", C.c_void_p), ("prev_k", C.c_int), ("frames_per_window", C.c_int),
        ("coef", C.c_void_p), ("t_arr", C.c_void_p), ("gamma_arr", C.c_void_p), ("x0_in", C.c_void_p),
        ("t", C.c_int), ("sqrt_one_minus_rho", C.c_float), ("sqrt_rho", C.c_float),
        ("sample", C.c_void_p), ("pred_xstart", C.c_void_p),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("clip_denoised", C.c_int),
    ]


class GNApplyParams(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("in_dtype", C.c_int), ("out", C.c_void_p), ("out_dtype", C.c_int),
        ("partial", C.c_void_p), ("nchunks", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("film_stride", C.c_int),
        ("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("groups", C.c_int),
        ("x_cstride", C.c_int), ("out_cstride", C.c_int),
        ("norm", C.c_int), ("silu", C.c_int), ("resample", C.c_int), ("eps", C.c_float),
    ]


class DeformConvParams(C.Structure):
    _fields_ = [
        ("xa", C.c_void_p), ("xa_gstride", C.c_longlong), ("xa_pstride", C.c_longlong), ("xa_nstride", C.c_longlong),
        ("xb", C.c_void_p), ("xb_gstride", C.c_longlong), ("xb_pstride", C.c_longlong), ("xb_nstride", C.c_longlong),
        ("om", C.c_void_p), ("om_cstride", C.c_int), ("flow1", C.c_void_p), ("flow2", C.c_void_p),
        ("wgt", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("out_cstride", C.c_int),
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("deform_groups", C.c_int),
        ("max_residue_magnitude", C.c_float), ("dtype", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run `python -m flair_b200.build` (nvcc, sm_100a). "
                "flair_b200 has no CPU or PyTorch fallback."
            )
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.flair_last_error.restype = C.c_char_p
        _lib.flair_version.restype = C.c_int
        _lib.flair_check_device.argtypes = [C.c_int]
        _lib.flair_conv_igemm.argtypes = [C.POINTER(ConvParams), C.c_void_p]
        vp, i, f, ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
        _lib.flair_sampler_update_f32.argtypes = [C.POINTER(UpdateParams), vp]
        _lib.flair_pred_xstart_f32.argtypes = [vp, vp, i, vp, vp, i, vp, i, i, i, i, vp]
        _lib.flair_dc_apply_f32.argtypes = [vp, vp, vp, f, vp, i, i, i, i, vp]
        _lib.flair_mean_variance_f32.argtypes = [vp, vp, vp, i, i, vp, vp, i, vp, vp, vp, i, i, i, vp]
        _lib.flair_axpby_f32.argtypes = [vp, vp, f, f, vp, ll, vp]
        _lib.flair_blur_down_f32.argtypes = [vp, vp, vp, i, i, i, i, i, i, vp]
        _lib.flair_filter_same_f32.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
        _lib.flair_blur_up_f32.argtypes = [vp, vp, vp, i, i, i, i, i, i, vp]
        _lib.flair_jpeg_f32.argtypes = [i, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, vp]
        _lib.flair_sandwich_f32.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp]
        _lib.flair_gn_stats_chunks.argtypes = [ll, i]
        _lib.flair_gn_stats.argtypes = [vp, i, i, ll, i, i, i, vp, i, vp, vp, f, vp]
        _lib.flair_gn_apply.argtypes = [C.POINTER(GNApplyParams), vp]
        if hasattr(_lib, "flair_gn_finalize"):  # (absent from an older A/B library given through FLAIR_B200_LIB)
            _lib.flair_conv_gn_tiles.argtypes = [i, i, i, i, i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
            _lib.flair_gn_finalize_splits.argtypes = [i]
            _lib.flair_gn_finalize.argtypes = [vp, i, i, i, i, ll, vp, vp, vp, f, vp]
        _lib.flair_copy_channels.argtypes = [vp, vp, ll, i, i, i, i, i, vp]
        _lib.flair_attn_spatial.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, vp]
        _lib.flair_attn_temporal.argtypes = [vp, vp, vp, vp, vp, i, i, ll, i, i, i, vp]
        _lib.flair_timestep_embedding_f32.argtypes = [vp, vp, vp, i, i, vp]
        _lib.flair_linear_f32.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp]
        _lib.flair_pack_im2col6.argtypes = [vp, vp, vp, i, i, i, i, vp]
        _lib.flair_flow_warp.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, vp]
        _lib.flair_flow_compose_f32.argtypes = [vp, vp, vp, i, i, i, vp]
        _lib.flair_planes_to_cl.argtypes = [vp, vp, i, i, i, i, i, i, i, vp]
        _lib.flair_deform_im2col.argtypes = [vp, vp, i, i, vp, i, i, vp, vp, vp, i, i, i, i, i, f, i, vp]
        _lib.flair_scale_pixels.argtypes = [vp, vp, ll, i, i, i, vp]
        _lib.flair_flow_warp2.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp]
        _lib.flair_deform_conv.argtypes = [C.POINTER(DeformConvParams), vp]
        if hasattr(_lib, "flair_warp_affine_cubic_f32"):  # (absent from an older A/B library)
            _lib.flair_warp_affine_cubic_f32.argtypes = [vp, vp, vp, i, i, i, i, i, i, C.POINTER(C.c_float), i, i, vp]
            _lib.flair_parse_mask_f32.argtypes = [vp, vp, i, i, i, i, C.c_uint, vp]
            _lib.flair_gaussian_blur_f32.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, f, vp]
            _lib.flair_aux_blend_f32.argtypes = [vp, vp, vp, vp, C.c_double, i, i, i, i, i, vp]
    return _lib


LAUNCHES = [0]  # C-ABI launch calls issued (graph replays add the launches they captured)


def check(rc: int) -> None:
    LAUNCHES[0] += 1
    if rc != 0:
        raise RuntimeError(f"flair_b200: {lib().flair_last_error().decode()} (rc={rc})")
