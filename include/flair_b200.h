/* flair_b200 — C ABI of the B200-native FLAIR hot path.
 *
 * The reference (wustl-cig/FLAIR) has no FFI: its hot path is Python calling
 * PyTorch ops.  Each entry point below therefore cites the reference *Python*
 * call it replaces (paths relative to the reference checkout).  The Python
 * boundary in `guided_diffusion/` binds these through ctypes (see
 * INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a device pointer owned by the caller unless noted;
 *   - `stream` is a cudaStream_t passed as void*; the library never
 *     synchronises and never allocates inside a launch call, so every call is
 *     CUDA-graph capturable;
 *   - return 0 on success, negative on failure (flair_last_error() explains);
 *   - no CPU fallback: a missing/incompatible GPU is an error.
 *
 * Activation layout: channels-last.  A feature map is [B][T][H][W][C] with a
 * per-pixel channel stride `cstride` (elements) that may exceed C, so a map can
 * be a channel slice of a wider concat buffer.  16-bit maps are bf16 (default)
 * or fp16 (FLAIR_F16); fp32 maps are used for the residual stream when asked.
 */
#ifndef FLAIR_B200_H
#define FLAIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLAIR_B200_VERSION 100

/* dtype tags */
#define FLAIR_BF16 0
#define FLAIR_F32 1
#define FLAIR_F16 2

/* conv epilogue activations */
#define FLAIR_ACT_NONE 0
#define FLAIR_ACT_RELU 1
#define FLAIR_ACT_LRELU01 2
#define FLAIR_ACT_SILU 3

/* output layouts of flair_conv_igemm */
#define FLAIR_OUT_NHWC 0
#define FLAIR_OUT_NCHW 1

const char* flair_last_error(void);
int flair_version(void);
/* 0 when device `dev` is sm_100 (B200); negative otherwise. */
int flair_check_device(int dev);

/* ------------------------------------------------------------------------
 * Implicit-GEMM convolution / GEMM on tcgen05 + TMEM, operands by TMA.
 * Replaces: every nn.Conv2d / nn.Conv3d / nn.Conv1d(k=1) / nn.Linear on the
 * UNet torso — unet_new.py:240-244,271-276,292-295,359,367,455-457,993,1220;
 * sr3.py:95,104,120,146; unet.py conv (3,1,1); mmedit conv3x3 inside
 * BasicVSR++ (unet_new.py:659-668,859-867).
 *
 *   out[b,t,h,w,n] = act( bias[n] + rowbias[b*T+t, n]
 *                         + sum_{taps,c} x[b,t+dt,h*sh+dh,w*sw+dw,c] * wgt[tap][n][c] )
 *                    (+ residual[b,t,h,w,n])
 *
 * `wgt` is packed [kt*kh*kw][Cout_pad][Cin_pad] (16-bit, Cin_pad = ceil64(Cin),
 * Cout_pad = ceil16(Cout)), zero padded.  Padding is k/2 in each dimension
 * (zero fill), stride 1 or 2 in H/W.
 * ---------------------------------------------------------------------- */
typedef struct flair_conv_params {
  const void* x;       /* [B][T][H][W][x_cstride] 16-bit                      */
  int B, T, H, W;      /* input extents                                       */
  int Cin;             /* real input channels (any; K is zero-padded to 64)   */
  int x_cstride;       /* elements between pixels (multiple of 8)             */
  const void* wgt;     /* packed weights, see above                           */
  int Cout;            /* real output channels                                */
  int kt, kh, kw;      /* kernel extents, each 1 or 3                         */
  int stride_hw;       /* 1 or 2                                              */
  const float* bias;   /* [Cout] or NULL                                      */
  const float* rowbias;/* [B*T][rowbias_stride] fp32 or NULL                  */
  int rowbias_stride;
  const void* residual;/* same geometry as out, or NULL                       */
  int residual_dtype;  /* FLAIR_BF16 / FLAIR_F32 / FLAIR_F16                  */
  int residual_cstride;
  void* out;
  int out_dtype;       /* FLAIR_BF16 / FLAIR_F32 / FLAIR_F16                  */
  int out_layout;      /* FLAIR_OUT_NHWC / FLAIR_OUT_NCHW                     */
  int out_cstride;     /* NHWC: elements between pixels                       */
  int act;             /* FLAIR_ACT_*; applied before the residual add        */
  int in_dtype;        /* FLAIR_BF16 or FLAIR_F16 (x and wgt)                 */
  float out_scale;     /* multiplies the result after act, before residual    */
  /* optional fused GroupNorm statistics of the OUTPUT (for the next norm):   */
  float* gn_partial;   /* NULL, or [B*T][gn_groups][2] fp32 sums, atomically  */
  int gn_groups;       /* accumulated (sum, sum of squares)                   */
} flair_conv_params;

int flair_conv_igemm(const flair_conv_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLAIR_B200_H */
