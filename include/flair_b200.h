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
 *   out[b,t,h,w,n] = act( bias[n] + rowbias[b*T+t, n] + preadd[b,t,h,w,n]
 *                         + sum_{taps,c} x[b,t+dt,h*sh+dh,w*sw+dw,c] * wgt[tap][n][c] )
 *                    * out_scale * rowscale[b*T+t, n]  (+ residual (+ residual2))
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
  const float* rowscale;/* [B*T][rowscale_stride] fp32 or NULL: multiplies the  */
  int rowscale_stride; /* activated result per (frame, channel) before the    */
                       /* residual add (gate of sr3.TemporalWrapper2)         */
  const void* residual;/* same geometry as out, or NULL                       */
  int residual_dtype;  /* FLAIR_BF16 / FLAIR_F32 / FLAIR_F16                  */
  int residual_cstride;
  const void* residual2; /* optional second residual (same dtype/stride rules)  */
  int residual2_dtype;
  int residual2_cstride;
  void* out;
  int out_dtype;       /* FLAIR_BF16 / FLAIR_F32 / FLAIR_F16                  */
  int out_layout;      /* FLAIR_OUT_NHWC / FLAIR_OUT_NCHW                     */
  int out_cstride;     /* NHWC: elements between pixels                       */
  int act;             /* FLAIR_ACT_*; applied before the residual add        */
  int in_dtype;        /* FLAIR_BF16 or FLAIR_F16 (x and wgt)                 */
  float out_scale;     /* multiplies the result after act, before residual    */
  /* optional fused GroupNorm statistics of the OUTPUT (for the next norm,    */
  /* nn_new.py:17-19): NULL, or [m_tiles][4][Cout/16][16] fp32 written by the  */
  /* epilogue — per M tile, 32-row quarter and 16-channel chunk the (sum, sum  */
  /* of squares) of every group of Cout/gn_groups channels in the chunk, of    */
  /* the values as stored (16-bit).  m_tiles from flair_conv_gn_tiles; reduced */
  /* to (mean, rstd) by flair_gn_finalize (deterministic, no atomics on data). */
  /* Needs the compact epilogue: 16-bit NHWC output, Cout % 16 == 0, no        */
  /* rowbias / rowscale, 16-bit addends; group size a power of two >= 2.       */
  float* gn_partial;
  int gn_groups;
  /* optional second copy of a 16-bit NHWC output as "pair planes"            */
  /* [Cout/gc][B*T*H*W][2][gc], gc = out2_group_channels: entry p holds pixel  */
  /* p and pixel p+1 (row-major) — the source layout flair_deform_conv gathers */
  /* from (BasicVSR++ feat_prop).  Slot 1 of the last entry is never written.  */
  /* out2_neighbor = d (0 means 1): slot 1 of entry p holds pixel p + d, i.e.  */
  /* d = 1 pairs (x, x+1) (what flair_deform_conv gathers from), d = W pairs   */
  /* (y, y+1); the last d entries' slot 1 is never written.                    */
  void* out2;
  int out2_group_channels;       /* 8 or 16                                   */
  long long out2_group_stride;   /* elements between group planes             */
  /* optional pre-activation addend with the geometry of `out` (NHWC):        */
  /*   out = act(bias + rowbias + preadd + conv(x)) ...                       */
  /* A convolution over a channel concat is linear in its input slices, so    */
  /* the slices known for all frames up front (BasicVSR++: the current frame  */
  /* and the flows in the offset net, unet_new.py:874-879; the spatial and    */
  /* backward features in the backbone input, :729-735) are convolved once,   */
  /* batched over T, and enter the per-frame recurrent launch here.           */
  const void* preadd;
  int preadd_dtype;    /* FLAIR_BF16 / FLAIR_F32 / FLAIR_F16                  */
  int preadd_cstride;
  int out2_neighbor;   /* see out2 (0 = 1)                                     */
} flair_conv_params;

int flair_conv_igemm(const flair_conv_params* p, void* stream);
/* M tiles flair_conv_igemm walks for these extents (sizes gn_partial), tiles per batch element of the call, and the
 * number of consecutive frames one tile spans (> 1 on maps smaller than 128 pixels). */
int flair_conv_gn_tiles(int B, int T, int H, int W, int kh, int kw, int stride_hw, int* m_tiles, int* tiles_per_batch,
                        int* frames_per_tile);
/* debug: in a library built with -DFLAIR_CONV_TRACE_BUILD and run with FLAIR_CONV_TRACE=1, CTA 0 of every conv
 * launch records clock64() at 15 points (see conv_igemm.cu); copies the 16 values of the most recent launch to
 * the host (synchronises).  All zeros in a normal build. */
int flair_debug_conv_trace(long long* host_out16);

/* ------------------------------------------------------------------------
 * Fused sampler update (fp32).  Replaces gaussian_diffusion.py:344-365,465-470,
 * 497-515 and the per-step `_extract_into_tensor` uploads (:692-705).
 * `coef` is a device table [steps][8] = {sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod, sqrt_alphas_cumprod_prev,
 * sqrt_one_minus_alphas_cumprod_prev, gamma, 0, 0, 0} (float64 tables cast to
 * fp32 exactly as the reference does after indexing).  The step index comes
 * from `t_arr` (device int64 per frame, graph-replay friendly) when non-NULL,
 * else from the host int `t`.
 * Data-consistency correction: `R` (N,3,H,W) as returned by any restore_fn, or
 * `q_lr` (N,3,H/sf,W/sf) for the blur operator, in which case
 * R = Upscale_OP(q_lr) is evaluated in-register (pseudoSR.py:196-225).
 * ---------------------------------------------------------------------- */
typedef struct flair_update_params {
  const float* x_t;        /* (N,3,H,W)                                       */
  const float* model_out;  /* (N,model_ch,H,W); eps = first 3 channels        */
  int model_ch;            /* 3 or 6                                          */
  const float* noise;      /* (N,3,H,W) standard normals (torch generator)    */
  const float* R;          /* or NULL                                         */
  const float* q_lr;       /* or NULL                                         */
  const float* up_taps;    /* (up_k,up_k) = ds_kernel * sf^2                  */
  int up_k, sf, pre_stride;
  const float* prev;       /* (B,prev_k,3,H,W) or NULL (prev_recon)           */
  int prev_k, frames_per_window;
  const float* coef;
  const long long* t_arr;  /* per-frame step index (device int64[N]) or NULL */
  const float* gamma_arr;  /* per-frame gamma (device fp32[N]) or NULL       */
  const float* x0_in;      /* final pred_xstart supplied by the caller (aux-  */
                           /* prior path): eps/DC/prev are skipped; or NULL   */
  int t;                   /* used when t_arr is NULL                         */
  float sqrt_one_minus_rho, sqrt_rho;
  float* sample;           /* (N,3,H,W) x_{t-1}                               */
  float* pred_xstart;      /* (N,3,H,W) or NULL                               */
  int N, H, W;
  int clip_denoised;
} flair_update_params;

int flair_sampler_update_f32(const flair_update_params* p, void* stream);
/* x0 = clamp(a_t x_t - b_t eps) — gaussian_diffusion.py:311-327,344-349. */
int flair_pred_xstart_f32(const float* x_t, const float* model_out, int model_ch, const float* coef,
                          const long long* t_arr, int t, float* x0, int N, int H, int W,
                          int clip_denoised, void* stream);
/* out = clamp(x0 - gamma R) — gaussian_diffusion.py:465-470 (generic restore_fn path). */
int flair_dc_apply_f32(const float* x0, const float* R, const float* gamma_arr, float gamma,
                       float* out, int N, int H, int W, int clip_denoised, void* stream);
/* posterior mean + model variance (API parity of p_mean_variance; unused by p_sample) —
 * gaussian_diffusion.py:226-248,278-309.  tab [steps][8] = {posterior_mean_coef1, coef2,
 * posterior_log_variance_clipped, log(beta), posterior_variance, posterior_log_variance_clipped,0,0}. */
int flair_mean_variance_f32(const float* x_t, const float* x0, const float* model_out, int model_ch,
                            int learned_range, const float* tab, const long long* t_arr, int t,
                            float* mean, float* variance, float* log_variance, int N, int H, int W,
                            void* stream);
/* out = alpha*x + beta*y (q_sample, gaussian_diffusion.py:206-224). */
int flair_axpby_f32(const float* x, const float* y, float alpha, float beta, float* out,
                    long long n, void* stream);

/* ------------------------------------------------------------------------
 * Blur + x4 data-consistency operator pieces (fp32, NCHW planes).
 * Replaces pseudoSR.py:180-244 (`DownscaleOP`, `Conv_LR_with_Inv_hTh_OP`,
 * `Upscale_OP`: depth-wise Filter_Layer convs with replication padding).
 * Taps are the fp32 casts the reference uploads (pseudoSR.py:25-33).
 * ---------------------------------------------------------------------- */
/* lr[n,c,m,k] = sum taps[u,v] * x[clamp(sf*m+pre+u-r), clamp(sf*k+pre+v-r)] */
int flair_blur_down_f32(const float* x, float* lr, const float* taps, int k, int sf, int pre,
                        int planes, int H, int W, void* stream);
/* out = replicate-padded k x k cross-correlation at the same resolution,
 * optionally minus `sub` (same shape): InvhTh(lr) - sub.                  */
int flair_filter_same_f32(const float* x, const float* sub, float* out, const float* taps, int k,
                          int planes, int H, int W, void* stream);
/* hr = Upscale_OP(lr): zero-insert at phase `pre`, replicate pad, k x k.   */
int flair_blur_up_f32(const float* lr, float* hr, const float* taps, int k, int sf, int pre,
                      int planes, int H, int W, void* stream);

/* ------------------------------------------------------------------------
 * DCT-domain JPEG (4:2:0 by decimation, 8x8 ortho DCT, QF-scaled tables).
 * Replaces jpeg.py:72-167 + dct.py:167-202.  `dct`/`idct` are the 8x8
 * LinearDCT weights built on the host the way the reference builds them.
 * mode 0: encode  x -> (luma, chroma) quantised coefficient planes
 * mode 1: decode  (luma, chroma) -> out
 * mode 2: decode(encode(x)) fused, nothing but `out` is written.
 * x/out: (N,3,h,w) in [-1,1]; luma (N,1,h,w); chroma (N,2,h/2,w/2); h,w % 16 == 0.
 * ---------------------------------------------------------------------- */
int flair_jpeg_f32(int mode, const float* x, float* luma, float* chroma, float* out,
                   const float* dct, const float* idct, const float* q_luma,
                   const float* q_chroma, int N, int h, int w, void* stream);

/* ------------------------------------------------------------------------
 * Separable-SVD bicubic operator (DDRM SRConv).  Replaces
 * restore_util.py:54-82,162-227 as used by scripts/video_sample.py:177-181:
 *   R = V1 (V1^T X V1 - G) V1^T  per channel image X (img x img),
 * V1 = first `small` columns of V (img x small, row-major), G (planes,small,small)
 * = S^-1 (U^T Y U) S^-1 precomputed per window (may be NULL).
 * flair_srconv_project: T = V1^T X V1 - G        (planes, small, small)
 * flair_srconv_expand : R = V1 T V1^T            (planes, img, img)
 * flair_sandwich_f32  : out = L X Rm (generic small dense product used for setup:
 *                       L (p,q), X (planes,q,r), Rm (r,s) -> (planes,p,s))
 * ---------------------------------------------------------------------- */
int flair_sandwich_f32(const float* L, const float* X, const float* Rm, const float* sub,
                       float* out, int planes, int p, int q, int r, int s, float* workspace,
                       void* stream);

/* ------------------------------------------------------------------------
 * GroupNorm32 on channels-last maps.  Replaces nn_new.py:17-19 + nn.py:359-367
 * (statistics over (C/G,T,H,W) per batch element), the SiLU that follows,
 * the scale-shift conditioning (unet_new.py:321-325) and the 2x resampling of
 * up/down ResBlocks (unet_new.py:249-254,310-315).
 *   flair_gn_stats: partial[b][chunk][g] = (sum, sum of squares), fp32.  With `counter`
 *        (device int[B], zero before the first use, left zero) and `final_stats` the last
 *        chunk to finish also writes final_stats[b][g] = (mean, rstd), reduced in chunk
 *        order in double (deterministic); launches sharing a counter must be stream-ordered.
 *   flair_gn_apply: finishes the statistics (nchunks > 0: `partial` = the partial sums;
 *        nchunks == 0: `partial` = final_stats) and writes
 *        out = resample( silu?( ((x-mean)*rstd*gamma+beta) * (1+scale) + shift ) )
 *   (norm = 0 skips the normalisation: plain resample / dtype cast of x).
 * ---------------------------------------------------------------------- */
int flair_gn_stats_chunks(long long pixels_per_batch, int C);
int flair_gn_stats(const void* x, int dtype, int B, long long pixels_per_batch, int C, int cstride,
                   int groups, float* partial, int nchunks, int* counter, float* final_stats, float eps,
                   void* stream);
typedef struct flair_gn_apply_params {
  const void* x; int in_dtype;   /* [B][T][H][W][x_cstride]                   */
  void* out; int out_dtype;      /* [B][T][H'][W'][out_cstride]               */
  const float* partial; int nchunks;
  const float* gamma; const float* beta;           /* [C]                     */
  const float* scale; const float* shift;          /* [B*T][film_stride] or NULL */
  int film_stride;
  int B, T, H, W, C, groups;
  int x_cstride, out_cstride;
  int norm, silu, resample;      /* resample: 0 none, 1 nearest x2, 2 avg 2x2 */
  float eps;
} flair_gn_apply_params;
int flair_gn_apply(const flair_gn_apply_params* p, void* stream);
/* Statistics left by flair_conv_igemm (gn_partial) -> final[b][g] = (mean, rstd) as flair_gn_apply reads them with
 * nchunks = 0.  scratch: [B][flair_gn_finalize_splits(tiles_per_batch)][C] doubles; counter: [B] ints, zero on entry
 * (self-cleaning, may be shared with flair_gn_stats).  pixels_per_batch = T*H*W of the normalised map. */
int flair_gn_finalize_splits(int tiles_per_batch);
int flair_gn_finalize(const float* partial, int B, int tiles_per_batch, int C, int groups, long long pixels_per_batch,
                      double* scratch, int* counter, float* final_stats, float eps, void* stream);
/* dst[p][dst_coffset + c] = src[p][c]: channel concat (th.cat(dim=2), unet_new.py:1359). */
int flair_copy_channels(const void* src, void* dst, long long pixels, int channels, int elem_bytes,
                        int src_cstride, int dst_cstride, int dst_coffset, void* stream);

/* ------------------------------------------------------------------------
 * Attention.  Spatial: QKVAttentionLegacy (unet_new.py:540-570), qkv channels
 * head-major (H,3,64), fp32 online softmax, optional per-frame channel bias
 * (AttentionbottleBlock emb term, unet_new.py:426-428).  Temporal: windowed
 * 1 x (F-1) attention of TemporalAttention (unet_new.py:473-517, nn.py:370-394)
 * over per-frame projections qkv = [q|k|v] with the positional constants
 * cq = Wq pe_mid + bq [C], ck = Wk pe_j + bk [F-1][C], bv [C] folded in.
 * ---------------------------------------------------------------------- */
int flair_attn_spatial(const void* qkv, void* out, const float* rowbias, int rowbias_stride, int N,
                       int L, int heads, int qkv_cstride, int out_cstride, int dtype, void* stream);
int flair_attn_temporal(const void* qkv, void* out, const float* cq, const float* ck, const float* bv,
                        int B, int T, long long pixels, int C, int frames, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * Conditioning path (fp32) and input packing.
 * flair_timestep_embedding_f32: nn_new.py:103-121 (freqs table from the host).
 * flair_linear_f32: y = act_out(bias + act_in(x) Wt), Wt is [K][N]; act codes 0 none,
 *   1 SiLU, 2 sigmoid (time_embed
 *   unet_new.py:979-984 and all ResBlock emb_layers :258-264 concatenated).
 * flair_pack_im2col6: cat([a,b], channel) (N,3,H,W) fp32 x2 -> [N][H][W][64]
 *   16-bit im2col (k = tap*6 + c) feeding the first conv (unet_new.py:993,1331).
 * ---------------------------------------------------------------------- */
int flair_timestep_embedding_f32(const float* t, const float* freqs, float* out, int N, int dim,
                                 void* stream);
int flair_linear_f32(const float* x, const float* Wt, const float* bias, float* y, int M, int K, int N,
                     int silu_in, int silu_out, void* stream);
int flair_pack_im2col6(const float* a, const float* b, void* out, int N, int H, int W, int dtype,
                       void* stream);

/* ------------------------------------------------------------------------
 * BasicVSR++ support (channels-last maps, fp32 flow planes [N][2][H][W]).
 * flair_flow_warp      : mmedit flow_warp / F.grid_sample bilinear, zeros padding,
 *                        align_corners=True (unet_new.py:706,719).
 * flair_flow_compose_f32: f1 + warp(f2, f1) (unet_new.py:718).
 * flair_planes_to_cl   : fp32 planes -> 16-bit channel slice (flows into the
 *                        offset-net input, unet_new.py:875).
 * flair_deform_im2col  : offset/mask post-processing (unet_new.py:877-887) +
 *                        sampling half of torchvision.ops.deform_conv2d (:889-898);
 *                        cols [N*H*W][9*2C], column = tap*2C + channel; the GEMM
 *                        half is flair_conv_igemm with a 1x1 kernel over 18C.
 * flair_scale_pixels   : feat_prop *= weight (unet_new.py:739).
 * ---------------------------------------------------------------------- */
int flair_flow_warp(const void* x, const float* flow, void* out, int N, int H, int W, int C,
                    int x_cstride, int out_cstride, int dtype, void* stream);
/* the first- and second-order warps of one propagation step (unet_new.py:706,719) in one launch */
int flair_flow_warp2(const void* xa, const void* xb, const float* flow_a, const float* flow_b, void* out_a, void* out_b,
                     int N, int H, int W, int C, int xa_cstride, int xb_cstride, int out_cstride, int dtype, void* stream);
int flair_flow_compose_f32(const float* f2, const float* f1, float* out, int N, int H, int W, void* stream);
int flair_planes_to_cl(const float* src, void* dst, int N, int Cs, int H, int W, int dst_cstride,
                       int dst_coffset, int dtype, void* stream);
int flair_deform_im2col(const void* xa, const void* xb, int xa_cstride, int xb_cstride, const void* om,
                        int om_cstride, int om_dtype, const float* flow1, const float* flow2, void* cols,
                        int N, int H, int W, int C, int deform_groups, float max_residue_magnitude,
                        int dtype, void* stream);
int flair_scale_pixels(void* x, const float* wmap, long long pixels, int C, int cstride, int dtype,
                       void* stream);

/* ------------------------------------------------------------------------
 * Fused second-order deformable alignment: offset/mask post-processing
 * (unet_new.py:874-887; unet.py:469-482) + the whole
 * torchvision.ops.deform_conv2d(cat(xa, xb), offset, weight, bias, padding=1,
 * mask=mask) (unet_new.py:889-898) in ONE launch: the bilinear gather writes the
 * tcgen05 A operand in shared memory, the im2col matrix is never materialised.
 * C = 64 or 128, 16 deform groups (the FLAIR configurations); other shapes use
 * flair_deform_im2col + flair_conv_igemm.
 *
 * Sources are pair planes [8 groups][N*H*W entries][2][C/8] (what flair_conv_igemm
 * writes through `out2`).  Element (image n, group g, entry (y,x), slot s, channel c) lives at
 *   base + n*nstride + g*gstride + (y*W + x)*pstride + s*C/8 + c,  pstride = 2*C/8.
 * Entry p = (pixel p, pixel p+1) of the row-major map (out2_neighbor 1): the two x-corners of a bilinear sample are one
 * aligned 32-byte (C = 64) / 64-byte (C = 128) entry; at C = 128 a lane pair fetches its two slots in one instruction.
 * Never-written slots (slot 1 of the last entry / last row) must hold finite values (read with weight 0).
 * `om`: [N*H*W][om_cstride] fp16 (always, also with bf16 features) output of the offset net with its 432 channels
 * permuted to tap-major order: channel = tap*48 + quad*12 + kind*4 + gi for deform
 * group quad*4 + gi, kind 0 = dy, 1 = dx, 2 = mask (reference order: dy/dx at
 * (group*9 + tap)*2 + {0,1}, mask at 288 + group*9 + tap) — permute the rows of the
 * last offset conv's weight/bias once at pack time.
 * `wgt`: flair pack of the (C, 18C) matrix in channel-block-major K order: k = (kbq*9 + tap)*64 + c, kbq = 64-channel
 *   block of cat(xa, xb) (all nine taps of a block are gathered back to back: its source planes stay in L1).
 * ---------------------------------------------------------------------- */
typedef struct flair_deform_conv_params {
  const void* xa; long long xa_gstride, xa_pstride, xa_nstride;
  const void* xb; long long xb_gstride, xb_pstride, xb_nstride;
  const void* om; int om_cstride;
  const float* flow1;  /* [N][2][H][W] fp32: added (flipped) to the offsets of groups 0..7  */
  const float* flow2;  /* same for groups 8..15                                           */
  const void* wgt;
  const float* bias;   /* [C] or NULL                                                     */
  void* out;           /* [N][H][W][out_cstride] 16-bit                                   */
  int out_cstride;
  int N, H, W, C;
  int deform_groups;   /* 16                                                              */
  float max_residue_magnitude;
  int dtype;           /* FLAIR_F16 / FLAIR_BF16: sources, wgt, out (om is fp16)          */
} flair_deform_conv_params;

int flair_deform_conv(const flair_deform_conv_params* p, void* stream);

/* ----------------------------------------------------------------------
 * Aux face-prior warps, fp32 NCHW planes.  Replace the OpenCV CPU round trips of
 * guided_diffusion/facelib/utils/face_restoration_helper.py:225-253 (get_crop_face_from_affine_matrices),
 * :264-345 (inverse_faces) and the blend of gaussian_diffusion.py:488-496.  Arithmetic = OpenCV's float path
 * (1/32-pixel fixed-point source positions, 32-entry bicubic table with A = -0.75, constant border).
 * ---------------------------------------------------------------------- */
/* dst[n,c] = cv2.warpAffine(src[n,c], M_n, (Wd, Hd), INTER_CUBIC, BORDER_CONSTANT, border[c]).
 * minv: [N][6] doubles on the DEVICE = destination->source map (M_n inverted as cv2.warpAffine /
 * cv2.invertAffineTransform do, in double, on the host).  border: [C] host floats or NULL (0).
 * in_mode 1: src holds [-1,1] images, sampled as clamp((x+1)/2,0,1)*255 (:229);  out_mode 1: result stored as
 * clamp((v/255-0.5)/0.5,-1,1) (:246-252).  C <= 4. */
int flair_warp_affine_cubic_f32(const float* src, float* dst, const double* minv, int N, int C, int Hs, int Ws,
                                int Hd, int Wd, const float* border, int in_mode, int out_mode, void* stream);
/* mask[n,y,x] = 255 if bit argmax_c(logits[n,c,y,x]) of lut_bits is set else 0 (:265-306: face_parse(...)[0].argmax(1)
 * + MASK_COLORMAP; first maximum wins like torch.argmax).  classes <= 32. */
int flair_parse_mask_f32(const float* logits, float* mask, int N, int classes, int H, int W, unsigned int lut_bits,
                         void* stream);
/* out = cv2.GaussianBlur(in, (101, 101), sigma) per image (separable, BORDER_REFLECT_101; taps = the 101 fp32 casts of
 * cv2.getGaussianKernel(101, sigma)).  tmp: N*H*W floats.  finish != 0: additionally clear a `thres`-pixel frame and
 * divide by `scale` (:309-318).  out may alias in. */
int flair_gaussian_blur_f32(const float* in, float* out, float* tmp, const float* taps, int ksize, int N, int H, int W,
                            int finish, int thres, float scale, void* stream);
/* out = w x0 + (1-w) clamp(x0 (1-mask) + face mask)  — gaussian_diffusion.py:488-496.  mask: (N,1,H,W). */
int flair_aux_blend_f32(const float* x0, const float* face, const float* mask, float* out, double w, int N, int C,
                        int H, int W, int clip, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLAIR_B200_H */
