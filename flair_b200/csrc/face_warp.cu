// Aux face-prior warps of the FLAIR sampler on the device (fp32, NCHW planes), SURVEY 8(f) f3.
//
// Replaces the OpenCV CPU round trips of guided_diffusion/facelib/utils/face_restoration_helper.py:
//   :225-253 get_crop_face_from_affine_matrices  (normalise -> D2H -> cv2.warpAffine INTER_CUBIC, constant border
//            (135, 133, 132) -> H2D -> normalise), called twice per sampling step (pred_xstart and x_t);
//   :264-345 inverse_faces (parse argmax -> D2H -> colormap -> two 101 x 101 sigma-26 GaussianBlur passes -> border
//            clear -> /255 -> cv2.warpAffine of face and mask with the inverse matrix -> H2D);
// and the blend of gaussian_diffusion.py:488-496.  The arithmetic follows OpenCV's float path (imgwarp.cpp
// WarpAffineInvoker + remapBicubic): destination -> source position in 1/1024 pixel fixed point (round half to even,
// offset 16), 5 fractional bits kept, 32-entry float bicubic table (A = -0.75), taps outside the source read the
// constant border value (border form cval + sum((S - cval) w), interior form sum(S w)).  The CodeFormer prior and the
// parsing network themselves stay reference PyTorch modules (BASELINE north_star).
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kAbBits = 10, kInterBits = 5, kTab = 32;

__constant__ float c_cubic[kTab * 4];

// imgwarp.cpp interpolateCubic, evaluated in float exactly like the library's table initialisation (host, no FMA)
void cubic_table_host(float* tab) {
  const float A = -0.75f;
  for (int i = 0; i < kTab; ++i) {
    volatile float x = static_cast<float>(i) * (1.0f / kTab);
    volatile float xp1 = x + 1.0f, omx = 1.0f - x;
    volatile float t0 = A * xp1; t0 = t0 - 5.0f * A; t0 = t0 * xp1; t0 = t0 + 8.0f * A; t0 = t0 * xp1; t0 = t0 - 4.0f * A;
    volatile float t1 = (A + 2.0f) * x; t1 = t1 - (A + 3.0f); t1 = t1 * x; t1 = t1 * x; t1 = t1 + 1.0f;
    volatile float t2 = (A + 2.0f) * omx; t2 = t2 - (A + 3.0f); t2 = t2 * omx; t2 = t2 * omx; t2 = t2 + 1.0f;
    volatile float t3 = 1.0f - t0; t3 = t3 - t1; t3 = t3 - t2;
    tab[4 * i] = t0; tab[4 * i + 1] = t1; tab[4 * i + 2] = t2; tab[4 * i + 3] = t3;
  }
}

struct WarpArgs {
  const float* src;      // (N, C, Hs, Ws)
  float* dst;            // (N, C, Hd, Wd)
  const double* minv;    // (N, 6): destination -> source map (already inverted like cv2.warpAffine does)
  int N, C, Hs, Ws, Hd, Wd;
  float border[4];
  int in_mode, out_mode;
};

__device__ __forceinline__ float load_px(const float* p, int in_mode) {
  float v = __ldg(p);
  if (in_mode == 1) {  // VF.normalize(x, [-1]*3, [2]*3).clamp(0, 1) * 255
    v = __fmul_rn(fminf(fmaxf(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 0.0f), 1.0f), 255.0f);
  }
  return v;
}

// one thread = one destination pixel, all channels (the 16 weights are shared)
__global__ void __launch_bounds__(256) warp_affine_cubic_kernel(const __grid_constant__ WarpArgs a) {
  pdl_sync();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (x >= a.Wd || y >= a.Hd) return;
  const double* M = a.minv + 6 * n;
  const double m0 = __ldg(M), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3), m4 = __ldg(M + 4), m5 = __ldg(M + 5);
  constexpr double kScale = static_cast<double>(1 << kAbBits);
  constexpr int kRound = (1 << kAbBits) / kTab / 2;
  // saturate_cast<int>(double) = round half to even; no FMA contraction (the library evaluates mul, add, mul)
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m0, static_cast<double>(x)), kScale));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m3, static_cast<double>(x)), kScale));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, static_cast<double>(y)), m2), kScale)) + kRound;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, static_cast<double>(y)), m5), kScale)) + kRound;
  const int X = (X0 + adelta) >> (kAbBits - kInterBits), Y = (Y0 + bdelta) >> (kAbBits - kInterBits);
  int sx = X >> kInterBits, sy = Y >> kInterBits;
  sx = min(max(sx, -32768), 32767) - 1;   // saturate_cast<short>, then the top-left tap
  sy = min(max(sy, -32768), 32767) - 1;
  const float* wx = c_cubic + 4 * (X & (kTab - 1));
  const float* wy = c_cubic + 4 * (Y & (kTab - 1));
  const long long splane = static_cast<long long>(a.Hs) * a.Ws, dplane = static_cast<long long>(a.Hd) * a.Wd;
  const bool outside = sx >= a.Ws || sx + 4 <= 0 || sy >= a.Hs || sy + 4 <= 0;
  const bool interior = sx >= 0 && sx < a.Ws - 3 && sy >= 0 && sy < a.Hs - 3;
  float w2[16];
#pragma unroll
  for (int ky = 0; ky < 4; ++ky)
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) w2[ky * 4 + kx] = __fmul_rn(wy[ky], wx[kx]);
  for (int c = 0; c < a.C; ++c) {
    const float cv = a.border[c < 4 ? c : 3];
    const float* sp = a.src + (static_cast<long long>(n) * a.C + c) * splane;
    float v;
    if (outside) {
      v = cv;
    } else if (interior) {
      float sum = 0.0f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float* rp = sp + static_cast<long long>(sy + ky) * a.Ws + sx;
        float row = __fmul_rn(load_px(rp, a.in_mode), w2[ky * 4]);
#pragma unroll
        for (int kx = 1; kx < 4; ++kx) row = __fadd_rn(row, __fmul_rn(load_px(rp + kx, a.in_mode), w2[ky * 4 + kx]));
        sum = __fadd_rn(sum, row);
      }
      v = sum;
    } else {
      float sum = 0.0f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int yy = sy + ky;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = sx + kx;
          const bool in = yy >= 0 && yy < a.Hs && xx >= 0 && xx < a.Ws;
          const float s = in ? load_px(sp + static_cast<long long>(yy) * a.Ws + xx, a.in_mode) : cv;
          sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(s, cv), w2[ky * 4 + kx]));
        }
      }
      v = __fadd_rn(sum, cv);
    }
    if (a.out_mode == 1)  // VF.normalize(v / 255, [0.5]*3, [0.5]*3).clamp(-1, 1)
      v = fminf(fmaxf(__fmul_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 2.0f), -1.0f), 1.0f);
    a.dst[(static_cast<long long>(n) * a.C + c) * dplane + static_cast<long long>(y) * a.Wd + x] = v;
  }
}

// argmax over the class logits (first maximum wins, like torch.argmax) -> colormap value (0 / 255)
__global__ void __launch_bounds__(256) parse_mask_kernel(const float* __restrict__ logits, float* __restrict__ mask,
                                                         int classes, long long hw, long long total,
                                                         unsigned int lut_bits) {
  pdl_sync();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const long long n = i / hw, p = i - n * hw;
  const float* lp = logits + n * classes * hw + p;
  float best = __ldg(lp);
  int arg = 0;
  for (int c = 1; c < classes; ++c) {
    const float v = __ldg(lp + c * hw);
    if (v > best) { best = v; arg = c; }
  }
  mask[i] = ((lut_bits >> arg) & 1u) ? 255.0f : 0.0f;
}

__device__ __forceinline__ int reflect101(int p, int len) {
  // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba  (len > 1; a kernel radius may exceed the image for tiny maps)
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
  }
  return p;
}

// separable Gaussian, horizontal pass: one CTA per (row, image); the reflected row sits in shared memory
template <int R>
__global__ void __launch_bounds__(256) gauss_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const float* __restrict__ taps, int H, int W) {
  pdl_sync();
  extern __shared__ float sm[];
  float* s_taps = sm;                 // 2R + 1
  float* s_row = sm + 2 * R + 4;      // W + 2R
  const long long base = (static_cast<long long>(blockIdx.y) * H + blockIdx.x) * W;
  for (int i = threadIdx.x; i < 2 * R + 1; i += blockDim.x) s_taps[i] = __ldg(taps + i);
  for (int i = threadIdx.x; i < W + 2 * R; i += blockDim.x) s_row[i] = __ldg(in + base + reflect101(i - R, W));
  __syncthreads();
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float acc = 0.0f;
#pragma unroll 4
    for (int t = 0; t < 2 * R + 1; ++t) acc = fmaf(s_taps[t], s_row[x + t], acc);
    out[base + x] = acc;
  }
}

// vertical pass: CTA = 32 columns x 64 rows of one image; thread (tx, ty) computes rows ty, ty + 8, ...
// finish != 0: clear a `thres`-pixel frame and scale (inverse_faces :311-318)
template <int R>
__global__ void __launch_bounds__(256) gauss_cols_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const float* __restrict__ taps, int H, int W, int finish,
                                                         int thres, float scale) {
  pdl_sync();
  extern __shared__ float sm[];
  constexpr int kRows = 64;
  float* s_taps = sm;                 // 2R + 1
  float* s_in = sm + 2 * R + 4;       // (kRows + 2R) x 32
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + tx, y0 = blockIdx.y * kRows;
  const long long base = static_cast<long long>(blockIdx.z) * H * W;
  for (int i = threadIdx.x; i < 2 * R + 1; i += blockDim.x) s_taps[i] = __ldg(taps + i);
  for (int r = ty; r < kRows + 2 * R; r += 8)
    s_in[r * 32 + tx] = (x < W) ? __ldg(in + base + static_cast<long long>(reflect101(y0 + r - R, H)) * W + x) : 0.0f;
  __syncthreads();
  if (x >= W) return;
  for (int r = ty; r < kRows; r += 8) {
    const int y = y0 + r;
    if (y >= H) break;
    float acc = 0.0f;
#pragma unroll 4
    for (int t = 0; t < 2 * R + 1; ++t) acc = fmaf(s_taps[t], s_in[(r + t) * 32 + tx], acc);
    if (finish) {
      const bool frame = y < thres || y >= H - thres || x < thres || x >= W - thres;
      acc = frame ? 0.0f : __fdiv_rn(acc, scale);
    }
    out[base + static_cast<long long>(y) * W + x] = acc;
  }
}

// x_with_face = x0 (1 - m) + face m; [clamp]; out = w x0 + (1 - w) x_with_face   (gaussian_diffusion.py:488-496)
__global__ void __launch_bounds__(256) aux_blend_kernel(const float* __restrict__ x0, const float* __restrict__ face,
                                                        const float* __restrict__ mask, float* __restrict__ out,
                                                        float w, float omw, int clip, long long hw, int C,
                                                        long long total) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long plane = i / hw, p = i - plane * hw;
    const long long n = plane / C;
    const float m = __ldg(mask + n * hw + p), a = __ldg(x0 + i), f = __ldg(face + i);
    float xw = __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, m)), __fmul_rn(f, m));
    if (clip) xw = fminf(fmaxf(xw, -1.0f), 1.0f);
    out[i] = __fadd_rn(__fmul_rn(w, a), __fmul_rn(omw, xw));
  }
}

}  // namespace

extern "C" int flair_warp_affine_cubic_f32(const float* src, float* dst, const double* minv, int N, int C, int Hs,
                                           int Ws, int Hd, int Wd, const float* border, int in_mode, int out_mode,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(src && dst && minv, "flair_warp_affine_cubic_f32: null pointer");
  FLAIR_REQUIRE(N > 0 && C > 0 && C <= 4 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && N <= 65535,
                "flair_warp_affine_cubic_f32: bad extents N=%d C=%d src %dx%d dst %dx%d", N, C, Hs, Ws, Hd, Wd);
  FLAIR_REQUIRE((in_mode == 0 || in_mode == 1) && (out_mode == 0 || out_mode == 1),
                "flair_warp_affine_cubic_f32: bad modes %d %d", in_mode, out_mode);
  static FlairPerDeviceOnce tab_once;
  if (tab_once.first()) {
    float tab[kTab * 4];
    cubic_table_host(tab);
    FLAIR_CHECK_CUDA(cudaMemcpyToSymbol(c_cubic, tab, sizeof(tab)));
  }
  WarpArgs a{};
  a.src = src; a.dst = dst; a.minv = minv;
  a.N = N; a.C = C; a.Hs = Hs; a.Ws = Ws; a.Hd = Hd; a.Wd = Wd;
  for (int c = 0; c < 4; ++c) a.border[c] = (border && c < C) ? border[c] : 0.0f;
  a.in_mode = in_mode; a.out_mode = out_mode;
  dim3 grid(ceil_div(Wd, 32), ceil_div(Hd, 8), N);
  FLAIR_CHECK_CUDA(flair_launch(warp_affine_cubic_kernel, grid, dim3(256), 0, stream, a));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_parse_mask_f32(const float* logits, float* mask, int N, int classes, int H, int W,
                                    unsigned int lut_bits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(logits && mask && N > 0 && classes > 0 && classes <= 32 && H > 0 && W > 0,
                "flair_parse_mask_f32: bad arguments N=%d classes=%d %dx%d", N, classes, H, W);
  const long long hw = static_cast<long long>(H) * W, total = hw * N;
  FLAIR_CHECK_CUDA(flair_launch(parse_mask_kernel, dim3(static_cast<unsigned>(ceil_div_ll(total, 256))), dim3(256), 0,
                                stream, logits, mask, classes, hw, total, lut_bits));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_gaussian_blur_f32(const float* in, float* out, float* tmp, const float* taps, int ksize, int N,
                                       int H, int W, int finish, int thres, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(in && out && tmp && taps, "flair_gaussian_blur_f32: null pointer");
  FLAIR_REQUIRE(ksize == 101, "flair_gaussian_blur_f32: ksize %d (the reference uses 101 only)", ksize);
  FLAIR_REQUIRE(N > 0 && N <= 65535 && H > 1 && W > 1 && H <= 65535, "flair_gaussian_blur_f32: bad extents N=%d %dx%d", N, H, W);
  constexpr int R = 50;
  const size_t smem_rows = sizeof(float) * (2 * R + 4 + W + 2 * R);
  const size_t smem_cols = sizeof(float) * (2 * R + 4 + (64 + 2 * R) * 32);
  FLAIR_REQUIRE(smem_rows <= 200 * 1024, "flair_gaussian_blur_f32: row too wide for shared memory (W=%d)", W);
  static FlairPerDeviceOnce attr;
  if (attr.first()) {
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(gauss_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(gauss_cols_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem_cols)));
  }
  FLAIR_CHECK_CUDA(flair_launch(gauss_rows_kernel<R>, dim3(H, N), dim3(256), smem_rows, stream, in, tmp, taps, H, W));
  FLAIR_CHECK_LAUNCH();
  FLAIR_CHECK_CUDA(flair_launch(gauss_cols_kernel<R>, dim3(ceil_div(W, 32), ceil_div(H, 64), N), dim3(256), smem_cols,
                                stream, static_cast<const float*>(tmp), out, taps, H, W, finish, thres, scale));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_aux_blend_f32(const float* x0, const float* face, const float* mask, float* out, double w, int N,
                                   int C, int H, int W, int clip, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x0 && face && mask && out && N > 0 && C > 0 && H > 0 && W > 0, "flair_aux_blend_f32: bad arguments");
  const long long hw = static_cast<long long>(H) * W, total = hw * C * N;
  long long blocks = ceil_div_ll(total, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  FLAIR_CHECK_CUDA(flair_launch(aux_blend_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, x0, face,
                                mask, out, static_cast<float>(w), static_cast<float>(1.0 - w), clip, hw, C, total));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
