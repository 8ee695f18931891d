// Data-consistency operators in fp32: blur x4 stencils (pseudoSR), DCT-domain JPEG,
// small dense "sandwich" products for the separable bicubic SRConv.
//
// Reference: guided_diffusion/pseudoSR.py:15-44,180-244 (Filter_Layer depth-wise convs with
// ReplicationPad2d), guided_diffusion/jpeg.py:7-167 + dct.py:167-202, guided_diffusion/
// restore_util.py:54-82,102-227.  All are HBM/latency-bound integer-free fp32 kernels; FMA
// contraction is avoided where the reference rounds (quantisation), accepted elsewhere.
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------ blur down (HR -> LR)
// One CTA = one 16x16 LR tile of one plane; the (16*sf + k - 1)^2 HR footprint is staged in
// shared memory once (each HR pixel feeds ~(k/sf)^2 outputs), replicate-clamped at the borders.
// Register tiling: every thread computes kOutPerThread ADJACENT outputs of one row, so the input window of a tap
// row is loaded from shared memory once and reused (one output per thread needed 2 shared loads per FMA and ran
// at 0.9 TB/s).  Each output still accumulates its taps in (u, v) order: bit-identical to the simple version.
constexpr int kTileH = 16, kTileW = 64, kOutPerThread = 4;

// stage rows [i0, i0 + span_h) x cols [j0, j0 + span_w) of one plane (replicate padding) into shared memory.
// The in-map part of every row is read with 16-byte loads, four rows per warp in flight (a 16 x 64 LR tile of the
// x4 blur needs 72 full 256-pixel rows: with one 4-byte load per lane and 4 in flight the CTA kept ~4 KB outstanding
// and blur_down ran at 1.06 TB/s, profiles/r02_hbm_probe.txt); the replicated border columns are filled afterwards.
//
// SWZ: the 16-byte chunks of a row are stored XOR-swizzled (chunk q at q ^ ((q >> 3) & 7); pitch a multiple of 32
// floats), see blur_down_tiled_kernel.
__device__ __forceinline__ int swz_col(int c) { return (((c >> 2) ^ ((c >> 5) & 7)) << 2) | (c & 3); }

template <bool SWZ = false>
__device__ __forceinline__ void stage_tile(const float* __restrict__ xp, float* __restrict__ s_in, int i0, int j0,
                                           int span_h, int span_w, int pitch, int H, int W) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(xp) & 15) == 0;
  // in-map columns [ja, jb) of the span; their 16-byte aligned core [ja4, jb4)
  const int ja = j0 > 0 ? j0 : 0, jb = (j0 + span_w < W) ? j0 + span_w : W;
  int ja4 = (ja + 3) & ~3, jb4 = jb & ~3;
  if (!vec_ok || jb4 <= ja4) { ja4 = ja; jb4 = ja; }   // no vector core
  const int nvec = (jb4 - ja4) >> 2;
  // whole-tile fast path (the x4 blur tile: 72 rows x 64 vectors, 8 warps): every load of the thread is issued before
  // the first store, 18 x 16 bytes in flight per thread (batches of 4 rows left the CTA latency-bound: 28 us)
  constexpr int kFastRows = 9, kFastVec = 2;
  if (SWZ && nvec <= 32 * kFastVec && span_h <= kFastRows * nwarps) {
    float4 raw[kFastRows][kFastVec];
#pragma unroll
    for (int q = 0; q < kFastRows; ++q) {
      const int ii = warp + q * nwarps;
      const float4* rp = reinterpret_cast<const float4*>(xp + static_cast<long long>(clampi(i0 + ii, 0, H - 1)) * W + ja4);
#pragma unroll
      for (int t = 0; t < kFastVec; ++t)
        if (ii < span_h && lane + 32 * t < nvec) raw[q][t] = __ldg(rp + lane + 32 * t);
    }
#pragma unroll
    for (int q = 0; q < kFastRows; ++q) {
      const int ii = warp + q * nwarps;
      float* sp = s_in + ii * pitch;
#pragma unroll
      for (int t = 0; t < kFastVec; ++t) {
        if (ii < span_h && lane + 32 * t < nvec) {
          const int c = (ja4 - j0) + 4 * (lane + 32 * t);
          sp[swz_col(c)] = raw[q][t].x; sp[swz_col(c + 1)] = raw[q][t].y;
          sp[swz_col(c + 2)] = raw[q][t].z; sp[swz_col(c + 3)] = raw[q][t].w;
        }
      }
    }
  } else
  for (int ii0 = warp * 4; ii0 < span_h; ii0 += nwarps * 4) {
    constexpr int kRows = 4;
    for (int v = lane; v < nvec; v += 32) {
      float4 raw[kRows];
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
        const int ii = ii0 + q;
        if (ii < span_h)
          raw[q] = __ldg(reinterpret_cast<const float4*>(xp + static_cast<long long>(clampi(i0 + ii, 0, H - 1)) * W + ja4) + v);
      }
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
        const int ii = ii0 + q;
        if (ii < span_h) {
          float* sp = s_in + ii * pitch;
          const int c = (ja4 - j0) + 4 * v;                      // (not 16-byte aligned in general)
          if (SWZ) {
            sp[swz_col(c)] = raw[q].x; sp[swz_col(c + 1)] = raw[q].y; sp[swz_col(c + 2)] = raw[q].z; sp[swz_col(c + 3)] = raw[q].w;
          } else {
            sp[c] = raw[q].x; sp[c + 1] = raw[q].y; sp[c + 2] = raw[q].z; sp[c + 3] = raw[q].w;
          }
        }
      }
    }
  }
  // everything outside the vector core: left part [0, ja4 - j0), right part [jb4 - j0, span_w)
  const int left = ja4 - j0, right0 = jb4 - j0;
  const int rest = left + (span_w - right0);
  for (int idx = threadIdx.x; idx < span_h * rest; idx += blockDim.x) {
    const int ii = idx / rest, k = idx - ii * rest;
    const int jj = k < left ? k : right0 + (k - left);
    s_in[ii * pitch + (SWZ ? swz_col(jj) : jj)] =
        __ldg(xp + static_cast<long long>(clampi(i0 + ii, 0, H - 1)) * W + clampi(j0 + jj, 0, W - 1));
  }
}

// Shared-memory layout: a thread's tap-row window is the 24 floats at column 16 * ln, so the lanes of a warp are 64 bytes
// apart and every scalar load hit ONE bank pair 16 ways (39 us for 50 MB at 64 frames: the kernel was bound by
// shared-memory wavefronts, not HBM).  The window is now read as six 16-byte loads from an XOR-swizzled row (chunk q
// at q ^ ((q >> 3) & 7)): the eight lanes of a quarter warp hit eight different bank groups.
constexpr int blur_down_pitch(int span_w) { return (((span_w + 3) / 4 + 7) & ~7) * 4; }

template <int K, int SF>
__global__ void __launch_bounds__(256)
blur_down_tiled_kernel(const float* __restrict__ x, float* __restrict__ lr, const float* __restrict__ taps, int pre,
                       int H, int W) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];
  constexpr int r = K / 2;
  constexpr int tile_w = 64, tile_h = 16;                       // LR outputs per CTA (16 threads x 4 per row)
  constexpr int span_h = tile_h * SF + K - 1, span_w = tile_w * SF + K - 1;
  constexpr int pitch = blur_down_pitch(span_w);
  constexpr int win = K + (kOutPerThread - 1) * SF;             // HR inputs of one tap row for 4 adjacent outputs
  constexpr int win4 = (win + 3) / 4;
  static_assert(kOutPerThread * SF == 16 && (tile_w / kOutPerThread - 1) * 16 + win4 * 4 <= pitch, "window layout");
  const int h = H / SF, w = W / SF;
  float* s_taps = sm;
  float* s_in = sm + ((K * K + 3) & ~3);
  const int plane = blockIdx.z;
  const int m0 = blockIdx.y * tile_h, n0 = blockIdx.x * tile_w;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) s_taps[i] = __ldg(taps + i);
  stage_tile<true>(x + static_cast<long long>(plane) * H * W, s_in, m0 * SF + pre - r, n0 * SF + pre - r, span_h,
                   span_w, pitch, H, W);
  __syncthreads();
  const int lm = threadIdx.x >> 4, ln = threadIdx.x & 15;
  float acc[kOutPerThread] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int u = 0; u < K; ++u) {
    const float4* row = reinterpret_cast<const float4*>(s_in + (lm * SF + u) * pitch);
    float wv[win4 * 4];
#pragma unroll
    for (int j = 0; j < win4; ++j) {
      const int q = 4 * ln + j;
      const float4 f = row[q ^ ((q >> 3) & 7)];
      wv[4 * j] = f.x; wv[4 * j + 1] = f.y; wv[4 * j + 2] = f.z; wv[4 * j + 3] = f.w;
    }
#pragma unroll
    for (int v = 0; v < K; ++v) {
      const float t = s_taps[u * K + v];
#pragma unroll
      for (int o = 0; o < kOutPerThread; ++o) acc[o] = fmaf(t, wv[o * SF + v], acc[o]);
    }
  }
  const int m = m0 + lm;
  if (m >= h) return;
#pragma unroll
  for (int o = 0; o < kOutPerThread; ++o) {
    const int n = n0 + ln * kOutPerThread + o;
    if (n < w) lr[(static_cast<long long>(plane) * h + m) * w + n] = acc[o];
  }
}

// generic (any k, sf): one output per thread
__global__ void __launch_bounds__(256)
blur_down_kernel(const float* __restrict__ x, float* __restrict__ lr, const float* __restrict__ taps,
                 int k, int sf, int pre, int H, int W) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ float sm[];
  const int h = H / sf, w = W / sf;
  const int r = k / 2;
  const int tile = 16;
  const int span = tile * sf + k - 1;  // HR rows/cols needed (a few unused at the far edge)
  float* s_taps = sm;
  float* s_in = sm + k * k;
  const int plane = blockIdx.z;
  const int m0 = blockIdx.y * tile, n0 = blockIdx.x * tile;
  const float* xp = x + static_cast<long long>(plane) * H * W;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) s_taps[i] = __ldg(taps + i);
  const int i0 = m0 * sf + pre - r, j0 = n0 * sf + pre - r;
  for (int idx = threadIdx.x; idx < span * span; idx += blockDim.x) {
    const int ii = idx / span, jj = idx % span;
    s_in[idx] = __ldg(xp + static_cast<long long>(clampi(i0 + ii, 0, H - 1)) * W + clampi(j0 + jj, 0, W - 1));
  }
  __syncthreads();
  const int lm = threadIdx.x / tile, ln = threadIdx.x % tile;
  const int m = m0 + lm, n = n0 + ln;
  if (m >= h || n >= w) return;
  float acc = 0.0f;
  for (int u = 0; u < k; ++u) {
    const float* row = s_in + (lm * sf + u) * span + ln * sf;
#pragma unroll 3
    for (int v = 0; v < k; ++v) acc = fmaf(s_taps[u * k + v], row[v], acc);
  }
  lr[(static_cast<long long>(plane) * h + m) * w + n] = acc;
}

// ------------------------------------------------------------------ same-size k x k filter
// K x K same-size filter, 4 adjacent outputs per thread, the (K + 3)-wide input window of a tap row held in
// registers and read with 16-byte shared loads: ~0.14 shared loads per FMA instead of 2.  The 39 x 39 inverse filter
// (1521 FMAs per output, LR maps) is FMA-bound, not HBM-bound.
template <int K>
__global__ void __launch_bounds__(256)
filter_same_tiled_kernel(const float* __restrict__ x, const float* __restrict__ sub, float* __restrict__ out,
                         const float* __restrict__ taps, int H, int W) {
  pdl_sync();
  extern __shared__ float sm[];
  constexpr int r = K / 2;
  constexpr int tile_w = 64, tile_h = 16;
  constexpr int span_h = tile_h + K - 1, span_w = tile_w + K - 1;
  constexpr int pitch = (span_w + 3) & ~3;                 // rows 16-byte aligned for float4 reads
  constexpr int win4 = (K + kOutPerThread - 1 + 3) / 4;    // float4 loads per tap row
  constexpr int kpad = (K + 3) & ~3;
  float* s_taps = sm;                                      // [K][kpad]
  float* s_in = sm + K * kpad;
  const int plane = blockIdx.z;
  const int m0 = blockIdx.y * tile_h, n0 = blockIdx.x * tile_w;
  for (int i = threadIdx.x; i < K * kpad; i += blockDim.x) {
    const int u = i / kpad, v = i - u * kpad;
    s_taps[i] = (v < K) ? __ldg(taps + u * K + v) : 0.0f;
  }
  stage_tile(x + static_cast<long long>(plane) * H * W, s_in, m0 - r, n0 - r, span_h, span_w, pitch, H, W);
  __syncthreads();
  const int lm = threadIdx.x >> 4, ln = threadIdx.x & 15;
  float acc[kOutPerThread] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int u = 0; u < K; ++u) {
    const float4* row = reinterpret_cast<const float4*>(s_in + (lm + u) * pitch + ln * kOutPerThread);
    const float4* trow = reinterpret_cast<const float4*>(s_taps + u * kpad);
    float wv[win4 * 4];
#pragma unroll
    for (int j = 0; j < win4; ++j) {
      const float4 f = row[j];
      wv[4 * j] = f.x; wv[4 * j + 1] = f.y; wv[4 * j + 2] = f.z; wv[4 * j + 3] = f.w;
    }
#pragma unroll
    for (int v4 = 0; v4 < kpad / 4; ++v4) {
      const float4 t4 = trow[v4];
      const float tt[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int v = v4 * 4 + e;
        if (v < K) {
#pragma unroll
          for (int o = 0; o < kOutPerThread; ++o) acc[o] = fmaf(tt[e], wv[v + o], acc[o]);
        }
      }
    }
  }
  const int m = m0 + lm;
  if (m >= H) return;
#pragma unroll
  for (int o = 0; o < kOutPerThread; ++o) {
    const int n = n0 + ln * kOutPerThread + o;
    if (n < W) {
      const long long oidx = (static_cast<long long>(plane) * H + m) * W + n;
      out[oidx] = sub ? acc[o] - __ldg(sub + oidx) : acc[o];
    }
  }
}

// generic (any odd k): one output per thread
__global__ void __launch_bounds__(256)
filter_same_kernel(const float* __restrict__ x, const float* __restrict__ sub, float* __restrict__ out,
                   const float* __restrict__ taps, int k, int H, int W) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ float sm[];
  const int r = k / 2;
  const int tile = 16;
  const int span = tile + k - 1;
  float* s_taps = sm;
  float* s_in = sm + k * k;
  const int plane = blockIdx.z;
  const int m0 = blockIdx.y * tile, n0 = blockIdx.x * tile;
  const float* xp = x + static_cast<long long>(plane) * H * W;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) s_taps[i] = __ldg(taps + i);
  for (int idx = threadIdx.x; idx < span * span; idx += blockDim.x) {
    const int ii = idx / span, jj = idx % span;
    s_in[idx] = __ldg(xp + static_cast<long long>(clampi(m0 - r + ii, 0, H - 1)) * W +
                      clampi(n0 - r + jj, 0, W - 1));
  }
  __syncthreads();
  const int lm = threadIdx.x / tile, ln = threadIdx.x % tile;
  const int m = m0 + lm, n = n0 + ln;
  if (m >= H || n >= W) return;
  float acc = 0.0f;
  for (int u = 0; u < k; ++u) {
    const float* row = s_in + (lm + u) * span + ln;
    for (int v = 0; v < k; ++v) acc = fmaf(s_taps[u * k + v], row[v], acc);
  }
  const long long o = (static_cast<long long>(plane) * H + m) * W + n;
  out[o] = sub ? acc - __ldg(sub + o) : acc;
}

// ------------------------------------------------------------------ blur up (LR -> HR), polyphase
__global__ void __launch_bounds__(256)
blur_up_kernel(const float* __restrict__ lr, float* __restrict__ hr, const float* __restrict__ taps,
               int k, int sf, int pre, int planes, int H, int W) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const int h = H / sf, w = W / sf, r = k / 2;
  const long long total = static_cast<long long>(planes) * H * W;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(e % W);
    const int i = static_cast<int>((e / W) % H);
    const long long plane = e / (static_cast<long long>(H) * W);
    const float* q = lr + plane * h * w;
    const int u0 = ((pre + r - i) % sf + sf) % sf, v0 = ((pre + r - j) % sf + sf) % sf;
    float acc = 0.0f;
    for (int u = u0; u < k; u += sf) {
      const int zi = i + u - r;
      if (zi < 0 || zi >= H) continue;
      const int m = (zi - pre) / sf;
      for (int v = v0; v < k; v += sf) {
        const int zj = j + v - r;
        if (zj < 0 || zj >= W) continue;
        const int n = (zj - pre) / sf;
        acc = fmaf(__ldg(taps + u * k + v), __ldg(q + m * w + n), acc);
      }
    }
    hr[e] = acc;
  }
}

// ------------------------------------------------------------------ JPEG
// One CTA (256 threads) = one 16x16 macroblock of one image: 4 luma blocks + Cb + Cr blocks.
// smem planes: 6 blocks of 8x8.  T = M b M^T done as two 8-term passes per coefficient.
__device__ __forceinline__ void xform8(const float* __restrict__ M, float (*blk)[8][8], float (*tmp)[8][8],
                                       int nblk) {
  // blk <- M blk M^T following apply_linear_2d (dct.py:194-202): rows first, then columns
  for (int idx = threadIdx.x; idx < nblk * 64; idx += blockDim.x) {
    const int b = idx >> 6, rr = (idx >> 3) & 7, kk = idx & 7;
    float acc = 0.0f;
#pragma unroll
    for (int n = 0; n < 8; ++n) acc = fmaf(blk[b][rr][n], M[kk * 8 + n], acc);
    tmp[b][rr][kk] = acc;  // X1[r,k] = sum_n b[r,n] M[k,n]
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nblk * 64; idx += blockDim.x) {
    const int b = idx >> 6, u = (idx >> 3) & 7, v = idx & 7;
    float acc = 0.0f;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) acc = fmaf(tmp[b][rr][v], M[u * 8 + rr], acc);
    blk[b][u][v] = acc;  // out[u,v] = sum_r M[u,r] X1[r,v]
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
jpeg_kernel(int mode, const float* __restrict__ x, float* __restrict__ luma, float* __restrict__ chroma,
            float* __restrict__ out, const float* __restrict__ dct, const float* __restrict__ idct,
            const float* __restrict__ q_luma, const float* __restrict__ q_chroma, int h, int w) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ float blk[6][8][8];
  __shared__ float tmp[6][8][8];
  __shared__ float s_d[64], s_di[64], s_q1[64], s_q2[64];
  const int n = blockIdx.z;
  const int by = blockIdx.y * 16, bx = blockIdx.x * 16;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const long long hw = static_cast<long long>(h) * w;
  const int h2 = h / 2, w2 = w / 2;
  if (threadIdx.x < 64) {
    s_d[threadIdx.x] = __ldg(dct + threadIdx.x);
    s_di[threadIdx.x] = __ldg(idct + threadIdx.x);
    s_q1[threadIdx.x] = __ldg(q_luma + threadIdx.x);
    s_q2[threadIdx.x] = __ldg(q_chroma + threadIdx.x);
  }
  const int lb = (ty >> 3) * 2 + (tx >> 3);  // luma block of this pixel
  if (mode != 1) {
    // ---- colour transform + 4:2:0 decimation (jpeg.py:7-33,75-79)
    const float* xp = x + static_cast<long long>(n) * 3 * hw + static_cast<long long>(by + ty) * w + bx + tx;
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      rgb[c] = __fmul_rn(__fdiv_rn(__fadd_rn(__ldg(xp + c * hw), 1.0f), 2.0f), 255.0f);
    const float m[3][3] = {{0.299f, 0.587f, 0.114f}, {-0.1687f, -0.3313f, 0.5f}, {0.5f, -0.4187f, -0.0813f}};
    float ycc[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
      ycc[kx] = __fadd_rn(__fadd_rn(__fmul_rn(rgb[0], m[kx][0]), __fmul_rn(rgb[1], m[kx][1])),
                          __fmul_rn(rgb[2], m[kx][2]));
    blk[lb][ty & 7][tx & 7] = __fsub_rn(ycc[0], 128.0f);
    if (((ty | tx) & 1) == 0) {
      blk[4][ty >> 1][tx >> 1] = __fsub_rn(__fadd_rn(ycc[1], 128.0f), 128.0f);
      blk[5][ty >> 1][tx >> 1] = __fsub_rn(__fadd_rn(ycc[2], 128.0f), 128.0f);
    }
    __syncthreads();
    xform8(s_d, blk, tmp, 6);
    // ---- quantise (true division, round-half-even)
    for (int idx = threadIdx.x; idx < 6 * 64; idx += blockDim.x) {
      const int b = idx >> 6, e = idx & 63;
      const float q = (b < 4) ? s_q1[e] : s_q2[e];
      (&blk[b][0][0])[e] = rintf(__fdiv_rn((&blk[b][0][0])[e], q));
    }
    __syncthreads();
    if (mode == 0) {
      luma[static_cast<long long>(n) * hw + static_cast<long long>(by + ty) * w + bx + tx] =
          blk[lb][ty & 7][tx & 7];
      if (threadIdx.x < 128) {
        const int c = threadIdx.x >> 6, e = threadIdx.x & 63;
        chroma[(static_cast<long long>(n) * 2 + c) * h2 * w2 +
               static_cast<long long>(by / 2 + (e >> 3)) * w2 + bx / 2 + (e & 7)] = (&blk[4 + c][0][0])[e];
      }
      return;
    }
  } else {
    __syncthreads();
    blk[lb][ty & 7][tx & 7] =
        __ldg(luma + static_cast<long long>(n) * hw + static_cast<long long>(by + ty) * w + bx + tx);
    if (threadIdx.x < 128) {
      const int c = threadIdx.x >> 6, e = threadIdx.x & 63;
      (&blk[4 + c][0][0])[e] = __ldg(chroma + (static_cast<long long>(n) * 2 + c) * h2 * w2 +
                                     static_cast<long long>(by / 2 + (e >> 3)) * w2 + bx / 2 + (e & 7));
    }
    __syncthreads();
  }
  // ---- dequantise + inverse transform (jpeg.py:129-143)
  for (int idx = threadIdx.x; idx < 6 * 64; idx += blockDim.x) {
    const int b = idx >> 6, e = idx & 63;
    const float q = (b < 4) ? s_q1[e] : s_q2[e];
    (&blk[b][0][0])[e] = __fmul_rn((&blk[b][0][0])[e], q);
  }
  __syncthreads();
  xform8(s_di, blk, tmp, 6);
  // ---- +128, nearest chroma up-sampling, YCbCr -> RGB, back to [-1,1] (jpeg.py:145-165)
  const float yv = __fadd_rn(blk[lb][ty & 7][tx & 7], 128.0f);
  const float cbv = __fsub_rn(__fadd_rn(blk[4][ty >> 1][tx >> 1], 128.0f), 128.0f);
  const float crv = __fsub_rn(__fadd_rn(blk[5][ty >> 1][tx >> 1], 128.0f), 128.0f);
  const float m2[3][3] = {{1.00000000e00f, -3.68199903e-05f, 1.40198758e00f},
                          {1.00000000e00f, -3.44113281e-01f, -7.14103821e-01f},
                          {1.00000000e00f, 1.77197812e00f, -1.34583413e-04f}};
  float* op = out + static_cast<long long>(n) * 3 * hw + static_cast<long long>(by + ty) * w + bx + tx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fadd_rn(__fadd_rn(__fmul_rn(yv, m2[c][0]), __fmul_rn(cbv, m2[c][1])),
                              __fmul_rn(crv, m2[c][2]));
    op[c * hw] = __fsub_rn(__fmul_rn(__fdiv_rn(v, 255.0f), 2.0f), 1.0f);
  }
}

// ------------------------------------------------------------------ batched fp32 GEMM (small)
// C[b] (M x N) = A[b] (M x K) * B[b] (K x N) [- Sub[b]], row-major, batch strides may be 0 (shared).
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sA, const float* __restrict__ B, long long sB,
                const float* __restrict__ Sub, float* __restrict__ Cc, int M, int N, int K) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int b = blockIdx.z;
  const float* Ap = A + b * sA;
  const float* Bp = B + b * sB;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = ty + i * 8;
      As[rr][tx] = (row0 + rr < M && k0 + tx < K) ? __ldg(Ap + static_cast<long long>(row0 + rr) * K + k0 + tx) : 0.f;
      Bs[rr][tx] = (k0 + rr < K && col0 + tx < N) ? __ldg(Bp + static_cast<long long>(k0 + rr) * N + col0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float bv = Bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + i * 8][kk], bv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = row0 + ty + i * 8, cc = col0 + tx;
    if (rr < M && cc < N) {
      const long long o = static_cast<long long>(b) * M * N + static_cast<long long>(rr) * N + cc;
      Cc[o] = Sub ? acc[i] - __ldg(Sub + o) : acc[i];
    }
  }
}

}  // namespace

extern "C" int flair_blur_down_f32(const float* x, float* lr, const float* taps, int k, int sf, int pre,
                                   int planes, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && lr && taps, "flair_blur_down_f32: null pointer");
  FLAIR_REQUIRE(k > 0 && (k & 1) && sf > 0 && H % sf == 0 && W % sf == 0 && pre >= 0 && pre < sf,
                "flair_blur_down_f32: bad geometry k=%d sf=%d pre=%d H=%d W=%d", k, sf, pre, H, W);
  if (k == 9 && sf == 4) {  // the FLAIR blur operator (pseudoSR.py: 9 x 9 ds_kernel, factor 4)
    const size_t smem = sizeof(float) * (84 + static_cast<size_t>(16 * 4 + 8) * blur_down_pitch(64 * 4 + 8));
    static FlairPerDeviceOnce attr;
    if (attr.first())
      FLAIR_CHECK_CUDA(cudaFuncSetAttribute(blur_down_tiled_kernel<9, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
    dim3 grid(ceil_div(W / sf, 64), ceil_div(H / sf, 16), planes);
    FLAIR_CHECK_CUDA(flair_launch(blur_down_tiled_kernel<9, 4>, dim3(grid), dim3(256), smem, stream, x, lr, taps, pre, H, W));
    FLAIR_CHECK_LAUNCH();
    return 0;
  }
  const int span = 16 * sf + k - 1;
  const size_t smem = sizeof(float) * (k * k + span * span);
  FLAIR_REQUIRE(smem <= 200 * 1024, "flair_blur_down_f32: footprint too large for shared memory");
  if (smem > 48 * 1024)
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(blur_down_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
  dim3 grid(ceil_div(W / sf, 16), ceil_div(H / sf, 16), planes);
  FLAIR_CHECK_CUDA(flair_launch(blur_down_kernel, dim3(grid), dim3(256), smem, stream, x, lr, taps, k, sf, pre, H, W));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_filter_same_f32(const float* x, const float* sub, float* out, const float* taps,
                                     int k, int planes, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && out && taps && k > 0 && (k & 1), "flair_filter_same_f32: bad arguments");
  if (k == 39) {  // the FLAIR inverse filter inv_hTh (pseudoSR.py:123-171)
    constexpr int K = 39, kpad = 40, pitch = (64 + K - 1 + 3) & ~3;
    const size_t smem = sizeof(float) * (K * kpad + static_cast<size_t>(16 + K - 1) * pitch);
    static FlairPerDeviceOnce attr;
    if (attr.first())
      FLAIR_CHECK_CUDA(cudaFuncSetAttribute(filter_same_tiled_kernel<39>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
    dim3 grid(ceil_div(W, 64), ceil_div(H, 16), planes);
    FLAIR_CHECK_CUDA(flair_launch(filter_same_tiled_kernel<39>, dim3(grid), dim3(256), smem, stream, x, sub, out, taps, H, W));
    FLAIR_CHECK_LAUNCH();
    return 0;
  }
  const int span = 16 + k - 1;
  const size_t smem = sizeof(float) * (k * k + span * span);
  FLAIR_REQUIRE(smem <= 200 * 1024, "flair_filter_same_f32: footprint too large for shared memory");
  if (smem > 48 * 1024)
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(filter_same_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
  dim3 grid(ceil_div(W, 16), ceil_div(H, 16), planes);
  FLAIR_CHECK_CUDA(flair_launch(filter_same_kernel, dim3(grid), dim3(256), smem, stream, x, sub, out, taps, k, H, W));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_blur_up_f32(const float* lr, float* hr, const float* taps, int k, int sf, int pre,
                                 int planes, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(lr && hr && taps && k > 0 && (k & 1) && sf > 0 && H % sf == 0 && W % sf == 0,
                "flair_blur_up_f32: bad arguments");
  FLAIR_REQUIRE(pre > 0 && pre < sf - 1,
                "flair_blur_up_f32: polyphase form needs 0 < pre < sf-1 (zero border rows), got %d", pre);
  const long long total = static_cast<long long>(planes) * H * W;
  long long blocks = ceil_div_ll(total, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  FLAIR_CHECK_CUDA(flair_launch(blur_up_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, lr, hr, taps, k, sf, pre, planes, H, W));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_jpeg_f32(int mode, const float* x, float* luma, float* chroma, float* out,
                              const float* dct, const float* idct, const float* q_luma,
                              const float* q_chroma, int N, int h, int w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(mode >= 0 && mode <= 2, "flair_jpeg_f32: mode must be 0, 1 or 2");
  FLAIR_REQUIRE(dct && idct && q_luma && q_chroma, "flair_jpeg_f32: null table pointer");
  FLAIR_REQUIRE(N > 0 && h > 0 && w > 0 && h % 16 == 0 && w % 16 == 0,
                "flair_jpeg_f32: h and w must be multiples of 16 (got %d x %d)", h, w);
  if (mode != 1) FLAIR_REQUIRE(x, "flair_jpeg_f32: x is NULL");
  if (mode != 2) FLAIR_REQUIRE(luma && chroma, "flair_jpeg_f32: coefficient planes are NULL");
  if (mode != 0) FLAIR_REQUIRE(out, "flair_jpeg_f32: out is NULL");
  dim3 grid(w / 16, h / 16, N);
  FLAIR_CHECK_CUDA(flair_launch(jpeg_kernel, dim3(grid), dim3(256), 0, stream, mode, x, luma, chroma, out, dct, idct, q_luma, q_chroma, h, w));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

// out[pl] (p x s) = L (p x q) * X[pl] (q x r) * Rm (r x s) - sub[pl]; workspace (planes, p, r)
extern "C" int flair_sandwich_f32(const float* L, const float* X, const float* Rm, const float* sub,
                                  float* out, int planes, int p, int q, int r, int s, float* workspace,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(L && X && Rm && out && workspace, "flair_sandwich_f32: null pointer");
  FLAIR_REQUIRE(planes > 0 && planes < 65536 && p > 0 && q > 0 && r > 0 && s > 0,
                "flair_sandwich_f32: bad sizes");
  dim3 g1(ceil_div(r, 32), ceil_div(p, 32), planes);
  FLAIR_CHECK_CUDA(flair_launch(gemm_f32_kernel, dim3(g1), dim3(256), 0, stream, L, 0, X, static_cast<long long>(q) * r, nullptr, workspace, p, r, q));
  FLAIR_CHECK_LAUNCH();
  dim3 g2(ceil_div(s, 32), ceil_div(p, 32), planes);
  FLAIR_CHECK_CUDA(flair_launch(gemm_f32_kernel, dim3(g2), dim3(256), 0, stream, workspace, static_cast<long long>(p) * r, Rm, 0, sub, out, p, s, r));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
