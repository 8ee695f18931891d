// BasicVSR++ support kernels on channels-last maps: flow-guided bilinear warping, flow
// composition, modulated deformable im2col (second-order alignment), per-pixel scaling.
//
// Replaces mmedit `flow_warp` (F.grid_sample bilinear, zeros padding, align_corners=True) as used
// at guided_diffusion/unet_new.py:706,718-719, the offset/mask post-processing of
// SecondOrderDeformableAlignment.forward (unet_new.py:874-887: 10*tanh residual + flipped flow,
// sigmoid mask) and the sampling half of torchvision.ops.deform_conv2d (unet_new.py:889-898;
// the contraction half runs on tcgen05 through flair_conv_igemm over the im2col map), and the
// in-place `feat_prop *= weight` (unet_new.py:739).
//
// Flows are fp32 planes [N][2][H][W] (x displacement, y displacement) as SPyNet produces them.
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

__device__ __forceinline__ void ld8(const uint16_t* p, int dtype, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f;
    if (dtype == FLAIR_F16) {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f = __half22float2(h);
    } else {
      f = unpack_bf16x2(w[i]);
    }
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(uint16_t* p, int dtype, const float (&v)[8]) {
  uint4 u;
  if (dtype == FLAIR_F16) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  } else {
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  }
  *reinterpret_cast<uint4*>(p) = u;
}

// out[n][h][w][coff + c] = bilinear(x[n], (w + fx, h + fy)), zeros outside (grid_sample semantics)
__global__ void __launch_bounds__(256)
flow_warp_kernel(const uint16_t* __restrict__ x, const float* __restrict__ flow, uint16_t* __restrict__ out, int N,
                 int H, int W, int C, int x_cstride, int out_cstride, int dtype) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  // 32-bit index arithmetic (host-checked: N*H*W*C/8 < 2^31); see flow_warp2_kernel
  const unsigned vecs = static_cast<unsigned>(C) / 8u;
  const unsigned hw = static_cast<unsigned>(H) * W;
  const unsigned items = static_cast<unsigned>(N) * hw * vecs;
  for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const unsigned pix = it / vecs;
    const int cv = static_cast<int>(it - pix * vecs);
    const unsigned n = pix / hw;
    const unsigned off = pix - n * hw;
    const unsigned hu = off / static_cast<unsigned>(W);
    const int h = static_cast<int>(hu), w = static_cast<int>(off - hu * static_cast<unsigned>(W));
    const float sx = w + __ldg(flow + static_cast<size_t>(n * 2 + 0) * hw + off);
    const float sy = h + __ldg(flow + static_cast<size_t>(n * 2 + 1) * hw + off);
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    // clamp before the int conversion: a wild flow (|v| > 2^31) must not overflow
    const int x0 = static_cast<int>(fmaxf(fminf(fx0, static_cast<float>(W)), -2.f));
    const int y0 = static_cast<int>(fmaxf(fminf(fy0, static_cast<float>(H)), -2.f));
    const float ax = sx - fx0, ay = sy - fy0;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = x0 + dx, yy = y0 + dy;
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
        const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay);
        float v[8];
        ld8(x + static_cast<size_t>(n * hw + static_cast<unsigned>(yy * W + xx)) * x_cstride + cv * 8, dtype, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
      }
    st8(out + static_cast<size_t>(pix) * out_cstride + cv * 8, dtype, acc);
  }
}

// Two independent warps in one launch (blockIdx.y selects the pair): the first- and second-order warps of one
// propagation step (unet_new.py:706,719) always come together, and on the serial BasicVSR++ chain a launch costs
// more than the 8 MB it moves.
struct Warp2Args {
  const uint16_t* x[2];
  const float* flow[2];
  uint16_t* out[2];
  int x_cstride[2];
};
// one bilinear sample of an 8-channel vector: corner loads are always issued (clamped address, weight 0 outside the
// map — identical bits to skipping the corner), so two items per thread can have their 8 loads in flight together
struct WarpItem {
  uint4 v[4];
  float w[4];
};
__device__ __forceinline__ void warp_item_load(WarpItem& it, const uint16_t* __restrict__ x, const float* __restrict__ flow,
                                               unsigned n, unsigned off, unsigned hw, int H, int W, int x_cstride, int cv) {
  const unsigned hu = off / static_cast<unsigned>(W);
  const int h = static_cast<int>(hu), w = static_cast<int>(off - hu * static_cast<unsigned>(W));
  const float sx = w + __ldg(flow + static_cast<size_t>(n * 2 + 0) * hw + off);
  const float sy = h + __ldg(flow + static_cast<size_t>(n * 2 + 1) * hw + off);
  const float fx0 = floorf(sx), fy0 = floorf(sy);
  const int x0 = static_cast<int>(fmaxf(fminf(fx0, static_cast<float>(W)), -2.f));
  const int y0 = static_cast<int>(fmaxf(fminf(fy0, static_cast<float>(H)), -2.f));
  const float ax = sx - fx0, ay = sy - fy0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int dy = c >> 1, dx = c & 1;
    const int xx = x0 + dx, yy = y0 + dy;
    const bool in = xx >= 0 && xx < W && yy >= 0 && yy < H;
    it.w[c] = in ? (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) : 0.f;
    const int xc = min(max(xx, 0), W - 1), yc = min(max(yy, 0), H - 1);
    it.v[c] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(n * hw + static_cast<unsigned>(yc * W + xc)) * x_cstride + cv * 8));
  }
}
__device__ __forceinline__ void warp_item_store(const WarpItem& it, uint16_t* __restrict__ out, unsigned pix, int out_cstride,
                                                int cv, int dtype) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t wd[4] = {it.v[c].x, it.v[c].y, it.v[c].z, it.v[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f;
      if (dtype == FLAIR_F16) {
        __half2 hh = *reinterpret_cast<const __half2*>(&wd[i]);
        f = __half22float2(hh);
      } else {
        f = unpack_bf16x2(wd[i]);
      }
      acc[2 * i] = fmaf(it.w[c], f.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(it.w[c], f.y, acc[2 * i + 1]);
    }
  }
  st8(out + static_cast<size_t>(pix) * out_cstride + cv * 8, dtype, acc);
}

__global__ void __launch_bounds__(256)
flow_warp2_kernel(const __grid_constant__ Warp2Args a, int N, int H, int W, int C, int out_cstride, int dtype) {
  pdl_sync();
  const int s = blockIdx.y;
  const uint16_t* __restrict__ x = a.x[s];
  const float* __restrict__ flow = a.flow[s];
  uint16_t* __restrict__ out = a.out[s];
  const int x_cstride = a.x_cstride[s];
  // 32-bit index arithmetic (host-checked: N*H*W*C/8 < 2^31): the 64-bit divisions of the first version were
  // ~500 emulated instructions per item and made this 34 MB gather take 19 us instead of ~6
  const unsigned vecs = static_cast<unsigned>(C) / 8u;   // power of two (host-checked)
  const unsigned lv = 31u - __clz(vecs);
  const unsigned hw = static_cast<unsigned>(H) * W;
  const unsigned items = static_cast<unsigned>(N) * hw * vecs;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < items; i0 += 2 * stride) {
    const unsigned i1 = i0 + stride;
    const bool two = i1 < items;
    WarpItem A, B;
    const int cv0 = static_cast<int>(i0 & (vecs - 1));
    const unsigned pix0 = i0 >> lv, n0 = pix0 / hw;
    warp_item_load(A, x, flow, n0, pix0 - n0 * hw, hw, H, W, x_cstride, cv0);
    const unsigned i1c = two ? i1 : i0;
    const int cv1 = static_cast<int>(i1c & (vecs - 1));
    const unsigned pix1 = i1c >> lv, n1 = pix1 / hw;
    if (two) warp_item_load(B, x, flow, n1, pix1 - n1 * hw, hw, H, W, x_cstride, cv1);
    warp_item_store(A, out, pix0, out_cstride, cv0, dtype);
    if (two) warp_item_store(B, out, pix1, out_cstride, cv1, dtype);
  }
}

// out = f1 + warp(f2, f1)   (fp32 planes [N][2][H][W]) — unet_new.py:718
__global__ void __launch_bounds__(256)
flow_compose_kernel(const float* __restrict__ f2, const float* __restrict__ f1, float* __restrict__ out, int N, int H,
                    int W) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long hw = static_cast<long long>(H) * W;
  const long long items = static_cast<long long>(N) * hw;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = it / hw, off = it - n * hw;
    const int h = static_cast<int>(off / W), w = static_cast<int>(off % W);
    const float u = __ldg(f1 + (n * 2) * hw + off), v = __ldg(f1 + (n * 2 + 1) * hw + off);
    const float sx = w + u, sy = h + v;
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
    const float ax = sx - fx0, ay = sy - fy0;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = x0 + dx, yy = y0 + dy;
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
        const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay);
        const long long o = static_cast<long long>(yy) * W + xx;
        a0 = fmaf(wgt, __ldg(f2 + (n * 2) * hw + o), a0);
        a1 = fmaf(wgt, __ldg(f2 + (n * 2 + 1) * hw + o), a1);
      }
    out[(n * 2) * hw + off] = u + a0;
    out[(n * 2 + 1) * hw + off] = v + a1;
  }
}

// dst[pix][coff + c] = (16-bit) src[n][c][h][w]    (few channels: flows into the offset-net input)
__global__ void __launch_bounds__(256)
planes_to_cl_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int N, int Cs, long long hw,
                    int dst_cstride, int dst_coff, int dtype) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long items = static_cast<long long>(N) * hw * Cs;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(it % Cs);
    const long long pix = it / Cs;
    const long long n = pix / hw, off = pix - n * hw;
    const float v = __ldg(src + (n * Cs + c) * hw + off);
    uint16_t bits;
    if (dtype == FLAIR_F16) {
      __half hv = __float2half_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&hv);
    } else {
      __nv_bfloat16 bv = __float2bfloat16_rn(v);
      bits = *reinterpret_cast<uint16_t*>(&bv);
    }
    dst[pix * dst_cstride + dst_coff + c] = bits;
  }
}

// Modulated deformable im2col, 3x3, pad 1, stride 1, dg deform groups over the 2C input channels
// (first C from xa, last C from xb).  om: [N][H][W][om_cstride] 16-bit raw offset-net output with
// channel blocks o1 (dg/2*18) | o2 (dg/2*18) | mask (dg*9)  (th.chunk(out, 3), unet_new.py:877).
// cols: [N*H*W][9 * 2C], column = tap*2C + ci.
//
// One CTA = kDefPix pixels.  Phase 1 stages the raw offset rows through shared memory with 16-byte
// loads and turns them into (dy, dx, mask) once per (pixel, group, tap); phase 2 gathers: consecutive
// threads write consecutive 16-byte column vectors, the four bilinear corners are 16-byte channel
// vectors of the channels-last source.
constexpr int kDefPix = 8;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256)
deform_im2col_kernel(const uint16_t* __restrict__ xa, const uint16_t* __restrict__ xb, int xa_cstride, int xb_cstride,
                     const uint16_t* __restrict__ om, int om_cstride, int om_dtype, const float* __restrict__ flow1,
                     const float* __restrict__ flow2, uint16_t* __restrict__ cols, int N, int H, int W, int C, int dg,
                     float mrm, int dtype) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ float dsm[];  // [kDefPix][3][dg*9]: dy | dx | mask
  const int pairs = dg * 9;
  const int nch = pairs * 3;  // channels of the offset-net output actually used
  const long long hw = static_cast<long long>(H) * W;
  const long long total_pix = static_cast<long long>(N) * hw;
  const long long pix0 = static_cast<long long>(blockIdx.x) * kDefPix;
  const int half_g = dg / 2;
  // ---- phase 1a: raw values -> smem (as fp32), 8 channels per 16-byte load
  const int vec_per_pix = nch / 8;
  for (int i = threadIdx.x; i < kDefPix * vec_per_pix; i += blockDim.x) {
    const int p = i / vec_per_pix, v = i % vec_per_pix;
    if (pix0 + p >= total_pix) continue;
    float f[8];
    ld8(om + (pix0 + p) * om_cstride + v * 8, om_dtype, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) dsm[p * nch + v * 8 + j] = f[j];
  }
  __syncthreads();
  // ---- phase 1b: (dy, dx, mask) per (pixel, group, tap), in place: reads raw[oc], raw[oc+1], raw[2*pairs+pair]
  float dyv[5], dxv[5], mkv[5];  // kDefPix*pairs / 256 <= 5 (checked on the host)
#pragma unroll
  for (int cnt = 0; cnt < 5; ++cnt) {
    const int i = threadIdx.x + cnt * 256;
    const int p = i / pairs, pair = i % pairs;  // pair = g*9 + tap
    const long long pix = pix0 + p;
    float dy = 0.f, dx = 0.f, mk = 0.f;
    if (i < kDefPix * pairs && pix < total_pix) {
      const int g = pair / 9;
      const long long n = pix / hw, off = pix - n * hw;
      const float* fl = (g >= half_g) ? flow2 : flow1;
      const float* raw = dsm + p * nch;
      // offset channel (g*9 + tap)*2 + {0: dy, 1: dx} inside cat(o1, o2); flow.flip(1) = (fy, fx)
      dy = mrm * tanh_fast(raw[pair * 2]) + __ldg(fl + (n * 2 + 1) * hw + off);
      dx = mrm * tanh_fast(raw[pair * 2 + 1]) + __ldg(fl + (n * 2 + 0) * hw + off);
      mk = 1.0f / (1.0f + __expf(-raw[2 * pairs + pair]));
    }
    dyv[cnt] = dy; dxv[cnt] = dx; mkv[cnt] = mk;
  }
  __syncthreads();
#pragma unroll
  for (int cnt = 0; cnt < 5; ++cnt) {
    const int i = threadIdx.x + cnt * 256;
    if (i < kDefPix * pairs) {
      const int p = i / pairs, pair = i % pairs;
      dsm[p * nch + pair] = dyv[cnt];
      dsm[p * nch + pairs + pair] = dxv[cnt];
      dsm[p * nch + 2 * pairs + pair] = mkv[cnt];
    }
  }
  __syncthreads();
  // ---- phase 2: gather.  item = ((p*9 + tap)*dg + g)*vpg + vi  -> consecutive 16-byte outputs
  const int cpg = 2 * C / dg, vpg = cpg / 8;
  const int items = kDefPix * 9 * dg * vpg;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    int r = it;
    const int vi = r % vpg; r /= vpg;
    const int g = r % dg; r /= dg;
    const int tap = r % 9;
    const int p = r / 9;
    const long long pix = pix0 + p;
    if (pix >= total_pix) continue;
    const long long n = pix / hw, off = pix - n * hw;
    const int h = static_cast<int>(off / W), w = static_cast<int>(off % W);
    const int pair = g * 9 + tap;
    const float sy = h + tap / 3 - 1 + dsm[p * nch + pair];
    const float sx = w + tap % 3 - 1 + dsm[p * nch + pairs + pair];
    const float mk = dsm[p * nch + 2 * pairs + pair];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sy > -1.f && sy < H && sx > -1.f && sx < W) {  // torchvision bilinear_interpolate bounds
      const float fy0 = floorf(sy), fx0 = floorf(sx);
      const int y0 = static_cast<int>(fy0), x0 = static_cast<int>(fx0);
      const float ay = sy - fy0, ax = sx - fx0;
      const bool second = g >= half_g;
      const int cg = g * cpg + vi * 8;  // channel inside the 2C concat
      const uint16_t* src = second ? xb : xa;
      const int cs = second ? xb_cstride : xa_cstride;
      const int cl = second ? cg - C : cg;
      const uint16_t* base = src + n * hw * cs + cl;
#pragma unroll
      for (int yy = 0; yy < 2; ++yy)
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
          const int py = y0 + yy, px = x0 + xx;
          if (py < 0 || py > H - 1 || px < 0 || px > W - 1) continue;
          const float wgt = (yy ? ay : 1.f - ay) * (xx ? ax : 1.f - ax) * mk;
          float v[8];
          ld8(base + (static_cast<long long>(py) * W + px) * cs, dtype, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
        }
    }
    st8(cols + pix * (18LL * C) + static_cast<long long>(tap) * 2 * C + g * cpg + vi * 8, dtype, acc);
  }
}

// x[pix][c] *= wmap[pix]   (wmap fp32 [N][H][W])
__global__ void __launch_bounds__(256)
scale_pixels_kernel(uint16_t* __restrict__ x, const float* __restrict__ wmap, long long pixels, int C, int cstride,
                    int dtype) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const int vecs = C / 8;
  const long long items = pixels * vecs;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = it / vecs;
    const int cv = static_cast<int>(it % vecs);
    const float s = __ldg(wmap + pix);
    float v[8];
    uint16_t* p = x + pix * cstride + cv * 8;
    ld8(p, dtype, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s;
    st8(p, dtype, v);
  }
}

int blocks_for(long long items) {
  long long b = ceil_div_ll(items, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

extern "C" int flair_flow_warp(const void* x, const float* flow, void* out, int N, int H, int W, int C,
                               int x_cstride, int out_cstride, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && flow && out && C % 8 == 0 && x_cstride % 8 == 0 && out_cstride % 8 == 0, "flair_flow_warp: bad arguments");
  FLAIR_REQUIRE(static_cast<long long>(N) * H * W * (C / 8) < (1ll << 31), "flair_flow_warp: map too large for 32-bit indexing");
  FLAIR_CHECK_CUDA(flair_launch(flow_warp_kernel, dim3(blocks_for(static_cast<long long>(N) * H * W * (C / 8))), dim3(256), 0, stream, 
      static_cast<const uint16_t*>(x), flow, static_cast<uint16_t*>(out), N, H, W, C, x_cstride, out_cstride, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_flow_warp2(const void* xa, const void* xb, const float* flow_a, const float* flow_b, void* out_a,
                                void* out_b, int N, int H, int W, int C, int xa_cstride, int xb_cstride, int out_cstride,
                                int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(xa && xb && flow_a && flow_b && out_a && out_b && C % 8 == 0 && xa_cstride % 8 == 0 && xb_cstride % 8 == 0 &&
                    out_cstride % 8 == 0,
                "flair_flow_warp2: bad arguments");
  FLAIR_REQUIRE(((C / 8) & (C / 8 - 1)) == 0 && static_cast<long long>(N) * H * W * (C / 8) < (1ll << 31),
                "flair_flow_warp2: C/8 must be a power of two and the map indexable with 32 bits (C=%d)", C);
  Warp2Args a;
  a.x[0] = static_cast<const uint16_t*>(xa); a.x[1] = static_cast<const uint16_t*>(xb);
  a.flow[0] = flow_a; a.flow[1] = flow_b;
  a.out[0] = static_cast<uint16_t*>(out_a); a.out[1] = static_cast<uint16_t*>(out_b);
  a.x_cstride[0] = xa_cstride; a.x_cstride[1] = xb_cstride;
  const long long items = static_cast<long long>(N) * H * W * (C / 8);
  int bx = blocks_for(items) / 2;
  if (bx < 1) bx = 1;
  FLAIR_CHECK_CUDA(flair_launch(flow_warp2_kernel, dim3(bx, 2), dim3(256), 0, stream, a, N, H, W, C, out_cstride, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_flow_compose_f32(const float* f2, const float* f1, float* out, int N, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(f2 && f1 && out, "flair_flow_compose_f32: null pointer");
  FLAIR_CHECK_CUDA(flair_launch(flow_compose_kernel, dim3(blocks_for(static_cast<long long>(N) * H * W)), dim3(256), 0, stream, f2, f1, out, N, H, W));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_planes_to_cl(const float* src, void* dst, int N, int Cs, int H, int W, int dst_cstride,
                                  int dst_coffset, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(src && dst, "flair_planes_to_cl: null pointer");
  const long long hw = static_cast<long long>(H) * W;
  FLAIR_CHECK_CUDA(flair_launch(planes_to_cl_kernel, dim3(blocks_for(N * hw * Cs)), dim3(256), 0, stream, src, static_cast<uint16_t*>(dst), N, Cs, hw,
                                                                  dst_cstride, dst_coffset, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_deform_im2col(const void* xa, const void* xb, int xa_cstride, int xb_cstride, const void* om,
                                   int om_cstride, int om_dtype, const float* flow1, const float* flow2, void* cols,
                                   int N, int H, int W, int C, int deform_groups, float max_residue_magnitude,
                                   int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(xa && xb && om && flow1 && flow2 && cols, "flair_deform_im2col: null pointer");
  FLAIR_REQUIRE(deform_groups > 0 && deform_groups % 2 == 0 && (2 * C) % deform_groups == 0 &&
                    ((2 * C) / deform_groups) % 8 == 0,
                "flair_deform_im2col: channels per deform group must be a multiple of 8 (C=%d, dg=%d)", C, deform_groups);
  FLAIR_REQUIRE((deform_groups * 27) % 8 == 0 && deform_groups * 9 * kDefPix <= 5 * 256 && om_cstride % 8 == 0,
                "flair_deform_im2col: unsupported deform_groups=%d", deform_groups);
  const long long total_pix = static_cast<long long>(N) * H * W;
  const size_t smem = sizeof(float) * kDefPix * deform_groups * 27;
  FLAIR_CHECK_CUDA(flair_launch(deform_im2col_kernel, dim3(static_cast<unsigned>(ceil_div_ll(total_pix, kDefPix))), dim3(256), smem, stream, 
      static_cast<const uint16_t*>(xa), static_cast<const uint16_t*>(xb), xa_cstride, xb_cstride,
      static_cast<const uint16_t*>(om), om_cstride, om_dtype, flow1, flow2, static_cast<uint16_t*>(cols), N, H, W, C,
      deform_groups, max_residue_magnitude, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_scale_pixels(void* x, const float* wmap, long long pixels, int C, int cstride, int dtype,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && wmap && C % 8 == 0, "flair_scale_pixels: bad arguments");
  FLAIR_CHECK_CUDA(flair_launch(scale_pixels_kernel, dim3(blocks_for(pixels * (C / 8))), dim3(256), 0, stream, static_cast<uint16_t*>(x), wmap, pixels, C,
                                                                       cstride, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
