// Spatial self-attention (QKVAttentionLegacy) and windowed temporal attention, channels-last.
//
// Spatial: replaces guided_diffusion/unet_new.py:540-570 — two einsum/bmm launches with a
// materialised (heads, L, L) fp32 weight tensor and a separate softmax — by one flash-style
// kernel (online softmax, fp32 accumulation, nothing L x L ever leaves the SM).  The qkv
// channel layout is the reference's head-major (H, 3, d) (:559); scale d^-1/4 on q and k == d^-1/2
// on the logits.  At the sizes of this model (L <= 1024 tokens, d = 64, <= 0.2 % of the UNet's
// FLOPs) the kernel is latency/occupancy-bound, so it runs on the FP32 pipes with warp shuffles.
//
// Temporal: replaces unet_new.py:473-517 — replicate-pad + unfold (5x data duplication), three
// Linears on the unfolded tensor and flash_attn_func on degenerate 1 x (F-1) problems
// (nn.py:370-394) — by a gather kernel over per-frame q/k/v projections: the positional terms
// are linear, so W(x + pe) = W x + W pe is folded into per-offset constant vectors at load time.
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kD = 64;  // head dim (num_head_channels = 64, scripts/video_sample.py:130)

__device__ __forceinline__ float2 cvt16(uint32_t u, int dtype) {
  if (dtype == FLAIR_F16) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
  }
  return unpack_bf16x2(u);
}
__device__ __forceinline__ uint32_t pk16(float a, float b, int dtype) {
  if (dtype == FLAIR_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(a, b);
}

// grid (ceil(L/16), heads, N), 128 threads; warp w handles queries 4w..4w+3 of the 16-query tile.
__global__ void __launch_bounds__(128)
attn_spatial_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, const float* __restrict__ rowbias,
                    int rowbias_stride, int L, int heads, int qkv_cstride, int out_cstride, int dtype, float scale) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ float Qs[16][kD];
  __shared__ float Ks[kD][33];
  __shared__ float Vs[32][kD];
  const int n = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint16_t* base = qkv + static_cast<long long>(n) * L * qkv_cstride + hd * 3 * kD;
  // stage the 16 queries (fp32)
  for (int i = threadIdx.x; i < 16 * (kD / 2); i += blockDim.x) {
    const int qi = i / (kD / 2), dp = i % (kD / 2);
    float2 f = make_float2(0.f, 0.f);
    if (q0 + qi < L)
      f = cvt16(__ldg(reinterpret_cast<const uint32_t*>(base + static_cast<long long>(q0 + qi) * qkv_cstride) + dp), dtype);
    Qs[qi][2 * dp] = f.x; Qs[qi][2 * dp + 1] = f.y;
  }
  float m[4], l[4], acc0[4], acc1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; acc0[i] = 0.f; acc1[i] = 0.f; }
  for (int k0 = 0; k0 < L; k0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * (kD / 2); i += blockDim.x) {
      const int ki = i / (kD / 2), dp = i % (kD / 2);
      float2 fk = make_float2(0.f, 0.f), fv = make_float2(0.f, 0.f);
      if (k0 + ki < L) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(base + static_cast<long long>(k0 + ki) * qkv_cstride);
        fk = cvt16(__ldg(row + kD / 2 + dp), dtype);
        fv = cvt16(__ldg(row + kD + dp), dtype);
      }
      Ks[2 * dp][ki] = fk.x; Ks[2 * dp + 1][ki] = fk.y;
      Vs[ki][2 * dp] = fv.x; Vs[ki][2 * dp + 1] = fv.y;
    }
    __syncthreads();
    const bool kvalid = (k0 + lane) < L;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = warp * 4 + i;
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < kD; ++d) s = fmaf(Qs[qi][d], Ks[d][lane], s);
      s = kvalid ? s * scale : -INFINITY;
      float mx = s;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m[i], mx);
      const float p = kvalid ? __expf(s - m_new) : 0.f;
      float ps = p;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      const float corr = __expf(m[i] - m_new);
      l[i] = l[i] * corr + ps;
      float a0 = acc0[i] * corr, a1 = acc1[i] * corr;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float pk = __shfl_sync(0xffffffffu, p, k);
        a0 = fmaf(pk, Vs[k][lane], a0);
        a1 = fmaf(pk, Vs[k][lane + 32], a1);
      }
      acc0[i] = a0; acc1[i] = a1; m[i] = m_new;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + warp * 4 + i;
    if (q >= L) continue;
    float o0 = acc0[i] / l[i], o1 = acc1[i] / l[i];
    const int c0 = hd * kD + lane, c1 = c0 + 32;
    if (rowbias != nullptr) {
      o0 += __ldg(rowbias + static_cast<long long>(n) * rowbias_stride + c0);
      o1 += __ldg(rowbias + static_cast<long long>(n) * rowbias_stride + c1);
    }
    uint16_t* orow = out + (static_cast<long long>(n) * L + q) * out_cstride;
    orow[c0] = static_cast<uint16_t>(pk16(o0, 0.f, dtype) & 0xFFFFu);
    orow[c1] = static_cast<uint16_t>(pk16(o1, 0.f, dtype) & 0xFFFFu);
  }
}

// qkv: [B][T][P][3C] with blocks q | k | v (each C = heads*64, channel = head*64 + i).
// 8 lanes per (pixel, head): each lane owns 8 channels.
__global__ void __launch_bounds__(256)
attn_temporal_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, const float* __restrict__ cq,
                     const float* __restrict__ ck, const float* __restrict__ bv, int B, int T, long long P, int C,
                     int frames, int dtype, float scale) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const int heads = C / kD;
  const int half = frames / 2;
  const long long items = static_cast<long long>(B) * T * P * heads * 8;
  const long long items32 = (items + 31) & ~31LL;  // warp-uniform trip count (shuffles below)
  for (long long it0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it0 < items32;
       it0 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool live = it0 < items;
    const long long it = live ? it0 : items - 8 + (it0 & 7);  // idle lanes redo the last octet, store nothing
    const int sub = static_cast<int>(it & 7);
    long long r = it >> 3;
    const int hd = static_cast<int>(r % heads); r /= heads;
    const long long p = r % P; r /= P;
    const int t = static_cast<int>(r % T);
    const int b = static_cast<int>(r / T);
    const int c0 = hd * kD + sub * 8;
    auto load_vec = [&](int tt, int block, float (&v)[8]) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(
          qkv + ((static_cast<long long>(b) * T + tt) * P + p) * (3LL * C) + block * C + c0));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = cvt16(w[i], dtype);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      }
    };
    float q[8];
    load_vec(t, 0, q);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] += __ldg(cq + c0 + i);
    float s[6], mx = -INFINITY;
    for (int j = 0; j < frames - 1; ++j) {
      const int off = (j < half) ? j - half : j - half + 1;
      int tt = t + off;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      float k[8];
      load_vec(tt, 1, k);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(q[i], k[i] + __ldg(ck + j * C + c0 + i), d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      s[j] = d * scale;
      mx = fmaxf(mx, s[j]);
    }
    float den = 0.f;
    for (int j = 0; j < frames - 1; ++j) { s[j] = __expf(s[j] - mx); den += s[j]; }
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < frames - 1; ++j) {
      const int off = (j < half) ? j - half : j - half + 1;
      int tt = t + off;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      float v[8];
      load_vec(tt, 2, v);
      const float pj = s[j] / den;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(pj, v[i], o[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += __ldg(bv + c0 + i);  // sum_j p_j = 1
    uint4 u;
    u.x = pk16(o[0], o[1], dtype); u.y = pk16(o[2], o[3], dtype);
    u.z = pk16(o[4], o[5], dtype); u.w = pk16(o[6], o[7], dtype);
    if (live) *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * T + t) * P + p) * C + c0) = u;
  }
}

}  // namespace

extern "C" int flair_attn_spatial(const void* qkv, void* out, const float* rowbias, int rowbias_stride, int N,
                                  int L, int heads, int qkv_cstride, int out_cstride, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(qkv && out, "flair_attn_spatial: null pointer");
  FLAIR_REQUIRE(N > 0 && N < 65536 && L > 0 && heads > 0 && qkv_cstride >= heads * 3 * kD && out_cstride >= heads * kD,
                "flair_attn_spatial: bad sizes");
  FLAIR_REQUIRE(dtype == FLAIR_BF16 || dtype == FLAIR_F16, "flair_attn_spatial: 16-bit maps only");
  dim3 grid(ceil_div(L, 16), heads, N);
  FLAIR_CHECK_CUDA(flair_launch(attn_spatial_kernel, dim3(grid), dim3(128), 0, stream, static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out),
                                                rowbias, rowbias_stride, L, heads, qkv_cstride, out_cstride, dtype,
                                                0.125f /* 64^-1/2 */));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_attn_temporal(const void* qkv, void* out, const float* cq, const float* ck, const float* bv,
                                   int B, int T, long long P, int C, int frames, int dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(qkv && out && cq && ck && bv, "flair_attn_temporal: null pointer");
  FLAIR_REQUIRE(C % kD == 0 && (frames == 5 || frames == 7) && T > 0 && B > 0 && P > 0,
                "flair_attn_temporal: unsupported C=%d frames=%d", C, frames);
  FLAIR_REQUIRE(dtype == FLAIR_BF16 || dtype == FLAIR_F16, "flair_attn_temporal: 16-bit maps only");
  const long long items = static_cast<long long>(B) * T * P * (C / kD) * 8;
  long long blocks = ceil_div_ll(items, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  FLAIR_CHECK_CUDA(flair_launch(attn_temporal_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, 
      static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), cq, ck, bv, B, T, P, C, frames, dtype, 0.125f));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
