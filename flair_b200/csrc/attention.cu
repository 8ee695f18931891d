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
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------
// Tensor-core spatial attention (L % 128 == 0): one CTA = 128 queries of one (frame, head), 128 threads.
//   S = Q K^T   tcgen05.mma M=128, N=128 keys, K=64: Q and K tiles are TMA boxes {64 ch, 128 tokens} of the
//               channels-last qkv map (head-major (H,3,64) channel layout) landing as K-major SWIZZLE_128B operands
//   softmax     thread t owns query row t = TMEM lane t: tcgen05.ld the 128 logits (two passes: max, then exp2),
//               fp32 online softmax across key blocks; P is written as 16-bit K-major SWIZZLE_128B A operand
//   O_blk = P V tcgen05.mma M=128, N=64, K=128 keys.  V is used as it sits in memory ([key][d], d contiguous),
//               i.e. as an MN-major B operand (instruction-descriptor bit 16); FLAIR_ATTN_VMODE=2 instead transposes V
//               through registers into a K-major tile (debug / cross-check path).
//   O accumulates in fp32 registers (64 per thread) with the usual exp2(m_old - m_new) correction, so TMEM holds only
//   the current block's S (128 columns) and O_blk (64 columns).  K/V blocks are double-buffered in shared memory.
// Replaces guided_diffusion/unet_new.py:556-566 (two einsums + materialised (N*heads, L, L) fp32 softmax).
// ---------------------------------------------------------------------------------------------------------
constexpr int kTcQ = 128;      // queries per CTA
constexpr int kTcKV = 128;     // keys per block
constexpr uint32_t kTcTile = kTcQ * kD * 2;  // 16 KB: one {64 ch x 128 tokens} box

__device__ __forceinline__ void attn_sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__global__ void __launch_bounds__(128, 1)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, uint16_t* __restrict__ out,
                       const float* __restrict__ rowbias, int rowbias_stride, int L, int out_cstride, int dtype,
                       float scale_log2, int vmode, const uint16_t* __restrict__ qkv, int qkv_cstride) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                         // 16 KB
  uint8_t* sK = smem + kTcTile;               // 2 x 16 KB
  uint8_t* sV = smem + 3 * kTcTile;           // 2 x 16 KB   ([key][d] as TMA wrote it)
  uint8_t* sP = smem + 5 * kTcTile;           // 32 KB: two K-major k-tiles (64 keys each) of 128 rows x 128 B
  uint8_t* sVt = smem + 7 * kTcTile;          // 16 KB (vmode 2 only): V^T as two K-major k-tiles of 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * kTcTile);
  uint64_t* q_bar = bars;          // Q landed
  uint64_t* kv_bar = bars + 1;     // [2] K/V block landed
  uint64_t* s_bar = bars + 3;      // S = Q K^T complete
  uint64_t* o_bar = bars + 4;      // O_blk = P V complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * kTcQ;
  const int nblk = L / kTcKV;
  const int row0 = n * L;                      // first token row of this frame in the [N*L][C3] map
  const int cq = hd * 3 * kD;                  // q | k | v channel origins of this head
  if (tid == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_bar, 1);
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);   // S: columns 0..127, O_blk: 128..191
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // qkv is the previous kernel's output

  auto load_kv = [&](int blk) {
    const int st = blk & 1;
    mbar_expect_tx(&kv_bar[st], 2 * kTcTile);
    tma_load_2d(sK + st * kTcTile, &tmQKV, &kv_bar[st], cq + kD, row0 + blk * kTcKV);
    tma_load_2d(sV + st * kTcTile, &tmQKV, &kv_bar[st], cq + 2 * kD, row0 + blk * kTcKV);
  };
  if (tid == 0) {
    mbar_expect_tx(q_bar, kTcTile);
    tma_load_2d(sQ, &tmQKV, q_bar, cq, row0 + q0);
    load_kv(0);
    if (nblk > 1) load_kv(1);
  }

  const uint32_t fmt = (dtype == FLAIR_BF16) ? 1u : 0u;
  const uint32_t idesc_s = umma_idesc_f16(kTcQ, kTcKV, fmt);
  // PV: N = 64; B (= V) MN-major unless the transposed copy is used
  const uint32_t idesc_o = umma_idesc_f16(kTcQ, kD, fmt) | ((vmode == 2) ? 0u : (1u << 16));
  const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;   // SBO 1024 B, version 1, SWIZZLE_128B
  const uint32_t desc_lo_flags = static_cast<uint32_t>(umma_desc_sw128(0));
  // MN-major V: 8-key groups are 1024 B apart (SBO); one 64-channel atom along N (LBO unused; vmode 1 swaps the roles)
  const uint64_t desc_hi_v = (vmode == 1) ? ((desc_hi & ~(static_cast<uint64_t>(0x3FFF) << 32)) | (static_cast<uint64_t>(1) << 32)) : desc_hi;
  const uint32_t desc_lo_v = (vmode == 1) ? ((desc_lo_flags & ~(0x3FFFu << 16)) | ((1024u >> 4) << 16)) : desc_lo_flags;
  auto lo = [](const void* p) { return (smem_u32(p) & 0x3FFFF) >> 4; };

  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);   // this warp's 32 TMEM lanes
  const int r = tid;                       // query row inside the tile
  const uint32_t p_row = smem_u32(sP) + (r >> 3) * 1024 + (r & 7) * 128;
  const int sw = r & 7;
  float m_run = -INFINITY, l_run = 0.f;
  float o[kD];
#pragma unroll
  for (int i = 0; i < kD; ++i) o[i] = 0.f;

  for (int blk = 0; blk < nblk; ++blk) {
    const int st = blk & 1;
    const uint32_t kv_phase = (blk >> 1) & 1, ph = blk & 1;
    if (vmode == 2) {
      // debug path: V^T through registers.  thread = key row, 64 channels -> 64 two-byte stores (K-major SW128 tile
      // of 64 rows (d) x 64 keys per k-tile)
      mbar_wait(&kv_bar[st], kv_phase);
      const uint16_t* vrow = qkv + static_cast<long long>(row0 + blk * kTcKV + tid) * qkv_cstride + cq + 2 * kD;
      const int kt = tid >> 6, kk = tid & 63;
      for (int d = 0; d < kD; ++d) {
        const uint32_t a = smem_u32(sVt) + kt * 8192 + (d >> 3) * 1024 + (d & 7) * 128 + ((((kk >> 3) ^ (d & 7))) << 4) + (kk & 7) * 2;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(vrow[d]) : "memory");
      }
    }
    if (warp == 0) {
      if (blk == 0) mbar_wait(q_bar, 0);
      mbar_wait(&kv_bar[st], kv_phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t alo = desc_lo_flags | lo(sQ), blo = desc_lo_flags | lo(sK + st * kTcTile);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_f16(tmem_base, desc_hi | (alo + 2u * k), desc_hi | (blo + 2u * k), idesc_s, k ? 1u : 0u);
        umma_commit(s_bar);
      }
      __syncwarp();
    }
    mbar_wait(s_bar, ph);
    tc_fence_after();
    // ---- pass 1: row maximum of the 128 logits
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kTcKV; c += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16(t_row + c, r0);
      tmem_ld16(t_row + c + 16, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
    }
    const float m_new = fmaxf(m_run, mx * scale_log2);   // running maximum in the exp2 domain
    const float corr = exp2f(m_run - m_new);             // 0 on the first block (m_run = -inf)
    // ---- pass 2: p = 2^(s*scale - m), row sum, P -> shared (16-bit, swizzled K-major)
    float psum = 0.f;
#pragma unroll 1
    for (int c = 0; c < kTcKV; c += 16) {
      uint32_t r0[16];
      tmem_ld16(t_row + c, r0);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = exp2f(fmaf(__uint_as_float(r0[2 * j]), scale_log2, -m_new));
        const float p1 = exp2f(fmaf(__uint_as_float(r0[2 * j + 1]), scale_log2, -m_new));
        pk[j] = pk16(p0, p1, dtype);
        // the row sum uses the ROUNDED probabilities: numerator (tensor core) and denominator see the same values
        const float2 pr = cvt16(pk[j], dtype);
        psum += pr.x + pr.y;
      }
      const uint32_t base = p_row + (c >> 6) * kTcTile;
      const int ch = (c & 63) >> 3;
      attn_sts128(base + (((ch) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
      attn_sts128(base + (((ch + 1) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    l_run = l_run * corr + psum;
    m_run = m_new;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // P (and V^T): generic-proxy writes -> tensor core reads
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_o = tmem_base + 128;
#pragma unroll
        for (int k = 0; k < kTcKV / 16; ++k) {
          const uint32_t alo = desc_lo_flags | (lo(sP + (k >> 2) * kTcTile) + 2u * (k & 3));
          if (vmode == 2) {
            const uint32_t blo = desc_lo_flags | (lo(sVt + (k >> 2) * 8192) + 2u * (k & 3));
            umma_f16(d_o, desc_hi | alo, desc_hi | blo, idesc_o, k ? 1u : 0u);
          } else {
            const uint32_t blo = desc_lo_v | (lo(sV + st * kTcTile) + static_cast<uint32_t>(k) * (2048u >> 4));
            umma_f16(d_o, desc_hi | alo, desc_hi_v | blo, idesc_o, k ? 1u : 0u);
          }
        }
        umma_commit(o_bar);
      }
      __syncwarp();
    }
    mbar_wait(o_bar, ph);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < kD; c += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16(t_row + 128 + c, r0);
      tmem_ld16(t_row + 128 + c + 16, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        o[c + j] = fmaf(o[c + j], corr, __uint_as_float(r0[j]));
        o[c + 16 + j] = fmaf(o[c + 16 + j], corr, __uint_as_float(r1[j]));
      }
    }
    tc_fence_before();
    __syncthreads();   // S / O_blk / P / this K,V stage are free again
    if (tid == 0 && blk + 2 < nblk) load_kv(blk + 2);
  }

  const float inv = 1.0f / l_run;
  const int c0 = hd * kD;
  uint16_t* orow = out + (static_cast<long long>(row0) + q0 + r) * out_cstride + c0;
  const float* rb = rowbias ? rowbias + static_cast<long long>(n) * rowbias_stride + c0 : nullptr;
#pragma unroll
  for (int c = 0; c < kD; c += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = o[c + j] * inv + (rb ? __ldg(rb + c + j) : 0.f);
    uint4 u;
    u.x = pk16(v[0], v[1], dtype); u.y = pk16(v[2], v[3], dtype);
    u.z = pk16(v[4], v[5], dtype); u.w = pk16(v[6], v[7], dtype);
    *reinterpret_cast<uint4*>(orow + c) = u;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
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
  // FLAIR_ATTN_TC: 1 (default) tensor-core kernel where it applies, 0 = SIMT kernel everywhere (A/B measurements);
  // FLAIR_ATTN_VMODE: 0 = V as MN-major operand (default), 1 = LBO/SBO roles swapped (probe), 2 = transposed copy
  static int use_tc = -1, vmode = 0;
  if (use_tc < 0) {
    const char* e = getenv("FLAIR_ATTN_TC");
    const char* v = getenv("FLAIR_ATTN_VMODE");
    vmode = v ? atoi(v) : 0;
    use_tc = e ? atoi(e) : 1;
  }
  if (use_tc && L % kTcQ == 0 && qkv_cstride % 8 == 0 && out_cstride % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      static_cast<long long>(N) * L < (1ll << 31)) {
    flair_tmap_encode_fn encode = flair_get_tmap_encode();
    FLAIR_REQUIRE(encode != nullptr, "flair_attn_spatial: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tm;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(heads) * 3 * kD, static_cast<cuuint64_t>(N) * L};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(qkv_cstride) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kD), static_cast<cuuint32_t>(kTcQ)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&tm, dtype == FLAIR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLAIR_REQUIRE(r == CUDA_SUCCESS, "flair_attn_spatial: tensor map rejected (%d)", static_cast<int>(r));
    const size_t smem_bytes = 8 * kTcTile + 1024 + 256;
    static FlairPerDeviceOnce attr_once;
    if (attr_once.first())
      FLAIR_CHECK_CUDA(cudaFuncSetAttribute(attn_spatial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem_bytes)));
    const float scale_log2 = 0.125f * 1.4426950408889634f;   // 64^-1/2 on the logits, exp2 domain
    FLAIR_CHECK_CUDA(flair_launch(attn_spatial_tc_kernel, dim3(L / kTcQ, heads, N), dim3(128), smem_bytes, stream, tm,
                                  static_cast<uint16_t*>(out), rowbias, rowbias_stride, L, out_cstride, dtype, scale_log2,
                                  vmode, static_cast<const uint16_t*>(qkv), qkv_cstride));
    FLAIR_CHECK_LAUNCH();
    return 0;
  }
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
