// Fused modulated deformable convolution (3x3, pad 1, stride 1, 16 deform groups) for sm_100a:
// the bilinear gather IS the A-operand producer of a tcgen05 GEMM, the im2col matrix never exists.
//
// Replaces SecondOrderDeformableAlignment.forward after the offset net (reference
// guided_diffusion/unet_new.py:874-898; unet.py:469-492): offset = mrm * tanh(cat(o1, o2)) + flipped
// flows, mask = sigmoid(.), torchvision.ops.deform_conv2d(cat(feat_prop, feat_n2), offset, weight,
// bias, padding 1, mask).  The two-kernel version (flair_deform_im2col + 1x1 GEMM, still the generic path) wrote
// and re-read 18C x 2 B per pixel (151 MB per 256x256 frame at C = 64): 217 + 34 us per frame; this kernel: 107 us
// (89 us inside the model), 59-63 us vs 93 + 21 us at C = 128 on a 128x128 frame.
//
// Roles (22 warps, one CTA per SM, persistent over 16 x 8-pixel tiles): warp 0 = TMA (one weight slab per
// k-block; the 48 offset-net channels of the current tap for the whole tile, one 4-D box), warp 1 = tcgen05.mma
// issuer (elect.sync), warps 2..5 = epilogue (TMEM -> +bias -> 16-bit NHWC, output may be a channel slice),
// warps 6..21 = gather producers.  M tile = 128 pixels, N = C, K = 9 taps x 2C walked in 64-channel k-blocks
// (k = tap * 2C + channel of cat(xa, xb), the layout flair_deform_im2col used, so the packed weight is unchanged).
// A producer warp owns 32 tile rows (lane = pixel) x 4 deform groups: per tap it reads its (dy, dx, mask) triples
// from the staged offset rows, computes the sample positions, loads the two x-corner pairs (one row each) of every
// group with 32-byte loads, blends (fp16: half2 FMAs like the reference's fp16 torchvision kernel; bf16: fp32) and
// writes the 16-byte chunks straight into the 128B-swizzled K-major operand tile the MMA consumes
// (st.shared + fence.proxy.async + mbarrier arrive by every lane).
//
// Source layout: pair planes [8 groups][pixel][2][C/8] written by the conv epilogue (`out2` of flair_conv_params):
// entry p = (pixel p, pixel p+1) of the row-major map, so both x-corners of a bilinear sample are one aligned
// 32-byte (C = 64) or two 32-byte (C = 128) loads: 2 L1 tag look-ups per sample instead of 4.  The kernel is bound
// by those look-ups (ncu: l1tex 88 % busy); shared memory is kept to 128-156 KB so that ~100 KB of L1 remain for the
// ~36x re-read of every source pixel.
//
// Ring invariants (both were violated once and produced rare wrong pixels, see profiles/r01_summary.md):
//  * the offset slot of a tap is released at the END of the tap, after its values were consumed — an arrive issued
//    right after the ld.shared does not wait for the loads;
//  * stages >= k-blocks per tap: the k-blocks of a tap are produced by different warps that throttle only on their
//    own stage, and a parity wait cannot distinguish "two phases ahead".
//
// Offset-net output layout.  The caller permutes the output channels of the last offset conv at
// weight-pack time so that the 48 values one tap needs are contiguous per pixel:
//   channel = tap*48 + quad*12 + kind*4 + gi,  group = quad*4 + gi,  kind 0 = dy, 1 = dx, 2 = mask
// (reference order: dy/dx at (group*9 + tap)*2 + {0,1}, mask at 288 + group*9 + tap).
#include <stdlib.h>

#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kProdWarps = 16;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps + kProdWarps);  // 704
constexpr uint32_t kABytes = kTileM * kBlockK * 2;  // 16 KB
constexpr int kOmStages = 2;
constexpr int kTileW = 16, kTileH = 8;
constexpr uint32_t kOmBytes = kTileM * 48 * 2;  // one tap: 128 pixels x 48 halves
constexpr int kGroups = 16;

struct DeformArgs {
  const uint16_t* src[2];
  long long src_gstride[2], src_nstride[2];
  int src_pstride[2];
  const float* flow1;
  const float* flow2;
  const float* bias;
  uint16_t* out;
  long long out_cstride;
  int N, H, W, C;
  int kpt;   // k-blocks per tap (2C / 64)
  int tiles, tiles_x, tiles_y;  // 16 x 8-pixel tiles
  int stages;
  uint32_t b_bytes, stage_bytes;
  float mrm;
  uint32_t fmt;
  int coop;  // lane-pair gather (see the producer loop): 1 = on
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Explicit shared-window accesses (32-bit addresses, STS/LDS instead of generic ST.E/LD.E).
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t x) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(x) : "memory");
}
template <bool BF16>
__device__ __forceinline__ uint32_t add2(uint32_t x, uint32_t y) {
  if (BF16) {
    __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __half2 r = __hadd2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&y));
  return *reinterpret_cast<uint32_t*>(&r);
}
// 32-byte read-only global load (LDG.E.256, sm_100+): two 16-byte vectors
__device__ __forceinline__ void ldg256(const uint16_t* p, uint32_t (&lo)[4], uint32_t (&hi)[4]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3])
               : "l"(p));
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if (BF16) return unpack_bf16x2(u);
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (BF16) return pack_bf16x2(lo, hi);
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Blend of the 4 bilinear corners of one 16-byte (8-channel) vector.
//   fp16: packed half2 FMAs.  The reference runs torchvision's deform_conv2d in fp16 here (bilinear_interpolate is
//         templated on scalar_t = Half: every product and sum is rounded to fp16), so fp16 accumulation is not a
//         precision loss against it; it halves the instruction count of this issue-bound loop.
//   bf16: fp32 accumulation (bf16 FMAs would add ~1e-2 relative error per sample).
template <bool BF16>
__device__ __forceinline__ uint4 blend4(const uint4 (&v)[4], const float (&w)[4]) {
  uint4 o;
  if (!BF16) {
    __half2 acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const __half2 wc = __float2half2_rn(w[c]);
      const __half2* hv = reinterpret_cast<const __half2*>(&v[c]);
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] = (c == 0) ? __hmul2(wc, hv[e]) : __hfma2(wc, hv[e], acc[e]);
    }
    o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
    o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
  } else {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t uu[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(uu[e]);
        acc[2 * e] = fmaf(w[c], f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(w[c], f.y, acc[2 * e + 1]);
      }
    }
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  }
  return o;
}

// tile index -> image n and top-left pixel of the 16 x 8 tile
__device__ __forceinline__ void tile_origin(const DeformArgs& a, int tile, int& n, int& h0, int& w0) {
  const int tx = tile % a.tiles_x;
  const int t2 = tile / a.tiles_x;
  const int ty = t2 % a.tiles_y;
  n = t2 / a.tiles_y;
  h0 = ty * kTileH;
  w0 = tx * kTileW;
}

template <int VPG, bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
deform_conv_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOM,
                   const __grid_constant__ DeformArgs a) {
  pdl_sync();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float s_bias[256];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_om = smem + static_cast<size_t>(a.stages) * a.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_om + kOmStages * kOmBytes);
  uint64_t* full_bar = bars;                    // [stages]  producers (+ weight TMA) -> MMA
  uint64_t* empty_bar = bars + a.stages;        // [stages]  MMA -> producers / TMA
  uint64_t* om_full = empty_bar + a.stages;     // [kOmStages]
  uint64_t* om_empty = om_full + kOmStages;     // [kOmStages]
  uint64_t* tfull_bar = om_empty + kOmStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int n_tile = 64 * VPG;          // = C
  constexpr int tmem_cols = 2 * n_tile;     // 128 or 256
  // producer warps = 4 row quarters x 4 group quads.  C = 64: a k-block (one tap, one source) holds 8 groups ->
  // filled by 2 quads x 4 quarters; C = 128: 4 groups per k-block -> 1 quad x 4 quarters.
  constexpr int prod_arrivals = kProdWarps;   // all 16 gather warps fill every k-block

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOM);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], prod_arrivals * 32 + 1);  // every producer lane arrives after its own proxy fence
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kOmStages; ++s) {
      mbar_init(&om_full[s], 1);
      mbar_init(&om_empty[s], kProdWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(tmem_cols));
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + n_tile) {
    const int i = threadIdx.x - 64;
    s_bias[i] = (a.bias != nullptr) ? __ldg(a.bias + i) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int my_tiles = (static_cast<int>(blockIdx.x) < a.tiles)
                           ? (a.tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  const int kb_per_tile = 9 * a.kpt;
  const int hw = a.H * a.W;

  if (warp == 0) {
    // ===================== TMA: offset rows per tap, weight slab per k-block =====================
    if (elect_one()) {
      int stage = 0, os = 0;
      uint32_t phase = 0, ophase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        int tn, th0, tw0;
        tile_origin(a, tile, tn, th0, tw0);
        // channel-block-major K order: all nine taps of one 64-channel block of cat(xa, xb), then the next block.
        // The gather of a block touches 8 (C = 64) / 4 (C = 128) source planes only, so its working set (~100 KB per
        // 16 x 8 tile) stays in L1 across the nine taps; in tap-major order every tap swept all 16 planes and the L1
        // hit rate was 26 % (the kernel then ran at the L2's random-sector throughput, ~4.5 TB/s, at both shapes).
        for (int it = 0; it < 9 * a.kpt; ++it) {
          const int tap = it % 9;
          mbar_wait(&om_empty[os], ophase ^ 1);
          mbar_expect_tx(&om_full[os], kOmBytes);
          tma_load_4d(smem_om + os * kOmBytes, &tmOM, &om_full[os], tap * 48, tw0, th0, tn);
          if (++os == kOmStages) { os = 0; ophase ^= 1; }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], a.b_bytes);
          tma_load_2d(smem + static_cast<size_t>(stage) * a.stage_bytes + kABytes, &tmB, &full_bar[stage], it * kBlockK, 0);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_f16(kTileM, static_cast<uint32_t>(n_tile), a.fmt);
    const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo_flags = static_cast<uint32_t>(umma_desc_sw128(0));
    const uint32_t stage0_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
    const uint32_t stage_step = a.stage_bytes >> 4;
    int stage = 0;
    uint32_t phase = 0;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * n_tile);
      uint32_t accum = 0;
      for (int it = 0; it < kb_per_tile; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t alo = desc_lo_flags | (stage0_lo + static_cast<uint32_t>(stage) * stage_step);
          const uint32_t blo = alo + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            umma_f16(d_tmem, desc_hi | (alo + 2u * k), desc_hi | (blo + 2u * k), idesc, accum);
            accum = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (it == kb_per_tile - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int tile = blockIdx.x + local * gridDim.x;
      int tn, th0, tw0;
      tile_origin(a, tile, tn, th0, tw0);
      const int eh = th0 + (row >> 4), ew = tw0 + (row & 15);
      const bool valid = eh < a.H && ew < a.W;
      const long long pix = static_cast<long long>(tn) * hw + static_cast<long long>(eh) * a.W + ew;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc * n_tile) + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < n_tile; cc += 32) {
        uint32_t r0[16], r1[16];
        __syncwarp();
        tmem_ld16(t_addr + static_cast<uint32_t>(cc), r0);
        tmem_ld16(t_addr + static_cast<uint32_t>(cc + 16), r1);
        tmem_ld_wait();
        if (valid) {
          uint16_t* op = a.out + pix * a.out_cstride + cc;
          uint4 u[4];
          uint32_t* uw = reinterpret_cast<uint32_t*>(u);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uw[j] = pack2<BF16>(__uint_as_float(r0[2 * j]) + s_bias[cc + 2 * j], __uint_as_float(r0[2 * j + 1]) + s_bias[cc + 2 * j + 1]);
            uw[8 + j] = pack2<BF16>(__uint_as_float(r1[2 * j]) + s_bias[cc + 16 + 2 * j],
                                    __uint_as_float(r1[2 * j + 1]) + s_bias[cc + 16 + 2 * j + 1]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(op)[q] = u[q];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  } else {
    // ===================== gather producers (warps 6..21) =====================
    // K is walked channel-block-major (see the TMA warp): k-block it = (kbq, tap), kbq = 64-channel block of
    // cat(xa, xb).  All 16 warps fill the SAME k-block: warp (rq, sub) owns tile rows rq*32..+31 and
    //   C = 64 : groups kbq*8 + sub*2 + {0, 1}  (two samples per lane, 16-byte chunks sub*2 + {0, 1} of the operand row)
    //   C = 128: group  kbq*4 + sub             (one sample of 16 channels, chunks sub*2 + {0, 1})
    // Sources are "pair planes" [group][pixel][2][C/8]: entry p holds pixels p and p+1 of the row-major map, so the
    // two x-corners of a bilinear sample are ONE aligned 32-byte (C = 64) / 64-byte (C = 128) read.
    //   C = 64 : lane = pixel, 2 samples x 2 rows = 4 LDG.256 in flight per lane.
    //   C = 128: lane pairs — lanes 2k / 2k+1 fetch slot 0 / slot 1 of the same 64-byte entry in one instruction (one
    //            L1 line access instead of two), blend their column over the two rows and exchange half of the 16
    //            channels, so each lane stores 8 finished channels; the sample geometry is computed once by the lane
    //            that owns the pixel and handed to the pair with 4 shuffles.
    const int pw = warp - (2 + kEpiWarps);
    const int rq = pw & 3;        // row quarter of the tile
    const int sub = pw >> 2;      // 0..3
    const int row = rq * 32 + lane;
    const int H = a.H, W = a.W;
    const float Hf = static_cast<float>(H), Wf = static_cast<float>(W);
    const int sw = row & 7;
    const uint32_t a_row = smem_u32(smem) + row * 128;
    int os = 0;
    uint32_t ophase = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      int n, th0, tw0;
      tile_origin(a, tile, n, th0, tw0);
      const int h = th0 + (row >> 4), w = tw0 + (row & 15);
      const bool valid = h < H && w < W;
      const int off = valid ? h * W + w : 0;
      float fyx[2][2];   // [source half][y, x] flow of this pixel
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const float* fl = hf ? a.flow2 : a.flow1;
        fyx[hf][0] = valid ? __ldg(fl + (static_cast<long long>(n) * 2 + 1) * hw + off) : 0.f;
        fyx[hf][1] = valid ? __ldg(fl + (static_cast<long long>(n) * 2 + 0) * hw + off) : 0.f;
      }
#pragma unroll 1
      for (int it = 0; it < kb_per_tile; ++it) {
        const int kbq = it / 9, tap = it - kbq * 9;
        // group(s) of this warp in this k-block, their source half / plane, and their offset quad
        const int g0 = (VPG == 1) ? kbq * 8 + sub * 2 : kbq * 4 + sub;
        const int half = g0 >> 3;                 // 0: xa (flow1), 1: xb (flow2)
        const int plane0 = g0 & 7;
        const int quad = g0 >> 2, gq = g0 & 3;    // offset quad and index inside it
        const uint16_t* img = a.src[half] + static_cast<long long>(plane0) * a.src_gstride[half] + n * static_cast<int>(a.src_nstride[half]);
        const long long gstride = a.src_gstride[half];
        const int pstride = a.src_pstride[half];
        const int wps = W * pstride;
        const float fy = half ? fyx[1][0] : fyx[0][0], fx = half ? fyx[1][1] : fyx[0][1];
        // ---- this tap's (dy x4 | dx x4 | mask x4) of the quad
        mbar_wait(&om_full[os], ophase);
        const uint32_t om_row = smem_u32(smem_om) + os * kOmBytes + row * 96 + quad * 24;
        const uint2 rdy = lds64(om_row), rdx = lds64(om_row + 8), rmk = lds64(om_row + 16);
        // NOTE: the slot is released at the END of the k-block, after the values were consumed (an arrive issued right
        // after the ld.shared does not wait for the loads).
        const int tdy = tap / 3 - 1, tdx = tap - (tap / 3) * 3 - 1;
        const float by = static_cast<float>(h + tdy) + fy;
        const float bx = static_cast<float>(w + tdx) + fx;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t stage_base = smem_u32(smem) + static_cast<uint32_t>(stage) * a.stage_bytes;
        // sample geometry of group index gi (inside the quad) at this lane's pixel
        auto geometry = [&](int gi, int& o0, int& o1, float (&wgt)[4]) {
          const float2 pdy = unpack2<false>(gi < 2 ? rdy.x : rdy.y), pdx = unpack2<false>(gi < 2 ? rdx.x : rdx.y),
                       pmk = unpack2<false>(gi < 2 ? rmk.x : rmk.y);
          const float sy = by + a.mrm * tanh_fast((gi & 1) ? pdy.y : pdy.x);
          const float sx = bx + a.mrm * tanh_fast((gi & 1) ? pdx.y : pdx.x);
          float mk = __fdividef(1.0f, 1.0f + __expf(-((gi & 1) ? pmk.y : pmk.x)));
          // torchvision bilinear_interpolate: zero outside (-1, H) x (-1, W); corners outside the map contribute 0
          if (!(valid && sy > -1.f && sy < Hf && sx > -1.f && sx < Wf)) mk = 0.f;
          const float fy0 = floorf(sy), fx0 = floorf(sx);
          const float ay = sy - fy0, ax = sx - fx0;
          const int y0 = static_cast<int>(fmaxf(fminf(fy0, Hf), -2.f)), x0 = static_cast<int>(fmaxf(fminf(fx0, Wf), -2.f));
          const float wy0 = (y0 >= 0) ? (1.f - ay) * mk : 0.f, wy1 = (y0 + 1 <= H - 1) ? ay * mk : 0.f;
          // pair entry xs holds pixels (xs, xs+1).  x0 = -1: the right corner (pixel 0) is slot 0 of entry 0.
          const float wxr = (x0 + 1 <= W - 1) ? ax : 0.f;
          const float wxa = (x0 >= 0) ? 1.f - ax : wxr, wxb = (x0 >= 0) ? wxr : 0.f;
          wgt[0] = wy0 * wxa; wgt[1] = wy0 * wxb; wgt[2] = wy1 * wxa; wgt[3] = wy1 * wxb;
          // clamped coordinates: a corner with zero weight may read any valid entry
          const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1), xs = min(max(x0, 0), W - 1);
          o0 = y0c * wps + xs * pstride;
          o1 = y1c * wps + xs * pstride;
        };
        if (VPG == 1) {
          // ---- C = 64: two samples (groups g0, g0 + 1) per lane, 4 x LDG.256 in flight.
          // (A lane-pair variant over VERTICAL pair planes — the two corner columns of a sample fetched by two lanes in
          // one instruction, 1.47 instead of 2 global data-pipe wavefronts per sample — was measured in round 2: the
          // six shuffles per 16 samples it needs go through the same l1tex data pipe and gave the wavefronts back
          // (170.8K -> 166.7K per SM, 101 -> 112 us in situ), see profiles/r02_summary.md.)
          uint32_t v[2][2][2][4];  // [sample][row][slot][words]
          float wgt[2][4];
#pragma unroll
          for (int sgi = 0; sgi < 2; ++sgi) {
            int o0, o1;
            geometry(gq + sgi, o0, o1, wgt[sgi]);
            const uint16_t* gb = img + sgi * gstride;
            ldg256(gb + o0, v[sgi][0][0], v[sgi][0][1]);
            ldg256(gb + o1, v[sgi][1][0], v[sgi][1][1]);
          }
#pragma unroll
          for (int sgi = 0; sgi < 2; ++sgi) {
            uint4 c4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t* pv = v[sgi][c >> 1][c & 1];
              c4[c] = make_uint4(pv[0], pv[1], pv[2], pv[3]);
            }
            const uint4 o = blend4<BF16>(c4, wgt[sgi]);
            sts128(stage_base + row * 128 + (((sub * 2 + sgi) ^ sw) << 4), o.x, o.y, o.z, o.w);
          }
        } else if (!a.coop) {
          // ---- C = 128, one lane per sample: 4 x LDG.256 (2 rows x 2 slots of 16 channels)
          uint32_t v[2][4][4];  // [row][16-byte vector: slot0 lo, slot0 hi, slot1 lo, slot1 hi][words]
          float wgt[4];
          int o0, o1;
          geometry(gq, o0, o1, wgt);
          ldg256(img + o0, v[0][0], v[0][1]);
          ldg256(img + o0 + 16, v[0][2], v[0][3]);
          ldg256(img + o1, v[1][0], v[1][1]);
          ldg256(img + o1 + 16, v[1][2], v[1][3]);
#pragma unroll
          for (int vi = 0; vi < 2; ++vi) {
            uint4 c4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t* pv = v[c >> 1][(c & 1) * 2 + vi];
              c4[c] = make_uint4(pv[0], pv[1], pv[2], pv[3]);
            }
            const uint4 o = blend4<BF16>(c4, wgt);
            sts128(stage_base + row * 128 + (((sub * 2 + vi) ^ sw) << 4), o.x, o.y, o.z, o.w);
          }
        } else {
          // ---- C = 128, lane pairs
          const int s_slot = lane & 1;
          float wgt[4];
          int off0, off1;
          geometry(gq, off0, off1, wgt);
          // weights as (slot 0, slot 1) half pairs per row
          const uint32_t wrow0 = pack2<false>(wgt[0], wgt[1]), wrow1 = pack2<false>(wgt[2], wgt[3]);
          const uint16_t* gplane = img + s_slot * 16;
          uint32_t v[2][2][8];   // [pixel half][row][8 words = 16 channels]
          uint32_t wsel[2][2];   // [pixel half][row]: this lane's weight, broadcast to both halves
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const int srcl = 16 * ph + (lane >> 1);
            const int o0 = __shfl_sync(0xffffffffu, off0, srcl), o1 = __shfl_sync(0xffffffffu, off1, srcl);
            const uint32_t w0 = __shfl_sync(0xffffffffu, wrow0, srcl), w1 = __shfl_sync(0xffffffffu, wrow1, srcl);
            wsel[ph][0] = __byte_perm(w0, 0, s_slot ? 0x3232 : 0x1010);
            wsel[ph][1] = __byte_perm(w1, 0, s_slot ? 0x3232 : 0x1010);
            uint32_t (&a0)[8] = v[ph][0];
            uint32_t (&a1)[8] = v[ph][1];
            ldg256(gplane + o0, *reinterpret_cast<uint32_t (*)[4]>(&a0[0]), *reinterpret_cast<uint32_t (*)[4]>(&a0[4]));
            ldg256(gplane + o1, *reinterpret_cast<uint32_t (*)[4]>(&a1[0]), *reinterpret_cast<uint32_t (*)[4]>(&a1[4]));
          }
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            uint32_t mine[4];
            if (!BF16) {
              const __half2 wA = *reinterpret_cast<const __half2*>(&wsel[ph][0]), wB = *reinterpret_cast<const __half2*>(&wsel[ph][1]);
              uint32_t part[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const __half2 r = __hfma2(wB, *reinterpret_cast<const __half2*>(&v[ph][1][j]),
                                          __hmul2(wA, *reinterpret_cast<const __half2*>(&v[ph][0][j])));
                part[j] = *reinterpret_cast<const uint32_t*>(&r);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t send = s_slot ? part[j] : part[4 + j];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
                mine[j] = add2<false>(s_slot ? part[4 + j] : part[j], recv);
              }
            } else {
              const float wA = __half2float(__ushort_as_half(static_cast<unsigned short>(wsel[ph][0] & 0xFFFFu)));
              const float wB = __half2float(__ushort_as_half(static_cast<unsigned short>(wsel[ph][1] & 0xFFFFu)));
              float part[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float2 f0 = unpack_bf16x2(v[ph][0][j]), f1 = unpack_bf16x2(v[ph][1][j]);
                part[2 * j] = fmaf(wB, f1.x, wA * f0.x);
                part[2 * j + 1] = fmaf(wB, f1.y, wA * f0.y);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float s0 = s_slot ? part[2 * j] : part[8 + 2 * j], s1 = s_slot ? part[2 * j + 1] : part[8 + 2 * j + 1];
                const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                const float k0 = s_slot ? part[8 + 2 * j] : part[2 * j], k1 = s_slot ? part[8 + 2 * j + 1] : part[2 * j + 1];
                mine[j] = pack_bf16x2(k0 + r0, k1 + r1);
              }
            }
            const int prow = rq * 32 + 16 * ph + (lane >> 1);   // tile row (pixel) this lane pair produced
            sts128(stage_base + prow * 128 + (((sub * 2 + s_slot) ^ (prow & 7)) << 4), mine[0], mine[1], mine[2], mine[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[stage]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&om_empty[os]);
        if (++os == kOmStages) { os = 0; ophase ^= 1; }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

template <int VPG, bool BF16>
cudaError_t launch_deform(int grid, size_t smem_bytes, cudaStream_t stream, const CUtensorMap& tmB, const CUtensorMap& tmOM,
                          const DeformArgs& a) {
  static FlairPerDeviceOnce attr_once;  // one per template instantiation
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(deform_conv_kernel<VPG, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
    if (e != cudaSuccess) return e;
    // ask for the smallest shared-memory carve-out that fits: the rest of the 256 KB stays L1 for the gather
    const int pct = static_cast<int>((smem_bytes + 2048) * 100 / (228 * 1024)) + 1;
    e = cudaFuncSetAttribute(deform_conv_kernel<VPG, BF16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
    if (e != cudaSuccess) return e;
  }
  return flair_launch(deform_conv_kernel<VPG, BF16>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmB, tmOM, a);
}

}  // namespace

extern "C" int flair_deform_conv(const flair_deform_conv_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(p != nullptr && p->xa && p->xb && p->om && p->flow1 && p->flow2 && p->wgt && p->out,
                "flair_deform_conv: null pointer");
  FLAIR_REQUIRE(p->deform_groups == kGroups, "flair_deform_conv: built for 16 deform groups (got %d)", p->deform_groups);
  FLAIR_REQUIRE(p->C == 64 || p->C == 128, "flair_deform_conv: C must be 64 or 128 (got %d); use flair_deform_im2col otherwise", p->C);
  FLAIR_REQUIRE(p->dtype == FLAIR_F16 || p->dtype == FLAIR_BF16, "flair_deform_conv: 16-bit operands only");
  FLAIR_REQUIRE(p->om_cstride >= 432 && p->om_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->om) & 15) == 0,
                "flair_deform_conv: offset map needs >= 432 channels, stride multiple of 8, 16-byte aligned");
  FLAIR_REQUIRE(p->out_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0, "flair_deform_conv: out alignment");
  const int cpg = 2 * p->C / kGroups;
  for (int s = 0; s < 2; ++s) {
    const void* ptr = s ? p->xb : p->xa;
    const long long gs = s ? p->xb_gstride : p->xa_gstride, ps = s ? p->xb_pstride : p->xa_pstride;
    const long long ns = s ? p->xb_nstride : p->xa_nstride;
    FLAIR_REQUIRE(static_cast<long long>(p->H) * p->W * ps + (p->N - 1) * ns < (1ll << 31) && ns % 8 == 0,
                  "flair_deform_conv: maps too large for 32-bit sample offsets");
    FLAIR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && gs % 16 == 0 && ps == 2 * cpg && ns % 16 == 0,
                  "flair_deform_conv: source %d must be 32-byte aligned pair planes (pixel stride 2 * C/8 = %d)", s, 2 * cpg);
  }
  const long long total_pix = static_cast<long long>(p->N) * p->H * p->W;
  FLAIR_REQUIRE(total_pix > 0 && total_pix < (1ll << 31) - kTileM, "flair_deform_conv: bad extents");

  DeformArgs a{};
  a.src[0] = static_cast<const uint16_t*>(p->xa); a.src[1] = static_cast<const uint16_t*>(p->xb);
  a.src_gstride[0] = p->xa_gstride; a.src_gstride[1] = p->xb_gstride;
  a.src_pstride[0] = static_cast<int>(p->xa_pstride); a.src_pstride[1] = static_cast<int>(p->xb_pstride);
  a.src_nstride[0] = p->xa_nstride; a.src_nstride[1] = p->xb_nstride;
  a.flow1 = p->flow1; a.flow2 = p->flow2; a.bias = p->bias;
  a.out = static_cast<uint16_t*>(p->out); a.out_cstride = p->out_cstride;
  a.N = p->N; a.H = p->H; a.W = p->W; a.C = p->C;
  a.kpt = 2 * p->C / kBlockK;
  a.tiles_x = ceil_div(p->W, kTileW); a.tiles_y = ceil_div(p->H, kTileH);
  a.tiles = a.tiles_x * a.tiles_y * p->N;
  a.b_bytes = static_cast<uint32_t>(p->C) * kBlockK * 2;
  a.stage_bytes = kABytes + a.b_bytes;
  a.mrm = p->max_residue_magnitude;
  a.fmt = (p->dtype == FLAIR_BF16) ? 1u : 0u;
  {
    static int coop = -1;   // FLAIR_DEFORM_COOP=0: one lane per sample everywhere (A/B measurements)
    if (coop < 0) { const char* e = getenv("FLAIR_DEFORM_COOP"); coop = e ? atoi(e) : 1; }
    a.coop = coop;
  }
  // Shared memory is kept to ~128 KB so that the 132 KB carve-out leaves ~120 KB of L1: the gather re-reads every
  // source pixel ~36 times (9 taps x 4 corners) and with 16 x 8 tiles its working set is ~140 KB; with all 227 KB
  // given to the pipeline the L1 hit rate was 5 % and the kernel ran at the L2 bandwidth limit (958 MB / launch).
  static int smem_kb = 0;
  if (smem_kb == 0) {
    const char* e = getenv("FLAIR_DEFORM_SMEM_KB");
    smem_kb = e ? atoi(e) : 128;
    if (smem_kb < 96) smem_kb = 96;
    if (smem_kb > 224) smem_kb = 224;
  }
  const int budget = smem_kb * 1024 - 1024 - 512 - 1024 /*static bias*/ - kOmStages * static_cast<int>(kOmBytes);
  int stages = budget / static_cast<int>(a.stage_bytes);
  if (stages > 8) stages = 8;
  // The ring must hold at least one whole tap (kpt k-blocks).  The k-blocks of a tap are filled by DIFFERENT warps,
  // which only throttle on their own stage: with fewer stages than k-blocks per tap a fast warp can be two phases
  // ahead of the MMA on a stage, and a parity wait cannot tell "two phases ago" from "now" (it then overwrote a
  // stage that had not been consumed: non-deterministic results at C = 128 with 3 stages).
  // (all 16 gather warps fill the same k-block now, so the ring no longer has to hold a whole tap)
  FLAIR_REQUIRE(stages >= 2, "flair_deform_conv: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem_bytes = static_cast<size_t>(stages) * a.stage_bytes + kOmStages * kOmBytes + 1024 + 512;

  flair_tmap_encode_fn encode = flair_get_tmap_encode();
  FLAIR_REQUIRE(encode != nullptr, "flair_deform_conv: cuTensorMapEncodeTiled unavailable");
  const CUtensorMapDataType dt16 = (p->dtype == FLAIR_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmB, tmOM;
  {
    // packed weight [1][C_pad16][18C] K-major, channel-block-major K: k = (kbq*9 + tap)*64 + c, kbq = 64-channel block of cat(xa, xb)
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(18 * p->C), static_cast<cuuint64_t>(p->C)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(18 * p->C) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(p->C)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&tmB, dt16, 2, const_cast<void*>(p->wgt), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLAIR_REQUIRE(r == CUDA_SUCCESS, "flair_deform_conv: weight tensor map rejected (%d)", static_cast<int>(r));
  }
  {
    const cuuint64_t px = static_cast<cuuint64_t>(p->om_cstride) * 2;
    cuuint64_t dims[4] = {432u, static_cast<cuuint64_t>(p->W), static_cast<cuuint64_t>(p->H), static_cast<cuuint64_t>(p->N)};
    cuuint64_t strides[3] = {px, px * p->W, px * p->W * p->H};
    cuuint32_t box[4] = {48u, static_cast<cuuint32_t>(kTileW), static_cast<cuuint32_t>(kTileH), 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = encode(&tmOM, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p->om), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLAIR_REQUIRE(r == CUDA_SUCCESS, "flair_deform_conv: offset tensor map rejected (%d)", static_cast<int>(r));
  }
  int grid = flair_num_sms();
  if (grid > a.tiles) grid = a.tiles;
  const bool bf16 = p->dtype == FLAIR_BF16;
  cudaError_t le;
  if (p->C == 64) le = bf16 ? launch_deform<1, true>(grid, smem_bytes, stream, tmB, tmOM, a) : launch_deform<1, false>(grid, smem_bytes, stream, tmB, tmOM, a);
  else le = bf16 ? launch_deform<2, true>(grid, smem_bytes, stream, tmB, tmOM, a) : launch_deform<2, false>(grid, smem_bytes, stream, tmB, tmOM, a);
  FLAIR_CHECK_CUDA(le);
  FLAIR_CHECK_LAUNCH();
  return 0;
}
