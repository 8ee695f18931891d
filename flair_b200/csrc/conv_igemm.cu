// Implicit-GEMM convolution / GEMM for sm_100a: TMA -> 128B-swizzled smem ->
// tcgen05.mma (UMMA, M=128) -> fp32 accumulators in TMEM -> fused epilogue.
//
// Replaces (reference, paths relative to the FLAIR checkout): every nn.Conv2d /
// nn.Conv3d / nn.Conv1d(k=1) / nn.Linear of the UNet torso, e.g.
// guided_diffusion/unet_new.py:240-244 (ResBlock in conv), :271-276 (out conv),
// :292-295 (1x1 skip), :359,367 (qkv / proj), :455-457 (temporal q/k/v),
// :859-867 (BasicVSR++ offset net); guided_diffusion/sr3.py:95-120.
//
// Data layout.  Activations are channels-last [B][T][H][W][C] (16-bit), so an
// output tile of 128 pixels x 64 input channels is exactly one TMA box
// {64, bw, bh, bt, 1}; it lands in smem as 128 rows of 128 B, i.e. the canonical
// K-major SWIZZLE_128B UMMA operand.  A filter tap (dt,dh,dw) is nothing but a
// shifted box origin, and TMA's out-of-bounds zero fill *is* the zero padding.
// Weights are packed [tap][Cout_pad][Cin_pad] so a (n_tile x 64) slab of one tap
// is one 2-D box, also K-major.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// alloc), warps 2..9 = epilogue (two per TMEM lane quarter, splitting the columns).  The epilogue
// is specialised at compile time on (output type/layout, activation): a runtime-generic version
// cost ~6k warp-instructions per tile and left the tensor pipe 6 % busy (profiles/r01_conv64_*).  The kernel is
// persistent: grid = min(#tiles, #SMs); accumulators are double-buffered in
// TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;                  // two per TMEM lane quarter (columns split in halves)
constexpr int kThreads = 64 + 32 * kEpiWarps;  // + TMA warp + MMA warp
constexpr int kMaxTaps = 27;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;

struct ConvKArgs {
  int B, T, Ho, Wo;
  int lbw, lbh, lbt;  // log2 of the tile box extents (bw*bh*bt == 128)
  int tiles_w, tiles_h, tiles_t;
  int m_tiles, n_tiles;
  int n_tile;   // accumulator columns per tile (multiple of 16, <= 256)
  int acc_cols; // TMEM column offset between the two accumulator buffers
  int tmem_cols;
  int Cout, Cout_pad;
  int kblocks, ntaps, stride;
  int trace;   // debug timeline on/off
  int k_last;  // UMMA K=16 steps that hold real channels in the LAST k-block (Cin = 196: 1 of 4)
  int stages;
  // mode 0: one (tap, k-block) per pipeline stage, A tile = 128 output pixels.
  // mode 1: "halo" — 16x8-pixel tiles; one (dt, dw, k-block) per stage loads a 16x10 halo slab of A
  //         once and issues the three dh taps from it (start address shifted by whole 2 KB row
  //         groups), cutting A traffic 2.4x.
  // mode 2: mode 1 + the weights of this CTA's N tile stay resident in shared memory.
  int mode, ntd;
  uint32_t a_bytes, b_bytes, stage_bytes, w_bytes;
  int8_t tap_dw[kMaxTaps], tap_dh[kMaxTaps], tap_dt[kMaxTaps];
  const float* bias;
  const float* rowbias;
  int rowbias_stride;
  const float* rowscale;
  int rowscale_stride;
  const void* preadd;  // added to the accumulator BEFORE the activation (precomputed partial convolution)
  int preadd_dtype;
  long long preadd_cstride;
  const void* residual;
  int residual_dtype;
  long long residual_cstride;
  const void* residual2;
  int residual2_dtype;
  long long residual2_cstride;
  void* out;
  int out_dtype, out_layout;
  long long out_cstride;
  int act;
  float out_scale;
  uint32_t fmt;
  float* gn_partial;
  int gn_groups;
  uint16_t* out2;          // optional copy of a 16-bit NHWC output as pair planes [Cout/out2_gs][pixels][2][out2_gs]:
  int out2_gs;             // entry p = (pixel p, pixel p+1), the source layout of flair_deform_conv
  long long out2_gstride;  // elements between group planes
};

// Debug timeline: in a library built with -DFLAIR_CONV_TRACE_BUILD (FLAIR_BUILD_TRACE=1 python -m flair_b200.build)
// and run with FLAIR_CONV_TRACE=1, CTA 0 records clock64() at a few points; read with flair_debug_conv_trace.
// Compiled out otherwise (the marks sit in the TMA / MMA issue loops).
__device__ long long g_conv_trace[16];
__device__ __forceinline__ void trace_mark(int on, int slot) {
#ifdef FLAIR_CONV_TRACE_BUILD
  if (on && blockIdx.x == 0) g_conv_trace[slot] = clock64();
#endif
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FLAIR_ACT_RELU: return fmaxf(v, 0.0f);
    case FLAIR_ACT_LRELU01: return v > 0.0f ? v : 0.1f * v;
    case FLAIR_ACT_SILU: return silu_f(v);
    default: return v;
  }
}

__device__ __forceinline__ uint32_t pack16(float lo, float hi, int dtype) {
  if (dtype == FLAIR_F16) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int dtype) {
  if (dtype == FLAIR_F16) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
  }
  return unpack_bf16x2(u);
}

// v[j] += residual[off + j] for the 16 columns of one accumulator chunk
__device__ __forceinline__ void add_residual(float (&v)[16], const void* base, int dtype, long long off,
                                             bool full16, int remaining) {
  if (dtype == FLAIR_F32) {
    const float* rp = static_cast<const float*>(base) + off;
    if (full16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(rp) + q);
        v[4 * q + 0] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < remaining) v[j] += __ldg(rp + j);
    }
  } else {
    const uint16_t* rp = static_cast<const uint16_t*>(base) + off;
    if (full16) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + q);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack16(uu[e], dtype);
          v[8 * q + 2 * e] += f.x;
          v[8 * q + 2 * e + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < remaining) {
          const uint32_t u = rp[j];
          v[j] += unpack16(u, dtype).x;
        }
      }
    }
  }
}

struct EpiPos {
  int w, h, n0;
  bool valid;
  long long frame, pix;
};

// One epilogue warp: accumulator columns [col_begin, col_end) of its 32 TMEM lanes, 16 at a time.
// KIND: 0 fp16 NHWC, 1 bf16 NHWC, 2 fp32 NHWC, 3 fp32 NCHW.  ACT: FLAIR_ACT_*.
// One 16-column chunk of one epilogue warp: bias, per-frame bias, activation, gate, residuals, store.
// KIND: 0 fp16 NHWC, 1 bf16 NHWC, 2 fp32 NHWC, 3 fp32 NCHW.  ACT: FLAIR_ACT_*.
template <int KIND, int ACT>
__device__ __forceinline__ void epilogue_chunk(const ConvKArgs& a, const uint32_t (&r)[16], int c0, const EpiPos& pos,
                                               const float* __restrict__ sb) {
    const int n = pos.n0 + c0;
    if (n >= a.Cout) return;  // warp-uniform: padded columns
    float v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bq = *reinterpret_cast<const float4*>(sb + c0 + 4 * q);
      v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
      v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
      v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
      v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
    }
    const bool full16 = (n + 16 <= a.Cout);
    if (a.rowbias != nullptr && pos.valid) {
      const float* rb = a.rowbias + pos.frame * a.rowbias_stride + n;
      if (full16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(rb) + q);
          v[4 * q + 0] += f.x; v[4 * q + 1] += f.y; v[4 * q + 2] += f.z; v[4 * q + 3] += f.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < a.Cout) v[j] += __ldg(rb + j);
      }
    }
    if (a.preadd != nullptr && pos.valid)
      add_residual(v, a.preadd, a.preadd_dtype, pos.pix * a.preadd_cstride + n, full16, a.Cout - n);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float y = v[j];
      if (ACT == FLAIR_ACT_RELU) y = fmaxf(y, 0.0f);
      else if (ACT == FLAIR_ACT_LRELU01) y = y > 0.0f ? y : 0.1f * y;
      else if (ACT == FLAIR_ACT_SILU) y = silu_f(y);
      v[j] = y * a.out_scale;
    }
    if (a.rowscale != nullptr && pos.valid) {  // per-(frame, channel) gate, e.g. sigmoid(g) of TemporalWrapper2
      const float* rs = a.rowscale + pos.frame * a.rowscale_stride + n;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < a.Cout) v[j] *= __ldg(rs + j);
    }
    if (a.residual != nullptr && pos.valid)
      add_residual(v, a.residual, a.residual_dtype, pos.pix * a.residual_cstride + n, full16, a.Cout - n);
    if (a.residual2 != nullptr && pos.valid)
      add_residual(v, a.residual2, a.residual2_dtype, pos.pix * a.residual2_cstride + n, full16, a.Cout - n);
    if (!pos.valid) return;
    if (KIND == 3) {
      // fp32 planar output: (frame, n, h, w); lanes walk w -> coalesced per channel
      float* op = static_cast<float*>(a.out);
      const long long plane = static_cast<long long>(a.Ho) * a.Wo;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < a.Cout)
          op[(pos.frame * a.Cout + n + j) * plane + static_cast<long long>(pos.h) * a.Wo + pos.w] = v[j];
    } else if (KIND == 2) {
      float* op = static_cast<float*>(a.out) + pos.pix * a.out_cstride + n;
      if (full16) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(op)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < a.Cout) op[j] = v[j];
      }
    } else {
      constexpr int DT = (KIND == 0) ? FLAIR_F16 : FLAIR_BF16;
      uint16_t* op = static_cast<uint16_t*>(a.out) + pos.pix * a.out_cstride + n;
      if (full16) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 u;
          u.x = pack16(v[8 * q + 0], v[8 * q + 1], DT);
          u.y = pack16(v[8 * q + 2], v[8 * q + 3], DT);
          u.z = pack16(v[8 * q + 4], v[8 * q + 5], DT);
          u.w = pack16(v[8 * q + 6], v[8 * q + 7], DT);
          reinterpret_cast<uint4*>(op)[q] = u;
          if (a.out2 != nullptr) {  // pair planes for the deformable gather: slot 0 of entry pix, slot 1 of entry pix-1
            const int ch = n + 8 * q;
            uint16_t* e = a.out2 + (ch / a.out2_gs) * a.out2_gstride + pos.pix * (2 * a.out2_gs) + (ch % a.out2_gs);
            *reinterpret_cast<uint4*>(e) = u;
            if (pos.pix > 0) *reinterpret_cast<uint4*>(e - a.out2_gs) = u;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < a.Cout) op[j] = static_cast<uint16_t>(pack16(v[j], 0.0f, DT) & 0xFFFFu);
      }
    }
}

// One epilogue warp: accumulator columns [col_begin, col_end) of its 32 TMEM lanes; two 16-column TMEM
// loads are in flight per wait.
template <int KIND, int ACT>
__device__ __forceinline__ void epilogue_cols(const ConvKArgs& a, uint32_t t_addr, int col_begin, int col_end,
                                              const EpiPos& pos, const float* __restrict__ sb) {
  for (int cc = col_begin; cc < col_end; cc += 32) {
    uint32_t r0[16], r1[16];
    __syncwarp();  // tcgen05.ld is warp-collective: reconverge after divergent stores
    tmem_ld16(t_addr + static_cast<uint32_t>(cc), r0);
    const bool second = cc + 16 < col_end;  // warp-uniform
    tmem_ld16(t_addr + static_cast<uint32_t>(cc + 16), r1);  // (columns past the tile are allocated, unused)
    tmem_ld_wait();
    epilogue_chunk<KIND, ACT>(a, r0, cc, pos, sb);
    if (second) epilogue_chunk<KIND, ACT>(a, r1, cc + 16, pos, sb);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvKArgs a) {
  // PDL: let the next launch's CTAs be scheduled as soon as SMs free up.  Nothing before pdl_wait() below touches
  // memory written by an earlier kernel: barrier init, TMEM allocation, descriptor prefetch and the loads of this
  // CTA's resident weights (constant since model load) all overlap the previous kernel's tail.
  pdl_trigger();
  if (threadIdx.x == 0) trace_mark(a.trace, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float s_bias[2][256];
  // carve: [resident weights (mode 2)] [stages][A][B] then barriers
  const uint32_t b_bytes = a.b_bytes;
  const uint32_t stage_bytes = a.stage_bytes;
  uint8_t* smem_w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
  uint8_t* smem = smem_w + a.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(a.stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + a.stages;
  uint64_t* tfull_bar = bars + 2 * a.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;  // [ntd * 3]: resident weights arrive per (dt, dw) group, in the order the MMAs use them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    for (int g = 0; g < 9; ++g) mbar_init(&w_bar[g], 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_mark(a.trace, 1);
  // (pdl_wait() is executed per role below: the MMA warp first issues the weight loads)

  // tile walk of this CTA: (m_idx, n_idx) = f(local iteration)
  //   modes 0/1: linear tile id = blockIdx.x + i*gridDim.x, n fastest (neighbouring CTAs share A in L2)
  //   mode 2   : n_idx fixed per CTA (its weights are resident), m strided by gridDim.x / n_tiles
  const int total_tiles = a.m_tiles * a.n_tiles;
  const bool resident = (a.mode == 2);
  const int my_n = resident ? static_cast<int>(blockIdx.x) % a.n_tiles : 0;
  const int m_first = resident ? static_cast<int>(blockIdx.x) / a.n_tiles : 0;
  const int m_step = resident ? static_cast<int>(gridDim.x) / a.n_tiles : 0;
  const int my_tiles = resident ? (m_first < a.m_tiles ? (a.m_tiles - 1 - m_first) / m_step + 1 : 0)
                                : (static_cast<int>(blockIdx.x) < total_tiles
                                       ? (total_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1
                                       : 0);
  auto tile_at = [&](int i, int& m_idx, int& n_idx) {
    if (resident) { m_idx = m_first + i * m_step; n_idx = my_n; }
    else { const int tile = blockIdx.x + i * gridDim.x; n_idx = tile % a.n_tiles; m_idx = tile / a.n_tiles; }
  };
  const int k_iters = (a.mode == 0) ? a.ntaps * a.kblocks : a.ntd * 3 * a.kblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // elect.sync instead of `lane == 0`: the compiler then keeps coordinates / addresses in uniform registers and
    // emits a bare UTMALDG; under `lane == 0` every TMA issue was a waterfall loop (ELECT + 7 R2UR.BROADCAST +
    // branch), ~300 cycles per issue — with 4 loads per stage that, not bandwidth, bounded the streamed-weight mode.
    pdl_wait();  // activations are the previous kernel's output
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      trace_mark(a.trace, 2);
      for (int i = 0; i < my_tiles; ++i) {
        int m_idx, n_idx;
        tile_at(i, m_idx, n_idx);
        if (i == 1) trace_mark(a.trace, 3);
        const int tw = m_idx % a.tiles_w; m_idx /= a.tiles_w;
        const int th = m_idx % a.tiles_h; m_idx /= a.tiles_h;
        const int tt = m_idx % a.tiles_t;
        const int b = m_idx / a.tiles_t;
        const int w0 = (tw << a.lbw) * a.stride;
        const int h0 = (th << a.lbh) * a.stride;
        const int t0 = tt << a.lbt;
        if (a.mode == 0) {
          for (int tap = 0; tap < a.ntaps; ++tap) {
            const int cw = w0 + a.tap_dw[tap];
            const int ch = h0 + a.tap_dh[tap];
            const int ct = t0 + a.tap_dt[tap];
            const int brow = tap * a.Cout_pad + n_idx * a.n_tile;
            for (int kb = 0; kb < a.kblocks; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
              uint8_t* sb = sa + a.a_bytes;
              mbar_expect_tx(&full_bar[stage], a.a_bytes + b_bytes);
              tma_load_5d(sa, &tmA, &full_bar[stage], kb * kBlockK, cw, ch, ct, b);
              tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBlockK, brow);
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          for (int dti = 0; dti < a.ntd; ++dti) {
            const int ct = t0 + dti - a.ntd / 2;
            for (int dwi = 0; dwi < 3; ++dwi) {
              for (int kb = 0; kb < a.kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
                mbar_expect_tx(&full_bar[stage], a.a_bytes + (resident ? 0u : 3u * b_bytes));
                // 16 x 10 halo slab: rows h0-1 .. h0+8 at column shift dw
                tma_load_5d(sa, &tmA, &full_bar[stage], kb * kBlockK, w0 + dwi - 1, h0 - 1, ct, b);
                if (!resident) {
                  for (int dhi = 0; dhi < 3; ++dhi) {
                    const int tap = (dti * 3 + dhi) * 3 + dwi;
                    tma_load_2d(sa + a.a_bytes + dhi * b_bytes, &tmB, &full_bar[stage], kb * kBlockK,
                                tap * a.Cout_pad + n_idx * a.n_tile);
                  }
                }
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One ELECTED lane issues (elect.sync: the compiler then knows the region is single-threaded and emits
    // UTCHMMA + 2 UMOV per MMA; under `if (lane == 0)` it wrapped every MMA in a 10-instruction waterfall loop).
    // The per-MMA instruction count is what bounds small-N tiles (a first version
    // rebuilt both 64-bit descriptors and did integer div/mod per MMA: ~150 cycles per issue, tensor pipe
    // 6 % busy).  Descriptors are (constant high word, low word = smem address >> 4): the loop only adds
    // small constants to the low words.
    const uint32_t idesc = umma_idesc_f16(kBlockM, static_cast<uint32_t>(a.n_tile), a.fmt);
    int stage = 0;
    uint32_t phase = 0;
    if (resident && elect_one()) {
      // this CTA's weights: every (tap, k-block) slab of its N tile, once, grouped per (dt, dw) in MMA order so the
      // first MMAs start after 3 * kblocks slabs, not after the whole filter.  Issued from this (otherwise idle) warp
      // while warp 0 issues the activation loads, and BEFORE pdl_wait: weights do not depend on the previous kernel.
      const uint32_t group_bytes = 3u * static_cast<uint32_t>(a.kblocks) * b_bytes;
      for (int dti = 0; dti < a.ntd; ++dti)
        for (int dwi = 0; dwi < 3; ++dwi) {
          uint64_t* bar = &w_bar[dti * 3 + dwi];
          mbar_expect_tx(bar, group_bytes);
          for (int kb = 0; kb < a.kblocks; ++kb)
            for (int dhi = 0; dhi < 3; ++dhi) {
              const int tap = (dti * 3 + dhi) * 3 + dwi;
              tma_load_2d(smem_w + static_cast<size_t>(tap * a.kblocks + kb) * b_bytes, &tmB, bar, kb * kBlockK,
                          tap * a.Cout_pad + my_n * a.n_tile);
            }
        }
    }
    __syncwarp();
    pdl_wait();
    if (lane == 0) trace_mark(a.trace, 4);
    const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;   // SBO, version, swizzle mode
    const uint32_t desc_lo_flags = static_cast<uint32_t>(umma_desc_sw128(0));  // LBO field
    const uint32_t stage0_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
    const uint32_t stage_step = stage_bytes >> 4;
    const uint32_t a_step = a.a_bytes >> 4;
    const uint32_t b_step = b_bytes >> 4;
    const uint32_t w_lo = (smem_u32(smem_w) & 0x3FFFF) >> 4;
    const uint32_t w_tap_step = static_cast<uint32_t>(a.kblocks) * b_step;  // resident weights: [tap][kb]
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * a.acc_cols);
      if (lane == 0 && local < 4) trace_mark(a.trace, 5 + local);  // MMA of tile `local` may start (TMEM buffer free)
      uint32_t accum = 0;  // first MMA of the tile overwrites the accumulator
      if (a.mode == 0) {
        int kb = 0;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // zero-padded channels of the last k-block are skipped in whole UMMA K steps (the TMA zero fill and the
          // zero weight columns make them no-ops; Cin = 196 / 388 would otherwise pay for 256 / 448 channels)
          const int ksteps = (kb == a.kblocks - 1) ? a.k_last : kBlockK / kUmmaK;
          if (++kb == a.kblocks) kb = 0;
          if (elect_one()) {
            const uint32_t alo = desc_lo_flags | (stage0_lo + static_cast<uint32_t>(stage) * stage_step);
            const uint32_t blo = alo + a_step;
            if (ksteps == kBlockK / kUmmaK) {
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                // +32 B per K step inside the 128 B swizzle row (encoded >> 4)
                umma_f16(d_tmem, desc_hi | (alo + 2u * k), desc_hi | (blo + 2u * k), idesc, accum);
                accum = 1;
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                umma_f16(d_tmem, desc_hi | (alo + 2u * k), desc_hi | (blo + 2u * k), idesc, accum);
                accum = 1;
              }
            }
            umma_commit(&empty_bar[stage]);
            if (it == k_iters - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      } else {
        int it = 0;
        for (int dti = 0; dti < a.ntd; ++dti) {
          for (int dwi = 0; dwi < 3; ++dwi) {
            // resident weights of tap (dti, dhi=0, dwi), k-block 0
            uint32_t wlo = desc_lo_flags | (w_lo + static_cast<uint32_t>(dti * 9 + dwi) * w_tap_step);
            if (resident && local == 0) mbar_wait(&w_bar[dti * 3 + dwi], 0);  // this group's weight slabs have landed
            for (int kb = 0; kb < a.kblocks; ++kb, ++it, wlo += b_step) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const int ksteps = (kb == a.kblocks - 1) ? a.k_last : kBlockK / kUmmaK;
              if (elect_one()) {
                const uint32_t alo = desc_lo_flags | (stage0_lo + static_cast<uint32_t>(stage) * stage_step);
#pragma unroll
                for (int dhi = 0; dhi < 3; ++dhi) {
                  // A: the dh tap is the same halo slab shifted by one 16-pixel row (2 KB); B: slab dhi of
                  // this stage, or the resident slab of tap (dti, dhi, dwi) = base + dhi * 3 taps
                  const uint32_t al = alo + static_cast<uint32_t>(dhi) * (2048u >> 4);
                  const uint32_t bl = resident ? wlo + static_cast<uint32_t>(dhi) * 3u * w_tap_step
                                               : alo + a_step + static_cast<uint32_t>(dhi) * b_step;
                  if (ksteps == kBlockK / kUmmaK) {
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                      umma_f16(d_tmem, desc_hi | (al + 2u * k), desc_hi | (bl + 2u * k), idesc, accum);
                      accum = 1;
                    }
                  } else {
                    for (int k = 0; k < ksteps; ++k) {
                      umma_f16(d_tmem, desc_hi | (al + 2u * k), desc_hi | (bl + 2u * k), idesc, accum);
                      accum = 1;
                    }
                  }
                }
                umma_commit(&empty_bar[stage]);
                if (it == k_iters - 1) umma_commit(&tfull_bar[acc]);
              }
              __syncwarp();
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    pdl_wait();  // residuals / per-frame biases may be the previous kernel's output
    const int ewarp = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int half = ewarp >> 2;           // which half of the accumulator columns
    const int row = quarter * 32 + lane;
    const int bw = 1 << a.lbw, bh = 1 << a.lbh;
    const int col_split = ((a.n_tile / 2 + 15) / 16) * 16;
    const int col_begin = half ? col_split : 0;
    const int col_end = half ? a.n_tile : col_split;
    const int etid = threadIdx.x - 64;     // 0..255
    // epilogue kind: 0 = fp16 NHWC, 1 = bf16 NHWC, 2 = fp32 NHWC, 3 = fp32 NCHW
    const int kind = (a.out_layout == FLAIR_OUT_NCHW) ? 3 : (a.out_dtype == FLAIR_F32 ? 2 : (a.out_dtype == FLAIR_F16 ? 0 : 1));
    int bias_n = -1, bias_buf = 0;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      int m_idx, n_idx;
      tile_at(local, m_idx, n_idx);
      const int tw = m_idx % a.tiles_w; m_idx /= a.tiles_w;
      const int th = m_idx % a.tiles_h; m_idx /= a.tiles_h;
      const int tt = m_idx % a.tiles_t;
      const int b = m_idx / a.tiles_t;
      EpiPos pos;
      pos.w = (tw << a.lbw) + (row & (bw - 1));
      pos.h = (th << a.lbh) + ((row >> a.lbw) & (bh - 1));
      const int t = (tt << a.lbt) + (row >> (a.lbw + a.lbh));
      pos.valid = (pos.w < a.Wo) && (pos.h < a.Ho) && (t < a.T);
      pos.frame = static_cast<long long>(b) * a.T + t;
      pos.pix = (pos.frame * a.Ho + pos.h) * a.Wo + pos.w;
      pos.n0 = n_idx * a.n_tile;
      // bias slice of this N tile -> shared, zero beyond Cout.  Restaged only when the N tile changes
      // (never, for one N tile or resident weights); two buffers so a restage cannot race the warps that
      // are still finishing the previous tile.
      if (n_idx != bias_n) {
        bias_buf ^= 1;
        bias_n = n_idx;
        if (etid < a.n_tile) {
          const int n = pos.n0 + etid;
          s_bias[bias_buf][etid] = (a.bias != nullptr && n < a.Cout) ? __ldg(a.bias + n) : 0.0f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      const float* sb = s_bias[bias_buf];

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (etid == 0 && local < 4) trace_mark(a.trace, 9 + local);  // accumulator of tile `local` complete
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc * a.acc_cols) +
                              (static_cast<uint32_t>(quarter * 32) << 16);
      switch (kind * 4 + a.act) {
#define EPI_CASE(K, A) case (K) * 4 + (A): epilogue_cols<K, A>(a, t_addr, col_begin, col_end, pos, sb); break;
        EPI_CASE(0, 0) EPI_CASE(0, 1) EPI_CASE(0, 2) EPI_CASE(0, 3)
        EPI_CASE(1, 0) EPI_CASE(1, 1) EPI_CASE(1, 2) EPI_CASE(1, 3)
        EPI_CASE(2, 0) EPI_CASE(2, 1) EPI_CASE(2, 2) EPI_CASE(2, 3)
        EPI_CASE(3, 0) EPI_CASE(3, 1) EPI_CASE(3, 2) EPI_CASE(3, 3)
#undef EPI_CASE
        default: break;
      }
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  if (threadIdx.x == 64) trace_mark(a.trace, 13);  // first epilogue thread done with all its tiles
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(a.trace, 14);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(a.tmem_cols));
  }
}

// FLAIR_CONV_MODE = 0 | 1 | 2 forces the tiling mode ceiling (debug / A-B measurements); default: auto
int flair_conv_mode_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("FLAIR_CONV_MODE");
    v = e ? atoi(e) : -1;
  }
  return v;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << (l + 1)) <= v) ++l;
  return l;
}
int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

extern "C" int flair_debug_conv_trace(long long* host_out16) {
  FLAIR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out16, g_conv_trace, sizeof(long long) * 16));
  return 0;
}

extern "C" int flair_conv_igemm(const flair_conv_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(p != nullptr, "flair_conv_igemm: null params");
  FLAIR_REQUIRE(p->x && p->wgt && p->out, "flair_conv_igemm: null tensor pointer");
  FLAIR_REQUIRE(p->B > 0 && p->T > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0,
                "flair_conv_igemm: bad extents B=%d T=%d H=%d W=%d Cin=%d Cout=%d", p->B, p->T,
                p->H, p->W, p->Cin, p->Cout);
  FLAIR_REQUIRE((p->kt == 1 || p->kt == 3) && (p->kh == 1 || p->kh == 3) &&
                    (p->kw == 1 || p->kw == 3),
                "flair_conv_igemm: kernel extents must be 1 or 3 (got %d,%d,%d)", p->kt, p->kh,
                p->kw);
  FLAIR_REQUIRE(p->stride_hw == 1 || p->stride_hw == 2, "flair_conv_igemm: stride must be 1 or 2");
  FLAIR_REQUIRE(p->x_cstride % 8 == 0 && p->x_cstride >= p->Cin,
                "flair_conv_igemm: x_cstride=%d must be a multiple of 8 and >= Cin=%d",
                p->x_cstride, p->Cin);
  FLAIR_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(p->wgt) & 15) == 0,
                "flair_conv_igemm: x / wgt must be 16-byte aligned");
  FLAIR_REQUIRE(p->in_dtype == FLAIR_BF16 || p->in_dtype == FLAIR_F16,
                "flair_conv_igemm: operands must be bf16 or fp16");
  if (p->out_layout == FLAIR_OUT_NCHW)
    FLAIR_REQUIRE(p->out_dtype == FLAIR_F32, "flair_conv_igemm: NCHW output is fp32 only");
  else
    FLAIR_REQUIRE(p->out_cstride >= p->Cout, "flair_conv_igemm: out_cstride < Cout");
  FLAIR_REQUIRE(p->gn_partial == nullptr, "flair_conv_igemm: fused GN partials not enabled yet");

  const int s = p->stride_hw;
  const int Ho = (p->H + s - 1) / s, Wo = (p->W + s - 1) / s;
  const int Cin_pad = ceil_div(p->Cin, kBlockK) * kBlockK;
  const int Cout_pad = ceil_div(p->Cout, 16) * 16;

  ConvKArgs a{};
  a.B = p->B; a.T = p->T; a.Ho = Ho; a.Wo = Wo;
  // mode selection (see ConvKArgs): halo tiles need a 3x3 spatial kernel, stride 1, and >= 16x8 maps
  const int env_mode = flair_conv_mode_override();
  int mode = 0;
  if (p->kh == 3 && p->kw == 3 && s == 1 && Wo >= 16 && Ho >= 8 && env_mode != 0) mode = 1;
  // tile box: (mode 0) as wide as possible, then tall, then across frames; (halo) 16 x 8 in one frame
  int bw = pow2_ceil(Wo); if (bw > kBlockM) bw = kBlockM;
  int bh = pow2_ceil(Ho); if (bh > kBlockM / bw) bh = kBlockM / bw;
  int bt = kBlockM / (bw * bh);
  if (mode != 0) { bw = 16; bh = 8; bt = 1; }
  a.lbw = ilog2(bw); a.lbh = ilog2(bh); a.lbt = ilog2(bt);
  a.tiles_w = ceil_div(Wo, bw); a.tiles_h = ceil_div(Ho, bh); a.tiles_t = ceil_div(p->T, bt);
  a.m_tiles = a.tiles_w * a.tiles_h * a.tiles_t * p->B;
  int n_tile = Cout_pad;
  const int n_cap = (mode != 0) ? 144 : 256;  // halo stages carry three weight slabs (>= 3 stages must fit)
  if (n_tile > n_cap) {
    n_tile = n_cap;
    while (Cout_pad % n_tile != 0) n_tile -= 16;
  }
  const int ntaps_total = p->kt * p->kh * p->kw;
  const int kblocks_ = Cin_pad / kBlockK;
  if (mode == 1 && env_mode != 1) {
    // resident weights: all slabs of one N tile (try the whole Cout first, then 64 columns) + >= 3 A stages
    const int budget = 227 * 1024 - 1024 - 256 - 2048 - 3 * 20480;
    // (shrinking the N tile to make the weights fit costs more in re-read A than residency saves:
    //  128->128@128x128: 77 us resident/n64 vs 55 us streamed/n128)
    int cand[1] = {n_tile};
    for (int ci = 0; ci < 1; ++ci) {
      const int nt_ = cand[ci];
      if (nt_ > n_tile || Cout_pad % nt_ != 0) continue;
      if (static_cast<long long>(ntaps_total) * kblocks_ * nt_ * 128 <= budget) { mode = 2; n_tile = nt_; break; }
    }
  }
  a.n_tile = n_tile;
  a.n_tiles = Cout_pad / n_tile;
  // the epilogue reads TMEM 32 columns at a time: leave 16 columns of slack per accumulator buffer
  // (n_tile = 256 splits into two 128-column halves and needs none)
  a.tmem_cols = (n_tile == 256) ? 512 : pow2_ceil(2 * (n_tile + 16));
  a.acc_cols = a.tmem_cols / 2;
  a.Cout = p->Cout; a.Cout_pad = Cout_pad;
  a.kblocks = Cin_pad / kBlockK;
  a.k_last = ceil_div(p->Cin - (a.kblocks - 1) * kBlockK, kUmmaK);
  {
    static int trace = -1;
    if (trace < 0) { const char* e = getenv("FLAIR_CONV_TRACE"); trace = (e && e[0] == '1') ? 1 : 0; }
    a.trace = trace;
  }
  a.stride = s;
  int nt = 0;
  for (int dt = -(p->kt / 2); dt <= p->kt / 2; ++dt)
    for (int dh = -(p->kh / 2); dh <= p->kh / 2; ++dh)
      for (int dw = -(p->kw / 2); dw <= p->kw / 2; ++dw) {
        a.tap_dt[nt] = static_cast<int8_t>(dt);
        a.tap_dh[nt] = static_cast<int8_t>(dh);
        a.tap_dw[nt] = static_cast<int8_t>(dw);
        ++nt;
      }
  a.ntaps = nt;
  a.bias = p->bias; a.rowbias = p->rowbias; a.rowbias_stride = p->rowbias_stride;
  a.rowscale = p->rowscale; a.rowscale_stride = p->rowscale_stride;
  a.preadd = p->preadd; a.preadd_dtype = p->preadd_dtype; a.preadd_cstride = p->preadd_cstride;
  if (p->preadd != nullptr)
    FLAIR_REQUIRE(p->preadd_cstride >= p->Cout && (p->preadd_dtype == FLAIR_F32 || p->preadd_dtype == FLAIR_F16 ||
                                                   p->preadd_dtype == FLAIR_BF16),
                  "flair_conv_igemm: bad preadd geometry / dtype");
  a.residual = p->residual; a.residual_dtype = p->residual_dtype;
  a.residual_cstride = p->residual_cstride;
  a.residual2 = p->residual2; a.residual2_dtype = p->residual2_dtype;
  a.residual2_cstride = p->residual2_cstride;
  a.out = p->out; a.out_dtype = p->out_dtype; a.out_layout = p->out_layout;
  a.out_cstride = p->out_cstride; a.act = p->act;
  a.out_scale = (p->out_scale == 0.0f) ? 1.0f : p->out_scale;
  a.fmt = (p->in_dtype == FLAIR_BF16) ? 1u : 0u;
  a.gn_partial = nullptr; a.gn_groups = 0;
  a.out2 = static_cast<uint16_t*>(p->out2); a.out2_gs = p->out2_group_channels; a.out2_gstride = p->out2_group_stride;
  if (p->out2 != nullptr)
    FLAIR_REQUIRE(p->out_layout == FLAIR_OUT_NHWC && p->out_dtype != FLAIR_F32 && p->Cout % 16 == 0 &&
                      (p->out2_group_channels == 8 || p->out2_group_channels == 16) && p->out2_group_stride % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(p->out2) & 15) == 0,
                  "flair_conv_igemm: out2 needs a 16-bit NHWC output, Cout %% 16 == 0, 8 or 16 channels per group");

  const uint32_t b_bytes = static_cast<uint32_t>(n_tile) * kBlockK * 2;  // multiple of 2 KB (n_tile % 16 == 0)
  a.mode = mode;
  a.ntd = p->kt;
  a.b_bytes = b_bytes;
  a.a_bytes = (mode == 0) ? kABytes : 160u * 128u;  // 16 x 10 halo rows of 128 B
  a.w_bytes = (mode == 2) ? static_cast<uint32_t>(ntaps_total) * a.kblocks * b_bytes : 0u;
  uint32_t stage_bytes = a.a_bytes + ((mode == 0) ? b_bytes : (mode == 1 ? 3u * b_bytes : 0u));
  stage_bytes = (stage_bytes + 1023u) & ~1023u;
  a.stage_bytes = stage_bytes;
  const int smem_budget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - 2048 /*static bias*/ - static_cast<int>(a.w_bytes);
  int stages = smem_budget / static_cast<int>(stage_bytes);
  if (stages > 8) stages = 8;
  FLAIR_REQUIRE(stages >= 2, "flair_conv_igemm: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem_bytes = a.w_bytes + static_cast<size_t>(stages) * stage_bytes + 1024 + 256;

  flair_tmap_encode_fn encode = flair_get_tmap_encode();
  FLAIR_REQUIRE(encode != nullptr, "flair_conv_igemm: cuTensorMapEncodeTiled unavailable");
  const CUtensorMapDataType dt16 =
      (p->in_dtype == FLAIR_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(p->Cin), static_cast<cuuint64_t>(p->W),
                          static_cast<cuuint64_t>(p->H), static_cast<cuuint64_t>(p->T),
                          static_cast<cuuint64_t>(p->B)};
    const cuuint64_t px = static_cast<cuuint64_t>(p->x_cstride) * 2;
    cuuint64_t strides[4] = {px, px * p->W, px * p->W * p->H, px * p->W * p->H * p->T};
    cuuint32_t box[5] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(bw * s),
                         static_cast<cuuint32_t>((mode != 0 ? bh + 2 : bh) * s), static_cast<cuuint32_t>(bt), 1u};
    if (s == 2) { box[1] -= 1; box[2] -= 1; }  // ceil(box/stride) == bw, no overreach
    cuuint32_t estr[5] = {1u, static_cast<cuuint32_t>(s), static_cast<cuuint32_t>(s), 1u, 1u};
    CUresult r = encode(&tmA, dt16, 5, const_cast<void*>(p->x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLAIR_REQUIRE(r == CUDA_SUCCESS, "flair_conv_igemm: activation tensor map rejected (%d)",
                  static_cast<int>(r));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(Cin_pad),
                          static_cast<cuuint64_t>(nt) * static_cast<cuuint64_t>(Cout_pad)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(Cin_pad) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(n_tile)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&tmB, dt16, 2, const_cast<void*>(p->wgt), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FLAIR_REQUIRE(r == CUDA_SUCCESS, "flair_conv_igemm: weight tensor map rejected (%d)",
                  static_cast<int>(r));
  }

  static FlairPerDeviceOnce attr_once;
  if (attr_once.first())
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  const int total_tiles = a.m_tiles * a.n_tiles;
  int grid = flair_num_sms();
  if (grid > total_tiles) grid = total_tiles;
  if (mode == 2) {  // every CTA keeps one N tile: grid must be a multiple of n_tiles
    grid = (grid / a.n_tiles) * a.n_tiles;
    if (grid < a.n_tiles) grid = a.n_tiles;
  }
  FLAIR_CHECK_CUDA(flair_launch(conv_igemm_kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, a));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
