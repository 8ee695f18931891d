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
// Roles (320 threads): warps 0..7 = epilogue (two per TMEM lane quarter, splitting the columns), warp 8 = TMA
// producer, warp 9 = MMA issuer (+TMEM alloc) — the two single-thread issuers carry the highest warp ids of their
// scheduler sub-partitions, see kWarpTma.  The epilogue
// is specialised at compile time on (output type/layout, activation): a runtime-generic version
// cost ~6k warp-instructions per tile and left the tensor pipe 6 % busy (profiles/r01_conv64_*).  The kernel is
// persistent: grid = min(#tiles, #SMs); accumulators are double-buffered in
// TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// 3x3 spatial kernels (stride 1, maps of at least 8 x 16 pixels) use the HALO mode: the output tile is 8 (w) x 16 (h)
// pixels and ONE 10 x 18-pixel halo slab per (dt, 64-channel block) feeds all nine (dh, dw) taps.  The slab lands
// row-major (180 rows of 128 B, SWIZZLE_128B); the A operand of tap (dh, dw) is the same slab described with start
// address + ((dh*10 + dw) rows) and a stride of 10 rows (1280 B) between the 8-row groups of the 128 tile rows:
// the 128-byte swizzle is a function of the physical shared-memory address only, so row-shifted starts and a
// non-1024 B group stride read exactly what TMA wrote (verified on B200: tests/gpu_probes/cu/desc_sbo.cu).  Each input
// pixel therefore crosses L2 -> SM 1.4 times per 64-channel block (round 1: 3.75 times, one 16 x 10 slab per dw).
// Weights of the CTA's N tile stay resident in shared memory when they fit; otherwise they stream through their
// own ring, one (tap, 64-channel block) slab per stage.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 x 16-bit = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;                  // two per TMEM lane quarter (columns split in halves)
constexpr int kThreads = 32 * kEpiWarps + 64;  // 8 epilogue warps, then the TMA warp and the MMA warp
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
  int debug;   // FLAIR_CONV_DEBUG bit mask (perf experiments only; results are wrong): 1 = epilogue does nothing,
               // 2 = epilogue reads TMEM but stores nothing, 4 = producer signals the stages without loading A,
               // 8 = the MMA warp does not wait for the accumulator buffer
  int k_last;  // UMMA K=16 steps that hold real channels in the LAST k-block (Cin = 196: 1 of 4)
  int stages;
  // mode 0: one (tap, k-block) per pipeline stage, A tile = 128 output pixels (any kernel / stride / map size).
  // mode 3: halo — 8 x 16-pixel tiles, one 10 x 18 halo slab of A per (dt, k-block) stage feeds the nine spatial
  //         taps; `resident` != 0: this CTA's weights stay in shared memory, else they stream through a second ring
  //         (b_stages slabs of one (tap, k-block) each).
  int mode, ntd, resident, b_stages;
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
  float* gn_partial;       // fused GroupNorm statistics of the OUTPUT: [m_tile][4 lane quarters][Cout_pad/16][16] floats
  int gn_groups;
  int gn_cpg;              // channels per group (power of two >= 2)
  int gn_nchunks;          // Cout_pad / 16
  int fast;                // compact epilogue applies (see epilogue_fast): 1 = no addends, 2 = addends prefetched
  uint16_t* out2;          // optional copy of a 16-bit NHWC output as pair planes [Cout/out2_gs][pixels][2][out2_gs]:
  int out2_gs;             // entry p = (pixel p, pixel p+1), the source layout of flair_deform_conv
  long long out2_gstride;  // elements between group planes
  long long out2_back;     // elements from slot 0 of entry pix back to slot 1 of entry pix - neighbor
  int out2_nb;             // neighbor distance in pixels: 1 (x, x+1 pairs) or W (y, y+1 pairs)
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
        uint4 uu[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uu[q].x = pack16(v[8 * q + 0], v[8 * q + 1], DT);
          uu[q].y = pack16(v[8 * q + 2], v[8 * q + 3], DT);
          uu[q].z = pack16(v[8 * q + 4], v[8 * q + 5], DT);
          uu[q].w = pack16(v[8 * q + 6], v[8 * q + 7], DT);
        }
        // the 16 columns of a chunk are one 32-byte sector: write it with ONE store when aligned (two 16-byte stores
        // were two partial-sector transactions each: ncu counted twice the sectors of the tensor at L1 and at L2)
        if ((reinterpret_cast<uintptr_t>(op) & 31) == 0) {
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(uu[0].x), "r"(uu[0].y),
                       "r"(uu[0].z), "r"(uu[0].w), "r"(uu[1].x), "r"(uu[1].y), "r"(uu[1].z), "r"(uu[1].w)
                       : "memory");
        } else {
          reinterpret_cast<uint4*>(op)[0] = uu[0];
          reinterpret_cast<uint4*>(op)[1] = uu[1];
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 u = uu[q];
          if (a.out2 != nullptr) {  // pair planes for the deformable gather: slot 0 of entry pix, slot 1 of entry pix-1
            const int ch = n + 8 * q;
            uint16_t* e = a.out2 + (ch / a.out2_gs) * a.out2_gstride + pos.pix * (2 * a.out2_gs) + (ch % a.out2_gs);
            *reinterpret_cast<uint4*>(e) = u;
            if (pos.pix >= a.out2_nb) *reinterpret_cast<uint4*>(e - a.out2_back) = u;
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

// ---------------------------------------------------------------------------------------------------------
// Compact epilogue for the common case: 16-bit NHWC output, Cout % 16 == 0, 32-byte aligned rows, no per-frame
// bias / gate.  The generic path above is ~2700 instructions per (output kind, activation) instance with all its
// partial-chunk and layout branches; on N = 64 tiles (every 64-channel conv of the 256 x 256 maps) the tensor pipe
// needs only ~1.5 us per tile and that epilogue did not fit behind it (profiles/r02_summary.md: 61 us per launch,
// 50 us with the epilogue's arithmetic and stores removed).  Here everything that varies per launch is a warp-uniform
// branch around a 16-element loop, and the 16-bit addends (preadd / residual / residual2) are loaded BEFORE the wait
// for the accumulator, so their L2 latency overlaps the tile's MMAs.
// ---------------------------------------------------------------------------------------------------------
constexpr int kPfChunks = 4;   // addend prefetch covers up to 4 chunks (64 columns) per warp, i.e. N tiles <= 128
struct EpiPrefetch {
  uint4 a[kPfChunks][2], b[kPfChunks][2];   // slot A = preadd, or residual2 when there is no preadd; slot B = residual
};

__device__ __forceinline__ void add16(float (&v)[16], const uint4 (&u2)[2], bool f16) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t uu[4] = {u2[q].x, u2[q].y, u2[q].z, u2[q].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f;
      if (f16) f = __half22float2(*reinterpret_cast<const __half2*>(&uu[e]));
      else f = unpack_bf16x2(uu[e]);
      v[8 * q + 2 * e] += f.x;
      v[8 * q + 2 * e + 1] += f.y;
    }
  }
}

// Sum N values per lane over the 32 lanes of a warp with N - 1 + (5 - log2 N) shuffles (halving butterfly): at every step
// a lane keeps one half of its values and receives the partner's copy of that half.  Afterwards lane l holds the total
// of value index (l >> (5 - log2 N)) in x[0].  Fixed order -> deterministic.
template <int N>
__device__ __forceinline__ float warp_sum_n(float (&x)[N], int lane) {
  int n = N;
#pragma unroll
  for (int mask = 16; mask >= 1; mask >>= 1) {
    if (n > 1) {
      const int half = n >> 1;
      const bool up = (lane & mask) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < half) {
          const float send = up ? x[i] : x[i + half];
          const float recv = __shfl_xor_sync(0xffffffffu, send, mask);
          x[i] = (up ? x[i + half] : x[i]) + recv;
        }
      }
      n = half;
    } else {
      x[0] += __shfl_xor_sync(0xffffffffu, x[0], mask);
    }
  }
  return x[0];
}

// GroupNorm statistics of one stored 16-channel chunk (values as rounded to the 16-bit output, zero for rows outside
// the map): per group of `cpg` channels inside the chunk (sum, sum of squares) over the warp's 32 pixels ->
// slot[2*g], slot[2*g+1] (cpg > 16: the chunk is part of one group -> slot[0], slot[1]).
template <int NG>
__device__ __forceinline__ void gn_chunk_stats(const float (&vr)[16], float* __restrict__ slot, int lane) {
  constexpr int CPG = 16 / NG;
  float x[2 * NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = vr[g * CPG + j];
      sm += t;
      sq = fmaf(t, t, sq);
    }
    x[2 * g] = sm;
    x[2 * g + 1] = sq;
  }
  const float tot = warp_sum_n<2 * NG>(x, lane);
  constexpr int SH = (NG == 8) ? 1 : (NG == 4) ? 2 : (NG == 2) ? 3 : 4;   // lane l holds value index l >> SH
  if ((lane & ((1 << SH) - 1)) == 0) slot[lane >> SH] = tot;
}

// one 16-column chunk: bias, [preadd], activation, [scale], [residuals], pack, one 32-byte store (+ pair planes)
// mask: bit 0 preadd (slot A), bit 1 residual (slot B), bit 2 residual2 (slot A)
template <bool GN>
__device__ __forceinline__ void fast_chunk(const ConvKArgs& a, const uint32_t (&r)[16], int c0, int n, long long pix,
                                           bool valid, const float* __restrict__ sb, const uint4 (&pa)[2],
                                           const uint4 (&pb)[2], int mask, bool f16, float* __restrict__ gn_tile) {
  float v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 bq = *reinterpret_cast<const float4*>(sb + c0 + 4 * q);
    v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
    v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
    v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
    v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
  }
  if (mask & 1) add16(v, pa, f16);
  if (a.act == FLAIR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (a.act == FLAIR_ACT_LRELU01) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);   // == v > 0 ? v : 0.1 v
  } else if (a.act == FLAIR_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
  }
  if (a.out_scale != 1.0f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= a.out_scale;
  }
  if (mask & 2) add16(v, pb, f16);
  if (mask & 4) add16(v, pa, f16);
  uint32_t u[8];
  if (f16) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      u[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
  }
  if (GN) {
    // fused GroupNorm statistics of the tensor being stored (a separate instantiation: with this code in the plain
    // path every N = 64 launch was 8 % slower, profiles/r02_summary.md) (nn_new.py:17-19 computes them on the stored 16-bit map)
    float vr[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float2 f;
      if (f16) f = __half22float2(*reinterpret_cast<const __half2*>(&u[j]));
      else f = unpack_bf16x2(u[j]);
      vr[2 * j] = valid ? f.x : 0.f;
      vr[2 * j + 1] = valid ? f.y : 0.f;
    }
    float* slot = gn_tile + (n >> 4) * 16;
    const int lane = threadIdx.x & 31;
    if (a.gn_cpg == 2) gn_chunk_stats<8>(vr, slot, lane);
    else if (a.gn_cpg == 4) gn_chunk_stats<4>(vr, slot, lane);
    else if (a.gn_cpg == 8) gn_chunk_stats<2>(vr, slot, lane);
    else gn_chunk_stats<1>(vr, slot, lane);
  }
  if (!valid) return;
  uint16_t* op = static_cast<uint16_t*>(a.out) + pix * a.out_cstride + n;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
               "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
  if (a.out2 != nullptr) {  // pair planes for the deformable gather: slot 0 of entry pix, slot 1 of entry pix-1
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int ch = n + 8 * q;
      uint16_t* e = a.out2 + (ch / a.out2_gs) * a.out2_gstride + pix * (2 * a.out2_gs) + (ch % a.out2_gs);
      const uint4 uq = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
      *reinterpret_cast<uint4*>(e) = uq;
      if (pix >= a.out2_nb) *reinterpret_cast<uint4*>(e - a.out2_back) = uq;
    }
  }
}

// PF: addends are in registers (<= 4 chunks per warp, chunk index compile-time)
template <bool PF, bool GN>
__device__ __forceinline__ void epilogue_fast(const ConvKArgs& a, uint32_t t_addr, int col_begin, int col_end, int n0,
                                              long long pix, bool valid, const float* __restrict__ sb,
                                              const EpiPrefetch& pf, int mask, bool f16, float* __restrict__ gn_tile) {
  if (PF) {
#pragma unroll
    for (int cp = 0; cp < kPfChunks / 2; ++cp) {
      const int cc = col_begin + 32 * cp;
      if (cc < col_end) {  // warp-uniform
        uint32_t r0[16], r1[16];
        __syncwarp();
        tmem_ld16(t_addr + static_cast<uint32_t>(cc), r0);
        tmem_ld16(t_addr + static_cast<uint32_t>(cc + 16), r1);
        tmem_ld_wait();
        fast_chunk<GN>(a, r0, cc, n0 + cc, pix, valid, sb, pf.a[2 * cp], pf.b[2 * cp], mask, f16, gn_tile);
        if (cc + 16 < col_end) fast_chunk<GN>(a, r1, cc + 16, n0 + cc + 16, pix, valid, sb, pf.a[2 * cp + 1], pf.b[2 * cp + 1], mask, f16, gn_tile);
      }
    }
  } else {
    const uint4 none[2] = {};
#pragma unroll 1
    for (int cc = col_begin; cc < col_end; cc += 32) {
      uint32_t r0[16], r1[16];
      __syncwarp();
      tmem_ld16(t_addr + static_cast<uint32_t>(cc), r0);
      tmem_ld16(t_addr + static_cast<uint32_t>(cc + 16), r1);
      tmem_ld_wait();
      fast_chunk<GN>(a, r0, cc, n0 + cc, pix, valid, sb, none, none, 0, f16, gn_tile);
      if (cc + 16 < col_end) fast_chunk<GN>(a, r1, cc + 16, n0 + cc + 16, pix, valid, sb, none, none, 0, f16, gn_tile);
    }
  }
}

// Three taps (one dh row: dw = 0, 1, 2) of the halo mode, issued by the elected lane.  Everything that varies is an
// immediate added to two uniform base registers, so an MMA costs the issuing thread a handful of uniform-datapath
// instructions: measured on B200 the tensor pipe needs 48 (N = 64) .. 128 (N = 256) cycles per K step, and a single
// thread that spends more than that per issue (the round-1 loop: ~85 cycles with its per-tap waits, elects and
// register->uniform moves) becomes the bottleneck of the whole kernel (profiles/r02_summary.md).
//   alo: descriptor low word of the slab start for (dh, dw = 0);  blo[j]: low word of the weight slab of tap (dh, j)
template <int KSTEPS>
__device__ __forceinline__ void issue_tap_row(uint32_t d_tmem, uint64_t hi_a, uint64_t hi_b, uint32_t alo, uint32_t blo0,
                                              uint32_t blo1, uint32_t blo2, uint32_t idesc, uint32_t accum_first, int ksteps) {
  const uint32_t bl[3] = {blo0, blo1, blo2};
#pragma unroll
  for (int dwi = 0; dwi < 3; ++dwi) {
    const uint32_t al = alo + static_cast<uint32_t>(dwi) * 8u;  // + dw rows of 128 B (>> 4)
    if (KSTEPS == 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(d_tmem, hi_a | (al + 2u * k), hi_b | (bl[dwi] + 2u * k), idesc, (dwi == 0 && k == 0) ? accum_first : 1u);
    } else {
      for (int k = 0; k < ksteps; ++k)
        umma_f16(d_tmem, hi_a | (al + 2u * k), hi_b | (bl[dwi] + 2u * k), idesc, (dwi == 0 && k == 0) ? accum_first : 1u);
    }
  }
}

constexpr uint32_t kSlabW = 10, kSlabH = 18;                      // halo slab of an 8 (w) x 16 (h) tile
constexpr uint32_t kSlabBytes = kSlabW * kSlabH * 128u;            // 180 rows of 128 B
constexpr uint32_t kSlabStage = (kSlabBytes + 1023u) & ~1023u;     // stage pitch (1024-B aligned for the swizzle)
constexpr int kMaxStages = 8;
// Warp roles.  The warp scheduler of an SM sub-partition prefers the HIGHEST warp id among its eligible warps
// (B300_MICROARCH.md, "arbiter priority"); warp w lives on sub-partition w % 4.  The two latency-critical
// single-thread issuers therefore get the highest ids of their sub-partitions (8 -> sub-partition 0, 9 -> 1), and
// the eight epilogue warps 0..7 keep warp % 4 = TMEM lane quarter.  With the issuers as warps 0 / 1 the bursts of
// the epilogue warps 4, 8 / 5, 9 pre-empted them and the tensor pipe drained (round 2: -30 % on N = 64 launches).
constexpr int kWarpTma = 8, kWarpMma = 9;

// EPI: which epilogues are compiled into the instance.  0 = all (generic (kind x activation) template + compact paths),
// 1 = compact paths only, 2 = compact paths with the fused GroupNorm statistics only.  The generic template is ~43 K of
// the kernel's ~50 K instructions; a per-frame launch fetches its code cold (conv, deformable conv and warp launches
// alternate on the BasicVSR++ chain), so the 1300 per-frame launches of a forward run a 4-6 K-instruction instance.
template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvKArgs a) {
  // PDL: let the next launch's CTAs be scheduled as soon as SMs free up.  Nothing before pdl_wait() below touches
  // memory written by an earlier kernel: barrier init, TMEM allocation, descriptor prefetch and the loads of this
  // CTA's resident weights (constant since model load) all overlap the previous kernel's tail.
  pdl_trigger();
  if (threadIdx.x == 32 * kWarpTma) trace_mark(a.trace, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float s_bias[2][256];
  // carve: [resident weights | weight ring (halo mode)] [A ring (mode 0: A + B per stage)] [barriers]
  const uint32_t b_bytes = a.b_bytes;
  const uint32_t stage_bytes = a.stage_bytes;
  uint8_t* smem_w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
  uint8_t* smem = smem_w + a.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(a.stages) * stage_bytes);
  uint64_t* full_bar = bars;                      // [kMaxStages] A ring (mode 0: A + B)
  uint64_t* empty_bar = bars + kMaxStages;        // [kMaxStages]
  uint64_t* bfull_bar = bars + 2 * kMaxStages;    // [kMaxStages] streamed-weight ring (halo mode)
  uint64_t* bempty_bar = bars + 3 * kMaxStages;   // [kMaxStages]
  uint64_t* tfull_bar = bars + 4 * kMaxStages;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2]
  uint64_t* w_bar = tempty_bar + 2;               // [9]: resident weights arrive per (dt, dh) group, in MMA order
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    for (int g = 0; g < 9; ++g) mbar_init(&w_bar[g], 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(a.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 32 * kWarpTma) trace_mark(a.trace, 1);
  // (pdl_wait() is executed per role below: the MMA warp first issues the weight loads)

  // tile walk of this CTA: (m_idx, n_idx) = f(local iteration)
  //   streamed weights: linear tile id = blockIdx.x + i*gridDim.x, n fastest (neighbouring CTAs share A in L2)
  //   resident weights: n_idx fixed per CTA, m strided by gridDim.x / n_tiles
  const int total_tiles = a.m_tiles * a.n_tiles;
  const bool halo = (a.mode == 3);
  const bool resident = halo && a.resident != 0;
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
  uint8_t* smem_bring = smem_w;  // streamed-weight ring lives where the resident weights would (w_bytes covers either)

  if (warp == kWarpTma) {
    // ===================== TMA producer =====================
    // elect.sync instead of `lane == 0`: the compiler then keeps coordinates / addresses in uniform registers and
    // emits a bare UTMALDG; under `lane == 0` every TMA issue was a waterfall loop (ELECT + 7 R2UR.BROADCAST +
    // branch), ~300 cycles per issue.
    pdl_wait();  // activations are the previous kernel's output
    if (elect_one()) {
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
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
        if (!halo) {
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
            for (int kb = 0; kb < a.kblocks; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (a.debug & 4) {
                mbar_arrive(&full_bar[stage]);
              } else {
              mbar_expect_tx(&full_bar[stage], kSlabBytes);
              // 10 x 18 halo slab: columns w0-1 .. w0+8, rows h0-1 .. h0+16 (out-of-bounds = zero padding)
              tma_load_5d(smem + static_cast<size_t>(stage) * stage_bytes, &tmA, &full_bar[stage], kb * kBlockK,
                          w0 - 1, h0 - 1, ct, b);
              }
              if (++stage == a.stages) { stage = 0; phase ^= 1; }
              if (!resident) {
                for (int tap9 = 0; tap9 < 9; ++tap9) {
                  mbar_wait(&bempty_bar[bstage], bphase ^ 1);
                  mbar_expect_tx(&bfull_bar[bstage], b_bytes);
                  tma_load_2d(smem_bring + static_cast<size_t>(bstage) * b_bytes, &tmB, &bfull_bar[bstage],
                              kb * kBlockK, (dti * 9 + tap9) * a.Cout_pad + n_idx * a.n_tile);
                  if (++bstage == a.b_stages) { bstage = 0; bphase ^= 1; }
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    // One ELECTED lane issues (elect.sync: the compiler then knows the region is single-threaded and emits
    // UTCHMMA + 2 UMOV per MMA; under `if (lane == 0)` it wrapped every MMA in a 10-instruction waterfall loop).
    // Descriptors are (constant high word, low word = smem address >> 4): the loop only adds small constants to
    // the low words.
    const uint32_t idesc = umma_idesc_f16(kBlockM, static_cast<uint32_t>(a.n_tile), a.fmt);
    int stage = 0, bstage = 0;
    uint32_t phase = 0, bphase = 0;
    if (resident && elect_one()) {
      // this CTA's weights: every (tap, k-block) slab of its N tile, once, grouped per (dt, dh) in MMA order so the
      // first MMAs start after 3 * kblocks slabs, not after the whole filter.  Issued from this (otherwise idle) warp
      // while warp 0 issues the activation loads, and BEFORE pdl_wait: weights do not depend on the previous kernel.
      const uint32_t group_bytes = 3u * static_cast<uint32_t>(a.kblocks) * b_bytes;
      for (int dti = 0; dti < a.ntd; ++dti)
        for (int dhi = 0; dhi < 3; ++dhi) {
          uint64_t* bar = &w_bar[dti * 3 + dhi];
          mbar_expect_tx(bar, group_bytes);
          for (int kb = 0; kb < a.kblocks; ++kb)
            for (int dwi = 0; dwi < 3; ++dwi) {
              const int tap = (dti * 3 + dhi) * 3 + dwi;
              tma_load_2d(smem_w + static_cast<size_t>(tap * a.kblocks + kb) * b_bytes, &tmB, bar, kb * kBlockK,
                          tap * a.Cout_pad + my_n * a.n_tile);
            }
        }
    }
    __syncwarp();
    pdl_wait();
    if (lane == 0) trace_mark(a.trace, 4);
    const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;   // SBO = 1024 B, version, swizzle mode
    // halo slab: the 8-row groups of the 128 tile rows are 10 slab rows (1280 B) apart
    const uint64_t desc_hi_slab = (desc_hi & ~(static_cast<uint64_t>(0x3FFF) << 32)) |
                                  (static_cast<uint64_t>((kSlabW * 128u) >> 4) << 32);
    const uint32_t desc_lo_flags = static_cast<uint32_t>(umma_desc_sw128(0));  // LBO field
    const uint32_t stage0_lo = (smem_u32(smem) & 0x3FFFF) >> 4;
    const uint32_t stage_step = stage_bytes >> 4;
    const uint32_t a_step = a.a_bytes >> 4;
    const uint32_t b_step = b_bytes >> 4;
    const uint32_t w_lo = (smem_u32(smem_w) & 0x3FFFF) >> 4;
    const uint32_t w_tap_step = static_cast<uint32_t>(a.kblocks) * b_step;  // resident weights: [tap][kb]
    const int k_iters = halo ? a.ntd * a.kblocks : a.ntaps * a.kblocks;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      if (!(a.debug & 8)) mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * a.acc_cols);
      if (lane == 0 && local < 4) trace_mark(a.trace, 5 + local);  // MMA of tile `local` may start (TMEM buffer free)
      uint32_t accum = 0;  // first MMA of the tile overwrites the accumulator
      if (!halo) {
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
          for (int kb = 0; kb < a.kblocks; ++kb, ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const int ksteps = (kb == a.kblocks - 1) ? a.k_last : kBlockK / kUmmaK;
            const bool full_k = ksteps == kBlockK / kUmmaK;
            const uint32_t alo = desc_lo_flags | (stage0_lo + static_cast<uint32_t>(stage) * stage_step);
            const bool last = it == k_iters - 1;
            if (resident) {
              // resident weights of tap (dti, 0, 0), this k-block; taps are w_tap_step apart
              const uint32_t wlo0 = desc_lo_flags | (w_lo + static_cast<uint32_t>(dti * 9) * w_tap_step +
                                                     static_cast<uint32_t>(kb) * b_step);
              if (local == 0 && kb == 0) {
                // first tile: the weight groups (one per dh row) are still landing — issue row by row
                for (int dhi = 0; dhi < 3; ++dhi) {
                  mbar_wait(&w_bar[dti * 3 + dhi], 0);
                  if (elect_one()) {
                    const uint32_t b0 = wlo0 + static_cast<uint32_t>(dhi * 3) * w_tap_step;
                    issue_tap_row<0>(d_tmem, desc_hi_slab, desc_hi, alo + static_cast<uint32_t>(dhi) * (kSlabW * 8u), b0,
                                     b0 + w_tap_step, b0 + 2u * w_tap_step, idesc, accum, ksteps);
                    if (dhi == 2) {
                      umma_commit(&empty_bar[stage]);
                      if (last) umma_commit(&tfull_bar[acc]);
                    }
                  }
                  __syncwarp();
                  accum = 1;
                }
              } else {
                if (elect_one()) {
                  uint32_t b0 = wlo0;
#pragma unroll
                  for (int dhi = 0; dhi < 3; ++dhi) {
                    const uint32_t ar = alo + static_cast<uint32_t>(dhi) * (kSlabW * 8u);
                    const uint32_t acc_first = (dhi == 0) ? accum : 1u;
                    if (full_k)
                      issue_tap_row<4>(d_tmem, desc_hi_slab, desc_hi, ar, b0, b0 + w_tap_step, b0 + 2u * w_tap_step, idesc,
                                       acc_first, 4);
                    else
                      issue_tap_row<0>(d_tmem, desc_hi_slab, desc_hi, ar, b0, b0 + w_tap_step, b0 + 2u * w_tap_step, idesc,
                                       acc_first, ksteps);
                    b0 += 3u * w_tap_step;
                  }
                  umma_commit(&empty_bar[stage]);
                  if (last) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                accum = 1;
              }
            } else {
              if (a.b_stages >= 6) {
                // streamed weights, deep ring (small N tiles): one dh row = three consecutive slots per elect region
                for (int dhi = 0; dhi < 3; ++dhi) {
                  int bs[3];
#pragma unroll
                  for (int j = 0; j < 3; ++j) {
                    bs[j] = bstage;
                    mbar_wait(&bfull_bar[bstage], bphase);
                    if (++bstage == a.b_stages) { bstage = 0; bphase ^= 1; }
                  }
                  tc_fence_after();
                  if (elect_one()) {
                    const uint32_t wl = desc_lo_flags | w_lo;
                    const uint32_t ar = alo + static_cast<uint32_t>(dhi) * (kSlabW * 8u);
                    if (full_k)
                      issue_tap_row<4>(d_tmem, desc_hi_slab, desc_hi, ar, wl + static_cast<uint32_t>(bs[0]) * b_step,
                                       wl + static_cast<uint32_t>(bs[1]) * b_step, wl + static_cast<uint32_t>(bs[2]) * b_step,
                                       idesc, accum, 4);
                    else
                      issue_tap_row<0>(d_tmem, desc_hi_slab, desc_hi, ar, wl + static_cast<uint32_t>(bs[0]) * b_step,
                                       wl + static_cast<uint32_t>(bs[1]) * b_step, wl + static_cast<uint32_t>(bs[2]) * b_step,
                                       idesc, accum, ksteps);
                    umma_commit(&bempty_bar[bs[0]]);
                    umma_commit(&bempty_bar[bs[1]]);
                    umma_commit(&bempty_bar[bs[2]]);
                    if (dhi == 2) {
                      umma_commit(&empty_bar[stage]);
                      if (last) umma_commit(&tfull_bar[acc]);
                    }
                  }
                  __syncwarp();
                  accum = 1;
                }
              } else {
                // streamed weights, shallow ring (N tile of 144..256 columns: a tap is >= 4 x 72 cycles of tensor work,
                // the per-tap hand-shake is hidden): one slot per elect region keeps the prefetch distance
                for (int tap9 = 0; tap9 < 9; ++tap9) {
                  mbar_wait(&bfull_bar[bstage], bphase);
                  tc_fence_after();
                  if (elect_one()) {
                    const uint32_t al = alo + static_cast<uint32_t>((tap9 / 3) * static_cast<int>(kSlabW) + tap9 % 3) * 8u;
                    const uint32_t bl = desc_lo_flags | (w_lo + static_cast<uint32_t>(bstage) * b_step);
                    if (full_k) {
#pragma unroll
                      for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        umma_f16(d_tmem, desc_hi_slab | (al + 2u * k), desc_hi | (bl + 2u * k), idesc, accum);
                        accum = 1;
                      }
                    } else {
                      for (int k = 0; k < ksteps; ++k) {
                        umma_f16(d_tmem, desc_hi_slab | (al + 2u * k), desc_hi | (bl + 2u * k), idesc, accum);
                        accum = 1;
                      }
                    }
                    umma_commit(&bempty_bar[bstage]);
                    if (tap9 == 8) {
                      umma_commit(&empty_bar[stage]);
                      if (last) umma_commit(&tfull_bar[acc]);
                    }
                  }
                  __syncwarp();
                  accum = 1;
                  if (++bstage == a.b_stages) { bstage = 0; bphase ^= 1; }
                }
              }
            }
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 0..7) =====================
    pdl_wait();  // residuals / per-frame biases may be the previous kernel's output
    const int ewarp = warp;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int half = ewarp >> 2;           // which half of the accumulator columns
    const int row = quarter * 32 + lane;
    const int bw = 1 << a.lbw, bh = 1 << a.lbh;
    const int col_split = ((a.n_tile / 2 + 15) / 16) * 16;
    const int col_begin = half ? col_split : 0;
    const int col_end = half ? a.n_tile : col_split;
    const int etid = threadIdx.x;          // 0..255
    // epilogue kind: 0 = fp16 NHWC, 1 = bf16 NHWC, 2 = fp32 NHWC, 3 = fp32 NCHW
    const int kind = (a.out_layout == FLAIR_OUT_NCHW) ? 3 : (a.out_dtype == FLAIR_F32 ? 2 : (a.out_dtype == FLAIR_F16 ? 0 : 1));
    int bias_n = -1, bias_buf = 0;
    const int pf_mask = (a.fast == 2) ? ((a.preadd ? 1 : 0) | (a.residual ? 2 : 0) | ((a.residual2 && !a.preadd) ? 4 : 0)) : 0;
    EpiPrefetch pf;
    for (int local = 0; local < my_tiles; ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      int m_idx, n_idx;
      tile_at(local, m_idx, n_idx);
      float* gn_tile = (a.gn_partial != nullptr)
                           ? a.gn_partial + (static_cast<long long>(m_idx) * 4 + quarter) * a.gn_nchunks * 16 : nullptr;
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

      if (a.fast == 2 && pos.valid) {   // addends of this warp's chunks -> registers, before the accumulator wait
        const uint16_t* pa = (pf_mask & 1) ? static_cast<const uint16_t*>(a.preadd) + pos.pix * a.preadd_cstride
                                           : static_cast<const uint16_t*>(a.residual2) + pos.pix * a.residual2_cstride;
        const uint16_t* pb = static_cast<const uint16_t*>(a.residual) + pos.pix * a.residual_cstride;
#pragma unroll
        for (int ci = 0; ci < kPfChunks; ++ci) {
          const int c = col_begin + 16 * ci;
          if (c < col_end) {
            const int n = pos.n0 + c;
            if (pf_mask & 5) {
              pf.a[ci][0] = __ldg(reinterpret_cast<const uint4*>(pa + n));
              pf.a[ci][1] = __ldg(reinterpret_cast<const uint4*>(pa + n) + 1);
            }
            if (pf_mask & 2) {
              pf.b[ci][0] = __ldg(reinterpret_cast<const uint4*>(pb + n));
              pf.b[ci][1] = __ldg(reinterpret_cast<const uint4*>(pb + n) + 1);
            }
          }
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (etid == 0 && local < 4) trace_mark(a.trace, 9 + local);  // accumulator of tile `local` complete
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc * a.acc_cols) +
                              (static_cast<uint32_t>(quarter * 32) << 16);
      if (a.debug & 3) {
        if (a.debug & 2) {
          for (int cc = col_begin; cc < col_end; cc += 16) {
            uint32_t r0[16];
            __syncwarp();
            tmem_ld16(t_addr + static_cast<uint32_t>(cc), r0);
            tmem_ld_wait();
            if (r0[0] == 0x7fc12345u && pos.valid) static_cast<uint16_t*>(a.out)[0] = 0;  // keep the load alive
          }
        }
      } else if (EPI == 2) {   // host: fast != 0 and gn_partial != NULL
        if (a.fast == 2)
          epilogue_fast<true, true>(a, t_addr, col_begin, col_end, pos.n0, pos.pix, pos.valid, sb, pf, pf_mask, a.out_dtype == FLAIR_F16, gn_tile);
        else
          epilogue_fast<false, true>(a, t_addr, col_begin, col_end, pos.n0, pos.pix, pos.valid, sb, pf, 0, a.out_dtype == FLAIR_F16, gn_tile);
      } else if (EPI == 1 || a.fast != 0) {   // (EPI == 1: host guarantees fast != 0, no statistics)
        if (a.fast == 2)
          epilogue_fast<true, false>(a, t_addr, col_begin, col_end, pos.n0, pos.pix, pos.valid, sb, pf, pf_mask, a.out_dtype == FLAIR_F16, nullptr);
        else
          epilogue_fast<false, false>(a, t_addr, col_begin, col_end, pos.n0, pos.pix, pos.valid, sb, pf, 0, a.out_dtype == FLAIR_F16, nullptr);
      } else if (EPI == 0)
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

  if (threadIdx.x == 0) trace_mark(a.trace, 13);  // first epilogue thread done with all its tiles
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 32 * kWarpTma) trace_mark(a.trace, 14);
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(a.tmem_cols));
  }
}

// FLAIR_CONV_MODE = 0 forces the generic tiling (debug / A-B measurements); FLAIR_CONV_RESIDENT=0 forces streamed
// weights in the halo mode; default: auto
int flair_conv_mode_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("FLAIR_CONV_MODE");
    v = e ? atoi(e) : -1;
  }
  return v;
}
int flair_conv_resident_override() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("FLAIR_CONV_RESIDENT");
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

// Number of M tiles (and M tiles per batch element) flair_conv_igemm walks for these extents: sizes the gn_partial
// buffer of the fused GroupNorm statistics ([m_tiles][4][Cout/16][16] floats).  Mirrors the tile selection below.
extern "C" int flair_conv_gn_tiles(int B, int T, int H, int W, int kh, int kw, int stride_hw, int* m_tiles,
                                   int* tiles_per_batch, int* frames_per_tile) {
  FLAIR_REQUIRE(m_tiles && tiles_per_batch && frames_per_tile && B > 0 && T > 0 && H > 0 && W > 0 &&
                    (stride_hw == 1 || stride_hw == 2),
                "flair_conv_gn_tiles: bad arguments");
  const int s = stride_hw;
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  int bw = pow2_ceil(Wo); if (bw > kBlockM) bw = kBlockM;
  int bh = pow2_ceil(Ho); if (bh > kBlockM / bw) bh = kBlockM / bw;
  int bt = kBlockM / (bw * bh);
  if (kh == 3 && kw == 3 && s == 1 && Wo >= 8 && Ho >= 16 && flair_conv_mode_override() != 0) { bw = 8; bh = 16; bt = 1; }
  *tiles_per_batch = ceil_div(Wo, bw) * ceil_div(Ho, bh) * ceil_div(T, bt);
  *m_tiles = *tiles_per_batch * B;
  *frames_per_tile = bt;
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

  const int s = p->stride_hw;
  const int Ho = (p->H + s - 1) / s, Wo = (p->W + s - 1) / s;
  const int Cin_pad = ceil_div(p->Cin, kBlockK) * kBlockK;
  const int Cout_pad = ceil_div(p->Cout, 16) * 16;

  ConvKArgs a{};
  a.B = p->B; a.T = p->T; a.Ho = Ho; a.Wo = Wo;
  // mode selection (see ConvKArgs): halo tiles need a 3x3 spatial kernel, stride 1, and maps of >= 8 x 16 pixels
  const int env_mode = flair_conv_mode_override();
  int mode = 0;
  if (p->kh == 3 && p->kw == 3 && s == 1 && Wo >= 8 && Ho >= 16 && env_mode != 0) mode = 3;
  // tile box: (mode 0) as wide as possible, then tall, then across frames; (halo) 8 (w) x 16 (h) in one frame
  int bw = pow2_ceil(Wo); if (bw > kBlockM) bw = kBlockM;
  int bh = pow2_ceil(Ho); if (bh > kBlockM / bw) bh = kBlockM / bw;
  int bt = kBlockM / (bw * bh);
  if (mode == 3) { bw = 8; bh = 16; bt = 1; }
  a.lbw = ilog2(bw); a.lbh = ilog2(bh); a.lbt = ilog2(bt);
  a.tiles_w = ceil_div(Wo, bw); a.tiles_h = ceil_div(Ho, bh); a.tiles_t = ceil_div(p->T, bt);
  a.m_tiles = a.tiles_w * a.tiles_h * a.tiles_t * p->B;
  int n_tile = Cout_pad;
  if (n_tile > 256) {
    n_tile = 256;
    while (Cout_pad % n_tile != 0) n_tile -= 16;
  }
  // few M tiles (low-resolution maps): split N further so that more SMs pull the (large) filter from L2 in parallel
  // — these launches are bound by streaming the weights, not by the tensor pipe
  {
    const int sms = flair_num_sms();
    while (n_tile % 32 == 0 && n_tile / 2 >= 64 && a.m_tiles * (Cout_pad / n_tile) * 2 <= sms) n_tile /= 2;
  }
  const int ntaps_total = p->kt * p->kh * p->kw;
  const int kblocks_ = Cin_pad / kBlockK;
  const int smem_cap = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - 2048 /*static bias*/;
  int resident = 0, b_stages = 0;
  if (mode == 3) {
    // resident weights: all (tap, k-block) slabs of the N tile + >= 2 halo slabs (a slab is 36 MMAs of work: two
    // stages hide its load).  The N tile is NOT shrunk to make the weights fit: N = 64 MMAs are bound by the
    // shared-memory operand reads (48 cycles per K step against 32 of tensor time, tests/gpu_probes/cu/mma_rate.cu)
    // and every extra N tile re-reads the activations.
    const long long w_all = static_cast<long long>(ntaps_total) * kblocks_ * 128 * n_tile;
    if (flair_conv_resident_override() != 0 && w_all + 2LL * kSlabStage <= smem_cap) resident = 1;
    if (!resident) {
      // streamed weights: ring of (tap, k-block) slabs next to >= 3 halo slabs
      const int b_bytes_ = n_tile * 128;
      b_stages = (smem_cap - 3 * static_cast<int>(kSlabStage)) / b_bytes_;
      if (b_stages > kMaxStages) b_stages = kMaxStages;
      FLAIR_REQUIRE(b_stages >= 3, "flair_conv_igemm: weight ring does not fit shared memory");  // a dh row waits for 3 slots
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
    static int debug = -1;
    if (debug < 0) { const char* e = getenv("FLAIR_CONV_DEBUG"); debug = e ? atoi(e) : 0; }
    a.debug = debug;
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
  {
    // compact epilogue: 16-bit NHWC rows of whole 32-byte chunks, no per-frame bias / gate; addends (if any) 16-bit,
    // vector-aligned, at most one of preadd / residual2 next to residual, and <= 4 chunks per epilogue warp
    auto vec_ok = [](const void* ptr, int dtype, long long cs) {
      return ptr == nullptr || (dtype != FLAIR_F32 && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && cs % 8 == 0);
    };
    const bool plain = p->out_layout == FLAIR_OUT_NHWC && p->out_dtype != FLAIR_F32 && p->Cout % 16 == 0 &&
                       (reinterpret_cast<uintptr_t>(p->out) & 31) == 0 && p->out_cstride % 16 == 0 &&
                       p->rowbias == nullptr && p->rowscale == nullptr;
    const bool any_add = p->preadd || p->residual || p->residual2;
    const bool add_ok = vec_ok(p->preadd, p->preadd_dtype, p->preadd_cstride) &&
                        vec_ok(p->residual, p->residual_dtype, p->residual_cstride) &&
                        vec_ok(p->residual2, p->residual2_dtype, p->residual2_cstride) &&
                        !(p->preadd && p->residual2) && n_tile <= 32 * kPfChunks;
    const bool same_dt = (!p->preadd || p->preadd_dtype == p->out_dtype) && (!p->residual || p->residual_dtype == p->out_dtype) &&
                         (!p->residual2 || p->residual2_dtype == p->out_dtype);
    static int fast_env = -1;
    if (fast_env < 0) { const char* e = getenv("FLAIR_CONV_FAST_EPI"); fast_env = e ? atoi(e) : 1; }
    a.fast = (!plain || !fast_env) ? 0 : (!any_add ? 1 : ((add_ok && same_dt) ? 2 : 0));
  }
  if (p->gn_partial != nullptr) {
    const int cpg = (p->gn_groups > 0 && p->Cout % p->gn_groups == 0) ? p->Cout / p->gn_groups : 0;
    FLAIR_REQUIRE(cpg >= 2 && (cpg & (cpg - 1)) == 0 && (cpg <= 16 || cpg % 16 == 0),
                  "flair_conv_igemm: fused GN statistics need a power-of-two group size >= 2 (Cout=%d groups=%d)", p->Cout,
                  p->gn_groups);
    if (a.fast == 0) {  // nothing was launched: the caller falls back to flair_gn_stats
      flair_set_error("flair_conv_igemm: fused GN statistics need the compact epilogue (16-bit NHWC output, Cout %% 16 == 0, "
                      "no per-frame bias / gate, 16-bit aligned addends)");
      return FLAIR_ERR_UNSUPPORTED;
    }
    a.gn_partial = p->gn_partial; a.gn_groups = p->gn_groups; a.gn_cpg = cpg; a.gn_nchunks = Cout_pad / 16;
  }
  a.out2 = static_cast<uint16_t*>(p->out2); a.out2_gs = p->out2_group_channels; a.out2_gstride = p->out2_group_stride;
  a.out2_nb = p->out2_neighbor > 0 ? p->out2_neighbor : 1;
  a.out2_back = static_cast<long long>(a.out2_nb) * 2 * a.out2_gs - a.out2_gs;
  if (p->out2 != nullptr)
    FLAIR_REQUIRE(p->out_layout == FLAIR_OUT_NHWC && p->out_dtype != FLAIR_F32 && p->Cout % 16 == 0 &&
                      (p->out2_group_channels == 8 || p->out2_group_channels == 16) && p->out2_group_stride % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(p->out2) & 15) == 0,
                  "flair_conv_igemm: out2 needs a 16-bit NHWC output, Cout %% 16 == 0, 8 or 16 channels per group");

  const uint32_t b_bytes = static_cast<uint32_t>(n_tile) * kBlockK * 2;  // multiple of 2 KB (n_tile % 16 == 0)
  a.mode = mode;
  a.ntd = p->kt;
  a.resident = resident;
  a.b_stages = b_stages;
  a.b_bytes = b_bytes;
  a.a_bytes = (mode == 0) ? kABytes : kSlabBytes;
  // w_bytes: the resident weights of this CTA's N tile, or the streamed-weight ring (halo mode)
  a.w_bytes = (mode == 3) ? (resident ? static_cast<uint32_t>(ntaps_total) * a.kblocks * b_bytes
                                      : static_cast<uint32_t>(b_stages) * b_bytes)
                          : 0u;
  uint32_t stage_bytes = (mode == 0) ? a.a_bytes + b_bytes : kSlabStage;
  stage_bytes = (stage_bytes + 1023u) & ~1023u;
  a.stage_bytes = stage_bytes;
  const int smem_budget = smem_cap - static_cast<int>(a.w_bytes);
  int stages = smem_budget / static_cast<int>(stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  FLAIR_REQUIRE(stages >= 2, "flair_conv_igemm: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem_bytes = a.w_bytes + static_cast<size_t>(stages) * stage_bytes + 1024 + 512;

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
    cuuint32_t box[5] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(mode == 3 ? bw + 2 : bw * s),
                         static_cast<cuuint32_t>(mode == 3 ? bh + 2 : bh * s), static_cast<cuuint32_t>(bt), 1u};
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
  if (attr_once.first()) {
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  }
  const int total_tiles = a.m_tiles * a.n_tiles;
  int grid = flair_num_sms();
  if (grid > total_tiles) grid = total_tiles;
  if (resident) {  // every CTA keeps one N tile: grid must be a multiple of n_tiles
    grid = (grid / a.n_tiles) * a.n_tiles;
    if (grid < a.n_tiles) grid = a.n_tiles;
  }
  static int slim = -1;   // FLAIR_CONV_SLIM=0: always launch the full instance (A/B measurements)
  if (slim < 0) { const char* e = getenv("FLAIR_CONV_SLIM"); slim = e ? atoi(e) : 1; }
  if (slim && a.fast != 0 && a.gn_partial != nullptr)
    FLAIR_CHECK_CUDA(flair_launch(conv_igemm_kernel<2>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, a));
  else if (slim && a.fast != 0)
    FLAIR_CHECK_CUDA(flair_launch(conv_igemm_kernel<1>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, a));
  else
    FLAIR_CHECK_CUDA(flair_launch(conv_igemm_kernel<0>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, a));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
