// GroupNorm32 statistics + fused normalise / FiLM / SiLU / 2x resample, channels-last.
//
// Replaces guided_diffusion/nn_new.py:17-19 (GroupNorm32: fp32 statistics, eps 1e-5) wrapped in
// LazyReshaper3D (nn.py:359-367: statistics over (C/G, T, H, W) of one batch element), the
// `nn.SiLU()` that follows every norm (unet_new.py:237-240,265-268), the scale-shift conditioning
// h = norm(h) * (1 + scale) + shift (unet_new.py:321-325) and the nearest-x2 / 2x2-average
// resampling applied to the activated tensor inside up/down ResBlocks (unet_new.py:249-254,310-315).
// One HBM pass for the statistics (2 B/element read), one for apply (2 B read + 2 B written);
// the reference spends >= 3 fp32 passes per norm (cast, native_group_norm, SiLU, cast).
//
// Layout: x[b][p][c], p = (t,h,w) pixel index, c fastest.  Group g owns channels [g*cpg,(g+1)*cpg).
#include <stdlib.h>

#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

constexpr int kGroupsMax = 32;

__device__ __forceinline__ void load8(const void* base, long long elem_off, int dtype, float (&v)[8]) {
  if (dtype == FLAIR_F32) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + elem_off));
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
}

__device__ __forceinline__ void store8(void* base, long long elem_off, int dtype, const float (&v)[8]) {
  if (dtype == FLAIR_F32) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
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
    *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + elem_off) = u;
  }
}

__device__ __forceinline__ void unpack8(const uint4& u, int dtype, float (&v)[8]) {
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

// ---------------------------------------------------------------- statistics (partial sums)
// grid (nchunks, B); block = vecs * ppb threads (vecs = C/8).  partial[b][chunk][g] = (sum, sumsq).
__global__ void __launch_bounds__(256, 4)
gn_stats_kernel(const void* __restrict__ x, int dtype, long long P, int C, int cstride, int groups,
                int ppb, long long pix_per_chunk, float2* __restrict__ partial, int* __restrict__ counter,
                float2* __restrict__ final_stats, float eps) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ float sm[];  // [2][ppb][C]
  const int vecs = C / 8;
  const int cv = threadIdx.x % vecs, pl = threadIdx.x / vecs;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const long long p0 = chunk * pix_per_chunk;
  long long p1 = p0 + pix_per_chunk;
  if (p1 > P) p1 = P;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  const long long base = static_cast<long long>(b) * P;
  // 8 independent 16-byte loads in flight per thread (one per iteration kept ~8 KB in flight per SM: 1.7 TB/s);
  // the accumulation order per thread is unchanged, so the statistics are bit-identical to the simple loop
  // (ncu: with 2 resident CTAs x 8 loads DRAM was 23 % busy and nothing else above 16 %: latency-bound.  The loads
  // are kept as raw 16-byte registers until consumed so that 4 CTAs fit per SM: 32 warps x 8 x 512 B in flight.)
  // Round 2: the batches are software-pipelined — the loads of batch i+1 are issued BEFORE batch i is accumulated, so a
  // warp always has 4-8 loads outstanding instead of alternating "8 in flight" / "none in flight" (a CTA walks ~9
  // dependent batches: that serialisation, not the byte count, set the 4.6 TB/s streaming rate).
  constexpr int kU = 4;
  long long p = p0 + pl;
  if (dtype != FLAIR_F32) {
    const uint16_t* xp = static_cast<const uint16_t*>(x);
    const long long step = static_cast<long long>(kU) * ppb;
    uint4 cur[kU], nxt[kU];
    auto load = [&](uint4 (&r)[kU], long long pp) {
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (pp + static_cast<long long>(u) * ppb < p1)
          r[u] = __ldg(reinterpret_cast<const uint4*>(xp + (base + pp + static_cast<long long>(u) * ppb) * cstride + cv * 8));
    };
    load(cur, p);
    for (; p < p1; p += step) {  // predicated batches, same accumulation order as the simple loop
      load(nxt, p + step);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (p + static_cast<long long>(u) * ppb < p1) {
          float v[8];
          unpack8(cur[u], dtype, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) cur[u] = nxt[u];
    }
  }
  for (; p < p1; p += ppb) {
    float v[8];
    load8(x, (base + p) * cstride + cv * 8, dtype, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
  }
  float* sS = sm;
  float* sQ = sm + ppb * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sS[pl * C + cv * 8 + j] = s[j];
    sQ[pl * C + cv * 8 + j] = q[j];
  }
  __syncthreads();
  // fixed-order (deterministic) reduction: per channel over ppb rows, then per group over channels
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, d = 0.f;
    for (int r = 0; r < ppb; ++r) { a += sS[r * C + c]; d += sQ[r * C + c]; }
    sS[c] = a; sQ[c] = d;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int cpg = C / groups;
    float a = 0.f, d = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { a += sS[c]; d += sQ[c]; }
    partial[(static_cast<long long>(b) * gridDim.x + chunk) * groups + threadIdx.x] = make_float2(a, d);
  }
  if (counter == nullptr) return;
  // ---- the LAST chunk of this batch element to finish reduces all partials once, in chunk order (deterministic),
  // and writes (mean, rstd): the apply kernels then read 2 floats per group instead of every CTA re-reducing
  // ~300 partials in its prologue (that prologue was 10-20 us of a 25-60 us apply launch).
  __shared__ int s_last;
  __shared__ double s_red[2][8][kGroupsMax];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter + b, 1) == static_cast<int>(gridDim.x) - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nch = gridDim.x;
  // 8 slots per group, each a contiguous range of chunks; combined in slot order (blockDim may be < 8 * groups)
  for (int idx = threadIdx.x; idx < 8 * groups; idx += blockDim.x) {
    const int g = idx % groups, part = idx / groups;
    const int per = (nch + 7) / 8;
    const int k0 = part * per, k1 = min(nch, k0 + per);
    double sd = 0.0, qd = 0.0;
    for (int k = k0; k < k1; k += 8) {  // 8 loads in flight, summed in chunk order
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = (k + u < k1) ? __ldcg(partial + (static_cast<long long>(b) * nch + k + u) * groups + g) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { sd += v[u].x; qd += v[u].y; }
    }
    s_red[0][part][g] = sd; s_red[1][part][g] = qd;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    double sd = 0.0, qd = 0.0;
    for (int part = 0; part < 8; ++part) { sd += s_red[0][part][threadIdx.x]; qd += s_red[1][part][threadIdx.x]; }
    const double n = static_cast<double>(P) * (C / groups);
    const double mean = sd / n;
    double var = qd / n - mean * mean;
    if (var < 0.0) var = 0.0;
    final_stats[static_cast<long long>(b) * groups + threadIdx.x] =
        make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
  if (threadIdx.x == 0) counter[b] = 0;  // self-cleaning: the next launch on this stream finds zeros
}

// ---------------------------------------------------------------- statistics fused into the producing conv
// flair_conv_igemm (gn_partial) leaves [B*tpb][4][nch][16] floats: per M tile, per 32-row quarter, per 16-channel chunk
// the (sum, sum of squares) of every group inside the chunk (cpg <= 16: floats 2g, 2g+1; cpg > 16: floats 0, 1 of each
// of the group's chunks).  Stage 1 (grid (nsplit, B)): column sums over a contiguous range of entries, in double,
// fixed order.  Stage 2: the LAST block of a batch element (ticket) adds the splits in order and writes (mean, rstd)
// in the format gn_apply reads (nchunks = 0).  Deterministic; replaces the 2 B/element statistics pass.
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partial, int entries, int cols, int cpg, int groups, double count,
                   double* __restrict__ scratch, int* __restrict__ counter, float2* __restrict__ final_stats, float eps) {
  pdl_sync();
  extern __shared__ double sred[];   // [rows][cols]
  const int b = blockIdx.y, sp = blockIdx.x, nsplit = gridDim.x;
  const int per = (entries + nsplit - 1) / nsplit;
  const int e0 = sp * per, e1 = min(entries, e0 + per);
  const int rows = (cols >= 256) ? 1 : 256 / cols;
  const float* base = partial + static_cast<long long>(b) * entries * cols;
  for (int f0 = 0; f0 < cols; f0 += 256) {
    const int f = f0 + static_cast<int>(threadIdx.x) % min(cols, 256);
    const int r = (cols >= 256) ? 0 : static_cast<int>(threadIdx.x) / cols;
    double acc = 0.0;
    if (f < cols && r < rows) {
      int e = e0 + r;
      for (; e + 7 * rows < e1; e += 8 * rows) {   // 8 loads in flight, summed in entry order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + static_cast<long long>(e + u * rows) * cols + f);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; e < e1; e += rows) acc += __ldcg(base + static_cast<long long>(e) * cols + f);
      sred[r * cols + f] = acc;
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < cols; f += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < rows; ++r) a += sred[r * cols + f];
    scratch[(static_cast<long long>(b) * nsplit + sp) * cols + f] = a;
  }
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter + b, 1) == nsplit - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // stage 2: column sums over the splits (8 loads in flight, split order), then per group
  __syncthreads();
  for (int f = threadIdx.x; f < cols; f += blockDim.x) {
    double a = 0.0;
    const double* col = scratch + static_cast<long long>(b) * nsplit * cols + f;
    int k = 0;
    for (; k + 8 <= nsplit; k += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(col + static_cast<long long>(k + u) * cols);
#pragma unroll
      for (int u = 0; u < 8; ++u) a += v[u];
    }
    for (; k < nsplit; ++k) a += __ldcg(col + static_cast<long long>(k) * cols);
    sred[f] = a;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double sd = 0.0, qd = 0.0;
    const int c_first = g * cpg;
    const int nparts = (cpg > 16) ? cpg / 16 : 1;
    for (int part = 0; part < nparts; ++part) {
      const int chunk = c_first / 16 + part;
      const int idx = (cpg > 16) ? 0 : (c_first % 16) / cpg;
      const int fs = chunk * 16 + 2 * idx;
      sd += sred[fs];
      qd += sred[fs + 1];
    }
    const double mean = sd / count;
    double var = qd / count - mean * mean;
    if (var < 0.0) var = 0.0;
    final_stats[static_cast<long long>(b) * groups + g] =
        make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
  if (threadIdx.x == 0) counter[b] = 0;  // self-cleaning
}

struct ApplyArgs {
  const void* x; int in_dtype;
  void* out; int out_dtype;
  const float2* partial; int nchunks;
  const float* gamma; const float* beta;
  const float* scale; const float* shift; int film_stride;  // per frame rows, or NULL
  int T, H, W, C, groups;
  int x_cstride, out_cstride;
  int norm, silu, resample;  // resample: 0 none, 1 nearest x2 up, 2 2x2 average down
  float eps;
};

// grid (blocks, B)
__global__ void __launch_bounds__(256) gn_apply_kernel(const __grid_constant__ ApplyArgs a) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ float s_mean[kGroupsMax], s_rstd[kGroupsMax];
  const int b = blockIdx.y;
  const int cpg = a.C / a.groups;
  if (a.norm && a.nchunks == 0 && threadIdx.x < a.groups) {  // statistics already finalised by gn_stats
    const float2 v = __ldg(a.partial + static_cast<long long>(b) * a.groups + threadIdx.x);
    s_mean[threadIdx.x] = v.x;
    s_rstd[threadIdx.x] = v.y;
  } else if (a.norm && threadIdx.x < a.groups) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < a.nchunks; ++k) {
      const float2 v = __ldg(a.partial + (static_cast<long long>(b) * a.nchunks + k) * a.groups + threadIdx.x);
      s += v.x; q += v.y;
    }
    const double n = static_cast<double>(a.T) * a.H * a.W * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  const int vecs = a.C / 8;
  // iteration space: input pixels for "none"/"up", output pixels for "down"
  const int Hi = (a.resample == 2) ? a.H / 2 : a.H;
  const int Wi = (a.resample == 2) ? a.W / 2 : a.W;
  const long long items = static_cast<long long>(a.T) * Hi * Wi * vecs;
  const long long in_frame = static_cast<long long>(a.H) * a.W;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(it % vecs);
    long long pix = it / vecs;
    const int w = static_cast<int>(pix % Wi); pix /= Wi;
    const int h = static_cast<int>(pix % Hi);
    const int t = static_cast<int>(pix / Hi);
    const int c0 = cv * 8;
    float g[8], be[8], sc[8], sh[8], mu[8], rs[8];
    if (a.norm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = __ldg(a.gamma + c0 + j); be[j] = __ldg(a.beta + c0 + j);
        const int grp = (c0 + j) / cpg;
        mu[j] = s_mean[grp]; rs[j] = s_rstd[grp];
      }
      if (a.scale != nullptr) {
        const long long row = (static_cast<long long>(b) * a.T + t) * a.film_stride;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = 1.0f + __ldg(a.scale + row + c0 + j); sh[j] = __ldg(a.shift + row + c0 + j); }
      }
    }
    auto xform = [&](float (&v)[8]) {
      if (a.norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = (v[j] - mu[j]) * rs[j] * g[j] + be[j];
          if (a.scale != nullptr) y = y * sc[j] + sh[j];
          v[j] = y;
        }
      }
      if (a.silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
      }
    };
    const long long in_base = (static_cast<long long>(b) * a.T + t) * in_frame;
    if (a.resample == 2) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          float v[8];
          load8(a.x, (in_base + static_cast<long long>(2 * h + dy) * a.W + 2 * w + dx) * a.x_cstride + c0,
                a.in_dtype, v);
          xform(v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
      const long long o = ((static_cast<long long>(b) * a.T + t) * Hi + h) * Wi + w;
      store8(a.out, o * a.out_cstride + c0, a.out_dtype, acc);
    } else {
      float v[8];
      load8(a.x, (in_base + static_cast<long long>(h) * a.W + w) * a.x_cstride + c0, a.in_dtype, v);
      xform(v);
      if (a.resample == 1) {
        const int Wo = 2 * a.W;
        const long long ob = (static_cast<long long>(b) * a.T + t) * (4 * in_frame);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx)
            store8(a.out, (ob + static_cast<long long>(2 * h + dy) * Wo + 2 * w + dx) * a.out_cstride + c0,
                   a.out_dtype, v);
      } else {
        store8(a.out, (in_base + static_cast<long long>(h) * a.W + w) * a.out_cstride + c0, a.out_dtype, v);
      }
    }
  }
}


// Fast path of gn_apply (no resampling, C/8 a power of two <= 256): one CTA row per frame, every thread
// keeps one 8-channel vector, so the whole per-channel affine (statistics, gamma/beta, FiLM) folds into
// y = x * A + B held in registers; the inner loop is one 16-byte load, 8 FMAs (+ SiLU), one 16-byte store.
// grid (blocks, B*T)
// Prologue shared by the fast apply kernels: statistics of this frame's batch element into shared memory, then the
// per-thread folded affine of the thread's 8 channels: y = x * A + B (statistics, gamma / beta, FiLM).
__device__ __forceinline__ void gn_fold_affine(const ApplyArgs& a, float* s_mean, float* s_rstd, int& c0_out, int& pl_out,
                                               int& ppb_out, float (&A)[8], float (&Bc)[8]) {
  const int frame = blockIdx.y;  // b*T + t
  const int b = frame / a.T;
  const int cpg = a.C / a.groups;
  if (a.norm && a.nchunks == 0 && threadIdx.x < a.groups) {  // statistics already finalised by gn_stats
    const float2 v = __ldg(a.partial + static_cast<long long>(b) * a.groups + threadIdx.x);
    s_mean[threadIdx.x] = v.x;
    s_rstd[threadIdx.x] = v.y;
  } else if (a.norm && threadIdx.x < a.groups) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < a.nchunks; ++k) {
      const float2 v = __ldg(a.partial + (static_cast<long long>(b) * a.nchunks + k) * a.groups + threadIdx.x);
      s += v.x; q += v.y;
    }
    const double n = static_cast<double>(a.T) * a.H * a.W * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
  __syncthreads();
  const int vecs = a.C / 8;
  const int cv = threadIdx.x % vecs, pl = threadIdx.x / vecs, ppb = blockDim.x / vecs;
  const int c0 = cv * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { A[j] = 1.0f; Bc[j] = 0.0f; }
  if (a.norm) {
    // per-channel constants with 16-byte loads (they were 32 scalar loads per thread: on the small maps this
    // prologue cost more than the pixels)
    float ga[8], be[8];
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + c0) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(a.beta + c0) + 1);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int grp = (c0 + j) / cpg;
      const float g = ga[j] * s_rstd[grp];
      A[j] = g;
      Bc[j] = be[j] - s_mean[grp] * g;
    }
    if (a.scale != nullptr) {
      const long long row = static_cast<long long>(frame) * a.film_stride + c0;
      float sc[8], sh[8];
      if ((a.film_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(a.scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.shift) & 15) == 0) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.scale + row)), s1 = __ldg(reinterpret_cast<const float4*>(a.scale + row) + 1);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(a.shift + row)), h1 = __ldg(reinterpret_cast<const float4*>(a.shift + row) + 1);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = __ldg(a.scale + row + j); sh[j] = __ldg(a.shift + row + j); }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float m = 1.0f + sc[j];
        A[j] *= m;
        Bc[j] = Bc[j] * m + sh[j];
      }
    }
  }
  c0_out = c0; pl_out = pl; ppb_out = ppb;
}

__global__ void __launch_bounds__(256, 3) gn_apply_fast_kernel(const __grid_constant__ ApplyArgs a) {
  pdl_sync();
  __shared__ float s_mean[kGroupsMax], s_rstd[kGroupsMax];
  const int frame = blockIdx.y;  // b*T + t
  int c0, pl, ppb;
  float A[8], Bc[8];
  gn_fold_affine(a, s_mean, s_rstd, c0, pl, ppb, A, Bc);
  const long long P = static_cast<long long>(a.H) * a.W;
  const long long base = static_cast<long long>(frame) * P;
  const long long stride = static_cast<long long>(gridDim.x) * ppb;
  auto xform = [&](float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(v[j], A[j], Bc[j]);
      if (a.silu) {  // y * sigmoid(y) = h + h * tanh(h), h = y/2: ONE MUFU op (tanh.approx, rel. error 2^-11) instead of ex2 + rcp
        const float h = 0.5f * y;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        y = fmaf(h, t, h);
      }
      v[j] = y;
    }
  };
  // 16-byte loads kept raw until consumed (register budget); software-pipelined like gn_stats: the loads of batch i+1
  // are in flight while batch i is transformed and stored
  constexpr int kU = 4;
  long long p = static_cast<long long>(blockIdx.x) * ppb + pl;
  if (a.in_dtype != FLAIR_F32) {
    const uint16_t* xp = static_cast<const uint16_t*>(a.x);
    const long long step = kU * stride;
    uint4 cur[kU], nxt[kU];
    auto load = [&](uint4 (&r)[kU], long long pp) {
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (pp + u * stride < P) r[u] = __ldg(reinterpret_cast<const uint4*>(xp + (base + pp + u * stride) * a.x_cstride + c0));
    };
    load(cur, p);
    for (; p < P; p += step) {  // predicated batches: no scalar tail (it ran with one load in flight)
      load(nxt, p + step);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (p + u * stride < P) {
          float v[8];
          unpack8(cur[u], a.in_dtype, v);
          xform(v);
          store8(a.out, (base + p + u * stride) * a.out_cstride + c0, a.out_dtype, v);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) cur[u] = nxt[u];
    }
  }
  for (; p < P; p += stride) {
    float v[8];
    load8(a.x, (base + p) * a.x_cstride + c0, a.in_dtype, v);
    xform(v);
    store8(a.out, (base + p) * a.out_cstride + c0, a.out_dtype, v);
  }
}

// Resampling variants of the fast path (ResBlock up / down, unet_new.py:249-254; sr3 Upsample): same per-thread folded
// affine, 32-bit pixel arithmetic, several independent 16-byte loads in flight.  RS = 1: nearest x2 up (one input pixel
// -> four stores), RS = 2: 2x2 average down (four transformed inputs -> one store, summed in (dy, dx) order like the
// generic kernel).  grid (blocks, B*T).
template <int RS>
__global__ void __launch_bounds__(256, 3) gn_apply_rs_kernel(const __grid_constant__ ApplyArgs a) {
  pdl_sync();
  __shared__ float s_mean[kGroupsMax], s_rstd[kGroupsMax];
  const int frame = blockIdx.y;
  int c0, pl, ppb;
  float A[8], Bc[8];
  gn_fold_affine(a, s_mean, s_rstd, c0, pl, ppb, A, Bc);
  auto xform = [&](float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(v[j], A[j], Bc[j]);
      if (a.silu) {
        const float h = 0.5f * y;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        y = fmaf(h, t, h);
      }
      v[j] = y;
    }
  };
  const int stride = gridDim.x * ppb;
  const long long in_base = static_cast<long long>(frame) * a.H * a.W;
  const uint16_t* xp = static_cast<const uint16_t*>(a.x);
  if (RS == 1) {
    const int P = a.H * a.W, Wo = 2 * a.W;
    const long long out_base = in_base * 4;
    constexpr int kU = 4;
    for (int p = blockIdx.x * ppb + pl; p < P; p += kU * stride) {
      uint4 raw[kU];   // 16-bit inputs only (host-checked): kept raw until consumed
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (p + u * stride < P) raw[u] = __ldg(reinterpret_cast<const uint4*>(xp + (in_base + p + u * stride) * a.x_cstride + c0));
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pp = p + u * stride;
        if (pp < P) {
          float vv[8];
          unpack8(raw[u], a.in_dtype, vv);
          xform(vv);
          const int h = pp / a.W, w = pp - h * a.W;
          const long long o = out_base + static_cast<long long>(2 * h) * Wo + 2 * w;
          store8(a.out, o * a.out_cstride + c0, a.out_dtype, vv);
          store8(a.out, (o + 1) * a.out_cstride + c0, a.out_dtype, vv);
          store8(a.out, (o + Wo) * a.out_cstride + c0, a.out_dtype, vv);
          store8(a.out, (o + Wo + 1) * a.out_cstride + c0, a.out_dtype, vv);
        }
      }
    }
  } else {
    const int Hi = a.H / 2, Wi = a.W / 2, Po = Hi * Wi;
    const long long out_base = static_cast<long long>(frame) * Po;
    constexpr int kU = 2;
    for (int p = blockIdx.x * ppb + pl; p < Po; p += kU * stride) {
      uint4 raw[kU][4];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pp = p + u * stride;
        if (pp < Po) {
          const int h = pp / Wi, w = pp - h * Wi;
          const long long i0 = in_base + static_cast<long long>(2 * h) * a.W + 2 * w;
          raw[u][0] = __ldg(reinterpret_cast<const uint4*>(xp + i0 * a.x_cstride + c0));
          raw[u][1] = __ldg(reinterpret_cast<const uint4*>(xp + (i0 + 1) * a.x_cstride + c0));
          raw[u][2] = __ldg(reinterpret_cast<const uint4*>(xp + (i0 + a.W) * a.x_cstride + c0));
          raw[u][3] = __ldg(reinterpret_cast<const uint4*>(xp + (i0 + a.W + 1) * a.x_cstride + c0));
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int pp = p + u * stride;
        if (pp < Po) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float vv[8];
            unpack8(raw[u][q], a.in_dtype, vv);
            xform(vv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += vv[j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
          store8(a.out, (out_base + pp) * a.out_cstride + c0, a.out_dtype, acc);
        }
      }
    }
  }
}

// dst[p][coff + c] = src[p][c]   (channel concat into a wider channels-last buffer)
__global__ void __launch_bounds__(256)
copy_channels_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long P, int vecs,
                     int src_vstride, int dst_vstride, int dst_voff) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long items = P * vecs;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int kU = 4;  // independent 16-byte loads in flight per thread
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items; it += kU * stride) {
    uint4 r[kU];
    long long d[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = it + u * stride;
      if (i < items) {
        const long long p = i / vecs;
        const int v = static_cast<int>(i - p * vecs);
        r[u] = __ldg(src + p * src_vstride + v);
        d[u] = p * dst_vstride + dst_voff + v;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (it + u * stride < items) dst[d[u]] = r[u];
  }
}

int ew_blocks(long long items, int per_sm) {
  long long blocks = ceil_div_ll(items, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

extern "C" int flair_gn_stats_chunks(long long pixels_per_batch, int C) {
  // enough CTAs to fill the machine, at least ~64 pixels per thread-row
  const int vecs = C / 8;
  int ppb = 256 / vecs;
  if (ppb < 1) ppb = 1;
  if (ppb > 32) ppb = 32;
  long long chunks = pixels_per_batch / (static_cast<long long>(ppb) * 16);
  if (chunks < 1) chunks = 1;
  static int cap = 0;
  if (cap == 0) {
    const char* e = getenv("FLAIR_GN_CHUNKS");
    cap = e ? atoi(e) : 296;
    if (cap < 1) cap = 296;
  }
  if (chunks > cap) chunks = cap;
  return static_cast<int>(chunks);
}

extern "C" int flair_gn_stats(const void* x, int dtype, int B, long long P, int C, int cstride, int groups,
                              float* partial, int nchunks, int* counter, float* final_stats, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && partial, "flair_gn_stats: null pointer");
  FLAIR_REQUIRE(groups > 0 && groups <= kGroupsMax && C % groups == 0 && C % 8 == 0 && C <= 2048,
                "flair_gn_stats: unsupported C=%d groups=%d", C, groups);
  FLAIR_REQUIRE(cstride % 8 == 0 && cstride >= C, "flair_gn_stats: bad channel stride %d", cstride);
  FLAIR_REQUIRE(nchunks > 0 && B > 0 && B < 65536, "flair_gn_stats: bad grid");
  FLAIR_REQUIRE((counter == nullptr) == (final_stats == nullptr), "flair_gn_stats: counter and final_stats go together");
  const int vecs = C / 8;
  int ppb = 256 / vecs;
  if (ppb < 1) ppb = 1;
  if (ppb > 32) ppb = 32;
  const long long ppc = ceil_div_ll(P, nchunks);
  const size_t smem = sizeof(float) * 2 * ppb * C;
  dim3 grid(nchunks, B);
  FLAIR_CHECK_CUDA(flair_launch(gn_stats_kernel, dim3(grid), dim3(vecs * ppb), smem, stream, x, dtype, P, C, cstride, groups, ppb, ppc,
                                reinterpret_cast<float2*>(partial), counter, reinterpret_cast<float2*>(final_stats),
                                eps > 0 ? eps : 1e-5f));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_gn_finalize_splits(int tiles_per_batch) {
  int n = (tiles_per_batch * 4) / 32;   // >= 32 entries per block, up to 2 blocks per SM
  if (n < 1) n = 1;
  if (n > 296) n = 296;
  return n;
}

extern "C" int flair_gn_finalize(const float* partial, int B, int tiles_per_batch, int C, int groups,
                                 long long pixels_per_batch, double* scratch, int* counter, float* final_stats,
                                 float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(partial && scratch && counter && final_stats, "flair_gn_finalize: null pointer");
  FLAIR_REQUIRE(B > 0 && B < 65536 && tiles_per_batch > 0 && C % 16 == 0 && C <= 2048 && groups > 0 && groups <= kGroupsMax &&
                    C % groups == 0,
                "flair_gn_finalize: bad arguments (C=%d groups=%d)", C, groups);
  const int cpg = C / groups;
  FLAIR_REQUIRE(cpg >= 2 && (cpg & (cpg - 1)) == 0 && (cpg <= 16 || cpg % 16 == 0), "flair_gn_finalize: unsupported group size %d", cpg);
  const int nsplit = flair_gn_finalize_splits(tiles_per_batch);
  const int rows = (C >= 256) ? 1 : 256 / C;
  const size_t smem = sizeof(double) * rows * C;
  FLAIR_CHECK_CUDA(flair_launch(gn_finalize_kernel, dim3(nsplit, B), dim3(256), smem, stream, partial, tiles_per_batch * 4, C, cpg,
                                groups, static_cast<double>(pixels_per_batch) * cpg, scratch, counter,
                                reinterpret_cast<float2*>(final_stats), eps > 0 ? eps : 1e-5f));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_gn_apply(const flair_gn_apply_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(p && p->x && p->out, "flair_gn_apply: null pointer");
  FLAIR_REQUIRE(p->C % 8 == 0 && p->x_cstride % 8 == 0 && p->out_cstride % 8 == 0, "flair_gn_apply: C must be a multiple of 8");
  if (p->norm)
    FLAIR_REQUIRE(p->partial && p->gamma && p->beta && p->groups > 0 && p->groups <= kGroupsMax &&
                      p->C % p->groups == 0 && p->nchunks >= 0,
                  "flair_gn_apply: bad normalisation arguments");
  FLAIR_REQUIRE(p->resample >= 0 && p->resample <= 2, "flair_gn_apply: resample must be 0, 1 or 2");
  if (p->resample == 2) FLAIR_REQUIRE(p->H % 2 == 0 && p->W % 2 == 0, "flair_gn_apply: odd size for 2x2 pooling");
  ApplyArgs a{};
  a.x = p->x; a.in_dtype = p->in_dtype; a.out = p->out; a.out_dtype = p->out_dtype;
  a.partial = reinterpret_cast<const float2*>(p->partial); a.nchunks = p->nchunks;
  a.gamma = p->gamma; a.beta = p->beta; a.scale = p->scale; a.shift = p->shift; a.film_stride = p->film_stride;
  a.T = p->T; a.H = p->H; a.W = p->W; a.C = p->C; a.groups = p->groups > 0 ? p->groups : 1;
  a.x_cstride = p->x_cstride; a.out_cstride = p->out_cstride;
  a.norm = p->norm; a.silu = p->silu; a.resample = p->resample; a.eps = p->eps > 0 ? p->eps : 1e-5f;
  const int vecs_ = p->C / 8;
  if (p->resample != 0 && p->in_dtype != FLAIR_F32 && vecs_ <= 256 && 256 % vecs_ == 0 &&
      static_cast<long long>(p->B) * p->T < 65536 &&
      static_cast<long long>(p->H) * p->W < (1ll << 30)) {
    const int frames = p->B * p->T;
    const int ppb = 256 / vecs_;
    const long long items = (p->resample == 2) ? static_cast<long long>(p->H / 2) * (p->W / 2) : static_cast<long long>(p->H) * p->W;
    long long bx = ceil_div_ll(items, static_cast<long long>(ppb) * 8);
    long long cap = static_cast<long long>(flair_num_sms()) * 3 / frames;   // one wave of 3 CTAs per SM
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (p->resample == 1)
      FLAIR_CHECK_CUDA(flair_launch(gn_apply_rs_kernel<1>, dim3(static_cast<unsigned>(bx), frames), dim3(256), 0, stream, a));
    else
      FLAIR_CHECK_CUDA(flair_launch(gn_apply_rs_kernel<2>, dim3(static_cast<unsigned>(bx), frames), dim3(256), 0, stream, a));
    FLAIR_CHECK_LAUNCH();
    return 0;
  }
  if (p->resample == 0 && vecs_ <= 256 && 256 % vecs_ == 0 && static_cast<long long>(p->B) * p->T < 65536) {
    const long long P = static_cast<long long>(p->H) * p->W;
    const int ppb = 256 / vecs_;
    const int frames = p->B * p->T;
    long long bx = ceil_div_ll(P, static_cast<long long>(ppb) * 16);  // >= 16 pixels per thread (prologue amortised)
    // ONE wave of 3 CTAs per SM: rounded DOWN (rounding up gave 450 CTAs for 444 slots at 10 frames: six CTAs ran
    // alone after the wave)
    long long cap = static_cast<long long>(flair_num_sms()) * 3 / frames;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    FLAIR_CHECK_CUDA(flair_launch(gn_apply_fast_kernel, dim3(static_cast<unsigned>(bx), frames), dim3(256), 0, stream, a));
    FLAIR_CHECK_LAUNCH();
    return 0;
  }
  const int Hi = (p->resample == 2) ? p->H / 2 : p->H, Wi = (p->resample == 2) ? p->W / 2 : p->W;
  const long long items = static_cast<long long>(p->T) * Hi * Wi * (p->C / 8);
  dim3 grid(ew_blocks(items, 8), p->B);
  FLAIR_CHECK_CUDA(flair_launch(gn_apply_kernel, dim3(grid), dim3(256), 0, stream, a));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_copy_channels(const void* src, void* dst, long long pixels, int channels, int elem_bytes,
                                   int src_cstride, int dst_cstride, int dst_coffset, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(src && dst, "flair_copy_channels: null pointer");
  const int epv = 16 / elem_bytes;  // elements per 16-byte vector
  FLAIR_REQUIRE((elem_bytes == 2 || elem_bytes == 4) && channels % epv == 0 && src_cstride % epv == 0 &&
                    dst_cstride % epv == 0 && dst_coffset % epv == 0,
                "flair_copy_channels: channel counts/offsets must be multiples of %d", epv);
  const int vecs = channels / epv;
  FLAIR_CHECK_CUDA(flair_launch(copy_channels_kernel, dim3(ew_blocks(pixels * vecs, 8)), dim3(256), 0, stream, 
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), pixels, vecs, src_cstride / epv,
      dst_cstride / epv, dst_coffset / epv));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
