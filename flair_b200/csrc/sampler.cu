// Fused FLAIR sampler update (fp32, HBM-bound, 128-bit vectorised).
//
// Replaces guided_diffusion/gaussian_diffusion.py:344-365 (_predict_xstart_from_eps,
// _predict_eps_from_xstart), :465-470 (data-consistency step), :497-506 (prev_recon
// overwrite), :507-515 (rho-mixed x_{t-1}) and the ~10 `_extract_into_tensor`
// host->device table uploads per step (:692-705): the coefficient table lives on the
// device and is indexed by a device-resident step counter, so a whole sampling step is
// CUDA-graph capturable.
//
//   x0   = clamp(a_t x_t - b_t eps, -1, 1)
//   x0   = clamp(x0 - gamma_t R, -1, 1)            R given at full resolution, or
//                                                  R = Up(q) for the blur x4 operator
//                                                  (pseudoSR.py:196-225, polyphase form)
//   x0[frames < k] = prev_recon
//   eps' = (a_t x_t - x0) / b_t
//   x_{t-1} = c_t x0 + [t != 0] (sqrt(1-rho) d_t eps' + sqrt(rho) d_t z)
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

struct UpdArgs {
  const float* x_t;
  const float* model_out;
  int model_ch;  // 3 or 6 (only the first 3 are eps)
  const float* noise;
  const float* R;      // (N,3,H,W) or NULL
  const float* q_lr;   // (N,3,H/sf,W/sf) or NULL: R = Up(q_lr)
  const float* up_taps;  // (kk,kk) fp32 = ds_kernel * sf^2
  int kk, sf, pre;
  const float* prev;  // (B,k,3,H,W) or NULL
  int prev_k, frames_per_window;
  const float* coef;  // [steps][8]: a,b,c,d,gamma,-,-,-
  const long long* t_arr;  // per-frame step index (device) or NULL
  const float* gamma_arr;  // per-frame gamma (device) or NULL -> coef[t][4]
  const float* x0_in;      // final pred_xstart given (skip eps / DC / prev) or NULL
  int t_host;
  float s1, s2;
  float* sample;
  float* pred_xstart;  // may be NULL
  int N, H, W;
  int clip;
};

__device__ __forceinline__ float clampf(float v, int clip) {
  return clip ? fminf(fmaxf(v, -1.0f), 1.0f) : v;
}

__device__ __forceinline__ float up_at(const UpdArgs& a, const float* q, int i, int j) {
  // Up(q)[i,j] = sum_{u,v} taps[u,v] * Z[i+u-r, j+v-r], Z non-zero only at (sf*m+pre, sf*n+pre)
  const int r = a.kk / 2;
  const int h = a.H / a.sf, w = a.W / a.sf;
  float acc = 0.0f;
  int u0 = ((a.pre + r - i) % a.sf + a.sf) % a.sf;
  int v0 = ((a.pre + r - j) % a.sf + a.sf) % a.sf;
  for (int u = u0; u < a.kk; u += a.sf) {
    const int zi = i + u - r;
    if (zi < 0 || zi >= a.H) continue;
    const int m = (zi - a.pre) / a.sf;
    if (m < 0 || m >= h) continue;
    for (int v = v0; v < a.kk; v += a.sf) {
      const int zj = j + v - r;
      if (zj < 0 || zj >= a.W) continue;
      const int n = (zj - a.pre) / a.sf;
      if (n < 0 || n >= w) continue;
      acc = fmaf(__ldg(a.up_taps + u * a.kk + v), __ldg(q + m * w + n), acc);
    }
  }
  return acc;
}

// Up(q) for the FLAIR blur operator (scale factor 4, 9 x 9 taps, 0 <= PRE < 4) for the FOUR adjacent outputs
// (i, j..j+3), j % 4 == 0, of one thread.  Same taps in the same (u, v) order as up_at (skipped taps enter as exact
// zeros), but: the four outputs share ONE 3 x 4 patch of q held in registers (12 loads per thread instead of 36
// predicated ones), and the <= 3 x 3 contributing taps of the four column phases come as one 16-byte shared load per
// (uu, vv) from a phase-major table s_tp[u0][uu][vv][k] = taps[u0 + 4 uu][v0(k) + 4 vv] (zero outside the 9 x 9
// support), v0(k) = (PRE - k) & 3.  PRE is a template parameter so that the patch column of every tap is a
// compile-time register index.  (The per-output version ran the fused update at 1.9 TB/s, issue-bound.)
template <int PRE>
__device__ __forceinline__ void up4_sf4k9(const UpdArgs& a, const float* __restrict__ s_tp, const float* __restrict__ q,
                                          int i, int j, float (&out)[4]) {
  const int h = a.H >> 2, w = a.W >> 2;
  const int u0 = (PRE + 4 - i) & 3;
  const int mb = (i + u0 - 4 - PRE) >> 2;  // LR row of tap row u0 (exact: the numerator is a multiple of 4)
  const int nb = (j >> 2) - 1;             // LR column of the left-most contributing tap of output j
  float Q[3][4];
#pragma unroll
  for (int uu = 0; uu < 3; ++uu) {
    const int m = mb + uu;
    const bool row_ok = m >= 0 && m < h;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int n = nb + cc;
      Q[uu][cc] = (row_ok && n >= 0 && n < w) ? __ldg(q + m * w + n) : 0.0f;
    }
  }
  const float4* tp = reinterpret_cast<const float4*>(s_tp) + u0 * 9;
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int uu = 0; uu < 3; ++uu) {
#pragma unroll
    for (int vv = 0; vv < 3; ++vv) {
      const float4 t = tp[uu * 3 + vv];
      const float tk[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int v0 = (PRE - k) & 3;
        const int off = (k + v0 - PRE) >> 2;   // first contributing LR column of output k: nb (+1); compile-time
        acc[k] = fmaf(tk[k], Q[uu][vv + off], acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = acc[k];
}

__global__ void __launch_bounds__(256) sampler_update_kernel(const __grid_constant__ UpdArgs a) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  __shared__ __align__(16) float s_tp[144];  // [u0][uu][vv][k] = taps[u0 + 4 uu][v0(k) + 4 vv] (0 outside the support)
  const bool fast_up = a.q_lr != nullptr && a.sf == 4 && a.kk == 9 && a.pre >= 0 && a.pre < 4 && (a.W & 3) == 0 &&
                       (a.H & 3) == 0;
  if (fast_up) {
    if (threadIdx.x < 144) {
      const int k = threadIdx.x & 3, r = (threadIdx.x >> 2) % 9, pu = threadIdx.x / 36;
      const int u = pu + 4 * (r / 3), v = ((a.pre - k) & 3) + 4 * (r % 3);
      s_tp[threadIdx.x] = (u < 9 && v < 9) ? __ldg(a.up_taps + u * 9 + v) : 0.0f;
    }
    __syncthreads();
  }
  // grid = (chunks of a plane, planes): no 64-bit division per element (plane = e / hw, i = off / W and j = off % W
  // were emulated 64-bit divisions, ~100 instructions each, in a kernel that otherwise needs ~60 per element)
  const int hw = a.H * a.W;
  const int plane = blockIdx.y;  // n*3 + c
  const int n = plane / 3, c = plane - 3 * n;
  for (int off = 4 * (blockIdx.x * blockDim.x + threadIdx.x); off < hw; off += 4 * gridDim.x * blockDim.x) {
    const long long e = static_cast<long long>(plane) * hw + off;
    const int t = a.t_arr ? static_cast<int>(__ldg(a.t_arr + n)) : a.t_host;
    const float ca = __ldg(a.coef + t * 8 + 0), cb = __ldg(a.coef + t * 8 + 1);
    const float cc = __ldg(a.coef + t * 8 + 2), cd = __ldg(a.coef + t * 8 + 3);
    const float gamma = a.gamma_arr ? __ldg(a.gamma_arr + n) : __ldg(a.coef + t * 8 + 4);
    const float nz = (t != 0) ? 1.0f : 0.0f;
    const float4 xt = __ldg(reinterpret_cast<const float4*>(a.x_t + e));
    float x[4] = {xt.x, xt.y, xt.z, xt.w};
    float x0[4];
    if (a.x0_in != nullptr) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.x0_in + e));
      x0[0] = g.x; x0[1] = g.y; x0[2] = g.z; x0[3] = g.w;
    } else {
      const float4 ep = __ldg(reinterpret_cast<const float4*>(
          a.model_out + (static_cast<long long>(n) * a.model_ch + c) * hw + off));
      const float eps[4] = {ep.x, ep.y, ep.z, ep.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)  // no FMA contraction: mirror the reference's mul, mul, sub
        x0[k] = clampf(__fsub_rn(__fmul_rn(ca, x[k]), __fmul_rn(cb, eps[k])), a.clip);
    }
    if (a.x0_in != nullptr) {
    } else if (a.R != nullptr) {
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.R + e));
      const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) x0[k] = clampf(__fsub_rn(x0[k], __fmul_rn(gamma, r[k])), a.clip);
    } else if (a.q_lr != nullptr) {
      const int i = off / a.W, j = off - i * a.W;
      const float* q = a.q_lr + static_cast<long long>(plane) * (hw / (a.sf * a.sf));
      if (fast_up) {
        float up[4];
        switch (a.pre) {
          case 0: up4_sf4k9<0>(a, s_tp, q, i, j, up); break;
          case 1: up4_sf4k9<1>(a, s_tp, q, i, j, up); break;
          case 2: up4_sf4k9<2>(a, s_tp, q, i, j, up); break;
          default: up4_sf4k9<3>(a, s_tp, q, i, j, up); break;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) x0[k] = clampf(__fsub_rn(x0[k], __fmul_rn(gamma, up[k])), a.clip);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          x0[k] = clampf(__fsub_rn(x0[k], __fmul_rn(gamma, up_at(a, q, i, j + k))), a.clip);
      }
    }
    if (a.prev != nullptr && a.x0_in == nullptr) {
      const int f = n % a.frames_per_window, b = n / a.frames_per_window;
      if (f < a.prev_k) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(
            a.prev + ((static_cast<long long>(b) * a.prev_k + f) * 3 + c) * hw + off));
        x0[0] = p4.x; x0[1] = p4.y; x0[2] = p4.z; x0[3] = p4.w;
      }
    }
    const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.noise + e));
    const float z[4] = {z4.x, z4.y, z4.z, z4.w};
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float eh = __fdiv_rn(__fsub_rn(__fmul_rn(ca, x[k]), x0[k]), cb);
      const float t1 = __fmul_rn(__fmul_rn(a.s1, cd), eh);
      const float t2 = __fmul_rn(__fmul_rn(a.s2, cd), z[k]);
      s[k] = __fadd_rn(__fmul_rn(cc, x0[k]), __fmul_rn(nz, __fadd_rn(t1, t2)));
    }
    *reinterpret_cast<float4*>(a.sample + e) = make_float4(s[0], s[1], s[2], s[3]);
    if (a.pred_xstart != nullptr)
      *reinterpret_cast<float4*>(a.pred_xstart + e) = make_float4(x0[0], x0[1], x0[2], x0[3]);
  }
}

__global__ void __launch_bounds__(256)
pred_xstart_kernel(const float* __restrict__ x_t, const float* __restrict__ model_out, int model_ch,
                   const float* __restrict__ coef, const long long* t_arr, int t_host,
                   float* __restrict__ x0, int N, long long hw, int clip) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long total4 = static_cast<long long>(N) * 3 * hw / 4;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total4;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = v * 4;
    const long long plane = e / hw;
    const int n = static_cast<int>(plane / 3), c = static_cast<int>(plane % 3);
    const long long off = e - plane * hw;
    const int t = t_arr ? static_cast<int>(__ldg(t_arr + n)) : t_host;
    const float ca = __ldg(coef + t * 8 + 0), cb = __ldg(coef + t * 8 + 1);
    const float4 xt = __ldg(reinterpret_cast<const float4*>(x_t + e));
    const float4 ep = __ldg(reinterpret_cast<const float4*>(
        model_out + (static_cast<long long>(n) * model_ch + c) * hw + off));
    float4 o;
    o.x = clampf(__fsub_rn(__fmul_rn(ca, xt.x), __fmul_rn(cb, ep.x)), clip);
    o.y = clampf(__fsub_rn(__fmul_rn(ca, xt.y), __fmul_rn(cb, ep.y)), clip);
    o.z = clampf(__fsub_rn(__fmul_rn(ca, xt.z), __fmul_rn(cb, ep.z)), clip);
    o.w = clampf(__fsub_rn(__fmul_rn(ca, xt.w), __fmul_rn(cb, ep.w)), clip);
    *reinterpret_cast<float4*>(x0 + e) = o;
  }
}

__global__ void __launch_bounds__(256)
axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float alpha, float beta,
             float* __restrict__ out, long long n4) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < n4;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + v);
    const float4 b = __ldg(reinterpret_cast<const float4*>(y) + v);
    reinterpret_cast<float4*>(out)[v] =
        make_float4(__fadd_rn(__fmul_rn(alpha, a.x), __fmul_rn(beta, b.x)),
                    __fadd_rn(__fmul_rn(alpha, a.y), __fmul_rn(beta, b.y)),
                    __fadd_rn(__fmul_rn(alpha, a.z), __fmul_rn(beta, b.z)),
                    __fadd_rn(__fmul_rn(alpha, a.w), __fmul_rn(beta, b.w)));
  }
}

int ew_grid(long long work_items) {
  long long blocks = ceil_div_ll(work_items, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

extern "C" int flair_sampler_update_f32(const flair_update_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(p && p->x_t && (p->model_out || p->x0_in) && p->noise && p->coef && p->sample,
                "flair_sampler_update_f32: null pointer");
  FLAIR_REQUIRE(p->x0_in || p->model_ch == 3 || p->model_ch == 6,
                "flair_sampler_update_f32: model_ch must be 3 or 6");
  FLAIR_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->W % 4 == 0,
                "flair_sampler_update_f32: W must be a positive multiple of 4");
  FLAIR_REQUIRE(!(p->R && p->q_lr), "flair_sampler_update_f32: give R or q_lr, not both");
  if (p->q_lr)
    FLAIR_REQUIRE(p->up_taps && p->up_k > 0 && p->sf > 0 && p->H % p->sf == 0 && p->W % p->sf == 0,
                  "flair_sampler_update_f32: bad blur-operator arguments");
  if (p->q_lr)  // the replicate padding of the zero-inserted image must be all-zero (pseudoSR.py:199-225)
    FLAIR_REQUIRE(p->pre_stride != 0 && p->pre_stride != p->sf - 1,
                  "flair_sampler_update_f32: polyphase Up needs 0 < pre_stride < sf-1");
  if (p->prev)
    FLAIR_REQUIRE(p->prev_k > 0 && p->frames_per_window > 0 && p->N % p->frames_per_window == 0,
                  "flair_sampler_update_f32: bad prev_recon arguments");
  UpdArgs a{};
  a.x_t = p->x_t; a.model_out = p->model_out; a.model_ch = p->model_ch; a.noise = p->noise;
  a.R = p->R; a.q_lr = p->q_lr; a.up_taps = p->up_taps; a.kk = p->up_k; a.sf = p->sf; a.pre = p->pre_stride;
  a.prev = p->prev; a.prev_k = p->prev_k; a.frames_per_window = p->frames_per_window;
  a.coef = p->coef; a.t_arr = p->t_arr; a.gamma_arr = p->gamma_arr; a.x0_in = p->x0_in; a.t_host = p->t;
  a.s1 = p->sqrt_one_minus_rho; a.s2 = p->sqrt_rho;
  a.sample = p->sample; a.pred_xstart = p->pred_xstart;
  a.N = p->N; a.H = p->H; a.W = p->W; a.clip = p->clip_denoised;
  FLAIR_REQUIRE(static_cast<long long>(p->H) * p->W < (1ll << 30) && p->N * 3 <= 65535,
                "flair_sampler_update_f32: plane too large / too many planes (H=%d W=%d N=%d)", p->H, p->W, p->N);
  const int chunks = ceil_div(p->H * p->W / 4, 256);
  FLAIR_CHECK_CUDA(flair_launch(sampler_update_kernel, dim3(chunks, p->N * 3), dim3(256), 0, stream, a));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_pred_xstart_f32(const float* x_t, const float* model_out, int model_ch,
                                     const float* coef, const long long* t_arr, int t, float* x0, int N,
                                     int H, int W, int clip_denoised, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x_t && model_out && coef && x0, "flair_pred_xstart_f32: null pointer");
  FLAIR_REQUIRE((static_cast<long long>(H) * W) % 4 == 0, "flair_pred_xstart_f32: H*W must be a multiple of 4");
  const long long hw = static_cast<long long>(H) * W;
  FLAIR_CHECK_CUDA(flair_launch(pred_xstart_kernel, dim3(ew_grid(static_cast<long long>(N) * 3 * hw / 4)), dim3(256), 0, stream, 
      x_t, model_out, model_ch, coef, t_arr, t, x0, N, hw, clip_denoised));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_axpby_f32(const float* x, const float* y, float alpha, float beta, float* out,
                               long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && y && out && n % 4 == 0, "flair_axpby_f32: null pointer or n not a multiple of 4");
  FLAIR_CHECK_CUDA(flair_launch(axpby_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, stream, x, y, alpha, beta, out, n / 4));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------
// Pieces needed only off the lean path (API parity of p_mean_variance, aux-prior branch).
namespace {

// x0 <- clamp(x0 - gamma[n] * R)            (gaussian_diffusion.py:465-470)
__global__ void __launch_bounds__(256)
dc_apply_kernel(const float* __restrict__ x0, const float* __restrict__ R, const float* __restrict__ gamma_arr,
                float gamma_host, float* __restrict__ out, long long hw3, long long total4, int clip) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < total4;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = v * 4;
    const float g = gamma_arr ? __ldg(gamma_arr + e / hw3) : gamma_host;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + e));
    const float4 r = __ldg(reinterpret_cast<const float4*>(R + e));
    float4 o;
    o.x = clampf(__fsub_rn(a.x, __fmul_rn(g, r.x)), clip);
    o.y = clampf(__fsub_rn(a.y, __fmul_rn(g, r.y)), clip);
    o.z = clampf(__fsub_rn(a.z, __fmul_rn(g, r.z)), clip);
    o.w = clampf(__fsub_rn(a.w, __fmul_rn(g, r.w)), clip);
    *reinterpret_cast<float4*>(out + e) = o;
  }
}

// posterior mean and (learned-range / fixed) variance  (gaussian_diffusion.py:226-248,278-309)
// tab: [steps][8] = {coef1, coef2, min_log (posterior_log_variance_clipped), max_log (log beta),
//                    fixed_var, fixed_log_var, 0, 0}
__global__ void __launch_bounds__(256)
mean_variance_kernel(const float* __restrict__ x_t, const float* __restrict__ x0,
                     const float* __restrict__ var_values, int model_ch, const float* __restrict__ tab,
                     const long long* __restrict__ t_arr, int t_host, float* __restrict__ mean,
                     float* __restrict__ variance, float* __restrict__ log_variance, int N, long long hw) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long total = static_cast<long long>(N) * 3 * hw;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long plane = e / hw;
    const int n = static_cast<int>(plane / 3), c = static_cast<int>(plane % 3);
    const int t = t_arr ? static_cast<int>(__ldg(t_arr + n)) : t_host;
    const float* row = tab + t * 8;
    mean[e] = __fadd_rn(__fmul_rn(row[0], x0[e]), __fmul_rn(row[1], x_t[e]));
    float lv, vv;
    if (var_values != nullptr) {
      const float v = __ldg(var_values + (static_cast<long long>(n) * model_ch + 3 + c) * hw + (e - plane * hw));
      const float frac = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
      lv = __fadd_rn(__fmul_rn(frac, row[3]), __fmul_rn(__fsub_rn(1.0f, frac), row[2]));
      vv = expf(lv);
    } else {
      vv = row[4];
      lv = row[5];
    }
    variance[e] = vv;
    log_variance[e] = lv;
  }
}

}  // namespace

extern "C" int flair_dc_apply_f32(const float* x0, const float* R, const float* gamma_arr, float gamma,
                                  float* out, int N, int H, int W, int clip_denoised, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x0 && R && out, "flair_dc_apply_f32: null pointer");
  const long long hw3 = 3LL * H * W;
  FLAIR_REQUIRE(hw3 % 4 == 0, "flair_dc_apply_f32: 3*H*W must be a multiple of 4");
  const long long total4 = static_cast<long long>(N) * hw3 / 4;
  FLAIR_CHECK_CUDA(flair_launch(dc_apply_kernel, dim3(ew_grid(total4)), dim3(256), 0, stream, x0, R, gamma_arr, gamma, out, hw3, total4,
                                                       clip_denoised));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_mean_variance_f32(const float* x_t, const float* x0, const float* model_out,
                                       int model_ch, int learned_range, const float* tab,
                                       const long long* t_arr, int t, float* mean, float* variance,
                                       float* log_variance, int N, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x_t && x0 && tab && mean && variance && log_variance, "flair_mean_variance_f32: null pointer");
  if (learned_range) FLAIR_REQUIRE(model_out && model_ch == 6, "flair_mean_variance_f32: learned range needs 6 channels");
  const long long hw = static_cast<long long>(H) * W;
  FLAIR_CHECK_CUDA(flair_launch(mean_variance_kernel, dim3(ew_grid(static_cast<long long>(N) * 3 * hw)), dim3(256), 0, stream, 
      x_t, x0, learned_range ? model_out : nullptr, model_ch, tab, t_arr, t, mean, variance, log_variance, N, hw));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
