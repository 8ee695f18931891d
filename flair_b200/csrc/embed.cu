// Timestep embedding, tiny fp32 Linears of the conditioning path, and the input packer.
//
// Replaces guided_diffusion/nn_new.py:103-121 (timestep_embedding), unet_new.py:979-984
// (time_embed MLP) and the 86 per-ResBlock `emb_layers` Linears (unet_new.py:258-264) — all
// fp32 in the reference, launched as 137 tiny GEMMs per step; here the per-block Linears are
// concatenated into ONE weight matrix and one launch — and the `th.cat([x, low_res_input])`
// + first 3x3 conv input path (unet_new.py:1330-1331,993): the 6-channel fp32 NCHW inputs are
// packed straight into a 64-wide channels-last im2col map so the first conv is a K=64 GEMM.
#include "common.cuh"
#include "../../include/flair_b200.h"

namespace {

__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ out, int N, int half) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * half) return;
  const int n = i / half, j = i % half;
  const float arg = __fmul_rn(__ldg(t + n), __ldg(freqs + j));
  out[static_cast<long long>(n) * 2 * half + j] = cosf(arg);
  out[static_cast<long long>(n) * 2 * half + half + j] = sinf(arg);
}

// y[m][n] = act_out( bias[n] + sum_k act_in(x[m][k]) * Wt[k][n] ),  M small (<= 64 per pass of 16)
template <int MB>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias,
                  float* __restrict__ y, int M, int K, int N, int silu_in, int silu_out) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  extern __shared__ float xs[];  // [MB][K]
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  for (int m0 = 0; m0 < M; m0 += MB) {
    const int mb = (M - m0 < MB) ? (M - m0) : MB;
    __syncthreads();
    for (int i = threadIdx.x; i < MB * K; i += blockDim.x) {
      const int mm = i / K;
      float v = (mm < mb) ? __ldg(x + static_cast<long long>(m0 + mm) * K + (i % K)) : 0.f;
      if (silu_in) v = v / (1.0f + expf(-v));
      xs[i] = v;
    }
    __syncthreads();
    if (n < N) {
      float acc[MB];
#pragma unroll
      for (int mm = 0; mm < MB; ++mm) acc[mm] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float w = __ldg(Wt + static_cast<long long>(k) * N + n);
#pragma unroll
        for (int mm = 0; mm < MB; ++mm) acc[mm] = fmaf(xs[mm * K + k], w, acc[mm]);
      }
      const float bb = bias ? __ldg(bias + n) : 0.f;
      for (int mm = 0; mm < mb; ++mm) {
        float v = acc[mm] + bb;
        if (silu_out == 1) v = v / (1.0f + expf(-v));
        else if (silu_out == 2) v = 1.0f / (1.0f + expf(-v));
        y[static_cast<long long>(m0 + mm) * N + n] = v;
      }
    }
  }
}

// out[n][h][w][k], k = tap*6 + c (tap = (dh+1)*3 + (dw+1)); channels c<3 from `a`, c>=3 from `b`.
__global__ void __launch_bounds__(256)
pack_im2col6_kernel(const float* __restrict__ a, const float* __restrict__ b, uint16_t* __restrict__ out, int N,
                    int H, int W, int dtype) {
  pdl_sync();  // PDL: release the next launch, then wait for the previous kernel's results
  const long long hw = static_cast<long long>(H) * W;
  const long long items = static_cast<long long>(N) * hw * 8;
  for (long long it = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(it & 7);
    const long long pix = it >> 3;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const long long n = pix / hw;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      float val = 0.f;
      if (k < 54) {
        const int tap = k / 6, c = k % 6;
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          const float* src = (c < 3) ? a : b;
          val = __ldg(src + (n * 3 + (c % 3)) * hw + static_cast<long long>(hh) * W + ww);
        }
      }
      f[j] = val;
    }
    uint4 u;
    if (dtype == FLAIR_F16) {
      __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
      __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    } else {
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
      u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    }
    *reinterpret_cast<uint4*>(out + pix * 64 + v * 8) = u;
  }
}

}  // namespace

extern "C" int flair_timestep_embedding_f32(const float* t, const float* freqs, float* out, int N, int dim,
                                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(t && freqs && out && N > 0 && dim > 0 && dim % 2 == 0, "flair_timestep_embedding_f32: bad arguments");
  const int half = dim / 2;
  FLAIR_CHECK_CUDA(flair_launch(timestep_embedding_kernel, dim3(ceil_div(N * half, 128)), dim3(128), 0, stream, t, freqs, out, N, half));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_linear_f32(const float* x, const float* Wt, const float* bias, float* y, int M, int K, int N,
                                int silu_in, int silu_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(x && Wt && y && M > 0 && K > 0 && N > 0, "flair_linear_f32: bad arguments");
  constexpr int MB = 16;
  const size_t smem = sizeof(float) * MB * K;
  FLAIR_REQUIRE(smem <= 96 * 1024, "flair_linear_f32: K=%d too large", K);
  static FlairPerDeviceOnce attr;
  if (attr.first())
    FLAIR_CHECK_CUDA(cudaFuncSetAttribute(linear_f32_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  FLAIR_CHECK_CUDA(flair_launch(linear_f32_kernel<MB>, dim3(ceil_div(N, 256)), dim3(256), smem, stream, x, Wt, bias, y, M, K, N, silu_in, silu_out));
  FLAIR_CHECK_LAUNCH();
  return 0;
}

extern "C" int flair_pack_im2col6(const float* a, const float* b, void* out, int N, int H, int W, int dtype,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FLAIR_REQUIRE(a && b && out && N > 0, "flair_pack_im2col6: bad arguments");
  const long long items = static_cast<long long>(N) * H * W * 8;
  long long blocks = ceil_div_ll(items, 256);
  const long long cap = static_cast<long long>(flair_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  FLAIR_CHECK_CUDA(flair_launch(pack_im2col6_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, a, b, static_cast<uint16_t*>(out), N, H, W, dtype));
  FLAIR_CHECK_LAUNCH();
  return 0;
}
