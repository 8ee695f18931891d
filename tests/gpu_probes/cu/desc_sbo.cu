// GPU probe (round 2): one halo slab for all nine taps of a 3x3 convolution.  The slab is an 18 (h) x 10 (w) pixel
// window stored row-major as 180 rows of 128 B in the physical SWIZZLE_128B layout TMA produces.  The A operand of
// tap (dh, dw) for the 16 (h) x 8 (w) output tile is described by ONE K-major descriptor: start address =
// slab + ((1+dh)*10 + (1+dw)) * 128 B, stride between 8-row groups (SBO) = 10 * 128 = 1280 B.  Exact comparison.
#include <cstdio>
#include <vector>
#include "../../../flair_b200/csrc/common.cuh"
void flair_set_error(const char*, ...) {}
__device__ __forceinline__ float av(int r, int k) { return float(((r * 5 + k * 3) % 11) - 5); }
__device__ __forceinline__ float bv(int n, int k) { return float(((n * 7 + k) % 9) - 4); }

__global__ void __launch_bounds__(128, 1) desc_sbo(int* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(smem);                 // 184 rows x 128 B
  __half* B = reinterpret_cast<__half*>(smem + 184 * 128);      // 64 rows x 128 B (23552 = 23 * 1024: aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 184 * 64; i += blockDim.x) {
    const int r = i / 64, k = i % 64;
    A[r * 64 + (((k / 8) ^ (r % 8)) * 8) + (k % 8)] = __float2half(av(r, k));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i / 64, k = i % 64;
    B[n * 64 + (((k / 8) ^ (n % 8)) * 8) + (k % 8)] = __float2half(bv(n, k));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = umma_idesc_f16(128, 64, 0);
  uint32_t phase = 0;
  for (int tap = 0; tap < 9; ++tap) {
    const int dh = tap / 3 - 1, dw = tap % 3 - 1;
    const int row0 = (1 + dh) * 10 + (1 + dw);
    if (warp == 0) {
      if (elect_one()) {
        uint64_t adesc = umma_desc_sw128(smem_u32(smem) + row0 * 128);
        adesc = (adesc & ~(uint64_t(0x3FFF) << 32)) | (uint64_t(1280 >> 4) << 32);   // SBO = 1280 B
        const uint64_t bdesc = umma_desc_sw128(smem_u32(B));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tm, adesc + 2u * k, bdesc + 2u * k, idesc, k != 0);
        umma_commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    const int i = warp * 32 + lane;
    const int srow = (i / 8 + 1 + dh) * 10 + (i % 8) + 1 + dw;
    int bad = 0;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tm + (uint32_t(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) {
        float e = 0.f;
        for (int k = 0; k < 64; ++k) e += av(srow, k) * bv(c0 + j, k);
        if (__uint_as_float(r[j]) != e) ++bad;
      }
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if (lane == 0) atomicAdd(&out[tap], bad);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  int* d_out;
  cudaMalloc(&d_out, 9 * sizeof(int));
  cudaMemset(d_out, 0, 9 * sizeof(int));
  cudaFuncSetAttribute(desc_sbo, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  desc_sbo<<<1, 128, 184 * 128 + 64 * 128 + 1024>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  int h[9];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("8x16 tile from ONE 10x18 halo slab, SBO = 1280 B: mismatching accumulator elements (of 8192) per tap\n");
  for (int t = 0; t < 9; ++t) printf("tap (dh=%+d, dw=%+d): %d\n", t / 3 - 1, t % 3 - 1, h[t]);
  return 0;
}
