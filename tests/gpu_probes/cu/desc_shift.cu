// GPU probe (round 2): can a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a multiple of 8 rows
// (1024 B) when the descriptor's "matrix base offset" field (bits 49-51) is set?  A 160-row slab is written in the
// physical layout TMA produces (16-byte chunk index XOR (row % 8)); for every row shift s = 0..9 and every
// base-offset value 0..7 the product D = A[s : s+128] * B^T is computed by tcgen05.mma and compared exactly
// (small integers) with a scalar evaluation.  Prints the mismatch count per (shift, base offset).
#include <cstdio>
#include <vector>
#include "../../../flair_b200/csrc/common.cuh"
void flair_set_error(const char*, ...) {}

__device__ __forceinline__ float av(int r, int k) { return float(((r * 5 + k * 3) % 11) - 5); }
__device__ __forceinline__ float bv(int n, int k) { return float(((n * 7 + k) % 9) - 4); }

__global__ void __launch_bounds__(128, 1) desc_shift(int* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(smem);                 // 160 rows x 128 B
  __half* B = reinterpret_cast<__half*>(smem + 160 * 128);      // 64 rows x 128 B (20480 = 20 * 1024: aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 64; i += blockDim.x) {
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
  for (int s = 0; s < 10; ++s) {
    for (int bo = 0; bo < 8; ++bo) {
      if (warp == 0) {
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem) + s * 128) | (uint64_t(bo) << 49);
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
      const int row = warp * 32 + lane;
      int bad = 0;
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tm + (uint32_t(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) {
          float e = 0.f;
          for (int k = 0; k < 64; ++k) e += av(row + s, k) * bv(c0 + j, k);
          if (__uint_as_float(r[j]) != e) ++bad;
        }
      }
      for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
      if (lane == 0) atomicAdd(&out[s * 8 + bo], bad);
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  int* d_out;
  cudaMalloc(&d_out, 80 * sizeof(int));
  cudaMemset(d_out, 0, 80 * sizeof(int));
  cudaFuncSetAttribute(desc_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  desc_shift<<<1, 128, 160 * 128 + 64 * 128 + 1024>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int> h(80);
  cudaMemcpy(h.data(), d_out, 80 * sizeof(int), cudaMemcpyDeviceToHost);
  printf("mismatching accumulator elements (of 8192) for A start = slab + shift rows, by descriptor base_offset\n");
  printf("shift | bo=0   bo=1   bo=2   bo=3   bo=4   bo=5   bo=6   bo=7\n");
  for (int s = 0; s < 10; ++s) {
    printf("%5d |", s);
    for (int bo = 0; bo < 8; ++bo) printf(" %5d ", h[s * 8 + bo]);
    printf("\n");
  }
  return 0;
}
