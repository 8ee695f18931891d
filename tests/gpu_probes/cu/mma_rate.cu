// GPU probe (round 2): issue-to-retire cycles per tcgen05.mma (kind::f16, K = 16) as a function of the instruction
// shape and of where the A operand lives (shared memory "SS" / tensor memory "TS"), plus the smem->TMEM copy rate
// (tcgen05.cp 128x256b).  Decides the operand orientation of the small-N convolutions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <vector>
#include "../../../flair_b200/csrc/common.cuh"

void flair_set_error(const char*, ...) {}

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void utccp_128x256b(uint32_t tmem_dst, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
}

// mode 0: SS MMAs; 1: TS MMAs (A in TMEM); 2: tcgen05.cp only; 3: per k-block 4 x cp + 4 x TS MMA (A staged through TMEM)
__global__ void __launch_bounds__(128, 1) mma_rate(int M, int N, int iters, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (16 + 32 + 16) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    const uint64_t adesc = umma_desc_sw128(smem_u32(smem));
    const uint64_t a2desc = umma_desc_sw128(smem_u32(smem + 48 * 1024));
    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + 16 * 1024));
    const uint32_t idesc = umma_idesc_f16(M, N, 0);
    const uint32_t a_tm = tm + 256;  // A operand in TMEM: 8 columns per K = 16 step
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (mode == 2 || mode == 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) utccp_128x256b(a_tm + ((it & 1) * 32 + 8 * k), ((it & 1) ? a2desc : adesc) + 2u * k);
        }
        if (mode == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tm, adesc + 2u * k, bdesc + 2u * k, idesc, (it | k) != 0);
        } else if (mode == 1 || mode == 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(tm, a_tm + ((it & 1) * 32 + 8 * k), bdesc + 2u * k, idesc, (it | k) != 0);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4000;
  struct Cfg { int M, N, mode; const char* what; };
  std::vector<Cfg> cfgs;
  for (int N : {32, 64, 96, 128, 144, 192, 224, 256}) cfgs.push_back({128, N, 0, "SS  M=128"});
  for (int N : {64, 128, 256}) cfgs.push_back({64, N, 0, "SS  M=64 "});
  for (int N : {32, 64, 128, 256}) cfgs.push_back({128, N, 1, "TS  M=128"});
  cfgs.push_back({128, 64, 2, "cp only  "});
  for (int N : {64, 128, 256}) cfgs.push_back({128, N, 3, "cp+TS    "});
  for (int grid : {1, 148}) {
    printf("---- grid = %d CTAs (one per SM), %d iterations x 4 K-steps\n", grid, iters);
    for (const Cfg& c : cfgs) {
      mma_rate<<<grid, 128, 66 * 1024 + 1024>>>(c.M, c.N, iters, c.mode, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s N=%d: %s\n", c.what, c.N, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
      const double cyc = double(mx) / (iters * 4.0);
      const double macs = double(c.M) * c.N * 16;
      printf("%s N=%3d: %7.1f cycles per K=16 step   %6.0f MAC/clk/SM  (%.0f %% of 4096)\n", c.what, c.N, cyc,
             c.mode == 2 ? 0.0 : macs / cyc, c.mode == 2 ? 0.0 : 100.0 * macs / cyc / 4096.0);
    }
  }
  return 0;
}
