// GPU probe (round 2): what slows the conv kernel's MMAs from 48 cycles per K step (isolated, N = 64) to ~78 in situ?
// One CTA per SM issues the conv's exact MMA sequence (9 taps x 4 K steps per tile from one 10 x 18 halo slab,
// resident weights [tap] of 8 KB) with optional concurrent activity:
//   variant 0: aligned A (start multiple of 1024 B, SBO 1024), B fixed            (the isolated-rate baseline)
//   variant 1: halo A (row-shifted start, SBO 1280), B fixed
//   variant 2: halo A, B per tap (resident weights)
//   variant 3: variant 2 + 8 "epilogue" warps reading the OTHER accumulator with tcgen05.ld in a loop
//   variant 4: variant 2 + one warp streaming TMA-like bulk copies (cp.async.bulk global->shared, 23 KB per tile)
//   variant 5: variant 2 with a tcgen05.commit + mbarrier wait per 36 MMAs (stage hand-shake cadence)
#include <cstdio>
#include <vector>
#include "../../../flair_b200/csrc/common.cuh"
void flair_set_error(const char*, ...) {}

__global__ void __launch_bounds__(320, 1) conv_like(int variant, int N, int tiles, const uint8_t* gsrc, long long* out, int randomize) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab = smem;                       // 2 x 23 KB halo slabs
  uint8_t* wts = smem + 2 * 23552;            // 9 x N x 128 B weights
  uint8_t* sink = wts + 9 * N * 128;          // 23 KB bulk-copy sink
  __shared__ uint64_t bar, bar2, cbar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < (2 * 23552 + 9 * N * 128) / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;  // fp16 (1.0, 1.0)
    if (randomize) {           // pseudo-random fp16 pairs in (-1, 1): realistic operand toggling
      uint32_t h = (i + 1) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const __half2 r = __floats2half2_rn(((h & 0xFFFF) / 32768.0f) - 1.0f, ((h >> 16) / 32768.0f) - 1.0f);
      v = *reinterpret_cast<const uint32_t*>(&r);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&cbar, 1); mbar_fence_init(); stop = 0; }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(128, N, 0);
    const uint64_t hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint64_t hi_slab = (hi & ~(uint64_t(0x3FFF) << 32)) | (uint64_t(1280 >> 4) << 32);
    const uint32_t lo_flags = uint32_t(umma_desc_sw128(0));
    const uint32_t slab_lo = (smem_u32(slab) & 0x3FFFF) >> 4, w_lo = (smem_u32(wts) & 0x3FFFF) >> 4;
    const uint32_t wstep = uint32_t(N * 128) >> 4;
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int t = 0; t < tiles; ++t) {
      const uint32_t d = tm + (t & 1) * 256;
      const uint32_t sl = slab_lo + (t & 1) * (23552 >> 4);
      if (elect_one()) {
        uint32_t accum = 0;
#pragma unroll 1
        for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const uint32_t al = lo_flags | (variant == 0 ? sl + uint32_t(dh * 3 + dw) * 64u : sl + uint32_t(dh * 10 + dw) * 8u);
            const uint32_t bl = lo_flags | (variant >= 2 ? w_lo + uint32_t(dh * 3 + dw) * wstep : w_lo);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d, (variant == 0 ? hi : hi_slab) | (al + 2u * k), hi | (bl + 2u * k), idesc, accum);
              accum = 1;
            }
          }
        }
        if (variant == 5) umma_commit(&cbar);
      }
      __syncwarp();
      if (variant == 5) { mbar_wait(&cbar, ph); ph ^= 1; tc_fence_after(); }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
  } else if (warp >= 2 && variant == 3) {
    // epilogue-like: read 64 columns of the other accumulator, over and over
    const uint32_t ta = tm + (uint32_t((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    while (!stop) {
      for (int c = 0; c < 64; c += 16) {
        uint32_t r[16];
        tmem_ld16(ta + 256 + c, r);
        tmem_ld_wait();
        acc += r[0] + r[7];
      }
    }
    if (acc == 0x12345) out[200] = acc;
  } else if (variant >= 6 && variant <= 9 && (warp >= 2 || warp == 0)) {
    // 9 warps wait on an mbarrier that is never signalled while the MMAs run (what the epilogue / producer warps of
    // the conv kernel do): 6 = plain try_wait spin, 7 = spin with __nanosleep(100), 8 = try_wait with a 2 us
    // suspend-time hint, 9 = one lane per warp spins, the others wait at __syncwarp
    const uint32_t addr = smem_u32(&bar2);
    while (!stop) {
      uint32_t ok = 0;
      if (variant == 8) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(0u), "r"(2000u) : "memory");
      } else if (variant == 9) {
        if (lane == 0) ok = mbar_try_wait(&bar2, 0) ? 1u : 0u;
        __syncwarp();
      } else {
        ok = mbar_try_wait(&bar2, 0) ? 1u : 0u;
        if (variant == 7) __nanosleep(100);
      }
      if (ok) break;
    }
  } else if (warp == 0 && variant == 4) {
    // TMA-like traffic: bulk async copies global -> shared, 23 KB per "tile", paced by their own mbarrier
    if (lane == 0) {
      uint32_t ph = 0;
      int i = 0;
      while (!stop) {
        mbar_expect_tx(&bar2, 23040);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sink)),
                     "l"(gsrc + (size_t(blockIdx.x) * 64 + (i & 63)) * 23040), "r"(23040u), "r"(smem_u32(&bar2))
                     : "memory");
        mbar_wait(&bar2, ph);
        ph ^= 1;
        ++i;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out;
  uint8_t* gsrc;
  cudaMalloc(&d_out, 256 * sizeof(long long));
  cudaMalloc(&gsrc, size_t(148) * 64 * 23040);
  cudaMemset(gsrc, 0, size_t(148) * 64 * 23040);
  cudaFuncSetAttribute(conv_like, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  const int tiles = 400;
  const char* names[10] = {"aligned A, B fixed", "halo A (SBO 1280), B fixed", "halo A, B per tap", "  + 8 warps tcgen05.ld",
                          "  + bulk-copy stream", "  + commit/wait per tile", "  + 9 warps try_wait spin", "  + 9 warps spin+nanosleep",
                          "  + 9 warps try_wait w/ hint", "  + 9 warps, 1 lane spins"};
  for (int randomize = 1; randomize < 2; ++randomize) {
    printf("---- operands: %s\n", randomize ? "pseudo-random fp16 in (-1, 1)" : "constant 1.0");
    for (int N : {64, 128}) {
      for (int v = 0; v < 10; ++v) {
        if (N == 128 && v != 2 && v < 6) continue;
        const size_t smem = 3 * 23552 + 9 * size_t(N) * 128 + 2048;
        conv_like<<<148, 320, smem>>>(v, N, tiles, gsrc, d_out, randomize);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d variant %d: %s\n", N, v, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1LL << 60; for (long long x : h) { mx = x > mx ? x : mx; mn = x < mn ? x : mn; }
        printf("N=%3d %-28s %6.1f (slowest SM) %6.1f (fastest SM) cycles per K=16 MMA (ideal %d)\n", N, names[v],
               double(mx) / (tiles * 36.0), double(mn) / (tiles * 36.0), N / 2 > (4096 + N * 32) / 128 ? N / 2 : (4096 + N * 32) / 128);
      }
    }
  }
  return 0;
}
