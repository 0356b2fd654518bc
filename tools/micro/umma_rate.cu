// Micro-benchmark: issue rate of small tcgen05.mma (M=128, K=16, bf16) instructions from shared-memory operands.
// Varies N, the number of TMEM accumulators the instruction stream alternates between, and the swizzle mode of
// the (synthetic) descriptors.  One CTA per SM-count option; prints clk / MMA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../causal-gen_b200/csrc -I../../include umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include "cg_common.cuh"

unsigned long long* cg_tl_ptr = nullptr;
void cg_set_error(const char*, ...) {}
int cg_require_sm100() { return 0; }

__device__ __forceinline__ uint64_t desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// mode 0: SWIZZLE_NONE K-major, A planes as in conv_tc (lbo = 2880 plane, sbo = 160), taps = shifted starts
// mode 1: SWIZZLE_NONE K-major dense (lbo = 2048, sbo = 128)
// mode 2: SWIZZLE_128B K-major (sbo = 1024)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int nacc, int mode, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(cg_smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(cg_smem_u32(&slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t a0 = cg_smem_u32(smem), b0 = cg_smem_u32(smem + 96 * 1024);
    uint64_t ad, bd;
    if (mode == 0) { ad = desc_sw(a0, 2880, 160, 0); bd = desc_sw(b0, N * 16, 128, 0); }
    else if (mode == 1) { ad = desc_sw(a0, 2048, 128, 0); bd = desc_sw(b0, N * 16, 128, 0); }
    else { ad = desc_sw(a0, 16, 1024, 2); bd = desc_sw(b0, 16, 1024, 2); }
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    // warm-up
    for (int r = 0; r < 16; ++r) tc_mma_bf16(tmem + (r & (nacc - 1)) * N, ad, bd, idesc, 1);
    tc_commit(cg_smem_u32(&bar));
    mbar_wait(cg_smem_u32(&bar), 0);
    long long t0 = clock64();
    const uint32_t step = mode == 2 ? 0u : 1u;
    for (int r = 0; r < reps; r += 8) {
      // advance the A start address like the taps do (units of 16 B), stay inside the buffer
#pragma unroll
      for (int u = 0; u < 8; ++u)
        tc_mma_bf16(tmem + ((r + u) & (nacc - 1)) * N, ad + (uint64_t)((u / 3) * 10 + (u % 3)) * step, bd, idesc, 1);
    }
    long long t1 = clock64();
    tc_commit(cg_smem_u32(&bar));
    mbar_wait(cg_smem_u32(&bar), 1);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 2048;
  printf("mode N nacc | clk/MMA(issue) clk/MMA(complete)\n");
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        rate_kernel<<<1, 128, 200 * 1024>>>(N, nacc, mode, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d nacc %d: %s\n", mode, N, nacc, cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%d %3d %d | %7.1f %7.1f\n", mode, N, nacc, (double)h[0] / reps, (double)h[1] / reps);
      }
  return 0;
}
