// Micro-benchmark 2: tcgen05.mma (M=128, K=16, bf16, SWIZZLE_NONE halo-tile A operand) issued the way conv_tc.cu issues it:
// 9 taps x 2 K-blocks per 11520-byte stage, stages walked round a ring, one tcgen05.commit per stage and per tile.
// Reports clk / MMA for N in {16, 32, 64}, on 1 CTA and on 148 CTAs at once (chip-wide effects), with and without 8
// epilogue-like warps draining the accumulator through tcgen05.ld, and reads %globaltimer to give the SM clock.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../causal-gen_b200/csrc -I../../include umma_rate2.cu -o umma_rate2
#include <cstdio>
#include <cstdlib>
#include "cg_common.cuh"

unsigned long long* cg_tl_ptr = nullptr;
void cg_set_error(const char*, ...) {}
int cg_require_sm100() { return 0; }

constexpr int kStage = 11520, kStages = 10;

__global__ void __launch_bounds__(320, 1) k2(int N, int tiles, int chunks, int drain, int sbo, int cmode, int vmode, int issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_stage[kStages], bar_end[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kStages * kStage + 81920) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(cg_smem_u32(&bar_full[i]), 1); mbar_init(cg_smem_u32(&bar_empty[i]), drain ? 8 : 1); }
    for (int i = 0; i < kStages; ++i) mbar_init(cg_smem_u32(&bar_stage[i]), 1);
    mbar_init(cg_smem_u32(&bar_end[0]), 1);
    mbar_init(cg_smem_u32(&bar_end[1]), 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(cg_smem_u32(&slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 8 || (warp == 9 && issuers == 2)) {
    const int iw = warp - 8;
    if (elect_one()) {
      const uint32_t a0 = cg_smem_u32(smem), b0 = cg_smem_u32(smem + kStages * kStage);
      const uint64_t a_d = umma_desc(a0, 2880, sbo), b_d = umma_desc(b0, N * 16, 128);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32), a_lo0 = (uint32_t)a_d, b_lo0 = (uint32_t)b_d;
      auto D64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0), bstep = N * 2, plane2 = (2 * 2880) >> 4, stage16 = kStage >> 4;
      unsigned long long g0, g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
      long long t0 = clock64();
      uint32_t stage = 0, as = 0, aph = 0, bwalk = 0;
      const uint32_t bwrap = (uint32_t)(81920 / (N * 32));  // K-blocks that fit the B region
      if (issuers == 2) as = (uint32_t)iw;                   // two issuer warps: warp w owns tiles t = w (mod 2) and accumulator w
      for (int t = iw; t < tiles; t += issuers) {
        if (issuers == 2) stage = (uint32_t)((t * chunks) % kStages);
        if (drain) { mbar_wait(cg_smem_u32(&bar_empty[as]), aph ^ 1u); tc_fence_after(); }
        uint32_t acc = (cmode >= 4) ? 1u : 0u;  // cmode 4: never overwrite (accumulate flag always set)
        const uint32_t d = tmem + as * N;
        for (int c = 0; c < chunks; ++c) {
          uint32_t alo = a_lo0 + ((vmode & 2) ? 0u : stage * stage16), blo = b_lo0 + (uint32_t)(c * 18) * bstep;
          if (vmode & 1) { if (bwalk + 18 > bwrap) bwalk = 0; blo = b_lo0 + bwalk * bstep; bwalk += 18; }
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
              const uint32_t sh = (vmode & 8) ? 0u : (uint32_t)((tp / 3) * 10 + (tp % 3));
              tc_mma_bf16(d, D64(a_hi, alo + sh), D64(b_hi, blo), idesc, acc);
              acc = 1;
              if (!(vmode & 16)) blo += bstep;
            }
            if (!(vmode & 4)) alo += plane2;
          }
          if (cmode == 0) tc_commit(cg_smem_u32(&bar_stage[stage]));  // cmode 0: one commit per A stage + one per tile
          if (++stage == kStages) stage = 0;
        }
        if (cmode <= 1 || drain) tc_commit(cg_smem_u32(&bar_full[as]));  // cmode 1: one commit per tile; 2: none
        if (issuers == 1 && cmode != 3) { if (++as == 2) { as = 0; aph ^= 1u; } }  // cmode 3: per-tile commit, same accumulator
      }
      long long t1 = clock64();
      // wait for the last accumulator
      tc_commit(cg_smem_u32(&bar_end[iw]));  // drain the pipe before the CTA exits
      mbar_wait(cg_smem_u32(&bar_end[iw]), 0);
      long long t2 = clock64();
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      if (blockIdx.x == 0 && iw == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = (long long)(g1 - g0); }
    }
  } else if (warp < 8 && drain) {
    uint32_t as = 0, aph = 0;
    const uint32_t row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float sink = 0.f;
    for (int t = 0; t < tiles; ++t) {
      if (lane == 0) mbar_wait(cg_smem_u32(&bar_full[as]), aph);
      __syncwarp();
      tc_fence_after();
      float v[16];
      tmem_ld16(row + as * N + (N > 16 ? (warp >> 2) * 16 : 0), v);
      if (drain == 2) {  // the epilogue zeroes the accumulator it just read, so the next tile can accumulate onto it
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(
                         row + as * N + (N > 16 ? (warp >> 2) * 16 : 0)), "r"(0) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(cg_smem_u32(&bar_empty[as]));
      sink += v[lane & 15];
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
    if (sink == 123.456f) out[3] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int smem = kStages * kStage + 81920;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 256;
  printf("vmode 0: one issuer warp | 100: two issuer warps on alternate tiles | 200: same with commits\n");
  printf("vmode bits: 1 B keeps walking (no per-tile restart) | 2 A stage fixed | 4 no K-block hop | 8 taps unshifted | 16 B fixed\n");
  printf("%4s %5s %6s %5s %4s %5s | %10s %12s %8s\n", "N", "grid", "chunks", "drain", "sbo", "vmode", "clk/MMA", "clk/tile", "SM MHz");
  for (int sbo : {160})
    for (int N : {16, 64})
      for (int grid : {1})
        for (int chunks : {1, 2})
          for (int drain : {0})
           for (int cmode : {4})
           for (int vmode : {0, 100, 200}) {
            const int issuers = vmode >= 100 ? 2 : 1;
            const int cm = vmode == 200 ? 0 : cmode;   // 200: two issuers with per-stage + per-tile commits
            k2<<<grid, 320, smem>>>(N, tiles, chunks, drain, sbo, cm, 0, issuers, d);
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            const double n = (double)tiles * chunks * 18;
            printf("%4d %5d %6d %5d %4d %5d | %10.1f %12.1f %8.0f\n", N, grid, chunks, drain, sbo, vmode, h[0] / n, (double)h[0] / tiles, h[1] / (h[2] / 1e3));
          }
  return 0;
}
