// Micro-benchmark: sustained TMA tile-load rate of the conv kernel's activation boxes, all SMs, no MMA / epilogue.
// Tensor = bf16 channel-octet planar (N, C8, H, W, 8) as in conv_tc.cu.  Every persistent CTA walks pixel tiles and
// loads, per tile, C8/box_c8 boxes of (box_w8 elements x box_h rows x box_c8 octets) into a ring of `depth` stages;
// a consumer warp releases each stage as soon as it has landed.  Prints GB/s of landed bytes and of useful bytes.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../causal-gen_b200/csrc -I../../include tma_rate.cu -o tma_rate -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "cg_common.cuh"

unsigned long long* cg_tl_ptr = nullptr;
void cg_set_error(const char*, ...) {}
int cg_require_sm100() { return 0; }

struct P {
  CUtensorMap map;
  int ntiles, tiles_x, tiles_per_img, tile_w, tile_h, halo, nbox, box_c8, depth;
  uint32_t box_bytes;
};

__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[32], empty[32];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.depth; ++i) { mbar_init(cg_smem_u32(&full[i]), 1); mbar_init(cg_smem_u32(&empty[i]), 1); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int n = t / p.tiles_per_img, r = t - n * p.tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        for (int b = 0; b < p.nbox; ++b) {
          mbar_wait(cg_smem_u32(&empty[st]), ph ^ 1u);
          mbar_expect_tx(cg_smem_u32(&full[st]), p.box_bytes);
          tma4(cg_smem_u32(smem + (size_t)st * p.box_bytes), &p.map, (tx * p.tile_w - p.halo) * 8, ty * p.tile_h - p.halo,
               b * p.box_c8, n, cg_smem_u32(&full[st]));
          if (++st == (uint32_t)p.depth) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x)
        for (int b = 0; b < p.nbox; ++b) {
          mbar_wait(cg_smem_u32(&full[st]), ph);
          mbar_arrive(cg_smem_u32(&empty[st]));
          if (++st == (uint32_t)p.depth) { st = 0; ph ^= 1u; }
        }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Case { const char* name; int N, H, C, tile_w, tile_h, halo, box_c8; };
  const Case cases[] = {
      {"r96 C64 3x3 halo  box(10px,18,4oct)  [conv c1 today]", 128, 96, 64, 8, 16, 1, 4},
      {"r96 C64 1x1       box( 8px,16,4oct)", 128, 96, 64, 8, 16, 0, 4},
      {"r96 C64 3x3 halo  box(10px,18,8oct)  one box per tile", 128, 96, 64, 8, 16, 1, 8},
      {"r96 C64 3x3 halo  box(18px,18,4oct)  tile pair", 128, 96, 64, 16, 16, 1, 4},
      {"r96 C64 3x3 halo  box(34px,18,4oct)  tile quad", 128, 96, 64, 32, 16, 1, 4},
      {"r96 C64 3x3 halo  box(34px,10,4oct)  32x8 tile", 128, 96, 64, 32, 8, 1, 4},
      {"r96 C16 3x3 halo  box(10px,18,2oct)  [conv c2 today]", 128, 96, 16, 8, 16, 1, 2},
      {"r96 C16 3x3 halo  box(34px,18,2oct)  tile quad", 128, 96, 16, 32, 16, 1, 2},
      {"r192 C32 3x3 halo box(10px,18,4oct)", 128, 192, 32, 8, 16, 1, 4},
      {"r192 C32 3x3 halo box(34px,18,4oct)  tile quad", 128, 192, 32, 32, 16, 1, 4},
      {"r48 C96 3x3 halo  box(10px,18,4oct)", 128, 48, 96, 8, 16, 1, 4},
      {"r24 C128 3x3 halo box(10px,18,4oct)", 128, 24, 128, 8, 16, 1, 4},
  };
  printf("%-58s %8s %8s %9s %9s %7s\n", "case", "depth", "us", "landGB/s", "usefGB/s", "GB/s/SM");
  for (const Case& c : cases) {
    const int C8 = c.C / 8;
    size_t elems = (size_t)c.N * C8 * c.H * c.H * 8;
    void* d;
    cudaMalloc(&d, elems * 2);
    cudaMemset(d, 0, elems * 2);
    P p;
    const cuuint64_t dims[4] = {(cuuint64_t)c.H * 8, (cuuint64_t)c.H, (cuuint64_t)C8, (cuuint64_t)c.N};
    const cuuint64_t strides[3] = {(cuuint64_t)c.H * 16, (cuuint64_t)c.H * c.H * 16, (cuuint64_t)C8 * c.H * c.H * 16};
    const cuuint32_t box[4] = {(cuuint32_t)(c.tile_w + 2 * c.halo) * 8, (cuuint32_t)(c.tile_h + 2 * c.halo), (cuuint32_t)c.box_c8, 1};
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    p.tile_w = c.tile_w; p.tile_h = c.tile_h; p.halo = c.halo; p.box_c8 = c.box_c8; p.nbox = C8 / c.box_c8;
    p.tiles_x = (c.H + c.tile_w - 1) / c.tile_w;
    p.tiles_per_img = p.tiles_x * ((c.H + c.tile_h - 1) / c.tile_h);
    p.ntiles = c.N * p.tiles_per_img;
    p.box_bytes = box[0] * box[1] * box[2] * 2;
    for (int depth : {4, 16}) {
      if ((size_t)depth * p.box_bytes > 200 * 1024) continue;
      p.depth = depth;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int w = 0; w < 2; ++w) tma_kernel<<<148, 64, depth * p.box_bytes>>>(p);
      cudaEventRecord(e0);
      const int reps = 5;
      for (int i = 0; i < reps; ++i) tma_kernel<<<148, 64, depth * p.box_bytes>>>(p);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / reps;
      const double landed = (double)p.ntiles * p.nbox * p.box_bytes, useful = (double)elems * 2;
      printf("%-58s %8d %8.1f %9.0f %9.0f %7.1f\n", c.name, depth, us, landed / us / 1e3, useful / us / 1e3, landed / us / 1e3 / 148);
    }
    cudaFree(d);
  }
  return 0;
}
