// Weight gradient of the 3x3 convolutions whose channel counts are too small to feed tcgen05 (sm_100a):
//   dW[co][ci][kh][kw] += sum_pixels dY[p][co] * act(X)[p + (kh-1, kw-1)][ci],   dbias[co] += sum_pixels dY[p][co]
//
// Why not tcgen05 here (measured on B200, profiles/r1e_timeline.txt): the reduction dimension is PIXELS, so each
// tcgen05.mma covers only K=16 pixels and re-reads a full M=128-row operand from shared memory although the
// layers have 16..64 real channels; one 128-pixel tile costs 72 instructions x ~75 clk (shared-memory operand
// rate ~60 B/clk) = 2.9 us against an HBM budget of 0.3 us.  Warp-level mma.sync (m16n8k16, bf16 -> fp32) has
// M/N granularity 16/8, takes its operands from registers, and lets each warp keep the WIDE operand's
// fragments in registers while it sweeps the three taps of one kernel row over the NARROW operand:
//
//   operands    both are TMA boxes of the planar bf16 layout (N, C/8, H, W, 8): the wide one as a 16x8-pixel
//               tile, the narrow one as the 18x10 halo tile (borders zero-filled by the TMA unit).  A pixel's 8
//               channels are 16 contiguous bytes = one ldmatrix row, so ldmatrix.trans with per-lane pixel
//               addresses yields the (channel x pixel) fragments directly; a tap is an address offset.
//   roles       M = dY channels (A fragments), N = X channels (B fragments), K = 16 pixels (two tile rows).
//               SHIFT_A: dY is narrow -> dY carries the halo, tile partitions X positions.
//               else   : X is narrow  -> X carries the halo, tile partitions dY positions.
//   warps       12 warps = 3 kernel rows (kh) x 4 pixel groups; a group owns every 4th tile of the CTA and issues its own TMA;
//               accumulators (3 taps x MT x NT fragments, <= 96 fp32 registers) persist over all tiles of the CTA.
//   activation  ReLU of the forward pre-activation is applied to the X fragments in registers.
//   dbias       one extra mma of the (unshifted) dY fragment against a ones fragment -- no second pass over dY.
//   flush       fragments -> fp32 shared-memory tile ordered [co][ci][tap] (= OIHW order) -> coalesced global
//               red.add into the flat gradient bucket.
// GELU layers and 1x1 / 1-pixel problems stay on the tcgen05 kernel of wgrad_tc.cu.
#include <cuda.h>

#include <cstdlib>

#include "cg_common.cuh"

namespace {

constexpr int kComputeWarps = 12;  // 3 kernel rows x 4 pixel groups (each group takes every 4th tile)
constexpr int kThreadsM = kComputeWarps * 32;  // 384: three warps per scheduler -> up to 168 registers/thread
constexpr int kMaxStagesM = 16;  // ring depth per launch: multiple of 4 (one slice per pixel group), from smem budget
constexpr int kPlaneHalo = 2880;  // 18 rows * 10 px * 16 B
constexpr int kPlaneFlat = 2048;  // 16 rows *  8 px * 16 B
constexpr int kMaxChunksM = 32;
constexpr int kHdrM = 256;
constexpr int kPixGroups = kComputeWarps / 3;  // 4

struct MChunk {
  int16_t src, c0, nc;  // chunk of the wide operand: X source index (or 0 for dY), first channel, channels
};

struct alignas(64) MParams {
  CUtensorMap x_map[CG_MAX_SRC];
  CUtensorMap dy_map;
  cg_wgrad_args a;
  MChunk chunk[kMaxChunksM];   // 3x3: chunks of the wide operand; 1x1: chunks of X
  MChunk ychunk[kMaxChunksM];  // 1x1 only: chunks of dY
  int ny;
  uint32_t x_bytes[CG_MAX_SRC], dy_bytes;  // TMA transaction bytes of one box
  int nchunks;
  int tiles_x, tiles_per_img, ntiles;
  int stage_bytes, wide_off, nst;
  unsigned long long* tl;  // debug timeline (CG_TIMELINE builds)
};

__device__ __forceinline__ void tma4m(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t relu2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  v = __hmax2(v, __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int MT, int NT, bool SHIFT_A>
__global__ void __launch_bounds__(kThreadsM, 1) wgrad_mma_kernel(const __grid_constant__ MParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0..15] full, [16..31] empty
  uint8_t* stages = smem + kHdrM;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (kMaxStagesM + i); };
  constexpr int CO = MT * 16, CI = NT * 8;
  const MChunk wc = P.chunk[blockIdx.y];
  if (threadIdx.x == 0) CG_TL(P.tl, 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nst; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(EMPTY(i), 3);  // the three warps of the owning pixel group
    }
    mbar_fence_init();
  }
  __syncthreads();

  float acc[3][MT][NT][4];
  float bacc[MT][4];
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[kw][mt][nt][q] = 0.f;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int q = 0; q < 4; ++q) bacc[mt][q] = 0.f;

  const int tg = warp % 3, pg = warp / 3;  // kernel row kh, pixel group
  const bool do_bias = P.a.dbias != nullptr && (SHIFT_A ? (blockIdx.y == 0 && tg == 1) : (tg == 0));

  // Each pixel group (3 warps = the three kernel rows) owns every 4th tile of the CTA and its own slice of the
  // stage ring (stages pg, pg+4, ...): a warp meets one barrier pair per WHOLE tile (8 k-steps), and the four groups
  // work on four different tiles at once.  TMA issue is folded into the group's first warp (one elected lane):
  // before computing its j-th tile it requests its (j + depth - 1)-th, i.e. the stage the group released last.
  const CUtensorMap* wide_map = SHIFT_A ? &P.x_map[wc.src] : &P.dy_map;
  const CUtensorMap* narrow_map = SHIFT_A ? &P.dy_map : &P.x_map[0];
  const uint32_t tx = SHIFT_A ? (P.x_bytes[wc.src] + P.dy_bytes) : (P.x_bytes[0] + P.dy_bytes);
  const int w_oct = wc.c0 >> 3;
  const int depth = P.nst / kPixGroups;  // stages per group (>= 2)
  uint32_t p_slot = 0, p_phase = 0;
  int p_tile = blockIdx.x + pg * gridDim.x;
  auto produce = [&]() {  // called by one elected lane of the group's first warp
    if (p_tile >= P.ntiles) return;
    const int n = p_tile / P.tiles_per_img;
    const int r = p_tile - n * P.tiles_per_img;
    const int ty = r / P.tiles_x;
    const int h0 = ty * 16, w0 = (r - ty * P.tiles_x) * 8;
    const int st = pg + kPixGroups * (int)p_slot;
    mbar_wait(EMPTY(st), p_phase ^ 1u);
    mbar_expect_tx(FULL(st), tx);
    const uint32_t sb = cg_smem_u32(stages + st * P.stage_bytes);
    tma4m(sb, narrow_map, (w0 - 1) * 8, h0 - 1, 0, n, FULL(st));
    tma4m(sb + P.wide_off, wide_map, w0 * 8, h0, w_oct, n, FULL(st));
    if (++p_slot == (uint32_t)depth) { p_slot = 0; p_phase ^= 1u; }
    p_tile += kPixGroups * gridDim.x;
  };
  if (tg == 0) {
    if (elect_one()) {
      for (int i = 0; i < depth - 1; ++i) produce();
    }
    __syncwarp();
  }
  {
    // ------------------------------------------------------------------ compute warps
    const int j = lane >> 3, i = lane & 7;  // ldmatrix: lane supplies row i of matrix j
    const bool relu = P.a.act == CG_ACT_RELU;
    const uint32_t ones = 0x3F803F80u;  // bf16 (1, 1)
    // per-lane byte offsets inside a stage (tile-invariant)
    //   A fragment x4: matrices (octet 0, rows lo) (octet 1, rows lo) (octet 0, rows hi) (octet 1, rows hi)
    //   B fragment x4: matrices (octet 0, rows lo) (octet 0, rows hi) (octet 1, rows lo) (octet 1, rows hi)
    uint32_t a_off, b_off;
    if (SHIFT_A) {
      a_off = (uint32_t)((j & 1) * kPlaneHalo + (((j >> 1) + 2 - tg) * 10 + (i + 2)) * 16);
      b_off = (uint32_t)(P.wide_off + (j >> 1) * kPlaneFlat + ((j & 1) * 8 + i) * 16);
    } else {
      a_off = (uint32_t)(P.wide_off + (j & 1) * kPlaneFlat + ((j >> 1) * 8 + i) * 16);
      b_off = (uint32_t)((j >> 1) * kPlaneHalo + (((j & 1) + tg) * 10 + i) * 16);
    }
    uint32_t slot = 0, phase = 0;
    int tl_i = 0;
    (void)tl_i;
    if (threadIdx.x == 0) CG_TL(P.tl, 0);
    for (int tile = blockIdx.x + pg * gridDim.x; tile < P.ntiles; tile += kPixGroups * gridDim.x) {
      if (tg == 0) {
        if (elect_one()) produce();
        __syncwarp();
      }
      const int stage = pg + kPixGroups * (int)slot;
      if (lane == 0) mbar_wait(FULL(stage), phase);
      __syncwarp();
      if (threadIdx.x == 0 && tl_i < 8) CG_TL(P.tl, 2 + 2 * tl_i);
      const uint32_t sb = cg_smem_u32(stages + stage * P.stage_bytes);
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) {
        const int r0 = 2 * ks;  // first of the two tile rows of this k-step
        if (SHIFT_A) {
          // X fragments of this k-step stay in registers while the three taps of kernel row tg sweep over dY:
          // dY position of tap (kh, kw) for tile pixel (r, c) is halo[(r + 2 - kh)][(c + 2 - kw)]
          uint32_t b[NT][2];
#pragma unroll
          for (int n2 = 0; n2 < NT / 2; ++n2) {
            uint32_t t4[4];
            ldsm_x4_t(t4, sb + b_off + (uint32_t)(n2 * 2 * kPlaneFlat + r0 * 8 * 16));
            b[2 * n2][0] = relu ? relu2(t4[0]) : t4[0];
            b[2 * n2][1] = relu ? relu2(t4[1]) : t4[1];
            b[2 * n2 + 1][0] = relu ? relu2(t4[2]) : t4[2];
            b[2 * n2 + 1][1] = relu ? relu2(t4[3]) : t4[3];
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            uint32_t a[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
              ldsm_x4_t(a[mt], sb + a_off + (uint32_t)(mt * 2 * kPlaneHalo + (r0 * 10 - kw) * 16));
            if (kw == 1 && do_bias) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) mma_bf16(bacc[mt], a[mt], ones, ones);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[kw][mt][nt], a[mt], b[nt][0], b[nt][1]);
          }
        } else {
          uint32_t a[MT][4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            ldsm_x4_t(a[mt], sb + a_off + (uint32_t)(mt * 2 * kPlaneFlat + r0 * 8 * 16));
          if (do_bias) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma_bf16(bacc[mt], a[mt], ones, ones);
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            // X fragments of tap (tg, kw): halo[(r + kh)][(c + kw)]
#pragma unroll
            for (int n2 = 0; n2 < NT / 2; ++n2) {
              uint32_t b[4];
              ldsm_x4_t(b, sb + b_off + (uint32_t)(n2 * 2 * kPlaneHalo + (r0 * 10 + kw) * 16));
              if (relu) {
#pragma unroll
                for (int q = 0; q < 4; ++q) b[q] = relu2(b[q]);
              }
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                mma_bf16(acc[kw][mt][2 * n2], a[mt], b[0], b[1]);
                mma_bf16(acc[kw][mt][2 * n2 + 1], a[mt], b[2], b[3]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (threadIdx.x == 0 && tl_i < 8) CG_TL(P.tl, 3 + 2 * tl_i);
      ++tl_i;
      if (lane == 0) mbar_arrive(EMPTY(stage));
      if (++slot == (uint32_t)depth) { slot = 0; phase ^= 1u; }
    }
  }

  // ------------------------------------------------------------------ flush
  __syncthreads();  // every stage has been consumed: the ring memory becomes the fp32 reduction tile
  if (threadIdx.x == 0) CG_TL(P.tl, 20);
  // [CO][CI*9 (+1 pad)]: OIHW order of this chunk; the three kernel-row groups own disjoint taps, so the only
  // overlap is between the pixel-row groups: group 0 stores, groups 1.. add in turn after a barrier -- no atomics.
  constexpr int ROW = CI * 9 + 1;  // +1: fragment rows (co) land in different banks
  float* sacc = reinterpret_cast<float*>(stages);
  float* sbias = sacc + CO * ROW;  // [CO]
  for (int e = threadIdx.x; e < CO; e += kThreadsM) sbias[e] = 0.f;
#pragma unroll
  for (int half = 0; half < kPixGroups; ++half) {
    if (pg == half) {
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tap = tg * 3 + kw;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            float* p0 = sacc + (mt * 16 + g) * ROW + (nt * 8 + 2 * t) * 9 + tap;
            float* p1 = p0 + 8 * ROW;
            if (half == 0) {
              p0[0] = acc[kw][mt][nt][0];
              p0[9] = acc[kw][mt][nt][1];
              p1[0] = acc[kw][mt][nt][2];
              p1[9] = acc[kw][mt][nt][3];
            } else {
              p0[0] += acc[kw][mt][nt][0];
              p0[9] += acc[kw][mt][nt][1];
              p1[0] += acc[kw][mt][nt][2];
              p1[9] += acc[kw][mt][nt][3];
            }
          }
      }
    }
    __syncthreads();
    if (half == 0 && do_bias && (lane & 3) == 0) {
      const int g = lane >> 2;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        atomicAdd(&sbias[mt * 16 + g], bacc[mt][0]);  // one warp per pixel group and row: cheap
        atomicAdd(&sbias[mt * 16 + g + 8], bacc[mt][2]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) CG_TL(P.tl, 21);
  {
    const int co0 = SHIFT_A ? 0 : wc.c0;
    const int xs = SHIFT_A ? wc.src : 0;
    const int ci0 = SHIFT_A ? wc.c0 : 0;
    const int co_n = SHIFT_A ? CO : wc.nc, ci_n = SHIFT_A ? wc.nc : CI;  // channels of the chunk actually loaded
    const int cin_l = P.a.cin_l, cout_l = P.a.cout_l, xlog = P.a.src_log[xs], xoff = P.a.src_off[xs];
    for (int e = threadIdx.x; e < CO * CI * 9; e += kThreadsM) {
      const int co = e / (CI * 9), rem = e - co * (CI * 9);
      const int ci = rem / 9;
      if (co < co_n && ci < ci_n && co0 + co < cout_l && ci0 + ci < xlog)
        atomicAdd(P.a.dw + ((long long)(co0 + co) * cin_l + (xoff + ci0)) * 9 + rem, sacc[co * ROW + rem]);
    }
    if (P.a.dbias != nullptr && (SHIFT_A ? blockIdx.y == 0 : true)) {
      for (int co = threadIdx.x; co < CO; co += kThreadsM)
        if (co < co_n && co0 + co < cout_l) atomicAdd(P.a.dbias + co0 + co, sbias[co]);
    }
  }
  if (threadIdx.x == 0) CG_TL(P.tl, 22);
}

// ---------------------------------------------------------------------------------------------------------
// 1x1 convolutions (z_proj, z_feat_proj, width_proj: src/vae.py:70-78,165-170): dW[co][ci] += sum_p dY[p][co] X[p][ci].
// Same machinery without taps: both operands are plain 16x8-pixel tiles.  12 warps = 4 pixel groups (every 4th
// tile, own ring slice) x 3 output-channel groups; the CTA owns CO = 3*MT*16 dY channels x CI = NT*8 X channels.
template <int MT, int NT>
__global__ void __launch_bounds__(kThreadsM, 1) wgrad1_mma_kernel(const __grid_constant__ MParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* stages = smem + kHdrM;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (kMaxStagesM + i); };
  constexpr int CO = 3 * MT * 16, CI = NT * 8;
  const int nx = P.nchunks;
  const MChunk yc = P.ychunk[blockIdx.y / nx];
  const MChunk xc = P.chunk[blockIdx.y % nx];
  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nst; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(EMPTY(i), 3);
    }
    mbar_fence_init();
  }
  __syncthreads();
  float acc[MT][NT][4];
  float bacc[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) bacc[mt][q] = 0.f;
  }
  const int cg = warp % 3, pg = warp / 3;  // output-channel group, pixel group
  const bool do_bias = P.a.dbias != nullptr && (blockIdx.y % nx) == 0;
  const uint32_t tx = P.dy_bytes + P.x_bytes[xc.src];
  const int depth = P.nst / kPixGroups;
  uint32_t p_slot = 0, p_phase = 0;
  int p_tile = blockIdx.x + pg * gridDim.x;
  auto produce = [&]() {
    if (p_tile >= P.ntiles) return;
    const int n = p_tile / P.tiles_per_img;
    const int r = p_tile - n * P.tiles_per_img;
    const int ty = r / P.tiles_x;
    const int h0 = ty * 16, w0 = (r - ty * P.tiles_x) * 8;
    const int st = pg + kPixGroups * (int)p_slot;
    mbar_wait(EMPTY(st), p_phase ^ 1u);
    mbar_expect_tx(FULL(st), tx);
    const uint32_t sb = cg_smem_u32(stages + st * P.stage_bytes);
    tma4m(sb, &P.dy_map, w0 * 8, h0, yc.c0 >> 3, n, FULL(st));
    tma4m(sb + P.wide_off, &P.x_map[xc.src], w0 * 8, h0, xc.c0 >> 3, n, FULL(st));
    if (++p_slot == (uint32_t)depth) { p_slot = 0; p_phase ^= 1u; }
    p_tile += kPixGroups * gridDim.x;
  };
  if (cg == 0) {
    if (elect_one()) {
      for (int i = 0; i < depth - 1; ++i) produce();
    }
    __syncwarp();
  }
  {
    const int j = lane >> 3, i = lane & 7;
    const bool relu = P.a.act == CG_ACT_RELU;
    const uint32_t ones = 0x3F803F80u;
    const uint32_t a_off = (uint32_t)((cg * MT * 2 + (j & 1)) * kPlaneFlat + ((j >> 1) * 8 + i) * 16);
    const uint32_t b_off = (uint32_t)(P.wide_off + (j >> 1) * kPlaneFlat + ((j & 1) * 8 + i) * 16);
    uint32_t slot = 0, phase = 0;
    for (int tile = blockIdx.x + pg * gridDim.x; tile < P.ntiles; tile += kPixGroups * gridDim.x) {
      if (cg == 0) {
        if (elect_one()) produce();
        __syncwarp();
      }
      const int stage = pg + kPixGroups * (int)slot;
      if (lane == 0) mbar_wait(FULL(stage), phase);
      __syncwarp();
      const uint32_t sb = cg_smem_u32(stages + stage * P.stage_bytes);
#pragma unroll 2
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t roff = (uint32_t)(2 * ks * 8 * 16);
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) ldsm_x4_t(a[mt], sb + a_off + (uint32_t)(mt * 2 * kPlaneFlat) + roff);
        if (do_bias) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) mma_bf16(bacc[mt], a[mt], ones, ones);
        }
#pragma unroll
        for (int n2 = 0; n2 < NT / 2; ++n2) {
          uint32_t b[4];
          ldsm_x4_t(b, sb + b_off + (uint32_t)(n2 * 2 * kPlaneFlat) + roff);
          if (relu) {
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = relu2(b[q]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(acc[mt][2 * n2], a[mt], b[0], b[1]);
            mma_bf16(acc[mt][2 * n2 + 1], a[mt], b[2], b[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(EMPTY(stage));
      if (++slot == (uint32_t)depth) { slot = 0; phase ^= 1u; }
    }
  }
  // flush: [CO][CI (+1 pad)] fp32 tile; pixel group 0 stores, groups 1..3 add in turn
  __syncthreads();
  constexpr int ROW = CI + 1;
  float* sacc = reinterpret_cast<float*>(stages);
  float* sbias = sacc + CO * ROW;
  for (int e = threadIdx.x; e < CO; e += kThreadsM) sbias[e] = 0.f;
#pragma unroll
  for (int half = 0; half < kPixGroups; ++half) {
    if (pg == half) {
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          float* p0 = sacc + ((cg * MT + mt) * 16 + g) * ROW + nt * 8 + 2 * t;
          float* p1 = p0 + 8 * ROW;
          if (half == 0) {
            p0[0] = acc[mt][nt][0]; p0[1] = acc[mt][nt][1]; p1[0] = acc[mt][nt][2]; p1[1] = acc[mt][nt][3];
          } else {
            p0[0] += acc[mt][nt][0]; p0[1] += acc[mt][nt][1]; p1[0] += acc[mt][nt][2]; p1[1] += acc[mt][nt][3];
          }
        }
    }
    __syncthreads();
    if (half == 0 && do_bias && (lane & 3) == 0) {
      const int g = lane >> 2;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        atomicAdd(&sbias[(cg * MT + mt) * 16 + g], bacc[mt][0]);
        atomicAdd(&sbias[(cg * MT + mt) * 16 + g + 8], bacc[mt][2]);
      }
    }
  }
  __syncthreads();
  {
    const int cin_l = P.a.cin_l, cout_l = P.a.cout_l, xlog = P.a.src_log[xc.src], xoff = P.a.src_off[xc.src];
    for (int e = threadIdx.x; e < CO * CI; e += kThreadsM) {
      const int co = e / CI, ci = e - co * CI;
      if (co < yc.nc && ci < xc.nc && yc.c0 + co < cout_l && xc.c0 + ci < xlog)
        atomicAdd(P.a.dw + (long long)(yc.c0 + co) * cin_l + xoff + xc.c0 + ci, sacc[co * ROW + ci]);
    }
    if (do_bias) {
      for (int co = threadIdx.x; co < CO; co += kThreadsM)
        if (co < yc.nc && yc.c0 + co < cout_l) atomicAdd(P.a.dbias + yc.c0 + co, sbias[co]);
    }
  }
}

template <int MT, int NT>
int launch_mma1(const MParams& kp, int gx, int gy, int smem_bytes, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad1_mma_kernel<MT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d_wgrad(mma 1x1): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  wgrad1_mma_kernel<MT, NT><<<dim3(gx, gy), kThreadsM, smem_bytes, st>>>(kp);
  return CG_OK;
}

template <int MT, int NT, bool SHIFT_A>
int launch_mma(const MParams& kp, int gx, int smem_bytes, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_mma_kernel<MT, NT, SHIFT_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d_wgrad(mma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  wgrad_mma_kernel<MT, NT, SHIFT_A><<<dim3(gx, kp.nchunks), kThreadsM, smem_bytes, st>>>(kp);
  return CG_OK;
}

}  // namespace

// Returns CG_OK with *handled = 1 when the problem was launched on the mma.sync kernel, *handled = 0 when the
// caller must use the tcgen05 kernel (1x1 / centre-tap / 1-pixel problems, GELU, both operands wide).
// Pixel tiles per CTA (lower bound): cg_wgrad_args.min_tiles, 24 when the caller gives none.  Weight gradients are off the
// critical path (low-priority pool streams) and the step is throughput-bound: every CTA of a weight gradient flushes its
// whole accumulator tile through atomics and holds an SM that a kernel of the dependent chain wants, so the low-resolution
// problems run on FEW CTAs with many tiles each -- as many as the step can hide (the policy lives with the caller:
// ops.wgrad_min_tiles, measurements in profiles/r4g_wgrad_grid_and_stem.txt).  CG_WGRAD_MIN_TILES overrides both.
static int wgrad_min_tiles(const cg_wgrad_args* a) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CG_WGRAD_MIN_TILES");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  if (v > 0) return v;
  return a->min_tiles > 0 ? a->min_tiles : 24;
}

static bool mma1_eligible(const cg_wgrad_args* a) {
  return a->ksize == 1 && a->taps == 1 && !(a->H == 1 && a->W == 1) &&
         (a->act == CG_ACT_NONE || a->act == CG_ACT_RELU);
}

static bool mma_eligible(const cg_wgrad_args* a) {
  if (mma1_eligible(a)) return true;
  if (a->ksize != 3 || a->taps != 9 || (a->H == 1 && a->W == 1)) return false;
  if (a->act != CG_ACT_NONE && a->act != CG_ACT_RELU) return false;
  int xtot = 0;
  for (int s = 0; s < a->nsrc; ++s) xtot += a->src[s].C;
  return (a->dy_c <= 48 && a->dy_c <= xtot) || (xtot <= 48 && a->nsrc == 1);
}

// kernels one cg_conv2d_wgrad call launches: the mma.sync kernel folds the bias gradient in, the tcgen05
// kernel is followed by a column-sum kernel when dbias is requested
extern "C" int32_t cg_conv2d_wgrad_launches(const cg_wgrad_args* a) {
  if (a == nullptr) return 0;
  const char* e = getenv("CG_WGRAD_TC_ONLY");
  const bool tc_only = e != nullptr && e[0] == '1';
  if (!tc_only && mma_eligible(a)) return 1;
  return a->dbias != nullptr ? 2 : 1;
}

static int wgrad_mma_1x1(const cg_wgrad_args* a, void* stream, int* handled) {
  int widest = 0;
  for (int s = 0; s < a->nsrc; ++s) widest = a->src[s].C > widest ? a->src[s].C : widest;
  // X chunk width (NT*8) follows the widest source; the dY chunk (3*MT*16) takes what the stage budget leaves
  const int NT = widest <= 16 ? 2 : (widest <= 32 ? 4 : 8);
  const int MT = NT == 2 ? 2 : 1;
  const int CO = 3 * MT * 16, CI = NT * 8;
  MParams kp;
  kp.a = *a;
  kp.tl = cg_tl_ptr;
  kp.nchunks = kp.ny = 0;
  for (int s = 0; s < a->nsrc; ++s)
    for (int c0 = 0; c0 < a->src[s].C; c0 += CI) {
      if (kp.nchunks >= kMaxChunksM) return CG_OK;
      const int nc = a->src[s].C - c0 < CI ? a->src[s].C - c0 : CI;
      kp.chunk[kp.nchunks++] = MChunk{(int16_t)s, (int16_t)c0, (int16_t)nc};
    }
  for (int c0 = 0; c0 < a->dy_c; c0 += CO) {
    if (kp.ny >= kMaxChunksM) return CG_OK;
    const int nc = a->dy_c - c0 < CO ? a->dy_c - c0 : CO;
    kp.ychunk[kp.ny++] = MChunk{0, (int16_t)c0, (int16_t)nc};
  }
  const int gy = kp.nchunks * kp.ny;
  if (gy > cg_device_sms()) return CG_OK;  // more channel blocks than SMs: leave it to the tcgen05 kernel
  for (int s = 0; s < a->nsrc; ++s) {
    const int c8 = a->src[s].C / 8, c8p = a->src[s].c8 > 0 ? a->src[s].c8 : c8;  // K-blocks / octets stored
    const int boct = c8 < NT ? c8 : NT;
    int rc = cg_make_planar_map(&kp.x_map[s], a->src[s].ptr, a->src[s].ns, a->N, a->H, a->W, c8p, 0, 64, 16, boct);
    if (rc != CG_OK) return rc;
    kp.x_bytes[s] = (uint32_t)boct * kPlaneFlat;
  }
  {
    const int c8 = a->dy_c / 8, c8p = a->dy_c8 > 0 ? a->dy_c8 : c8;
    const int boct = c8 < CO / 8 ? c8 : CO / 8;
    int rc = cg_make_planar_map(&kp.dy_map, a->dy, a->dy_ns, a->N, a->H, a->W, c8p, 0, 64, 16, boct);
    if (rc != CG_OK) return rc;
    kp.dy_bytes = (uint32_t)boct * kPlaneFlat;
  }
  kp.tiles_x = (a->W + 7) / 8;
  kp.tiles_per_img = kp.tiles_x * ((a->H + 15) / 16);
  kp.ntiles = a->N * kp.tiles_per_img;
  kp.wide_off = (CO / 8) * kPlaneFlat;  // X tile sits after the dY planes
  kp.stage_bytes = kp.wide_off + NT * kPlaneFlat;
  kp.nst = (232448 - kHdrM) / kp.stage_bytes / 4 * 4;
  if (kp.nst > kMaxStagesM) kp.nst = kMaxStagesM;
  if (kp.nst < 8) return CG_OK;
  const int smem_bytes = kHdrM + kp.nst * kp.stage_bytes;
  int gx = cg_device_sms() / gy;
  if (gx > (kp.ntiles + wgrad_min_tiles(a) - 1) / wgrad_min_tiles(a)) gx = (kp.ntiles + wgrad_min_tiles(a) - 1) / wgrad_min_tiles(a);
  if (gx < 1) gx = 1;
  cudaStream_t st = cg_stream(stream);
  int rc;
  if (NT == 2) rc = launch_mma1<2, 2>(kp, gx, gy, smem_bytes, st);
  else if (NT == 4) rc = launch_mma1<1, 4>(kp, gx, gy, smem_bytes, st);
  else rc = launch_mma1<1, 8>(kp, gx, gy, smem_bytes, st);
  if (rc != CG_OK) return rc;
  CG_LAUNCH_CHECK("cg_conv2d_wgrad(mma 1x1)");
  *handled = 1;
  return CG_OK;
}

int cg_wgrad_mma_try(const cg_wgrad_args* a, void* stream, int* handled) {
  *handled = 0;
  if (!mma_eligible(a)) return CG_OK;
  if (mma1_eligible(a)) return wgrad_mma_1x1(a, stream, handled);
  int xtot = 0;
  for (int s = 0; s < a->nsrc; ++s) xtot += a->src[s].C;
  const int dyc = a->dy_c;
  int MT, NT;
  bool shift_a;
  if (dyc <= 48 && dyc <= xtot) {  // dY is the narrow operand: it carries the halo
    shift_a = true;
    MT = dyc / 16;
    const int nt_max = MT == 1 ? 8 : (MT == 2 ? 4 : 2);
    int widest = 0;
    for (int s = 0; s < a->nsrc; ++s) widest = a->src[s].C > widest ? a->src[s].C : widest;
    NT = widest <= 16 ? 2 : (widest <= 32 ? 4 : 8);
    if (NT > nt_max) NT = nt_max;
  } else if (xtot <= 48 && a->nsrc == 1) {  // X is the narrow operand
    shift_a = false;
    NT = xtot / 8;
    const int mt_max = NT == 2 ? 4 : (NT == 4 ? 2 : 1);
    MT = dyc <= 16 ? 1 : (dyc <= 32 ? 2 : 4);
    if (MT > mt_max) MT = mt_max;
  } else {
    return CG_OK;
  }
  MParams kp;
  kp.a = *a;
  kp.tl = cg_tl_ptr;
  kp.ny = 0;
  const int wide_planes = shift_a ? NT : MT * 2, narrow_planes = shift_a ? MT * 2 : NT;
  const int wch = wide_planes * 8;
  kp.nchunks = 0;
  if (shift_a) {
    for (int s = 0; s < a->nsrc; ++s)
      for (int c0 = 0; c0 < a->src[s].C; c0 += wch) {
        if (kp.nchunks >= kMaxChunksM) return CG_OK;
        const int nc = a->src[s].C - c0 < wch ? a->src[s].C - c0 : wch;
        kp.chunk[kp.nchunks++] = MChunk{(int16_t)s, (int16_t)c0, (int16_t)nc};
      }
  } else {
    for (int c0 = 0; c0 < dyc; c0 += wch) {
      if (kp.nchunks >= kMaxChunksM) return CG_OK;
      const int nc = dyc - c0 < wch ? dyc - c0 : wch;
      kp.chunk[kp.nchunks++] = MChunk{0, (int16_t)c0, (int16_t)nc};
    }
  }
  // tensor maps: the narrow operand carries the halo
  for (int s = 0; s < a->nsrc; ++s) {
    const int c8 = a->src[s].C / 8, c8p = a->src[s].c8 > 0 ? a->src[s].c8 : c8;  // K-blocks / octets stored
    int boct, rc;
    if (shift_a) {
      boct = c8 < wide_planes ? c8 : wide_planes;
      rc = cg_make_planar_map(&kp.x_map[s], a->src[s].ptr, a->src[s].ns, a->N, a->H, a->W, c8p, 0, 64, 16, boct);
      kp.x_bytes[s] = (uint32_t)boct * kPlaneFlat;
    } else {
      boct = narrow_planes;  // == c8
      rc = cg_make_planar_map(&kp.x_map[s], a->src[s].ptr, a->src[s].ns, a->N, a->H, a->W, c8p, 0, 80, 18, boct);
      kp.x_bytes[s] = (uint32_t)boct * kPlaneHalo;
    }
    if (rc != CG_OK) return rc;
  }
  {
    const int c8 = dyc / 8, c8p = a->dy_c8 > 0 ? a->dy_c8 : c8;
    int boct, rc;
    if (shift_a) {
      boct = narrow_planes;  // == c8
      rc = cg_make_planar_map(&kp.dy_map, a->dy, a->dy_ns, a->N, a->H, a->W, c8p, 0, 80, 18, boct);
      kp.dy_bytes = (uint32_t)boct * kPlaneHalo;
    } else {
      boct = c8 < wide_planes ? c8 : wide_planes;
      rc = cg_make_planar_map(&kp.dy_map, a->dy, a->dy_ns, a->N, a->H, a->W, c8p, 0, 64, 16, boct);
      kp.dy_bytes = (uint32_t)boct * kPlaneFlat;
    }
    if (rc != CG_OK) return rc;
  }
  kp.tiles_x = (a->W + 7) / 8;
  kp.tiles_per_img = kp.tiles_x * ((a->H + 15) / 16);
  kp.ntiles = a->N * kp.tiles_per_img;
  kp.wide_off = narrow_planes * kPlaneHalo;  // multiple of 128 (even plane count)
  kp.stage_bytes = (kp.wide_off + wide_planes * kPlaneFlat + 127) / 128 * 128;
  kp.nst = (232448 - kHdrM) / kp.stage_bytes / 4 * 4;
  if (kp.nst > kMaxStagesM) kp.nst = kMaxStagesM;
  const int smem_bytes = kHdrM + kp.nst * kp.stage_bytes;
  // CTAs along the pixel axis share the chunk's gradient through coalesced atomics
  int gx = cg_device_sms() / kp.nchunks;
  if (gx > (kp.ntiles + wgrad_min_tiles(a) - 1) / wgrad_min_tiles(a)) gx = (kp.ntiles + wgrad_min_tiles(a) - 1) / wgrad_min_tiles(a);
  if (gx < 1) gx = 1;
  cudaStream_t st = cg_stream(stream);
  int rc = CG_ERR_UNSUPPORTED;
#define CG_MMA_CASE(mt, nt, sa) \
  if (MT == mt && NT == nt && shift_a == sa) rc = launch_mma<mt, nt, sa>(kp, gx, smem_bytes, st);
  CG_MMA_CASE(1, 8, true) CG_MMA_CASE(1, 4, true) CG_MMA_CASE(1, 2, true)
  CG_MMA_CASE(2, 4, true) CG_MMA_CASE(2, 2, true) CG_MMA_CASE(3, 2, true)
  CG_MMA_CASE(4, 2, false) CG_MMA_CASE(2, 2, false) CG_MMA_CASE(1, 2, false)
  CG_MMA_CASE(2, 4, false) CG_MMA_CASE(1, 4, false) CG_MMA_CASE(1, 6, false)
#undef CG_MMA_CASE
  if (rc == CG_ERR_UNSUPPORTED) {
    cg_set_error("cg_conv2d_wgrad(mma): no kernel instance for MT=%d NT=%d shift_a=%d", MT, NT, (int)shift_a);
    return rc;
  }
  if (rc != CG_OK) return rc;
  CG_LAUNCH_CHECK("cg_conv2d_wgrad(mma)");
  *handled = 1;
  return CG_OK;
}
