// Weight-gradient of the convolutions on tcgen05 (sm_100a):
//   dW[co][ci][tap] += sum_pixels dY[p][co] * act(X)[p + tap][ci]        (reference: autograd of nn.Conv2d)
//
// GEMM view: the reduction (K) dimension is PIXELS, so both operands are "MN-major" (channels contiguous):
//   D[M][N] (+)= A[M][pixels] * B[N][pixels]
// One operand is a 128-pixel tile of dY, the other the halo tile of the (re-activated) forward input; the
// halo tile is staged exactly like in conv_tc.cu, so the 9 taps are again 9 descriptor start offsets
// into one shared-memory tile, and each tap accumulates into its own TMEM column range.
// The wide side (<=128 channels per CTA) sits on M, the narrow side (<=48 channels for 3x3, <=64 for 1x1)
// on N.  Each persistent CTA accumulates over all of its pixel tiles in TMEM and flushes once with fp32
// atomics straight into the OIHW gradient tensor.
#include "cg_common.cuh"

namespace {

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kLoadWarp0 = 5;
constexpr int kLoadWarps = 8;
constexpr int kThreads = (kLoadWarp0 + kLoadWarps) * 32;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kStages = 3;
constexpr int kPlaneHalo = 2976;  // 18*10*16 padded (see conv_tc.cu)
constexpr int kPlaneFlat = 2080;  // 16*8*16 padded
constexpr int kStageBytes = 16 * kPlaneHalo + 8 * kPlaneFlat;  // 64256, covers both operand assignments
constexpr int kHdrBytes = 256;
constexpr int kMaxChunks = 24;

struct WChunk {
  int16_t src;  // 0..2: forward input source, 3: dY
  int16_t c0, nc;
};

struct WParams {
  cg_wgrad_args a;
  WChunk pch[kMaxChunks], qch[kMaxChunks];
  int nP, nQ;
  int x_on_m, ntaps, halo;
  int tiles_x, ntiles, Hp;
  long long P;
  uint32_t tmem_cols;
};

struct Geom {
  int v0, w0;
  long long p0;
};

__device__ __forceinline__ Geom geom_of(const WParams& P, int tile) {
  Geom g;
  if (P.halo) {
    int tv = tile / P.tiles_x;
    g.v0 = tv * 16;
    g.w0 = (tile - tv * P.tiles_x) * 8;
    g.p0 = 0;
  } else {
    g.v0 = g.w0 = 0;
    g.p0 = (long long)tile * 128;
  }
  return g;
}

// Stage one operand tile as channel-octet planes.  `with_halo`: 18x10 pixels around the tile (3x3 forward
// input), else the tile's own 128 pixels in [16][8] order.
__device__ __forceinline__ void stage_tile(const WParams& P, uint8_t* dst, int plane, const void* ptr, int ld, int bcast,
                                           int act, int c0, int nc8, bool with_halo, const Geom& g, int lt) {
  const int H = P.a.H, W = P.a.W, N = P.a.N;
  const int npix = with_halo ? 180 : 128;
  const int items = npix * nc8;
  for (int it = lt; it < items; it += kLoadThreads) {
    const int c8 = it % nc8;
    const int pix = it / nc8;
    bool valid;
    long long off;
    if (P.halo) {
      int rr, cc;
      if (with_halo) { rr = pix / 10; cc = pix - rr * 10; } else { rr = (pix >> 3) + 1; cc = (pix & 7) + 1; }
      const int v = g.v0 - 1 + rr, w = g.w0 - 1 + cc;
      const int n = v / P.Hp, h = v - n * P.Hp;
      valid = (v >= 0) && (w >= 0) && (w < W) && (n < N) && (h < H);
      off = bcast ? (long long)n * ld : ((long long)(n * H + h) * W + w) * ld;
    } else {
      const long long p = g.p0 + pix;
      valid = p < P.P;
      off = bcast ? (p / ((long long)H * W)) * ld : p * ld;
    }
    uint4 u = make_uint4(0, 0, 0, 0);
    if (valid) {
      u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ptr) + off + c0 + c8 * 8));
      if (act != CG_ACT_NONE) {
        float f[8];
        cg_unpack8(u, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = cg_act(f[i], act);
        u = cg_pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(dst + c8 * plane + pix * 16) = u;
  }
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0..2] full, [3..5] empty, [6] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  uint8_t* stages = smem + kHdrBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const WChunk pc = P.pch[blockIdx.y / P.nQ];
  const WChunk qc = P.qch[blockIdx.y % P.nQ];
  const int Nq = qc.nc;
  // which operand carries the halo (the forward input X of a 3x3 conv)
  const bool p_is_x = P.x_on_m != 0;
  const bool p_halo = P.halo && p_is_x, q_halo = P.halo && !p_is_x;
  const int planeP = p_halo ? kPlaneHalo : kPlaneFlat;
  const int planeQ = q_halo ? kPlaneHalo : kPlaneFlat;
  const int q_off = 16 * planeP;  // Q tile sits after the 16 P planes of the stage

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(BAR(i), kLoadWarps);
      mbar_init(BAR(3 + i), 1);
    }
    mbar_init(BAR(6), 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(cg_smem_u32(tmem_slot), P.tmem_cols);
  // rows of A beyond the chunk's channels are never loaded: clear them once so no NaN bit patterns
  // ever enter the tensor core (their D rows are discarded anyway)
  for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, Nq, 1, 1);
      const uint32_t pitchP = p_halo ? 160u : 128u, pitchQ = q_halo ? 160u : 128u;
      uint32_t stage = 0, phase = 0, accum_any = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(BAR(stage), phase);
        tc_fence_after();
        const uint32_t sP = cg_smem_u32(stages) + stage * kStageBytes, sQ = sP + q_off;
        for (int t = 0; t < P.ntaps; ++t) {
          const uint32_t toff = P.halo ? (uint32_t)((t / 3) * 10 + (t % 3)) * 16u : 0u;
          // the un-shifted operand of a 3x3 problem is staged without halo: start at its own pixel 0
          const uint32_t offP = p_halo ? toff : 0u, offQ = q_halo ? toff : 0u;
          for (int ks = 0; ks < 8; ++ks) {
            uint64_t ad = umma_desc(sP + offP + (uint32_t)ks * 2u * pitchP, pitchP, (uint32_t)planeP);
            uint64_t bd = umma_desc(sQ + offQ + (uint32_t)ks * 2u * pitchQ, pitchQ, (uint32_t)planeQ);
            tc_mma_bf16(tmem_base + (uint32_t)(t * Nq), ad, bd, idesc, accum_any | (uint32_t)(ks > 0));
          }
        }
        accum_any = 1;
        tc_commit(BAR(3 + stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(BAR(6));
    }
  } else if (warp >= kLoadWarp0) {
    const int lt = threadIdx.x - kLoadWarp0 * 32;
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const Geom g = geom_of(P, tile);
      mbar_wait(BAR(3 + stage), phase ^ 1u);
      uint8_t* sP = stages + stage * kStageBytes;
      uint8_t* sQ = sP + q_off;
      // P operand
      if (p_is_x) {
        const cg_src& s = P.a.src[pc.src];
        stage_tile(P, sP, planeP, s.ptr, s.ld, s.bcast, P.a.act, pc.c0, pc.nc / 8, p_halo, g, lt);
        stage_tile(P, sQ, planeQ, P.a.dy, P.a.dy_ld, 0, CG_ACT_NONE, qc.c0, qc.nc / 8, false, g, lt);
      } else {
        const cg_src& s = P.a.src[qc.src];
        stage_tile(P, sP, planeP, P.a.dy, P.a.dy_ld, 0, CG_ACT_NONE, pc.c0, pc.nc / 8, false, g, lt);
        stage_tile(P, sQ, planeQ, s.ptr, s.ld, s.bcast, P.a.act, qc.c0, qc.nc / 8, q_halo, g, lt);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(stage));
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else {
    // epilogue: lane m of TMEM = channel m of the P chunk
    mbar_wait(BAR(6), 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    const int kk = P.a.ksize * P.a.ksize;
    const WChunk& xc = p_is_x ? pc : qc;  // chunk that indexes input channels
    const int xs = xc.src;
    const int ncols = P.ntaps * Nq;
    for (int col = 0; col < ncols; col += 16) {
      float acc[16];
      __syncwarp();
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, acc);
      if (m >= pc.nc) continue;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = col + i;
        const int t = c / Nq, nq = c - t * Nq;
        const int xch = xc.c0 + (p_is_x ? m : nq);   // channel inside the X source
        const int ych = (p_is_x ? qc.c0 + nq : pc.c0 + m);  // dY channel
        if (xch >= P.a.src_log[xs] || ych >= P.a.cout_l) continue;
        const int ci = P.a.src_off[xs] + xch;
        const int tap = (kk == 9) ? (P.ntaps == 9 ? t : 4) : 0;
        atomicAdd(P.a.dw + ((long long)ych * P.a.cin_l + ci) * kk + tap, acc[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

}  // namespace

extern "C" int cg_conv2d_wgrad(const cg_wgrad_args* a, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(a != nullptr && (a->ksize == 1 || a->ksize == 3), "cg_conv2d_wgrad: ksize");
  CG_REQUIRE(a->nsrc >= 1 && a->nsrc <= CG_MAX_SRC, "cg_conv2d_wgrad: nsrc %d", a->nsrc);
  CG_REQUIRE(a->dy != nullptr && ((uintptr_t)a->dy & 15) == 0 && a->dy_c % 16 == 0 && a->dy_ld % 8 == 0,
             "cg_conv2d_wgrad: dy (c=%d ld=%d)", a->dy_c, a->dy_ld);
  CG_REQUIRE(a->taps == 1 || a->taps == a->ksize * a->ksize, "cg_conv2d_wgrad: taps %d", a->taps);
  WParams kp;
  kp.a = *a;
  kp.ntaps = a->taps;
  kp.halo = (a->ksize == 3 && a->taps == 9) ? 1 : 0;
  int xtot = 0;
  for (int s = 0; s < a->nsrc; ++s) {
    CG_REQUIRE(a->src[s].ptr != nullptr && ((uintptr_t)a->src[s].ptr & 15) == 0 && a->src[s].C % 16 == 0 &&
                   a->src[s].ld % 8 == 0,
               "cg_conv2d_wgrad: src %d", s);
    xtot += a->src[s].C;
  }
  const int nmax = kp.halo ? 48 : 64;
  // narrow side on N: dY if it fits in one N chunk and X does not, else X
  kp.x_on_m = (a->dy_c <= nmax && xtot > a->dy_c) ? 1 : 0;
  auto split = [&](WChunk* out, int& n, int src, int C, int step) -> bool {
    for (int c0 = 0; c0 < C; c0 += step) {
      if (n >= kMaxChunks) return false;
      out[n++] = WChunk{(int16_t)src, (int16_t)c0, (int16_t)((C - c0) < step ? (C - c0) : step)};
    }
    return true;
  };
  kp.nP = kp.nQ = 0;
  bool ok = true;
  if (kp.x_on_m) {
    for (int s = 0; s < a->nsrc; ++s) ok = ok && split(kp.pch, kp.nP, s, a->src[s].C, 128);
    ok = ok && split(kp.qch, kp.nQ, 3, a->dy_c, nmax);
  } else {
    ok = ok && split(kp.pch, kp.nP, 3, a->dy_c, 128);
    for (int s = 0; s < a->nsrc; ++s) ok = ok && split(kp.qch, kp.nQ, s, a->src[s].C, nmax);
  }
  CG_REQUIRE(ok, "cg_conv2d_wgrad: too many channel chunks");
  kp.Hp = a->H + 1;
  kp.P = (long long)a->N * a->H * a->W;
  if (kp.halo) {
    kp.tiles_x = (a->W + 7) / 8;
    kp.ntiles = ((a->N * kp.Hp + 15) / 16) * kp.tiles_x;
  } else {
    kp.tiles_x = 1;
    kp.ntiles = (int)((kp.P + 127) / 128);
  }
  uint32_t cols = 32;
  while (cols < (uint32_t)(kp.ntaps * nmax)) cols <<= 1;
  kp.tmem_cols = cols;
  const int smem_bytes = kHdrBytes + kStages * kStageBytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  const int combos = kp.nP * kp.nQ;
  int gx = cg_device_sms() / combos;
  if (gx < 1) gx = 1;
  if (gx > kp.ntiles) gx = kp.ntiles;
  wgrad_tc_kernel<<<dim3(gx, combos), kThreads, smem_bytes, cg_stream(stream)>>>(kp);
  CG_LAUNCH_CHECK("cg_conv2d_wgrad");
  if (a->dbias != nullptr) {
    int rc = cg_colsum(a->dy, a->dbias, kp.P, a->cout_l, a->dy_ld, stream);
    if (rc != CG_OK) return rc;
  }
  return CG_OK;
}
