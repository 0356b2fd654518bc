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
#include <cmath>

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

// Issue the cp.async copies (zero-fill for padding / out-of-image pixels) of one operand tile as
// channel-octet planes.  `with_halo`: 18x10 pixels around the tile (3x3 forward input), else the tile's own
// 128 pixels in [16][8] order.
__device__ __forceinline__ void stage_tile_async(const WParams& P, uint8_t* dst, int plane, const void* ptr, int ld,
                                                 int bcast, int c0, int nc8, bool with_halo, const Geom& g, int lt) {
  const int H = P.a.H, W = P.a.W, N = P.a.N;
  const int npix = with_halo ? 180 : 128;
  const int items = npix * nc8;
  const uint32_t dst_u = cg_smem_u32(dst);
  const bf16* base = reinterpret_cast<const bf16*>(ptr) + c0;
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
    cp_async16(dst_u + c8 * plane + pix * 16, valid ? base + off + c8 * 8 : base, valid ? 16u : 0u);
  }
}

// in-place activation of the slots this thread copied (same index walk as stage_tile_async)
__device__ __forceinline__ void act_tile_inplace(uint8_t* dst, int plane, int act, int nc8, bool with_halo, int lt) {
  const int items = (with_halo ? 180 : 128) * nc8;
  for (int it = lt; it < items; it += kLoadThreads) {
    uint4* p = reinterpret_cast<uint4*>(dst + (it % nc8) * plane + (it / nc8) * 16);
    float f[8];
    cg_unpack8(*p, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = cg_act(f[i], act);
    *p = cg_pack8(f);
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
      // descriptors built once; per MMA only the start-address field advances (16-byte units)
      const uint64_t p_desc0 = umma_desc(cg_smem_u32(stages), pitchP, (uint32_t)planeP);
      const uint64_t q_desc0 = umma_desc(cg_smem_u32(stages) + q_off, pitchQ, (uint32_t)planeQ);
      const uint32_t stage16 = (uint32_t)kStageBytes >> 4;
      const uint32_t ksP = (2u * pitchP) >> 4, ksQ = (2u * pitchQ) >> 4;
      uint32_t stage = 0, phase = 0, accum_any = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(BAR(stage), phase);
        tc_fence_after();
        const uint64_t pd = p_desc0 + stage * stage16, qd = q_desc0 + stage * stage16;
        for (int t = 0; t < P.ntaps; ++t) {
          // the un-shifted operand of a 3x3 problem is staged without halo: it starts at its own pixel 0
          const uint32_t toff = P.halo ? (uint32_t)((t / 3) * 10 + (t % 3)) : 0u;
          uint64_t ad = pd + (p_halo ? toff : 0u), bd = qd + (q_halo ? toff : 0u);
          const uint32_t d = tmem_base + (uint32_t)(t * Nq);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            tc_mma_bf16(d, ad, bd, idesc, accum_any | (uint32_t)(ks > 0));
            ad += ksP;
            bd += ksQ;
          }
        }
        accum_any = 1;
        tc_commit(BAR(3 + stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(BAR(6));
    }
  } else if (warp >= kLoadWarp0) {
    // loaders: cp.async copies issued kStages-1 tiles ahead; activation applied in place afterwards
    const int lt = threadIdx.x - kLoadWarp0 * 32;
    constexpr int D = kStages - 1;
    uint32_t stage = 0, phase = 0;
    uint32_t q_stage[kStages];
    int q_head = 0, q_len = 0;
    const cg_src& xs = P.a.src[p_is_x ? pc.src : qc.src];
    auto finalize = [&](uint32_t st) {
      if (P.a.act != CG_ACT_NONE) {
        uint8_t* sP = stages + st * kStageBytes;
        if (p_is_x) act_tile_inplace(sP, planeP, P.a.act, pc.nc / 8, p_halo, lt);
        else act_tile_inplace(sP + q_off, planeQ, P.a.act, qc.nc / 8, q_halo, lt);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(st));
    };
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const Geom g = geom_of(P, tile);
      mbar_wait(BAR(3 + stage), phase ^ 1u);
      uint8_t* sP = stages + stage * kStageBytes;
      uint8_t* sQ = sP + q_off;
      if (p_is_x) {
        stage_tile_async(P, sP, planeP, xs.ptr, xs.ld, xs.bcast, pc.c0, pc.nc / 8, p_halo, g, lt);
        stage_tile_async(P, sQ, planeQ, P.a.dy, P.a.dy_ld, 0, qc.c0, qc.nc / 8, false, g, lt);
      } else {
        stage_tile_async(P, sP, planeP, P.a.dy, P.a.dy_ld, 0, pc.c0, pc.nc / 8, false, g, lt);
        stage_tile_async(P, sQ, planeQ, xs.ptr, xs.ld, xs.bcast, qc.c0, qc.nc / 8, q_halo, g, lt);
      }
      cp_async_commit();
      q_stage[(q_head + q_len) % kStages] = stage;
      ++q_len;
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
      if (q_len == D) {
        cp_async_wait<D - 1>();
        finalize(q_stage[q_head]);
        q_head = (q_head + 1) % kStages;
        --q_len;
      }
    }
    cp_async_wait<0>();
    while (q_len > 0) {
      finalize(q_stage[q_head]);
      q_head = (q_head + 1) % kStages;
      --q_len;
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
    // Nq is a multiple of 16, so a 16-column chunk never straddles two taps: tap / channel bases hoisted
    const bool row_ok = m < pc.nc;
    const int x_m = xc.c0 + m, y_m = pc.c0 + m;
    int t = 0, nq0 = 0;
    for (int col = 0; col < ncols; col += 16) {
      float acc[16];
      __syncwarp();
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, acc);
      const int tap = (kk == 9) ? (P.ntaps == 9 ? t : 4) : 0;
      if (row_ok) {
        if (p_is_x) {  // lane = input channel, columns = dY channels
          const bool x_ok = x_m < P.a.src_log[xs];
          float* base = P.a.dw + ((long long)(qc.c0 + nq0) * P.a.cin_l + (P.a.src_off[xs] + x_m)) * kk + tap;
          const long long step = (long long)P.a.cin_l * kk;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (x_ok && qc.c0 + nq0 + i < P.a.cout_l) atomicAdd(base + i * step, acc[i]);
        } else {  // lane = dY channel, columns = input channels
          const bool y_ok = y_m < P.a.cout_l;
          float* base = P.a.dw + ((long long)y_m * P.a.cin_l + (P.a.src_off[xs] + xc.c0 + nq0)) * kk + tap;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (y_ok && xc.c0 + nq0 + i < P.a.src_log[xs]) atomicAdd(base + i * kk, acc[i]);
        }
      }
      nq0 += 16;
      if (nq0 >= Nq) { nq0 = 0; ++t; }
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
  // CTAs along the pixel axis: every CTA ends with one fp32 atomic per accumulator element, so few pixel
  // tiles spread over many CTAs is dominated by the flush.  Balance tile time against flush time:
  //   T(gx) ~ ntiles/gx * t_tile + gx * A / R   ->  gx* = sqrt(t_tile * ntiles * R / A)
  // (t_tile ~ 2 us, R ~ 1e5 atomics/us chip-wide, A = atomics per CTA; measured on B200, profiles/r1a)
  const int combos = kp.nP * kp.nQ;
  const double atoms = 128.0 * kp.ntaps * nmax;
  int gx = (int)(sqrt(2.0 * kp.ntiles * 1.0e5 / atoms) + 0.5);
  const int cap = cg_device_sms() / combos > 0 ? cg_device_sms() / combos : 1;
  if (gx > cap) gx = cap;
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
