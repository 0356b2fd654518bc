// Weight-gradient of the convolutions on tcgen05 (sm_100a):
//   dW[co][ci][tap] += sum_pixels dY[p][co] * act(X)[p + tap][ci]        (reference: autograd of nn.Conv2d)
//
// GEMM view: the reduction (K) dimension is PIXELS, so both operands are "MN-major" (channels contiguous):
//   D[M][N] (+)= A[M][pixels] * B[N][pixels]
// One operand is a 16x8-pixel tile of dY, the other the halo tile of the (re-activated) forward input.  Both are
// TMA boxes of the planar bf16 layout (N, C/8, H, W, 8) and land in shared memory as channel-octet planes, so
// -- exactly like conv_tc.cu -- the 9 taps are 9 descriptor start offsets into one halo tile, and each tap
// accumulates into its own TMEM column range.  Out-of-image pixels are zero-filled by the TMA unit.
// The wide side (<=128 channels per CTA) sits on M, the narrow side (<=48 channels for 3x3, <=64 for 1x1)
// on N.  Each persistent CTA accumulates over all of its pixel tiles in TMEM and flushes once with fp32
// atomics straight into the OIHW gradient tensor.
//
// Warp roles: 0-3 flush (TMEM lane quarters), 4 MMA issuer + TMEM owner, 5 TMA producer, 6-9 transform
// (re-apply the forward pre-activation to the X tile in place).
//
// Scope: this kernel serves the problems the mma.sync kernels of wgrad_mma.cu decline -- GELU pre-activations on wide
// inputs, centre-tap (3x3 on a 1x1 image) and 1-pixel problems, both operands wider than 48 channels.  ReLU / linear
// 3x3 and 1x1 problems (all of UKBB) are dispatched to wgrad_mma.cu first (cg_conv2d_wgrad below), where the reasons
// are written down: K = pixels gives tcgen05 16 pixels per instruction against a mandatory 128-row operand.
#include <cuda.h>

#include <cmath>
#include <cstdlib>

#include "cg_common.cuh"

namespace {

constexpr int kMmaWarp = 4;
constexpr int kTmaWarp = 5;
constexpr int kXfWarp0 = 6;
constexpr int kXfWarps = 4;
constexpr int kThreads = (kXfWarp0 + kXfWarps) * 32;  // 320
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kStages = 3;
constexpr int kPlaneHalo = 2880;  // 18*10*16
constexpr int kPlaneFlat = 2048;  // 16*8*16
constexpr int kStageBytes = 16 * kPlaneHalo + 8 * kPlaneFlat;  // 62464, covers both operand assignments
constexpr int kHdrBytes = 256;
constexpr int kMaxChunks = 24;

struct WChunk {
  int16_t src;  // 0..2: forward input source, 3: dY
  int16_t c0, nc;
};

struct alignas(64) WParams {
  CUtensorMap x_map[CG_MAX_SRC];
  CUtensorMap dy_map;
  cg_wgrad_args a;
  WChunk pch[kMaxChunks], qch[kMaxChunks];
  uint32_t x_bytes[CG_MAX_SRC], dy_bytes;  // TMA transaction bytes per box
  int nP, nQ;
  int x_on_m, ntaps, halo, flat;
  int tiles_x, tiles_per_img, ntiles;
  uint32_t tmem_cols;
  unsigned long long* tl;  // debug timeline (CG_TIMELINE builds)
};

struct Geom {
  int n, h0, w0;
};

__device__ __forceinline__ Geom geom_of(const WParams& P, int tile) {
  Geom g;
  if (P.flat) {
    g.n = tile * 128;
    g.h0 = g.w0 = 0;
  } else {
    g.n = tile / P.tiles_per_img;
    const int r = tile - g.n * P.tiles_per_img;
    const int ty = r / P.tiles_x;
    g.h0 = ty * 16;
    g.w0 = (r - ty * P.tiles_x) * 8;
  }
  return g;
}

__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

__device__ __forceinline__ uint4 wact8(uint4 u, int act) {
  if (act == CG_ACT_RELU) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
    return u;
  }
  float f[8];
  cg_unpack8(u, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = cg_gelu(f[i]);
  return cg_pack8(f);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [0..2] landed, [3..5] full, [6..8] empty, [9] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  uint8_t* stages = smem + kHdrBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  if (threadIdx.x == 0) CG_TL(P.tl, 1);

  const WChunk pc = P.pch[blockIdx.y / P.nQ];
  const WChunk qc = P.qch[blockIdx.y % P.nQ];
  const int Nq = qc.nc;
  // which operand carries the halo (the forward input X of a 3x3 conv)
  const bool p_is_x = P.x_on_m != 0;
  const bool p_halo = P.halo && p_is_x, q_halo = P.halo && !p_is_x;
  const int planeP = p_halo ? kPlaneHalo : kPlaneFlat;
  const int planeQ = q_halo ? kPlaneHalo : kPlaneFlat;
  const int q_off = 16 * planeP;  // Q tile sits after the 16 P planes of the stage
  const int act = P.a.act;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(BAR(i), 1);
      mbar_init(BAR(3 + i), kXfWarps);
      mbar_init(BAR(6 + i), 1);
    }
    mbar_init(BAR(9), 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(cg_smem_u32(tmem_slot), P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CG_TL(P.tl, 0);

  if (warp == kMmaWarp) {
    if (elect_one()) {
      int tl_i = 0;
      (void)tl_i;
      const uint32_t idesc = umma_idesc_bf16(128, Nq, 1, 1);
      const uint32_t pitchP = p_halo ? 160u : 128u, pitchQ = q_halo ? 160u : 128u;
      // descriptors built once; per MMA only the start-address field advances (16-byte units)
      const uint64_t p_d = umma_desc(cg_smem_u32(stages), pitchP, (uint32_t)planeP);
      const uint64_t q_d = umma_desc(cg_smem_u32(stages) + q_off, pitchQ, (uint32_t)planeQ);
      const uint32_t p_hi = (uint32_t)(p_d >> 32), q_hi = (uint32_t)(q_d >> 32);
      const uint32_t p_lo0 = (uint32_t)p_d, q_lo0 = (uint32_t)q_d;
      auto D64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t stage16 = (uint32_t)kStageBytes >> 4;
      const uint32_t ksP = (2u * pitchP) >> 4, ksQ = (2u * pitchQ) >> 4;
      const int ready0 = (act == CG_ACT_NONE) ? 0 : 3;
      uint32_t stage = 0, phase = 0, accum_any = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(BAR(ready0 + stage), phase);
        tc_fence_after();
        if (tl_i < 8) CG_TL(P.tl, 2 + 2 * tl_i);
        const uint32_t plo = p_lo0 + stage * stage16, qlo = q_lo0 + stage * stage16;
        for (int t = 0; t < P.ntaps; ++t) {
          // the un-shifted operand of a 3x3 problem is staged without halo: it starts at its own pixel 0
          const uint32_t toff = P.halo ? (uint32_t)((t / 3) * 10 + (t % 3)) : 0u;
          uint32_t alo = plo + (p_halo ? toff : 0u), blo = qlo + (q_halo ? toff : 0u);
          const uint32_t d = tmem_base + (uint32_t)(t * Nq);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            tc_mma_bf16(d, D64(p_hi, alo), D64(q_hi, blo), idesc, accum_any | (uint32_t)(ks > 0));
            alo += ksP;
            blo += ksQ;
          }
        }
        accum_any = 1;
        tc_commit(BAR(6 + stage));
        if (tl_i < 8) CG_TL(P.tl, 3 + 2 * tl_i);
        ++tl_i;
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      tc_commit(BAR(9));
    }
  } else if (warp == kTmaWarp) {
    if (elect_one()) {
      const int xs_i = p_is_x ? pc.src : qc.src;
      const CUtensorMap* xmap = &P.x_map[xs_i];
      const uint32_t tx = P.x_bytes[xs_i] + P.dy_bytes;
      const int x_oct = (p_is_x ? pc.c0 : qc.c0) >> 3, y_oct = (p_is_x ? qc.c0 : pc.c0) >> 3;
      const uint32_t x_off = p_is_x ? 0u : (uint32_t)q_off, y_off = p_is_x ? (uint32_t)q_off : 0u;
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const Geom g = geom_of(P, tile);
        mbar_wait(BAR(6 + stage), phase ^ 1u);
        mbar_expect_tx(BAR(stage), tx);
        const uint32_t sbase = cg_smem_u32(stages + stage * kStageBytes);
        if (P.flat) {
          tma3(sbase + x_off, xmap, 0, g.n, x_oct, BAR(stage));
          tma3(sbase + y_off, &P.dy_map, 0, g.n, y_oct, BAR(stage));
        } else {
          tma4(sbase + x_off, xmap, (g.w0 - P.halo) * 8, g.h0 - P.halo, x_oct, g.n, BAR(stage));
          tma4(sbase + y_off, &P.dy_map, g.w0 * 8, g.h0, y_oct, g.n, BAR(stage));
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kXfWarp0) {
    if (act != CG_ACT_NONE) {
      const int xt = threadIdx.x - kXfWarp0 * 32;
      const int xs_i = p_is_x ? pc.src : qc.src;
      const int n16 = (int)(P.x_bytes[xs_i] >> 4);
      const uint32_t x_off = p_is_x ? 0u : (uint32_t)q_off;
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        if (lane == 0) mbar_wait(BAR(stage), phase);
        __syncwarp();
        uint4* base = reinterpret_cast<uint4*>(stages + stage * kStageBytes + x_off);
        for (int i = xt; i < n16; i += kXfThreads) base[i] = wact8(base[i], act);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(3 + stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // flush: lane m of TMEM = channel m of the P chunk
    if (lane == 0) mbar_wait(BAR(9), 0);
    __syncwarp();
    tc_fence_after();
    if (threadIdx.x == 0) CG_TL(P.tl, 20);
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
    if (threadIdx.x == 0) CG_TL(P.tl, 21);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) CG_TL(P.tl, 22);
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
  CG_REQUIRE(a->dy != nullptr && ((uintptr_t)a->dy & 15) == 0 && a->dy_c % 16 == 0 && a->dy_ns % 8 == 0,
             "cg_conv2d_wgrad: dy (c=%d ns=%lld)", a->dy_c, (long long)a->dy_ns);
  CG_REQUIRE(a->taps == 1 || a->taps == a->ksize * a->ksize, "cg_conv2d_wgrad: taps %d", a->taps);
  {
    // small-channel 3x3 problems run on the warp-level mma.sync kernel (wgrad_mma.cu explains why);
    // CG_WGRAD_TC_ONLY=1 forces the tcgen05 kernel (A/B measurements)
    static int tc_only = -1;
    if (tc_only < 0) {
      const char* e = getenv("CG_WGRAD_TC_ONLY");
      tc_only = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (!tc_only) {
      for (int s = 0; s < a->nsrc; ++s)
        CG_REQUIRE(a->src[s].ptr != nullptr && ((uintptr_t)a->src[s].ptr & 15) == 0 && a->src[s].C % 16 == 0 &&
                       a->src[s].C > 0 && a->src[s].ns % 8 == 0,
                   "cg_conv2d_wgrad: src %d", s);
      int handled = 0;
      int rc = cg_wgrad_mma_try(a, stream, &handled);
      if (rc != CG_OK) return rc;
      if (handled) return CG_OK;
    }
  }
  WParams kp;
  kp.a = *a;
  kp.tl = cg_tl_ptr;
  kp.ntaps = a->taps;
  kp.halo = (a->ksize == 3 && a->taps == 9) ? 1 : 0;
  kp.flat = (a->H == 1 && a->W == 1) ? 1 : 0;
  CG_REQUIRE(!(kp.flat && kp.halo), "cg_conv2d_wgrad: 3x3 on a 1x1 image must use taps=1");
  int xtot = 0;
  for (int s = 0; s < a->nsrc; ++s) {
    CG_REQUIRE(a->src[s].ptr != nullptr && ((uintptr_t)a->src[s].ptr & 15) == 0 && a->src[s].C % 16 == 0 &&
                   a->src[s].ns % 8 == 0,
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
  // tensor maps: the X operand carries the halo for 3x3; box octet counts = chunk size of its role
  const int x_box_oct = kp.x_on_m ? 16 : nmax / 8, y_box_oct = kp.x_on_m ? nmax / 8 : 16;
  const int xw = kp.halo ? 80 : 64, xh = kp.halo ? 18 : 16;
  const int xplane = kp.halo ? kPlaneHalo : kPlaneFlat;
  for (int s = 0; s < a->nsrc; ++s) {
    const int c8 = a->src[s].C / 8, c8p = a->src[s].c8 > 0 ? a->src[s].c8 : c8;  // K-blocks / octets stored
    const int boct = c8 < x_box_oct ? c8 : x_box_oct;
    int rc = cg_make_planar_map(&kp.x_map[s], a->src[s].ptr, a->src[s].ns, a->N, a->H, a->W, c8p, kp.flat, xw, xh, boct);
    if (rc != CG_OK) return rc;
    kp.x_bytes[s] = (uint32_t)boct * xplane;
  }
  {
    const int c8 = a->dy_c / 8, c8p = a->dy_c8 > 0 ? a->dy_c8 : c8;
    const int boct = c8 < y_box_oct ? c8 : y_box_oct;
    int rc = cg_make_planar_map(&kp.dy_map, a->dy, a->dy_ns, a->N, a->H, a->W, c8p, kp.flat, 64, 16, boct);
    if (rc != CG_OK) return rc;
    kp.dy_bytes = (uint32_t)boct * kPlaneFlat;
  }
  if (kp.flat) {
    kp.tiles_x = kp.tiles_per_img = 1;
    kp.ntiles = (a->N + 127) / 128;
  } else {
    kp.tiles_x = (a->W + 7) / 8;
    kp.tiles_per_img = kp.tiles_x * ((a->H + 15) / 16);
    kp.ntiles = a->N * kp.tiles_per_img;
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
    int rc = cg_colsum(a->dy, a->dbias, a->N, a->H * a->W, a->cout_l, a->dy_ns, stream);
    if (rc != CG_OK) return rc;
  }
  return CG_OK;
}
