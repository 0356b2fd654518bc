// Implicit-GEMM convolution on tcgen05 (sm_100a): forward and data-gradient pass of every
// nn.Conv2d in Block / DecoderBlock (reference src/vae.py:49-84,165-170).
//
//   GEMM view   D[M=128 pixels][N=Cout chunk] += A[pixels][K] * B[Cout][K],  K = taps * Cin
//   A operand   one halo tile of the NHWC bf16 input per K-chunk of 32 channels, staged by the loader
//               warps (activation applied on the way in) as channel-octet planes
//               [c8][18 rows][10 px][8 ch]; because 8 consecutive pixels of a plane row are one
//               128-byte UMMA core matrix (SWIZZLE_NONE, K-major), every one of the 9 taps is just a
//               different descriptor start address into the SAME tile -- im2col without copies.
//   rows        the batch is viewed as one tall image of N*(H+1) "virtual rows" (one shared zero row
//               between images) so a 16x8-pixel tile has a constant row pitch for any H.
//   B operand   packed weights for the CTA's Cout chunk, bulk-copied (cp.async.bulk) into shared
//               memory once and kept resident while the persistent CTA walks its pixel tiles.
//   D           fp32 in TMEM, double buffered (2 x Nc columns) so the epilogue of tile i overlaps the
//               MMAs of tile i+1.  Epilogue: bias, channel-split segments, act'(x) multiply (backward),
//               residual / accumulate add, bf16 or fp32 stores.
//
// Warp roles: 0-3 epilogue (TMEM lane quarters), 4 MMA issuer + TMEM owner, 5-12 loaders.
#include "cg_common.cuh"

namespace {

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kLoadWarp0 = 5;
constexpr int kLoadWarps = 8;
constexpr int kThreads = (kLoadWarp0 + kLoadWarps) * 32;  // 416
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kStages = 4;
constexpr int kPlane3 = 2976;  // 18*10*16 = 2880, padded so the 4 planes of a stage hit distinct banks
constexpr int kPlane1 = 2080;  // 128*16   = 2048, same padding rule
constexpr int kStageBytes = 4 * kPlane3;
constexpr int kHdrBytes = 1280;  // barriers (<=128B) | tmem slot | bias[256]
constexpr int kSmemMax = 232448;  // 227 KB
constexpr int kMaxChunks = 40;

struct Chunk {
  uint16_t src, c0, nc16, kbase;
};

struct KParams {
  cg_conv_args a;
  Chunk chunk[kMaxChunks];
  int nchunks, ntaps, Nc, nN, ktot16;
  int tiles_x, ntiles, Hp, V;
  long long P;  // N*H*W
  uint32_t idesc, tmem_cols, slab_bytes;
};

struct TileGeom {
  int v0, w0;       // 3x3: first virtual row / column of the tile
  long long p0;     // 1x1: first flat pixel
};

__device__ __forceinline__ TileGeom tile_geom(const KParams& P, int tile) {
  TileGeom g;
  if (P.a.ksize == 3) {
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

__device__ __forceinline__ uint4 load_act8(const KParams& P, const cg_src& s, long long elem_off, bool valid) {
  uint4 u = make_uint4(0, 0, 0, 0);
  if (valid) {
    u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.ptr) + elem_off));
    if (P.a.act != CG_ACT_NONE) {
      float f[8];
      cg_unpack8(u, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = cg_act(f[i], P.a.act);
      u = cg_pack8(f);
    }
  }
  return u;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // barrier map: [0..3] a_full, [4..7] a_empty, [8] b_full, [9,10] acc_full, [11,12] acc_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
  float* s_bias = reinterpret_cast<float*>(smem + 256);
  uint8_t* sA = smem + kHdrBytes;
  uint8_t* sB = sA + kStages * kStageBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunkN = blockIdx.y;
  const int Nc = P.Nc;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(BAR(i), kLoadWarps);
      mbar_init(BAR(4 + i), 1);
    }
    mbar_init(BAR(8), 1);
    mbar_init(BAR(9), 1);
    mbar_init(BAR(10), 1);
    mbar_init(BAR(11), kEpiWarps);
    mbar_init(BAR(12), kEpiWarps);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(cg_smem_u32(tmem_slot), P.tmem_cols);
  for (int i = threadIdx.x; i < Nc; i += kThreads) {
    int c = nchunkN * Nc + i;
    s_bias[i] = (P.a.bias != nullptr && c < P.a.bias_n) ? P.a.bias[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool k3 = P.a.ksize == 3;
  const int plane = k3 ? kPlane3 : kPlane1;

  if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.a.wpack) + (size_t)nchunkN * P.slab_bytes;
      mbar_expect_tx(BAR(8), P.slab_bytes);
      for (uint32_t off = 0; off < P.slab_bytes; off += 32768u) {
        uint32_t n = min(32768u, P.slab_bytes - off);
        bulk_g2s(cg_smem_u32(sB) + off, wsrc + off, n, BAR(8));
      }
      mbar_wait(BAR(8), 0);
      const uint32_t sB_addr = cg_smem_u32(sB);
      const uint32_t b_lbo = (uint32_t)Nc * 16u, b_step = (uint32_t)Nc * 32u;
      const uint32_t a_sbo = k3 ? 160u : 128u;
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(BAR(11 + as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)Nc;
        uint32_t accum = 0;
        for (int c = 0; c < P.nchunks; ++c) {
          const Chunk ch = P.chunk[c];
          mbar_wait(BAR(stage), phase);
          tc_fence_after();
          const uint32_t a_base = cg_smem_u32(sA) + stage * kStageBytes;
          for (int j = 0; j < ch.nc16; ++j) {
            for (int t = 0; t < P.ntaps; ++t) {
              uint32_t toff = k3 ? (uint32_t)((t / 3) * 10 + (t % 3)) * 16u : 0u;
              uint64_t ad = umma_desc(a_base + (uint32_t)(2 * j) * plane + toff, (uint32_t)plane, a_sbo);
              uint64_t bd = umma_desc(sB_addr + (uint32_t)(ch.kbase + j * P.ntaps + t) * b_step, b_lbo, 128u);
              tc_mma_bf16(d_tmem, ad, bd, P.idesc, accum);
              accum = 1;
            }
          }
          tc_commit(BAR(4 + stage));  // frees the A stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(BAR(9 + as));  // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= kLoadWarp0) {
    // ------------------------------------------------------------------ A-tile loaders
    const int lt = threadIdx.x - kLoadWarp0 * 32;
    uint32_t stage = 0, phase = 0;
    const int H = P.a.H, W = P.a.W, N = P.a.N;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const TileGeom g = tile_geom(P, tile);
      for (int c = 0; c < P.nchunks; ++c) {
        const Chunk ch = P.chunk[c];
        const cg_src& s = P.a.src[ch.src];
        const int nc8 = ch.nc16 * 2;
        const int sh = (nc8 == 4) ? 2 : 1;
        mbar_wait(BAR(4 + stage), phase ^ 1u);
        uint8_t* dst = sA + stage * kStageBytes;
        const int npix = k3 ? 180 : 128;
        const int items = npix << sh;
        for (int it = lt; it < items; it += kLoadThreads) {
          const int c8 = it & (nc8 - 1);
          const int pix = it >> sh;
          bool valid;
          long long off;
          if (k3) {
            const int rr = pix / 10, cc = pix - rr * 10;
            const int v = g.v0 - 1 + rr, w = g.w0 - 1 + cc;
            const int n = v / P.Hp, h = v - n * P.Hp;
            valid = (v >= 0) && (w >= 0) && (w < W) && (n < N) && (h < H);
            off = s.bcast ? (long long)n * s.ld : ((long long)(n * H + h) * W + w) * s.ld;
          } else {
            const long long p = g.p0 + pix;
            valid = p < P.P;
            off = s.bcast ? (p / ((long long)H * W)) * s.ld : p * s.ld;
          }
          uint4 u = load_act8(P, s, off + ch.c0 + c8 * 8, valid);
          *reinterpret_cast<uint4*>(dst + c8 * plane + pix * 16) = u;
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    uint32_t as = 0, aphase = 0;
    const int m = warp * 32 + lane;
    const int H = P.a.H, W = P.a.W, N = P.a.N;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const TileGeom g = tile_geom(P, tile);
      bool valid;
      long long pix;
      if (k3) {
        const int v = g.v0 + (m >> 3), w = g.w0 + (m & 7);
        const int n = v / P.Hp, h = v - n * P.Hp;
        valid = (n < N) && (h < H) && (w < W);
        pix = (long long)(n * H + h) * W + w;
      } else {
        pix = g.p0 + m;
        valid = pix < P.P;
      }
      mbar_wait(BAR(9 + as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + as * (uint32_t)Nc + ((uint32_t)(warp * 32) << 16);
      for (int col = 0; col < Nc; col += 16) {
        float acc[16];
        __syncwarp();  // .aligned TMEM load needs the whole warp converged
        tmem_ld16(t_row + (uint32_t)col, acc);
        const int cg0 = nchunkN * Nc + col;
        if (cg0 >= P.a.cout) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += s_bias[col + i];
        if (!valid) continue;
        for (int sgi = 0; sgi < P.a.nseg; ++sgi) {
          const cg_seg& sg = P.a.seg[sgi];
          const int lc = cg0 - sg.c0;
          if (lc < 0 || lc >= sg.cn) continue;
          const int cnt = min(16, sg.cn - lc);  // 8 or 16
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = acc[i];
          if (sg.mul != nullptr) {
            const bf16* mp = reinterpret_cast<const bf16*>(sg.mul) + pix * sg.mul_ld + lc;
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(__ldg(reinterpret_cast<const uint4*>(mp + h8)), x);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] *= cg_dact(x[i], sg.mul_act);
            }
          }
          if (sg.add != nullptr) {
            const bf16* ap = reinterpret_cast<const bf16*>(sg.add) + pix * sg.add_ld + lc;
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(*reinterpret_cast<const uint4*>(ap + h8), x);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] += x[i];
            }
          }
          if (sg.add2 != nullptr) {
            const bf16* ap = reinterpret_cast<const bf16*>(sg.add2) + pix * sg.add2_ld + lc;
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(*reinterpret_cast<const uint4*>(ap + h8), x);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] += x[i];
            }
          }
          if (sg.dtype == CG_F32) {
            float* op = reinterpret_cast<float*>(sg.ptr) + pix * sg.ld + lc;
            for (int q = 0; q < cnt; q += 4)
              *reinterpret_cast<float4*>(op + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          } else {
            bf16* op = reinterpret_cast<bf16*>(sg.ptr) + pix * sg.ld + lc;
            for (int h8 = 0; h8 < cnt; h8 += 8) *reinterpret_cast<uint4*>(op + h8) = cg_pack8(v + h8);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(11 + as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

int pick_nc(int ktot16, int cout) {
  const int budget = kSmemMax - kHdrBytes - kStages * kStageBytes;
  int nc_max = (budget / (ktot16 * 32)) / 16 * 16;
  if (nc_max > 256) nc_max = 256;
  if (nc_max < 16) return 0;
  int nN = (cout + nc_max - 1) / nc_max;
  int nc = ((cout + nN - 1) / nN + 15) / 16 * 16;
  return nc;
}

}  // namespace

extern "C" int32_t cg_conv_nchunk(int32_t ktot16, int32_t cout) { return pick_nc(ktot16, cout); }

extern "C" int64_t cg_packed_weight_bytes(int32_t ktot16, int32_t cout) {
  int nc = pick_nc(ktot16, cout);
  if (nc <= 0) return -1;
  int nN = (cout + nc - 1) / nc;
  return (int64_t)nN * ktot16 * nc * 32;
}

extern "C" int cg_conv2d(const cg_conv_args* a, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(a != nullptr, "cg_conv2d: null args");
  CG_REQUIRE(a->ksize == 1 || a->ksize == 3, "cg_conv2d: ksize %d not in {1,3}", a->ksize);
  CG_REQUIRE(a->nsrc >= 1 && a->nsrc <= CG_MAX_SRC, "cg_conv2d: nsrc %d", a->nsrc);
  CG_REQUIRE(a->nseg >= 1 && a->nseg <= CG_MAX_SEG, "cg_conv2d: nseg %d", a->nseg);
  CG_REQUIRE(a->cout > 0 && a->cout % 16 == 0, "cg_conv2d: cout %d must be a positive multiple of 16", a->cout);
  CG_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0, "cg_conv2d: empty tensor");
  CG_REQUIRE(a->wpack != nullptr && ((uintptr_t)a->wpack & 15) == 0, "cg_conv2d: wpack null or unaligned");
  KParams kp;
  kp.a = *a;
  kp.ntaps = a->ksize * a->ksize;
  int nchunks = 0, c16 = 0;
  for (int s = 0; s < a->nsrc; ++s) {
    const cg_src& src = a->src[s];
    CG_REQUIRE(src.ptr != nullptr && ((uintptr_t)src.ptr & 15) == 0, "cg_conv2d: src %d null/unaligned", s);
    CG_REQUIRE(src.C > 0 && src.C % 16 == 0 && src.ld % 8 == 0 && src.ld >= src.C,
               "cg_conv2d: src %d C=%d ld=%d (C multiple of 16, ld multiple of 8)", s, src.C, src.ld);
    for (int c0 = 0; c0 < src.C; c0 += 32) {
      CG_REQUIRE(nchunks < kMaxChunks, "cg_conv2d: too many K chunks");
      int n16 = (src.C - c0 >= 32) ? 2 : 1;
      kp.chunk[nchunks++] = Chunk{(uint16_t)s, (uint16_t)c0, (uint16_t)n16, (uint16_t)(c16 * kp.ntaps)};
      c16 += n16;
    }
  }
  kp.nchunks = nchunks;
  kp.ktot16 = c16 * kp.ntaps;
  kp.Nc = pick_nc(kp.ktot16, a->cout);
  CG_REQUIRE(kp.Nc >= 16, "cg_conv2d: K=%d too large for a resident weight slab", kp.ktot16 * 16);
  kp.nN = (a->cout + kp.Nc - 1) / kp.Nc;
  kp.slab_bytes = (uint32_t)kp.ktot16 * kp.Nc * 32u;
  for (int s = 0; s < a->nseg; ++s) {
    const cg_seg& sg = a->seg[s];
    CG_REQUIRE(sg.ptr != nullptr && ((uintptr_t)sg.ptr & 15) == 0, "cg_conv2d: seg %d null/unaligned", s);
    CG_REQUIRE(sg.c0 % 16 == 0 && sg.cn % 8 == 0 && sg.cn > 0 && sg.ld % (sg.dtype == CG_F32 ? 4 : 8) == 0,
               "cg_conv2d: seg %d c0=%d cn=%d ld=%d", s, sg.c0, sg.cn, sg.ld);
    CG_REQUIRE(sg.add == nullptr || (((uintptr_t)sg.add & 15) == 0 && sg.add_ld % 8 == 0), "cg_conv2d: seg %d add", s);
    CG_REQUIRE(sg.add2 == nullptr || (((uintptr_t)sg.add2 & 15) == 0 && sg.add2_ld % 8 == 0), "cg_conv2d: seg %d add2", s);
    CG_REQUIRE(sg.mul == nullptr || (((uintptr_t)sg.mul & 15) == 0 && sg.mul_ld % 8 == 0), "cg_conv2d: seg %d mul", s);
  }
  kp.Hp = a->H + 1;
  kp.V = a->N * kp.Hp;
  kp.P = (long long)a->N * a->H * a->W;
  if (a->ksize == 3) {
    kp.tiles_x = (a->W + 7) / 8;
    kp.ntiles = ((kp.V + 15) / 16) * kp.tiles_x;
  } else {
    kp.tiles_x = 1;
    kp.ntiles = (int)((kp.P + 127) / 128);
  }
  kp.idesc = umma_idesc_bf16(128, kp.Nc, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * kp.Nc) cols <<= 1;
  kp.tmem_cols = cols;
  const int smem_bytes = kHdrBytes + kStages * kStageBytes + (int)kp.slab_bytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  const int sms = cg_device_sms();
  int gx = sms / kp.nN;
  if (gx < 1) gx = 1;
  if (gx > kp.ntiles) gx = kp.ntiles;
  conv_tc_kernel<<<dim3(gx, kp.nN), kThreads, smem_bytes, cg_stream(stream)>>>(kp);
  CG_LAUNCH_CHECK("cg_conv2d");
  return CG_OK;
}
