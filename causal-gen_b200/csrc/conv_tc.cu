// Implicit-GEMM convolution on tcgen05 (sm_100a): forward and data-gradient pass of every
// nn.Conv2d in Block / DecoderBlock (reference src/vae.py:49-84,165-170).
//
//   GEMM view   D[M=128 pixels][N=Cout chunk] += A[pixels][K] * B[Cout][K],  K = taps * Cin
//   A operand   one halo tile of the NHWC bf16 input per K-chunk of 32 channels, staged as channel-octet
//               planes [c8][18 rows][10 px][8 ch]; because 8 consecutive pixels of a plane row are one
//               128-byte UMMA core matrix (SWIZZLE_NONE, K-major), every one of the 9 taps is just a
//               different descriptor start address into the SAME tile -- im2col without copies.
//   rows        the batch is viewed as one tall image of N*(H+1) "virtual rows" (one shared zero row
//               between images) so a 16x8-pixel tile has a constant row pitch for any H.
//   B operand   packed weights for the CTA's Cout chunk, bulk-copied (cp.async.bulk) into shared
//               memory once and kept resident while the persistent CTA walks its pixel tiles.
//   D           fp32 in TMEM, double buffered (2 x Nc columns, Nc <= 64; wider outputs are split over
//               blockIdx.y) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue    bias, channel-split segments, act'(x) multiply (backward), residual / accumulate adds,
//               bf16 or fp32 stores.  The tensors the epilogue READS (residual, pre-activation) are staged
//               through shared memory ("E stages") by the same producer pipeline, up to 3 tiles ahead.
//
// Producer pipeline.  Measured on B200 (profiles/r1c): mbarrier arrivals are the scarce resource -- 256
// per-thread arrivals per stage cost ~0.6 us -- so every hand-off here is ONE arrival:
//   stage-owner warps  warp w owns ring stage w: it alone issues all cp.async (LDGSTS, zero-fill padding) of a
//                      K-chunk, waits for its own copies (wait_group 0), applies the pre-activation in place,
//                      fence.proxy.async, then one lane publishes the stage.  kStages warps work on kStages
//                      different chunks concurrently, so global latency overlaps across warps.
//   E-owner warps      same idea for the epilogue-operand ring (one warp per E stage)
//   MMA warp           one thread issues tcgen05.mma; tcgen05.commit frees the stage / signals the epilogue
//   epilogue warps     TMEM -> registers -> global
//
// Warp roles: 0-3 epilogue (TMEM lane quarters), 4 MMA issuer + TMEM owner, 5-10 A-stage owners,
//             11-13 E-stage owners.
#include <cstdlib>

#include "cg_common.cuh"

namespace {

constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kStages = 6;
constexpr int kOwnWarp0 = 5;                 // warps 5..10 own A stages 0..5
constexpr int kEStages = 3;                  // epilogue-operand ring depth
constexpr int kEWarp0 = kOwnWarp0 + kStages; // warps 11..13 own E stages 0..2
constexpr int kThreads = (kEWarp0 + kEStages) * 32;  // 448
constexpr int kPlane3 = 2976;  // 18*10*16 = 2880, padded so the 4 planes of a stage hit distinct banks
constexpr int kPlane1 = 2080;  // 128*16   = 2048, same padding rule
constexpr int kStageBytes = 4 * kPlane3;
constexpr int kHdrBytes = 1408;   // barriers (<=256B) | tmem slot @256 | bias[256] @320
constexpr int kSmemMax = 232448;  // 227 KB
constexpr int kMaxChunks = 40;
constexpr int kMaxNc = 64;        // GEMM-N per CTA
constexpr int kESlots = 2;        // staged operands per tile
// a staged operand tile is [128 pixel rows][Nc channels] bf16 with a 16-byte row pad (bank spread)
__host__ __device__ constexpr int e_pitch(int nc) { return nc * 2 + 16; }
__host__ __device__ constexpr int e_slot_bytes(int nc) { return 128 * e_pitch(nc); }
__host__ __device__ constexpr int e_bytes(int nc) { return kEStages * kESlots * e_slot_bytes(nc); }

// barrier indices
constexpr int B_AFULL = 0, B_AEMPTY = kStages, B_BFULL = 2 * kStages,
              B_ACCFULL = B_BFULL + 1, B_ACCEMPTY = B_ACCFULL + 2, B_EFULL = B_ACCEMPTY + 2,
              B_EEMPTY = B_EFULL + kEStages, B_COUNT = B_EEMPTY + kEStages;
static_assert(B_COUNT * 8 <= 256, "barrier block overflows the header");

struct Chunk {
  uint16_t src, c0, nc16, kbase;
};

struct EOp {           // one epilogue input staged through shared memory
  const void* ptr;     // bf16 tensor
  int ld, seg, kind;   // kind: 0 add, 1 add2, 2 mul
};

struct KParams {
  cg_conv_args a;
  Chunk chunk[kMaxChunks];
  EOp eop[kESlots];
  int nE, emode;
  int nchunks, ntaps, Nc, nN, ktot16;
  int tiles_x, ntiles, Hp, V;
  uint32_t hp_magic;            // floor(2^32/(H+1))+1: exact n = v/(H+1) for v < 2^24, H < 256
  long long P;                  // N*H*W
  uint32_t idesc, tmem_cols, slab_bytes;
  int dbg;  // CG_DEBUG_SKIP bit mask (profiling experiments only): 1 no cp.async, 2 no act, 4 no mma, 8 no epilogue stores
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

// warp-level wait: one lane polls the mbarrier, the warp re-converges behind it
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

__device__ __forceinline__ uint4 act8(uint4 u, int act) {
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

__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  float* s_bias = reinterpret_cast<float*>(smem + 320);
  uint8_t* sA = smem + kHdrBytes;
  uint8_t* sE = sA + kStages * kStageBytes;
  uint8_t* sB = sE + (P.emode ? e_bytes(P.Nc) : 0);
  const int epitch = e_pitch(P.Nc), eslot = e_slot_bytes(P.Nc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunkN = blockIdx.y;
  const int Nc = P.Nc;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(BAR(B_AFULL + i), 1);
      mbar_init(BAR(B_AEMPTY + i), 1);
    }
    mbar_init(BAR(B_BFULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_ACCFULL + i), 1);
      mbar_init(BAR(B_ACCEMPTY + i), kEpiWarps);
    }
    for (int i = 0; i < kEStages; ++i) {
      mbar_init(BAR(B_EFULL + i), 1);
      mbar_init(BAR(B_EEMPTY + i), kEpiWarps);
    }
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
  const int npix = k3 ? 180 : 128;
  const int H = P.a.H, W = P.a.W, N = P.a.N;

  if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.a.wpack) + (size_t)nchunkN * P.slab_bytes;
      mbar_expect_tx(BAR(B_BFULL), P.slab_bytes);
      for (uint32_t off = 0; off < P.slab_bytes; off += 32768u) {
        uint32_t n = min(32768u, P.slab_bytes - off);
        bulk_g2s(cg_smem_u32(sB) + off, wsrc + off, n, BAR(B_BFULL));
      }
      mbar_wait(BAR(B_BFULL), 0);
      // Descriptors are built once; per MMA only the 14-bit start-address field (low word) advances
      // (all offsets are multiples of 16 B): ~4 instructions per tcgen05.mma for the issuing thread.
      const uint32_t a_sbo = k3 ? 160u : 128u;
      const uint64_t a_d = umma_desc(cg_smem_u32(sA), (uint32_t)plane, a_sbo);
      const uint64_t b_d = umma_desc(cg_smem_u32(sB), (uint32_t)Nc * 16u, 128u);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32);
      const uint32_t a_lo0 = (uint32_t)a_d, b_lo0 = (uint32_t)b_d;
      auto D64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t b_step16 = (uint32_t)Nc * 2u;  // (Nc*32 B per K-block) >> 4
      const uint32_t plane2_16 = (uint32_t)(2 * plane) >> 4;
      const uint32_t stage16 = (uint32_t)kStageBytes >> 4;
      const uint32_t idesc = P.idesc;
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        mbar_wait(BAR(B_ACCEMPTY + as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)Nc;
        uint32_t accum = 0;
        for (int c = 0; c < P.nchunks; ++c) {
          const Chunk ch = P.chunk[c];
          mbar_wait(BAR(B_AFULL + stage), phase);
          tc_fence_after();
          uint32_t alo = a_lo0 + stage * stage16;
          uint32_t blo = b_lo0 + (uint32_t)ch.kbase * b_step16;
          for (int j = 0; j < ch.nc16 && !(P.dbg & 4); ++j) {
            if (k3) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                tc_mma_bf16(d_tmem, D64(a_hi, alo + (uint32_t)((t / 3) * 10 + (t % 3))), D64(b_hi, blo), idesc, accum);
                accum = 1;
                blo += b_step16;
              }
            } else {
              tc_mma_bf16(d_tmem, D64(a_hi, alo), D64(b_hi, blo), idesc, accum);
              accum = 1;
              blo += b_step16;
            }
            alo += plane2_16;
          }
          tc_commit(BAR(B_AEMPTY + stage));  // frees the A stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(BAR(B_ACCFULL + as));  // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= kEWarp0) {
    // ------------------------------------------------------------------ E-stage owner warps
    // warp e stages the tensors the epilogue of tiles e, e+3, e+6 ... (of this CTA) will read: residual /
    // pre-activation rows of the 128 output pixels x Nc channels
    const int nE = P.nE;
    if (nE > 0) {
      const int es = warp - kEWarp0;
      const int ncE8 = Nc >> 3;  // channel octets per staged operand row
      const int items = 128 * ncE8;
      uint32_t ephase = 0;
      int tseq = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++tseq) {
        if (tseq % kEStages != es) continue;
        const TileGeom g = tile_geom(P, tile);
        warp_wait(BAR(B_EEMPTY + es), ephase ^ 1u, lane);
        for (int k = 0; k < nE; ++k) {
          const EOp& op = P.eop[k];
          const cg_seg& sg = P.a.seg[op.seg];
          const uint32_t dstE = cg_smem_u32(sE + (es * kESlots + k) * eslot);
          for (int i = lane; i < items; i += 32) {
            const int row = i / ncE8, oc = i - row * ncE8;
            const int lc = nchunkN * Nc + oc * 8 - sg.c0;  // channel inside the segment
            bool ok;
            long long px;
            if (k3) {
              const int v = g.v0 + (row >> 3), w = g.w0 + (row & 7);
              const int n = (int)__umulhi((uint32_t)v, P.hp_magic), h = v - n * P.Hp;
              ok = (n < N) && (h < H) && (w < W);
              px = (long long)(n * H + h) * W + w;
            } else {
              px = g.p0 + row;
              ok = px < P.P;
            }
            if (ok && lc >= 0 && lc < sg.cn && !(P.dbg & 1))
              cp_async16(dstE + row * epitch + oc * 16, reinterpret_cast<const bf16*>(op.ptr) + px * op.ld + lc, 16u);
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_EFULL + es));
        ephase ^= 1u;
      }
    }
  } else if (warp >= kOwnWarp0) {
    // ------------------------------------------------------------------ A-stage owner warps
    const int st = warp - kOwnWarp0;  // the ring stage this warp owns
    const int act = P.a.act;
    uint8_t* const sbase = sA + st * kStageBytes;
    const uint32_t sbase_u = cg_smem_u32(sbase);
    const int nslots = npix * 4;
    uint32_t phase = 0;
    const long long nwork = (long long)((P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) * P.nchunks;
    for (long long gidx = st; gidx < nwork; gidx += kStages) {
      const int tseq = (int)(gidx / P.nchunks);
      const Chunk ch = P.chunk[(int)(gidx - (long long)tseq * P.nchunks)];
      const TileGeom g = tile_geom(P, (int)blockIdx.x + tseq * (int)gridDim.x);
      const cg_src& s = P.a.src[ch.src];
      const int nc8 = ch.nc16 * 2;
      const bf16* base = reinterpret_cast<const bf16*>(s.ptr) + ch.c0;
      warp_wait(BAR(B_AEMPTY + st), phase ^ 1u, lane);
      if (!(P.dbg & 1)) {
        for (int sl = lane; sl < nslots; sl += 32) {
          const int c8 = sl & 3, pix = sl >> 2;
          if (c8 >= nc8) continue;
          bool valid;
          long long off;
          if (k3) {
            const int rr = pix / 10, cc = pix - rr * 10;
            const int v = g.v0 - 1 + rr, w = g.w0 - 1 + cc;
            const int n = (int)__umulhi((uint32_t)max(v, 0), P.hp_magic), h = v - n * P.Hp;
            valid = (v >= 0) && (w >= 0) && (w < W) && (n < N) && (h < H);
            off = s.bcast ? (long long)n * s.ld : ((long long)(n * H + h) * W + w) * s.ld;
          } else {
            const long long p = g.p0 + pix;
            valid = p < P.P;
            off = s.bcast ? (long long)((uint32_t)min(p, P.P - 1) / (uint32_t)(H * W)) * s.ld : p * s.ld;
          }
          cp_async16(sbase_u + c8 * plane + pix * 16, valid ? base + off + c8 * 8 : base, valid ? 16u : 0u);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();  // this warp's copies only; the other owner warps keep their chunks in flight
      if (act != CG_ACT_NONE && !(P.dbg & 2)) {
        for (int sl = lane; sl < nslots; sl += 32) {
          const int c8 = sl & 3;
          if (c8 >= nc8) continue;
          uint4* p = reinterpret_cast<uint4*>(sbase + c8 * plane + (sl >> 2) * 16);
          *p = act8(*p, act);
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_AFULL + st));
      phase ^= 1u;
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    uint32_t as = 0, aphase = 0, es = 0, ephase = 0;
    const int m = warp * 32 + lane;
    const int nE = P.nE;
    // staged-operand slot of (segment, kind) or -1 -> direct global load
    auto slot_of = [&](int sgi, int kind) {
      for (int k = 0; k < nE; ++k)
        if (P.eop[k].seg == sgi && P.eop[k].kind == kind) return k;
      return -1;
    };
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const TileGeom g = tile_geom(P, tile);
      bool valid;
      long long pix;
      if (k3) {
        const int v = g.v0 + (m >> 3), w = g.w0 + (m & 7);
        const int n = (int)__umulhi((uint32_t)v, P.hp_magic), h = v - n * P.Hp;
        valid = (n < N) && (h < H) && (w < W);
        pix = (long long)(n * H + h) * W + w;
      } else {
        pix = g.p0 + m;
        valid = pix < P.P;
      }
      if (nE > 0) warp_wait(BAR(B_EFULL + es), ephase, lane);
      warp_wait(BAR(B_ACCFULL + as), aphase, lane);
      tc_fence_after();
      const uint8_t* e_row = sE + (size_t)(es * kESlots) * eslot + (size_t)m * epitch;
      const uint32_t t_row = tmem_base + as * (uint32_t)Nc + ((uint32_t)(warp * 32) << 16);
      for (int col = 0; col < Nc; col += 16) {
        float acc[16];
        __syncwarp();  // .aligned TMEM load needs the whole warp converged
        tmem_ld16(t_row + (uint32_t)col, acc);
        const int cg0 = nchunkN * Nc + col;
        if (cg0 >= P.a.cout) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += s_bias[col + i];
        if (!valid || (P.dbg & 8)) continue;
        for (int sgi = 0; sgi < P.a.nseg; ++sgi) {
          const cg_seg& sg = P.a.seg[sgi];
          const int lc = cg0 - sg.c0;
          if (lc < 0 || lc >= sg.cn) continue;
          const int cnt = min(16, sg.cn - lc);  // 8 or 16
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = acc[i];
          // operand fetch: staged tile row (shared memory) if the producers brought it, else global
          auto fetch = [&](int kind, const void* gptr, int gld, int h8) -> uint4 {
            const int k = slot_of(sgi, kind);
            if (k >= 0) return *reinterpret_cast<const uint4*>(e_row + (size_t)k * eslot + (col + h8) * 2);
            return *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(gptr) + pix * gld + lc + h8);
          };
          if (sg.mul != nullptr) {
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(fetch(2, sg.mul, sg.mul_ld, h8), x);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] *= cg_dact(x[i], sg.mul_act);
            }
          }
          if (sg.add != nullptr) {
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(fetch(0, sg.add, sg.add_ld, h8), x);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] += x[i];
            }
          }
          if (sg.add2 != nullptr) {
            for (int h8 = 0; h8 < cnt; h8 += 8) {
              float x[8];
              cg_unpack8(fetch(1, sg.add2, sg.add2_ld, h8), x);
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
      if (lane == 0) {
        mbar_arrive(BAR(B_ACCEMPTY + as));
        if (nE > 0) mbar_arrive(BAR(B_EEMPTY + es));
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
      if (nE > 0 && ++es == kEStages) { es = 0; ephase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// GEMM-N per CTA (<= 64, multiple of 16) such that the resident weight slab fits.  Prefer a size that also
// leaves room for the epilogue-operand ring (*emode = 1); huge-K layers fall back to no ring.
int pick_nc(int ktot16, int cout, int* emode) {
  const int base = kSmemMax - kHdrBytes - kStages * kStageBytes;
  for (int mode = 1; mode >= 0; --mode) {
    for (int nc_max = kMaxNc; nc_max >= 16; nc_max -= 16) {
      if (ktot16 * nc_max * 32 + (mode ? e_bytes(nc_max) : 0) > base) continue;
      const int nN = (cout + nc_max - 1) / nc_max;
      const int nc = ((cout + nN - 1) / nN + 15) / 16 * 16;
      if (emode) *emode = mode;
      return nc;
    }
  }
  return 0;
}

uint32_t magic_of(long long d) { return (uint32_t)((1ull << 32) / (unsigned long long)d + 1ull); }

}  // namespace

extern "C" int32_t cg_conv_nchunk(int32_t ktot16, int32_t cout) { return pick_nc(ktot16, cout, nullptr); }

extern "C" int64_t cg_packed_weight_bytes(int32_t ktot16, int32_t cout) {
  int nc = pick_nc(ktot16, cout, nullptr);
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
  CG_REQUIRE((long long)a->N * (a->H + 1) < (1ll << 24) && (long long)a->N * a->H * a->W < (1ll << 31),
             "cg_conv2d: tensor too large for 32-bit pixel arithmetic");
  CG_REQUIRE(a->wpack != nullptr && ((uintptr_t)a->wpack & 15) == 0, "cg_conv2d: wpack null or unaligned");
  CG_REQUIRE(a->ksize == 1 || a->H < 256, "cg_conv2d: 3x3 path supports H < 256 (got %d)", a->H);
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
  kp.Nc = pick_nc(kp.ktot16, a->cout, &kp.emode);
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
  kp.nE = 0;
  if (kp.emode) {
    for (int s = 0; s < a->nseg && kp.nE < kESlots; ++s) {
      const cg_seg& sg = a->seg[s];
      const void* ptrs[3] = {sg.add, sg.add2, sg.mul};
      const int lds[3] = {sg.add_ld, sg.add2_ld, sg.mul_ld};
      for (int kind = 0; kind < 3 && kp.nE < kESlots; ++kind)
        if (ptrs[kind] != nullptr) kp.eop[kp.nE++] = EOp{ptrs[kind], lds[kind], s, kind};
    }
  }
  kp.Hp = a->H + 1;
  kp.V = a->N * kp.Hp;
  kp.P = (long long)a->N * a->H * a->W;
  kp.hp_magic = magic_of(kp.Hp);
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
  const int smem_bytes = kHdrBytes + kStages * kStageBytes + (kp.emode ? e_bytes(kp.Nc) : 0) + (int)kp.slab_bytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("CG_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; }
  kp.dbg = dbg;
  const int sms = cg_device_sms();
  int gx = sms / kp.nN;
  if (gx < 1) gx = 1;
  if (gx > kp.ntiles) gx = kp.ntiles;
  conv_tc_kernel<<<dim3(gx, kp.nN), kThreads, smem_bytes, cg_stream(stream)>>>(kp);
  CG_LAUNCH_CHECK("cg_conv2d");
  return CG_OK;
}
