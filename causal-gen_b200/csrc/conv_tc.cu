// Implicit-GEMM convolution on tcgen05 (sm_100a): forward and data-gradient pass of every
// nn.Conv2d in Block / DecoderBlock (reference src/vae.py:49-84,165-170).
//
//   layout      activations are bf16 "channel-octet planar": (N, C/8, H, W, 8).  One TMA box
//               (8*10 elements x 18 rows x 4 octets) of that tensor lands in shared memory EXACTLY as the
//               UMMA SWIZZLE_NONE K-major operand: channel-octet planes [c8][18 rows][10 px][8 ch] in which 8
//               consecutive pixels of a row are one 128-byte core matrix.  Image borders are zero-filled by
//               the TMA unit (out-of-bounds box coordinates), so "same" padding costs nothing.
//   GEMM view   D[M=128 pixels (16x8 tile)][N=Cout chunk] += A[pixels][K] * B[Cout][K],  K = taps * Cin.
//               Every one of the 9 taps is just a different descriptor start address (+(kh*10+kw)*16 B)
//               into the SAME halo tile -- im2col without copies.
//   B operand   packed weights for the CTA's Cout chunk, bulk-copied (cp.async.bulk) into shared memory once
//               and kept resident while the persistent CTA walks its pixel tiles.
//   D           fp32 in TMEM, double buffered (2 x Nc columns, Nc <= 64; wider outputs are split over
//               blockIdx.y) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   epilogue    bias, channel-split segments, act'(x) multiply (backward), residual / accumulate adds, stored
//               activation (out_act) and raw + activated dual store (act_copy), bf16 (planar, 128-byte coalesced per
//               octet) or fp32 (row) stores.  The tensors the epilogue READS (residual, pre-activation) are TMA-staged
//               too ("E stages").  Which segment / operands a 16-column chunk maps to is resolved once per warp
//               (ChunkPlan); the common cases run a straight-line templated path (epi_chunk_fast) with only the
//               store predicated.
//
// Pipeline (every hand-off is one mbarrier arrival or a TMA transaction count):
//   producer warp    one elected thread: wait stage free -> expect_tx -> cp.async.bulk.tensor (A chunk / E operands)
//   transform warps  wait "landed", apply the pre-activation (ReLU/GELU) in place, fence.proxy.async, publish
//                    (skipped entirely when the conv has no input activation: the MMA waits on "landed")
//   MMA warp         one elected thread issues tcgen05.mma; tcgen05.commit frees the stage / signals the epilogue
//   epilogue warps   TMEM -> registers -> global
// Ring depths (A stages, E stages) are chosen per launch from the shared memory left after the weight slab.
// Single-thread regions are guarded by elect.sync, NOT `lane == 0`: ptxas then emits tcgen05.mma / TMA straight
// instead of inside a per-active-thread ELECT/BRA.U.ANY loop (measured: ~80 -> ~40 clk per MMA, profiles/r1e_*).
// The kernel is launched with programmatic stream serialization: everything before griddepcontrol.wait (barriers,
// TMEM allocation, bias, weight-slab request) overlaps the tail of the previous kernel in the stream.
//
//   folded mode (template parameter FOLD, KParams::fold): wide-input 3x3 layers with <= 32 outputs put the three kernel
//               COLUMNS on the GEMM-N axis (tile = 8 rows x 16 px, 3 MMAs per K-block, shuffle-add epilogue).
//
// Warp roles: 0-15 epilogue (TMEM lane quarter = warp & 3, 16-column chunk = warp >> 2; chunk-less ones join the transform
//             role), 16 MMA issuer (even tiles) + TMEM owner, 17 TMA producer (even tiles), 18-21 transform, 22 MMA issuer
//             (odd tiles), 23 TMA producer (odd tiles).
#include <cuda.h>

#include <cstdlib>

#include "cg_common.cuh"

namespace {

// Sixteen epilogue warps: warp = quarter + 4*chunk owns TMEM lanes [32*quarter, +32) and ONE 16-column chunk (Nc <= 64).
// ncu on the 8-warp version: 9 cycles of latency per issued instruction and ~250 dependent instructions per warp and tile
// = the whole per-tile time (issue slots 44 % busy, nothing else above 40 %): the epilogue warps' serial instruction
// chains were the bound.  One chunk per warp halves each chain and the accumulator registers (-> 768 threads fit).
constexpr int kEpiWarps = 16;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kTmaWarp = kEpiWarps + 1;
constexpr int kXfWarp0 = kEpiWarps + 2;
constexpr int kXfWarps = 4;
constexpr int kMmaWarp2 = kXfWarp0 + kXfWarps;        // second MMA issuer (tiles 1, 3, 5, ... of the CTA)
constexpr int kTmaWarp2 = kMmaWarp2 + 1;              // second TMA producer (feeds ring 1 = the tiles of the second issuer)
constexpr int kThreads = (kTmaWarp2 + 1) * 32;        // 768
constexpr int kXfThreads = kXfWarps * 32;
constexpr int kMaxStages = 32;  // A-ring depth is chosen per launch from the shared memory left over
constexpr int kMinStages = 4;   //   after the resident weight slab (pick_nc guarantees this many)
constexpr int kPlane3 = 2880;  // 18 rows * 10 px * 16 B
constexpr int kPlane1 = 2048;  // 16 rows *  8 px * 16 B
constexpr int kPlaneF = 2560;  // column-folded 3x3: 10 rows * 16 px * 16 B (row halo only, see KParams::fold)
constexpr int kFoldW = 14;     // valid output pixels per 16-pixel tile row in folded mode
constexpr int kStageBytes = 4 * kPlane3;  // 11520: largest A stage (32 channels with halo); multiple of 128
static_assert(kStageBytes / 16 <= 6 * 128, "transform warps cover a stage in six 16-byte slots per thread");
constexpr int kHdrBytes = 2176;   // barriers (<=1024B) | tmem slot @1024 | bias[256] @1088
constexpr int kSmemMax = 232448;  // 227 KB
constexpr int kMaxChunks = 40;
constexpr int kMaxNc = 64;        // GEMM-N per CTA
constexpr int kMaxEStages = 8;    // epilogue-operand ring depth (chosen per launch, >= kMinEStages)
constexpr int kMinEStages = 2;
constexpr int kESlots = 2;        // staged operands per tile
// a staged operand tile is [Nc/8 octets][16 rows][8 px][8 ch] bf16 = Nc/8 planes of 2048 B
__host__ __device__ constexpr int e_slot_bytes(int nc) { return (nc / 8) * kPlane1; }
__host__ __device__ constexpr int e_bytes_min(int nc) { return kMinEStages * kESlots * e_slot_bytes(nc); }

// barrier indices (fixed slots sized for the maximum ring depths)
constexpr int B_LANDED = 0, B_AFULL = kMaxStages, B_AEMPTY = 2 * kMaxStages, B_BFULL = 3 * kMaxStages,
              B_ACCFULL = B_BFULL + 1, B_ACCEMPTY = B_ACCFULL + 2, B_EFULL = B_ACCEMPTY + 2,
              B_EEMPTY = B_EFULL + kMaxEStages, B_COUNT = B_EEMPTY + kMaxEStages;
static_assert(B_COUNT * 8 <= 1024, "barrier block overflows the header");

struct Chunk {
  uint16_t src, oct0, nc16, kbase;  // source index, first channel octet, K-blocks of 16 (1 or 2), first K-block
};

struct EOp {  // one epilogue input staged through shared memory
  int seg, kind;  // kind: 0 add, 1 add2, 2 mul
  int oct_off;    // (first output channel of the segment) / 8
};

struct alignas(64) KParams {
  CUtensorMap src_map[CG_MAX_SRC];
  CUtensorMap e_map[kESlots];
  cg_conv_args a;
  Chunk chunk[kMaxChunks];
  EOp eop[kESlots];
  uint32_t src_bytes[CG_MAX_SRC];  // TMA transaction bytes of one A box per source
  int nE, emode;
  int nchunks, ntaps, Nc, nN, ktot16;
  int flat;  // 1: H=W=1, samples are the GEMM rows (128 per tile)
  // Column-folded 3x3 (wide input, cout <= 32).  A tcgen05.mma with small N is bound by its 4 KB A-operand read from shared
  // memory (39 clk at N=16, 44 at N=48), and those reads take the shared-memory bandwidth the TMA writes and the activation
  // pass need: on the final kernel a wide-input layer costs loads + MMAs, not max(loads, MMAs) (profiles/r3q_*: 32 -> 8
  // @192^2 168 us, 87 us without the MMAs).  Folded, the GEMM-N axis carries the three kernel COLUMNS (N = 3*Nc):
  // D[p][kx*Nc+co] = sum_{ky,ci} X[p + (ky-1) rows][ci] * W[co][ci][ky][kx], three MMAs (one per kernel row = +256 B start
  // offset) per K-block instead of nine, and the epilogue forms out(x) = D_0(x-1) + D_1(x) + D_2(x+1) with two lane shuffles
  // per channel.  Tile = 8 rows x 16 px (TMEM lane = row*16 + px; core matrix = 8 px of a row, SBO 128 B, no column halo in
  // shared memory), pixels 1..14 of a row are valid outputs (tiles advance by 14 px).
  int fold;
  int Ng;    // GEMM-N per CTA: Nc, or 3*Nc when folded
  int nst, nest;       // A-ring / E-ring depths of this launch
  int nst0;            // stages of ring 0 (even local tiles); ring 1 (odd tiles) has nst - nst0; nst0 == nst: one ring
  int stage_bytes;     // bytes of one A stage (largest source box)
  int tiles_x, tiles_per_img, ntiles;
  long long HW8;  // H*W*8: elements per channel-octet plane
  uint32_t idesc, tmem_cols, slab_bytes, e_tx_bytes;
  unsigned long long* tl;  // debug timeline (CG_TIMELINE builds)
  int dbg;  // CG_DEBUG_SKIP bit mask (profiling experiments only): 2 no act, 4 no mma, 8 no epilogue stores
};

struct TileGeom {
  int n, h0, w0;
};

// Tile walk of a persistent CTA (tile, tile + gridDim.x, ...) without per-tile integer divisions: the stride is
// decomposed once into (images, tile rows, tile columns) and the coordinates advance with carries.  (The producer is one
// thread; two dependent divisions per tile sat on its critical path next to the TMA issue.)
struct TileWalk {
  int n, ty, tx;        // current tile: image, tile row, tile column (flat: n = tile index, ty = tx = 0)
  int dn, dty, dtx;     // gridDim.x decomposed
  int tiles_x, tiles_y;
  __device__ __forceinline__ TileWalk(const KParams& P, int first, int step) {
    tiles_x = P.tiles_x;
    tiles_y = P.tiles_per_img / P.tiles_x;
    n = first / P.tiles_per_img;
    int r = first - n * P.tiles_per_img;
    ty = r / tiles_x;
    tx = r - ty * tiles_x;
    dn = step / P.tiles_per_img;
    r = step - dn * P.tiles_per_img;
    dty = r / tiles_x;
    dtx = r - dty * tiles_x;
  }
  __device__ __forceinline__ void next() {
    tx += dtx;
    if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
    ty += dty;
    if (ty >= tiles_y) { ty -= tiles_y; ++n; }
    n += dn;
  }
  __device__ __forceinline__ TileGeom geom(const KParams& P, bool fold) const {  // fold: compile-time at every call site
    TileGeom g;
    if (P.flat) { g.n = n * 128; g.h0 = g.w0 = 0; }
    else if (fold) { g.n = n; g.h0 = ty * 8; g.w0 = tx * kFoldW - 1; }
    else { g.n = n; g.h0 = ty * 16; g.w0 = tx * 8; }
    return g;
  }
};

// warp-level wait: one lane polls the mbarrier, the warp re-converges behind it
__device__ __forceinline__ void warp_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
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

// Straight-line epilogue of one 16-column chunk for the common case (bf16 destination, all 16 columns live,
// every fused operand staged in shared memory, ReLU-type activations): no data-dependent branches, invalid
// (out-of-image) lanes compute on whatever their shared-memory slot holds and only the store is predicated.
template <bool MUL, bool ADD, bool ADD2, int OACT, bool COPY>
__device__ __forceinline__ void epi_chunk_fast(float (&v)[16], const uint8_t* e0, int eslot, int k_mul, int k_add,
                                               int k_add2, const float* bias, bf16* op, bf16* op2, long long HW8,
                                               bool valid, int noct) {
  if (bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 16; q += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + q);
      v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h >= noct) break;  // a destination with 8 (mod 16) channels stores one octet of its last chunk
    float x[8];
    if (MUL) {  // ReLU'(x): pass the gradient where the saved forward input is positive
      cg_unpack8(*reinterpret_cast<const uint4*>(e0 + k_mul * eslot + h * kPlane1), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] = x[i] > 0.f ? v[8 * h + i] : 0.f;
    }
    if (ADD) {
      cg_unpack8(*reinterpret_cast<const uint4*>(e0 + k_add * eslot + h * kPlane1), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] += x[i];
    }
    if (ADD2) {
      cg_unpack8(*reinterpret_cast<const uint4*>(e0 + k_add2 * eslot + h * kPlane1), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] += x[i];
    }
    if (COPY) {  // raw value to the primary destination, activated value to the copy
      const uint4 o = cg_pack8(v + 8 * h);
      if (valid) *reinterpret_cast<uint4*>(op + h * HW8) = o;
    }
    if (OACT == CG_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] = fmaxf(v[8 * h + i], 0.f);
    } else if (OACT == CG_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * h + i] = cg_gelu(v[8 * h + i]);
    }
    const uint4 o = cg_pack8(v + 8 * h);
    if (valid) *reinterpret_cast<uint4*>((COPY ? op2 : op) + h * HW8) = o;
  }
}

// DUAL: compiled with the act_copy (raw + activated) destinations of GELU blocks; the lean instance serves every
// launch without them (all of UKBB) and keeps the epilogue code small.
// FOLD (column-folded 3x3, KParams::fold) is a template parameter, not a run-time flag: with the flag read from the
// parameters the nine-tap instantiation carried the folded epilogue's registers and branches on its per-tile path and
// lost 4 % of the whole step (2955 against 3070..3089 images/s over six sessions, profiles/r4n_*) -- more than folding
// won back.  (The same lesson as the 16x16-step mode of round 2)
template <bool DUAL, bool FOLD>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 1024);
  float* s_bias = reinterpret_cast<float*>(smem + 1088);
  const int eslot = e_slot_bytes(P.Nc);
  const int nst = P.nst, nest = P.nest, stage_bytes = P.stage_bytes;
  const int estage = P.nE * eslot;  // bytes of one E stage (only the operands this launch stages)
  uint8_t* sA = smem + kHdrBytes;
  uint8_t* sE = sA + nst * stage_bytes;
  uint8_t* sB = sE + nest * estage;

  // warp index through a shuffle so the compiler knows it is warp-uniform (role branches, epilogue plans)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nchunkN = blockIdx.y;
  const int Nc = P.Nc;
  const uint32_t bar0 = cg_smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  const int act = (P.dbg & 2) ? CG_ACT_NONE : P.a.act;
  // Only 4 * (Nc/16) of the sixteen epilogue warps own a chunk.  The others do not idle at the hand-off barriers: they join
  // the four transform warps (the per-stage activation pass is the same kind of latency chain: with Nc = 16 sixteen warps
  // share a stage, two 16-byte slots per thread instead of six), or leave when the conv has no input activation.
  const int n_epi = 4 * ((Nc + 15) >> 4);              // epilogue warps with a chunk
  // warps on the transform role: about two 16-byte slots of the largest stage per thread, at least the four dedicated ones
  // (more warps than that only add barrier polling: 8 -> 32 @192^2 lost 4 % with twelve warps on a 360-slot stage)
  int n_xfw = (P.stage_bytes / 16 + 31) / 32;
  n_xfw = n_xfw < kXfWarps ? kXfWarps : (n_xfw > kXfWarps + kEpiWarps - n_epi ? kXfWarps + kEpiWarps - n_epi : n_xfw);
  // With two A rings the transform warps form two TEAMS, one per ring (= per tile parity): a stage's pass is a chain of
  // wait -> loads -> stores -> fence -> arrive whose fixed part outweighs the per-slot part, so two chains in flight beat
  // one chain with twice the threads.
  const int n_xf0 = (P.nst0 < P.nst) ? (n_xfw + 1) / 2 : n_xfw;  // warps of team 0 (ring 0); team 1 has the rest
  if (threadIdx.x == 0) CG_TL(P.tl, 32);

  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(BAR(B_LANDED + i), 1);
      mbar_init(BAR(B_AFULL + i), i < P.nst0 ? n_xf0 : n_xfw - n_xf0);
      mbar_init(BAR(B_AEMPTY + i), 1);
    }
    mbar_init(BAR(B_BFULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_ACCFULL + i), 1);
      mbar_init(BAR(B_ACCEMPTY + i), n_epi);
    }
    for (int i = 0; i < nest; ++i) {
      mbar_init(BAR(B_EFULL + i), 1);
      mbar_init(BAR(B_EEMPTY + i), n_epi);
    }
    mbar_fence_init();
    // resident weight slab of this CTA's Cout chunk.  Packed weights were written at the start of the step, not
    // by the kernel right before this one, so the copy may start before the grid dependency is resolved (below).
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.a.wpack) + (size_t)nchunkN * P.slab_bytes;
    mbar_expect_tx(BAR(B_BFULL), P.slab_bytes);
    for (uint32_t off = 0; off < P.slab_bytes; off += 32768u) {
      uint32_t n = min(32768u, P.slab_bytes - off);
      bulk_g2s(cg_smem_u32(sB) + off, wsrc + off, n, BAR(B_BFULL));
    }
  }
  if (warp == kMmaWarp) tmem_alloc(cg_smem_u32(tmem_slot), P.tmem_cols);
  for (int i = threadIdx.x; i < Nc; i += kThreads) {
    int c = nchunkN * Nc + i;
    s_bias[i] = (P.a.bias != nullptr && c < P.a.bias_n) ? P.a.bias[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Programmatic dependent launch: everything above (barriers, TMEM, bias, weight slab request) overlapped the tail
  // of the previous kernel in the stream.  From here on we touch its outputs: wait for it, then let OUR dependent
  // start its own prologue (after our wait, so everything older than this kernel is complete for it).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CG_TL(P.tl, 33);
  const bool k3 = P.a.ksize == 3;
  constexpr bool fold = FOLD;
  const int plane = fold ? kPlaneF : (k3 ? kPlane3 : kPlane1);
  const int Ng = fold ? P.Ng : Nc;
  const int H = P.a.H, W = P.a.W, N = P.a.N;

  if (warp == kMmaWarp || warp == kMmaWarp2) {
    // ------------------------------------------------------------------ MMA issuers
    // TWO issuing warps on alternate tiles (issuer w: local tiles w, w+2, ... into accumulator w).  The tensor pipe's
    // instruction queue is shallow: with one issuer, the scalar work between two tiles (barrier polls, fences, descriptor
    // set-up, commits: ~300 clk) drains it and every tile pays that bubble on top of its MMAs -- measured 1084 -> 711 clk
    // per 18-MMA tile and 1854 -> 1409 clk per 36-MMA tile with the second issuer (tools/micro/umma_rate2.cu,
    // profiles/r2i_umma_rate2.txt).  tcgen05.commit tracks the MMAs of the executing thread, so each issuer's commits
    // cover exactly its own tile.
    const uint32_t iw = warp == kMmaWarp ? 0u : 1u;
    if (elect_one()) {
      mbar_wait(BAR(B_BFULL), 0);  // weight slab (requested by thread 0 in the prologue)
      if (iw == 0) CG_TL(P.tl, 34);
      int tl_i = iw == 0 ? 0 : 1000;  // timeline marks: issuer 0 only (its tiles 0, 2, 4, ...)
      (void)tl_i;
      // Descriptors are built once; per MMA only the 14-bit start-address field (low word) advances
      // (all offsets are multiples of 16 B).
      const uint32_t a_sbo = (k3 && !fold) ? 160u : 128u;
      const uint64_t a_d = umma_desc(cg_smem_u32(sA), (uint32_t)plane, a_sbo);
      const uint64_t b_d = umma_desc(cg_smem_u32(sB), (uint32_t)Ng * 16u, 128u);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), b_hi = (uint32_t)(b_d >> 32);
      const uint32_t a_lo0 = (uint32_t)a_d, b_lo0 = (uint32_t)b_d;
      auto D64 = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      const uint32_t b_step16 = (uint32_t)Ng * 2u;  // (Ng*32 B per K-block) >> 4
      const uint32_t plane2_16 = (uint32_t)(2 * plane) >> 4;
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t idesc = P.idesc;
      const int ready0 = (act == CG_ACT_NONE) ? B_LANDED : B_AFULL;  // no activation: consume TMA data directly
      // Ring split: even local tiles live in ring 0 (stages [0, nst0)), odd tiles in ring 1 (stages [nst0, nst)), and issuer
      // w only ever touches ring w.  Every waiter therefore observes EVERY fill of the stages it waits on, in order --
      // an mbarrier parity wait is only valid for the phase right after the last one the waiter saw, so an issuer that
      // skipped the other's fills of a shared stage could read a stale parity (that version deadlocked).
      // One tile per CTA (nst0 == nst): a single ring, issuer 1 idles.
      const bool two = P.nst0 < nst;
      const uint32_t rbase = (two && iw) ? (uint32_t)P.nst0 : 0u, rn = two ? (iw ? (uint32_t)(nst - P.nst0) : (uint32_t)P.nst0) : (uint32_t)nst;
      uint32_t stage = 0, phase = 0, aphase = 0, as = two ? iw : 0u;
      const int tstep = two ? 2 * (int)gridDim.x : (int)gridDim.x;
      for (int tile = blockIdx.x + (two ? (int)iw * (int)gridDim.x : 0); tile < P.ntiles && (two || iw == 0); tile += tstep) {
        mbar_wait(BAR(B_ACCEMPTY + as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)Ng;
        uint32_t accum = 0;
        for (int c = 0; c < P.nchunks; ++c) {
          const Chunk ch = P.chunk[c];
          const uint32_t gs = rbase + stage;
          mbar_wait(BAR(ready0 + gs), phase);
          tc_fence_after();
          if (c == 0 && tl_i < 6) CG_TL(P.tl, 35 + 2 * tl_i);
          uint32_t alo = a_lo0 + gs * stage16;
          uint32_t blo = b_lo0 + (uint32_t)ch.kbase * b_step16;
          for (int j = 0; j < ch.nc16 && !(P.dbg & 4); ++j) {
            if (fold) {
#pragma unroll
              for (int t = 0; t < 3; ++t) {  // kernel rows: +256 B (one 16-pixel row) per tap, columns live on GEMM-N
                tc_mma_bf16(d_tmem, D64(a_hi, alo + (uint32_t)(t * 16)), D64(b_hi, blo), idesc, accum);
                accum = 1;
                blo += b_step16;
              }
            } else if (k3) {
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
          tc_commit(BAR(B_AEMPTY + gs));  // frees the A stage once these MMAs retire
          if (++stage == rn) { stage = 0; phase ^= 1u; }
        }
        tc_commit(BAR(B_ACCFULL + as));  // accumulator ready for the epilogue
        if (tl_i < 6) CG_TL(P.tl, 36 + 2 * tl_i);
        ++tl_i;
        if (two) aphase ^= 1u;
        else if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp == kTmaWarp || warp == kTmaWarp2) {
    // ------------------------------------------------------------------ TMA producers
    // TWO producer threads, one per A ring (producer w: local tiles w, w+2, ... -- the tiles issuer w consumes).  One thread
    // issues a box every ~0.3 us (barrier poll + expect_tx + UTMALDG), which made the producer, not the TMA unit (0.22 us per
    // box of this shape, tools/micro/tma_rate.cu) or HBM, the bound of the high-resolution layers (profiles/
    // r2p_timeline_producer_marks.txt).  The epilogue-operand ring is walked in tile order by the epilogue, so with two
    // producers its depth is even and producer w only ever owns the stages of parity w (it sees every phase of them).
    const uint32_t pw = warp == kTmaWarp ? 0u : 1u;
    const bool two = P.nst0 < nst;
    if (elect_one() && (two || pw == 0)) {
      if (pw == 0) {
        for (int s = 0; s < P.a.nsrc; ++s) tma_prefetch_desc(&P.src_map[s]);
        for (int k = 0; k < P.nE; ++k) tma_prefetch_desc(&P.e_map[k]);
      }
      const uint32_t rbase = (two && pw) ? (uint32_t)P.nst0 : 0u;
      const uint32_t rlen = two ? (pw ? (uint32_t)(nst - P.nst0) : (uint32_t)P.nst0) : (uint32_t)nst;
      const uint32_t lstep = two ? 2u : 1u;
      uint32_t es = two ? pw : 0u, ephase = 0, lt = two ? pw : 0u, st = 0, ph = 0;
      const int halo_y = k3 ? 1 : 0, halo_x = (k3 && !fold) ? 1 : 0;
      const int first = blockIdx.x + (int)lt * (int)gridDim.x, tstep = (int)lstep * (int)gridDim.x;
      TileWalk walk(P, first, tstep);
      for (int tile = first; tile < P.ntiles; tile += tstep, lt += lstep, walk.next()) {
        const TileGeom g = walk.geom(P, fold);
        if (pw == 0 && lt < 12) CG_TL(P.tl, 90 + 2 * lt);
        if (P.nE > 0) {
          mbar_wait(BAR(B_EEMPTY + es), ephase ^ 1u);
          mbar_expect_tx(BAR(B_EFULL + es), P.e_tx_bytes);
          for (int k = 0; k < P.nE; ++k) {
            const uint32_t dst = cg_smem_u32(sE + es * estage + k * eslot);
            const int oct = nchunkN * (Nc >> 3) - P.eop[k].oct_off;  // may be out of range: zero-filled
            if (P.flat) tma_load_3d(dst, &P.e_map[k], 0, g.n, oct, BAR(B_EFULL + es));
            else tma_load_4d(dst, &P.e_map[k], g.w0 * 8, g.h0, oct, g.n, BAR(B_EFULL + es));
          }
          es += lstep;
          if (es >= (uint32_t)nest) { es -= (uint32_t)nest; ephase ^= 1u; }
        }
        for (int c = 0; c < P.nchunks; ++c) {
          const Chunk ch = P.chunk[c];
          const uint32_t stage = rbase + st;
          mbar_wait(BAR(B_AEMPTY + stage), ph ^ 1u);
          mbar_expect_tx(BAR(B_LANDED + stage), P.src_bytes[ch.src]);
          const uint32_t dst = cg_smem_u32(sA + stage * stage_bytes);
          if (P.flat) tma_load_3d(dst, &P.src_map[ch.src], 0, g.n, ch.oct0, BAR(B_LANDED + stage));
          else tma_load_4d(dst, &P.src_map[ch.src], (g.w0 - halo_x) * 8, g.h0 - halo_y, ch.oct0, g.n, BAR(B_LANDED + stage));
          if (++st == rlen) { st = 0; ph ^= 1u; }
        }
        if (pw == 0 && lt < 12) CG_TL(P.tl, 91 + 2 * lt);
      }
    }
  } else if ((warp >= kXfWarp0 && warp < kMmaWarp2) || (warp < kEpiWarps && warp >= n_epi && warp - n_epi < n_xfw - kXfWarps)) {
    // ------------------------------------------------------------------ transform warps (+ chunk-less epilogue warps)
    if (act != CG_ACT_NONE) {
      const int xw = warp < kEpiWarps ? warp - n_epi : (n_xfw - kXfWarps) + warp - kXfWarp0;  // index among the transform warps
      const bool two = P.nst0 < nst;
      const int team = (two && xw >= n_xf0) ? 1 : 0;
      const int nthr = (team ? n_xfw - n_xf0 : n_xf0) * 32;
      const int xt = (xw - (team ? n_xf0 : 0)) * 32 + lane;
      const uint32_t rbase = team ? (uint32_t)P.nst0 : 0u;
      const uint32_t rlen = two ? (team ? (uint32_t)(nst - P.nst0) : (uint32_t)P.nst0) : (uint32_t)nst;
      const int tstep = (two ? 2 : 1) * (int)gridDim.x;
      uint32_t st = 0, ph = 0;
      for (int tile = blockIdx.x + team * (int)gridDim.x; tile < P.ntiles; tile += tstep) {
        for (int c = 0; c < P.nchunks; ++c) {
          const int n16 = (int)(P.src_bytes[P.chunk[c].src] >> 4);  // 16-byte slots the TMA box filled
          const uint32_t stage = rbase + st;
          warp_wait(BAR(B_LANDED + stage), ph, lane);
          uint4* base = reinterpret_cast<uint4*>(sA + stage * stage_bytes) + xt;
          // `per` 16-byte slots per thread (a stage holds <= 720, a team has >= 64 threads).  Usual case per <= 3: straight
          // code, all loads first, then activate + store (one round trip of shared-memory latency); else rounds of six.
          const int per = (n16 + nthr - 1) / nthr;  // warp-uniform
          if (per <= 3) {
            uint4 v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
              if (k < per && xt + k * nthr < n16) v[k] = base[k * nthr];
#pragma unroll
            for (int k = 0; k < 3; ++k)
              if (k < per && xt + k * nthr < n16) base[k * nthr] = act8(v[k], act);
          } else {
            for (int k0 = 0; k0 < per; k0 += 6) {
              uint4 v[6];
#pragma unroll
              for (int k = 0; k < 6; ++k)
                if (k0 + k < per && xt + (k0 + k) * nthr < n16) v[k] = base[(k0 + k) * nthr];
#pragma unroll
              for (int k = 0; k < 6; ++k)
                if (k0 + k < per && xt + (k0 + k) * nthr < n16) base[(k0 + k) * nthr] = act8(v[k], act);
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_AFULL + stage));
          if (++st == rlen) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < n_epi) {
    // ------------------------------------------------------------------ epilogue
    // warp = quarter + 4*chunk: TMEM lanes [32*quarter, +32) (one pixel per lane) and the 16-column chunk `chunk` of the
    // CTA's GEMM-N range (Nc <= 64 -> <= 4 chunks; warps whose chunk lies beyond Nc help the transform role or leave).
    // Which segment / staged operands the chunk maps to is the same for every tile, so it is resolved ONCE here
    // (ChunkPlan); per tile the warp issues its TMEM load, releases the accumulator and then does bias / act' / residual
    // / store from registers.
    const int quarter = warp & 3, chunk = warp >> 2;
    const int m = quarter * 32 + lane;
    const int nE = P.nE;
    const bool has_bias = P.a.bias != nullptr;
    struct ChunkPlan {
      int col, sgi, lc, cnt;        // sgi < 0: chunk has no destination (padding / beyond cout)
      int k_add, k_add2, k_mul;     // -2 absent, -1 read from global memory, >= 0 staged E slot
      int dtype, mul_act, out_act;
      int fast;                     // -1: generic path; else bit0 mul, bit1 add, bit2 add2, bit3 relu, bit4 gelu, bit5 copy
      uint8_t* out2;                // act_copy destination (+ plane offset) or nullptr
      long long ns2;                // its sample stride
      uint8_t* out;                 // bf16: ptr + (lc/8) planes; fp32: ptr + lc floats
      long long ns;
    } plan[1];
#pragma unroll
    for (int j = 0; j < 1; ++j) {
      ChunkPlan& pl = plan[j];
      pl.col = 16 * chunk;
      pl.sgi = -1;
      pl.lc = pl.cnt = 0;
      pl.k_add = pl.k_add2 = pl.k_mul = -2;
      pl.dtype = CG_BF16;
      pl.mul_act = CG_ACT_NONE;
      pl.out_act = CG_ACT_NONE;
      pl.fast = -1;
      pl.out = nullptr;
      pl.out2 = nullptr;
      pl.ns2 = 0;
      pl.ns = 0;
      const int cg0 = nchunkN * Nc + pl.col;
      if (pl.col >= Nc || cg0 >= P.a.cout || (P.dbg & 8)) continue;
      for (int sgi = 0; sgi < P.a.nseg; ++sgi) {
        const cg_seg& sg = P.a.seg[sgi];
        const int lc = cg0 - sg.c0;
        if (lc < 0 || lc >= sg.cn) continue;
        pl.sgi = sgi;
        pl.lc = lc;
        pl.cnt = min(16, sg.cn - lc);
        pl.dtype = sg.dtype;
        pl.mul_act = sg.mul_act;
        pl.out_act = sg.out_act;
        pl.ns = sg.ns;
        pl.out = reinterpret_cast<uint8_t*>(sg.ptr) +
                 (sg.dtype == CG_F32 ? (long long)lc * 4 : (long long)(lc >> 3) * P.HW8 * 2);
        if (DUAL && sg.act_copy != nullptr) {
          pl.out2 = reinterpret_cast<uint8_t*>(sg.act_copy) + (long long)(lc >> 3) * P.HW8 * 2;
          pl.ns2 = sg.act_copy_ns;
        }
        if (sg.add != nullptr) pl.k_add = -1;
        if (sg.add2 != nullptr) pl.k_add2 = -1;
        if (sg.mul != nullptr) pl.k_mul = -1;
        for (int k = 0; k < nE; ++k)
          if (P.eop[k].seg == sgi) {
            if (P.eop[k].kind == 0) pl.k_add = k;
            else if (P.eop[k].kind == 1) pl.k_add2 = k;
            else pl.k_mul = k;
          }
        if (sg.dtype == CG_BF16 && (pl.cnt == 16 || pl.cnt == 8) && pl.k_add != -1 && pl.k_add2 != -1 && pl.k_mul != -1 &&
            sg.out_act <= CG_ACT_GELU &&  // LeakyReLU (predictors) takes the generic path
            (pl.k_mul == -2 || sg.mul_act == CG_ACT_RELU) && !(pl.k_mul >= 0 && pl.k_add2 >= 0))
          pl.fast = (pl.k_mul >= 0 ? 1 : 0) | (pl.k_add >= 0 ? 2 : 0) | (pl.k_add2 >= 0 ? 4 : 0) |
                    (sg.out_act == CG_ACT_RELU ? 8 : 0) | (sg.out_act == CG_ACT_GELU ? 16 : 0) |
                    (sg.act_copy != nullptr ? 32 : 0);
        break;
      }
    }
    const ChunkPlan& pl = plan[0];
    // per-thread part of the pixel offset / validity: the per-tile part is warp-uniform
    const int mrow = fold ? (m >> 4) : (m >> 3), mpx = fold ? (m & 15) : (m & 7);  // position of the lane inside the tile
    const int moff = P.flat ? 0 : mrow * W + mpx;
    const bool all_valid = !P.flat && !fold && (H & 15) == 0 && (W & 7) == 0;
    const bool px_ok = !fold || (mpx >= 1 && mpx <= kFoldW);  // folded: pixels 0 and 15 of a tile row are halo
    uint32_t as = 0, aphase = 0, es = 0, ephase = 0;
    int tl_i = 0;
    (void)tl_i;
    // The tile loop, parameterised by what happens to the chunk once it is in registers: the dispatch on the chunk's
    // epilogue variant is done ONCE per kernel (below), not per tile -- every instruction of this loop sits on the warp's
    // serial per-tile chain, which is what bounds the kernel (ncu: ~210 instructions per warp and tile, 9 cycles each).
    auto tile_loop = [&](auto&& chunk_fn) {
      TileWalk walk(P, blockIdx.x, gridDim.x);
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, walk.next()) {
        const TileGeom g = walk.geom(P, fold);
        bool valid;
        int n;
        long long hw;  // pixel index inside the image
        if (P.flat) {
          n = g.n + m;
          hw = 0;
          valid = n < N;
        } else {
          n = g.n;
          hw = (long long)(g.h0 * W + g.w0 + moff);
          valid = all_valid || (px_ok && (g.h0 + mrow < H) && (g.w0 + mpx < W));
        }
        if (nE > 0) warp_wait(BAR(B_EFULL + es), ephase, lane);
        warp_wait(BAR(B_ACCFULL + as), aphase, lane);
        tc_fence_after();
        const uint8_t* e_row = sE + (size_t)es * estage + (size_t)m * 16;
        const uint32_t t_row = tmem_base + as * (uint32_t)Ng + ((uint32_t)(quarter * 32) << 16);
        float acc[16];
        __syncwarp();  // .aligned TMEM loads need the whole warp converged
        if (fold) {
          // out(x) = D_0(x-1) + D_1(x) + D_2(x+1): the three column partials sit Nc columns apart, lanes of a 16-pixel row
          // are neighbours (pixels 0 and 15 are halo).  One partial at a time keeps 32 values live, not 48.
          float side[16];
          tmem_ld16_nowait(t_row + (uint32_t)(Nc + pl.col), acc);
          tmem_ld16_nowait(t_row + (uint32_t)pl.col, side);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] += __shfl_up_sync(0xffffffffu, side[i], 1);
          tmem_ld16_nowait(t_row + (uint32_t)(2 * Nc + pl.col), side);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] += __shfl_down_sync(0xffffffffu, side[i], 1);
        } else {
          tmem_ld16_nowait(t_row + (uint32_t)pl.col, acc);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACCEMPTY + as));  // accumulator is in registers: MMA may reuse it
        chunk_fn(acc, n, hw, valid, e_row);
        if (nE > 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_EEMPTY + es));
          if (++es == nest) { es = 0; ephase ^= 1u; }
        }
        if (threadIdx.x == 0 && tl_i < 8) CG_TL(P.tl, 50 + tl_i);
        ++tl_i;
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    };
    auto generic_fn = [&](float (&acc)[16], int n, long long hw, bool valid, const uint8_t* e_row) {
        if (!valid) return;
        float* v = acc;
        if (has_bias) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + pl.col + q);
            v[q] += b4.x; v[q + 1] += b4.y; v[q + 2] += b4.z; v[q + 3] += b4.w;
          }
        }
        auto fetch = [&](int k, int kind, int h8, float* x) {
          uint4 u;
          if (k >= 0) {
            u = *reinterpret_cast<const uint4*>(e_row + (size_t)k * eslot + ((pl.col + h8) >> 3) * kPlane1);
          } else {
            const cg_seg& sg = P.a.seg[pl.sgi];
            const void* gp = kind == 0 ? sg.add : (kind == 1 ? sg.add2 : sg.mul);
            const long long gns = kind == 0 ? sg.add_ns : (kind == 1 ? sg.add2_ns : sg.mul_ns);
            u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(gp) + n * gns +
                                                ((pl.lc + h8) >> 3) * P.HW8 + hw * 8);
          }
          cg_unpack8(u, x);
        };
#pragma unroll
        for (int h8 = 0; h8 < 16; h8 += 8) {
          if (h8 >= pl.cnt) continue;
          float x[8];
          if (pl.k_mul != -2) {
            fetch(pl.k_mul, 2, h8, x);
            if (pl.mul_act == CG_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] = x[i] > 0.f ? v[h8 + i] : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[h8 + i] *= cg_dact(x[i], pl.mul_act);
            }
          }
          if (pl.k_add != -2) {
            fetch(pl.k_add, 0, h8, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h8 + i] += x[i];
          }
          if (pl.k_add2 != -2) {
            fetch(pl.k_add2, 1, h8, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h8 + i] += x[i];
          }
          if (DUAL && pl.out2 != nullptr) {  // raw value first, the activated copy goes to act_copy
            bf16* orw = reinterpret_cast<bf16*>(pl.out) + n * pl.ns + (long long)(h8 >> 3) * P.HW8 + hw * 8;
            *reinterpret_cast<uint4*>(orw) = cg_pack8(v + h8);
          }
          if (pl.out_act == CG_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h8 + i] = fmaxf(v[h8 + i], 0.f);
          } else if (pl.out_act == CG_ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h8 + i] = cg_gelu(v[h8 + i]);
          } else if (pl.out_act == CG_ACT_LRELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[h8 + i] = v[h8 + i] > 0.f ? v[h8 + i] : 0.01f * v[h8 + i];
          }
          if (DUAL && pl.out2 != nullptr) {
            bf16* oc = reinterpret_cast<bf16*>(pl.out2) + n * pl.ns2 + (long long)(h8 >> 3) * P.HW8 + hw * 8;
            *reinterpret_cast<uint4*>(oc) = cg_pack8(v + h8);
          } else if (pl.dtype == CG_F32) {  // fp32 statistics: row layout (pixel, channel), pitch ns
            float* op = reinterpret_cast<float*>(pl.out) + ((long long)n * H * W + hw) * pl.ns + h8;
            *reinterpret_cast<float4*>(op) = make_float4(v[h8], v[h8 + 1], v[h8 + 2], v[h8 + 3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(v[h8 + 4], v[h8 + 5], v[h8 + 6], v[h8 + 7]);
          } else {  // bf16 planar: 8 lanes of a tile row write one contiguous 128-byte line per octet
            bf16* op = reinterpret_cast<bf16*>(pl.out) + n * pl.ns + (long long)(h8 >> 3) * P.HW8 + hw * 8;
            *reinterpret_cast<uint4*>(op) = cg_pack8(v + h8);
          }
        }
    };
    if (pl.sgi < 0) {
      tile_loop([](float (&)[16], int, long long, bool, const uint8_t*) {});  // padding chunk: hand-offs only
    } else if (pl.fast < 0) {
      tile_loop(generic_fn);
    } else {
      const float* bp = has_bias ? s_bias + pl.col : nullptr;
#define CG_EPI(code, M, A, A2, OA, CP)                                                                                   \
  case code:                                                                                                             \
    tile_loop([&](float (&acc)[16], int n, long long hw, bool valid, const uint8_t* e_row) {                             \
      epi_chunk_fast<M, A, A2, OA, CP>(acc, e_row + (pl.col >> 3) * kPlane1, eslot, pl.k_mul, pl.k_add, pl.k_add2, bp,   \
                                       reinterpret_cast<bf16*>(pl.out) + n * pl.ns + hw * 8,                             \
                                       reinterpret_cast<bf16*>(pl.out2) + n * pl.ns2 + hw * 8, P.HW8, valid, pl.cnt >> 3); \
    });                                                                                                                  \
    break;
      switch (pl.fast) {
        CG_EPI(0, false, false, false, 0, false) CG_EPI(8, false, false, false, 1, false)
        CG_EPI(2, false, true, false, 0, false) CG_EPI(10, false, true, false, 1, false)
        CG_EPI(6, false, true, true, 0, false) CG_EPI(14, false, true, true, 1, false)
        CG_EPI(1, true, false, false, 0, false) CG_EPI(3, true, true, false, 0, false)
        default:
          if (DUAL && pl.fast == 48) {
            tile_loop([&](float (&acc)[16], int n, long long hw, bool valid, const uint8_t* e_row) {
              epi_chunk_fast<false, false, false, 2, true>(acc, e_row + (pl.col >> 3) * kPlane1, eslot, pl.k_mul, pl.k_add,
                                                           pl.k_add2, bp, reinterpret_cast<bf16*>(pl.out) + n * pl.ns + hw * 8,
                                                           reinterpret_cast<bf16*>(pl.out2) + n * pl.ns2 + hw * 8, P.HW8,
                                                           valid, pl.cnt >> 3);
            });
          } else if (DUAL && pl.fast == 40) {
            tile_loop([&](float (&acc)[16], int n, long long hw, bool valid, const uint8_t* e_row) {
              epi_chunk_fast<false, false, false, 1, true>(acc, e_row + (pl.col >> 3) * kPlane1, eslot, pl.k_mul, pl.k_add,
                                                           pl.k_add2, bp, reinterpret_cast<bf16*>(pl.out) + n * pl.ns + hw * 8,
                                                           reinterpret_cast<bf16*>(pl.out2) + n * pl.ns2 + hw * 8, P.HW8,
                                                           valid, pl.cnt >> 3);
            });
          } else if (DUAL && pl.fast == 16) {
            tile_loop([&](float (&acc)[16], int n, long long hw, bool valid, const uint8_t* e_row) {
              epi_chunk_fast<false, false, false, 2, false>(acc, e_row + (pl.col >> 3) * kPlane1, eslot, pl.k_mul, pl.k_add,
                                                            pl.k_add2, bp, reinterpret_cast<bf16*>(pl.out) + n * pl.ns + hw * 8,
                                                            reinterpret_cast<bf16*>(pl.out2) + n * pl.ns2 + hw * 8, P.HW8,
                                                            valid, pl.cnt >> 3);
            });
          } else {
            tile_loop(generic_fn);
          }
          break;
      }
#undef CG_EPI
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) CG_TL(P.tl, 60);
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// GEMM-N per CTA (<= 64, multiple of 16) such that the resident weight slab fits.  want_e: the launches of this pack
// stage epilogue operands (residual / act' mask), so prefer a size that leaves room for the operand ring (*emode = 1);
// packs whose launches never have such operands (first convs of a Block: huge K, narrow N) take the largest chunk.
// Huge-K layers fall back to no ring (operands are then read straight from global memory by the epilogue).
// Wider chunks (<= 128 columns per CTA) were built and measured in round 2 (profiles/r2c_microbench_b128.txt): with K
// this small the epilogue, not the A-tile traffic, bounds wide outputs, and two CTAs beat one wide CTA.
int pick_nc(int ktot16, int cout, int* emode, int want_e) {
  const int base = kSmemMax - kHdrBytes - kMinStages * kStageBytes;
  for (int mode = want_e ? 1 : 0; mode >= 0; --mode) {
    for (int nc_max = kMaxNc; nc_max >= 16; nc_max -= 16) {
      if (ktot16 * nc_max * 32 + (mode ? e_bytes_min(nc_max) : 0) > base) continue;
      const int nN = (cout + nc_max - 1) / nc_max;
      const int nc = ((cout + nN - 1) / nN + 15) / 16 * 16;
      if (emode) *emode = mode;
      return nc;
    }
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// Tensor map over a planar bf16 view (N, C8, H, W, 8) whose sample stride is `ns` elements.
//   spatial: dims (W*8, H, C8, N), box (box_w8, box_h, box_c8, 1)
//   flat (H=W=1): dims (8, N, C8), box (8, 128, box_c8)   -- samples are the GEMM rows
int cg_make_planar_map(void* map_out, const void* ptr, long long ns, int N, int H, int W, int C8, int flat, int box_w8,
                       int box_h, int box_c8) {
  CUtensorMap* map = reinterpret_cast<CUtensorMap*>(map_out);
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) {
    cg_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return CG_ERR_CUDA;
  }
  CUresult r;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  if (flat) {
    const cuuint64_t dims[3] = {8, (cuuint64_t)N, (cuuint64_t)C8};
    const cuuint64_t strides[2] = {(cuuint64_t)ns * 2, 16};
    const cuuint32_t box[3] = {8, 128, (cuuint32_t)box_c8};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)C8, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)ns * 2};
    const cuuint32_t box[4] = {(cuuint32_t)box_w8, (cuuint32_t)box_h, (cuuint32_t)box_c8, 1};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, ones,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    cg_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p ns=%lld N=%d H=%d W=%d C8=%d flat=%d box=(%d,%d,%d)", (int)r,
                 ptr, ns, N, H, W, C8, flat, box_w8, box_h, box_c8);
    return CG_ERR_CUDA;
  }
  return CG_OK;
}

extern "C" int32_t cg_conv_nchunk_ex(int32_t ktot16, int32_t cout, int32_t want_e) {
  return pick_nc(ktot16, cout, nullptr, want_e);
}
extern "C" int32_t cg_conv_nchunk(int32_t ktot16, int32_t cout) { return pick_nc(ktot16, cout, nullptr, 1); }
// 1 when a 3x3 conv with `ktot16` K-blocks (9 taps * sum(C)/16) and `cout` padded output channels can run column-folded
// (cg_conv_args.fold): one GEMM-N chunk (nc == cout <= 32) whose weight slab -- same bytes as the 9-tap image -- and, with
// want_e, the epilogue-operand ring fit next to the minimum input ring.
extern "C" int32_t cg_conv_fold_ok(int32_t ktot16, int32_t cout, int32_t want_e) {
  if (cout <= 0 || cout > 32 || cout % 16 != 0 || ktot16 % 9 != 0) return 0;
  const int room = kSmemMax - kHdrBytes - kMinStages * kStageBytes;
  return ktot16 * cout * 32 + (want_e ? e_bytes_min(cout) : 0) <= room ? 1 : 0;
}

extern "C" int64_t cg_packed_weight_bytes_nc(int32_t ktot16, int32_t cout, int32_t nc) {
  if (nc <= 0 || nc % 16 != 0 || nc > kMaxNc) return -1;
  int nN = (cout + nc - 1) / nc;
  return (int64_t)nN * ktot16 * nc * 32;
}
extern "C" int64_t cg_packed_weight_bytes(int32_t ktot16, int32_t cout) {
  return cg_packed_weight_bytes_nc(ktot16, cout, pick_nc(ktot16, cout, nullptr, 1));
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
  const bool flat = (a->H == 1 && a->W == 1);
  CG_REQUIRE(!(flat && a->ksize == 3), "cg_conv2d: a 3x3 conv on a 1x1 image must be passed as its centre tap (ksize=1)");
  KParams kp;
  kp.a = *a;
  kp.flat = flat ? 1 : 0;
  kp.fold = a->fold ? 1 : 0;
  CG_REQUIRE(!kp.fold || (a->ksize == 3 && !flat && a->nc == a->cout && a->cout <= 32),
             "cg_conv2d: fold needs ksize 3, a spatial image and nc == cout <= 32 (ksize %d, nc %d, cout %d)", a->ksize, a->nc,
             a->cout);
  kp.ntaps = kp.fold ? 3 : a->ksize * a->ksize;  // taps on the K axis
  kp.HW8 = (long long)a->H * a->W * 8;
  const int halo = a->ksize == 3 ? 2 : 0;
  int nchunks = 0, c16 = 0;
  for (int s = 0; s < a->nsrc; ++s) {
    const cg_src& src = a->src[s];
    CG_REQUIRE(src.ptr != nullptr && ((uintptr_t)src.ptr & 15) == 0, "cg_conv2d: src %d null/unaligned", s);
    CG_REQUIRE(src.C > 0 && src.C % 16 == 0 && src.ns % 8 == 0, "cg_conv2d: src %d C=%d ns=%lld", s, src.C,
               (long long)src.ns);
    const int box_c8 = src.C / 8 < 4 ? src.C / 8 : 4;
    const int c8_phys = src.c8 > 0 ? src.c8 : src.C / 8;  // octets stored; the box may run past them (zero fill)
    CG_REQUIRE(c8_phys <= src.C / 8 && c8_phys * 8 + 8 > src.C - 8, "cg_conv2d: src %d c8=%d vs C=%d", s, c8_phys, src.C);
    int rc = kp.fold ? cg_make_planar_map(&kp.src_map[s], src.ptr, src.ns, a->N, a->H, a->W, c8_phys, 0, 16 * 8, 10, box_c8)
                     : cg_make_planar_map(&kp.src_map[s], src.ptr, src.ns, a->N, a->H, a->W, c8_phys, kp.flat, (8 + halo) * 8,
                                          16 + halo, box_c8);
    if (rc != CG_OK) return rc;
    kp.src_bytes[s] = (uint32_t)box_c8 * (kp.fold ? kPlaneF : (flat ? kPlane1 : (a->ksize == 3 ? kPlane3 : kPlane1)));
    for (int c0 = 0; c0 < src.C; c0 += 32) {
      CG_REQUIRE(nchunks < kMaxChunks, "cg_conv2d: too many K chunks");
      int n16 = (src.C - c0 >= 32) ? 2 : 1;
      kp.chunk[nchunks++] = Chunk{(uint16_t)s, (uint16_t)(c0 / 8), (uint16_t)n16, (uint16_t)(c16 * kp.ntaps)};
      c16 += n16;
    }
  }
  kp.nchunks = nchunks;
  kp.ktot16 = c16 * kp.ntaps;
  if (a->nc > 0) {  // chunk the caller packed the weights for (cg_conv_nchunk_ex)
    CG_REQUIRE(a->nc % 16 == 0 && a->nc <= kMaxNc, "cg_conv2d: nc=%d", a->nc);
    kp.Nc = a->nc;
    const int ng = kp.fold ? 3 * kp.Nc : kp.Nc;
    const int room = kSmemMax - kHdrBytes - kMinStages * kStageBytes;
    kp.emode = (kp.ktot16 * ng * 32 + e_bytes_min(kp.Nc) <= room) ? 1 : 0;
    CG_REQUIRE(kp.ktot16 * ng * 32 <= room, "cg_conv2d: weight slab of K=%d x N=%d does not fit", kp.ktot16 * 16, ng);
  } else {
    kp.Nc = pick_nc(kp.ktot16, a->cout, &kp.emode, 1);
  }
  CG_REQUIRE(kp.Nc >= 16, "cg_conv2d: K=%d too large for a resident weight slab", kp.ktot16 * 16);
  kp.nN = (a->cout + kp.Nc - 1) / kp.Nc;
  kp.Ng = kp.fold ? 3 * kp.Nc : kp.Nc;
  kp.slab_bytes = (uint32_t)kp.ktot16 * kp.Ng * 32u;
  for (int s = 0; s < a->nseg; ++s) {
    const cg_seg& sg = a->seg[s];
    CG_REQUIRE(sg.ptr != nullptr && ((uintptr_t)sg.ptr & 15) == 0, "cg_conv2d: seg %d null/unaligned", s);
    CG_REQUIRE(sg.c0 % 16 == 0 && sg.cn % 8 == 0 && sg.cn > 0 && sg.ns % (sg.dtype == CG_F32 ? 4 : 8) == 0,
               "cg_conv2d: seg %d c0=%d cn=%d ns=%lld", s, sg.c0, sg.cn, (long long)sg.ns);
    CG_REQUIRE(sg.add == nullptr || (((uintptr_t)sg.add & 15) == 0 && sg.add_ns % 8 == 0), "cg_conv2d: seg %d add", s);
    CG_REQUIRE(sg.add2 == nullptr || (((uintptr_t)sg.add2 & 15) == 0 && sg.add2_ns % 8 == 0), "cg_conv2d: seg %d add2", s);
    CG_REQUIRE(sg.mul == nullptr || (((uintptr_t)sg.mul & 15) == 0 && sg.mul_ns % 8 == 0), "cg_conv2d: seg %d mul", s);
    CG_REQUIRE(sg.act_copy == nullptr || (sg.dtype == CG_BF16 && ((uintptr_t)sg.act_copy & 15) == 0 && sg.act_copy_ns % 8 == 0),
               "cg_conv2d: seg %d act_copy needs a bf16 segment and 16-byte alignment", s);
    for (int t = 0; t < s; ++t)
      CG_REQUIRE(sg.c0 >= a->seg[t].c0 + a->seg[t].cn || a->seg[t].c0 >= sg.c0 + sg.cn,
                 "cg_conv2d: segments %d and %d overlap (output channel ranges must be disjoint)", t, s);
  }
  kp.nE = 0;
  if (kp.emode) {
    for (int s = 0; s < a->nseg && kp.nE < kESlots; ++s) {
      const cg_seg& sg = a->seg[s];
      const void* ptrs[3] = {sg.add, sg.add2, sg.mul};
      const long long nss[3] = {sg.add_ns, sg.add2_ns, sg.mul_ns};
      for (int kind = 0; kind < 3 && kp.nE < kESlots; ++kind) {
        if (ptrs[kind] == nullptr) continue;
        int rc = cg_make_planar_map(&kp.e_map[kp.nE], ptrs[kind], nss[kind], a->N, a->H, a->W, sg.cn / 8, kp.flat,
                                    kp.fold ? 128 : 64, kp.fold ? 8 : 16, kp.Nc / 8);
        if (rc != CG_OK) return rc;
        kp.eop[kp.nE++] = EOp{s, kind, sg.c0 / 8};
      }
    }
  }
  kp.e_tx_bytes = (uint32_t)(kp.nE * e_slot_bytes(kp.Nc));
  if (flat) {
    kp.tiles_x = 1;
    kp.tiles_per_img = 1;
    kp.ntiles = (a->N + 127) / 128;
  } else {
    kp.tiles_x = kp.fold ? (a->W + kFoldW - 1) / kFoldW : (a->W + 7) / 8;
    kp.tiles_per_img = kp.tiles_x * (kp.fold ? (a->H + 7) / 8 : (a->H + 15) / 16);
    kp.ntiles = a->N * kp.tiles_per_img;
  }
  kp.idesc = umma_idesc_bf16(128, kp.Ng, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * kp.Ng) cols <<= 1;
  kp.tmem_cols = cols;
  // ring depths: all shared memory left after the weight slab is prefetch distance.  E stages cover as many
  // tiles ahead as the A ring does (bytes per tile: nchunks A stages + nE staged operands).
  kp.stage_bytes = 0;
  for (int s = 0; s < a->nsrc; ++s)
    if ((int)kp.src_bytes[s] > kp.stage_bytes) kp.stage_bytes = (int)kp.src_bytes[s];
  {
    const int budget = kSmemMax - kHdrBytes - (int)kp.slab_bytes;
    const int e_tile = kp.nE * e_slot_bytes(kp.Nc), a_tile = kp.nchunks * kp.stage_bytes;
    kp.nest = 0;
    if (e_tile > 0) {
      int d = budget / (a_tile + e_tile);
      kp.nest = d < kMinEStages ? kMinEStages : (d > kMaxEStages ? kMaxEStages : d);
    }
    kp.nst = (budget - kp.nest * e_tile) / kp.stage_bytes;
    if (kp.nst > kMaxStages) kp.nst = kMaxStages;
    CG_REQUIRE(kp.nst >= 2, "cg_conv2d: no shared memory left for the input ring (K=%d)", kp.ktot16 * 16);
  }
  kp.nst0 = kp.nst;  // one ring unless the grid below gives a CTA more than one tile
  const int smem_bytes = kHdrBytes + kp.nst * kp.stage_bytes + kp.nest * kp.nE * e_slot_bytes(kp.Nc) + (int)kp.slab_bytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
    if (e != cudaSuccess) {
      cg_set_error("cg_conv2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CG_ERR_CUDA;
    }
    attr_done = true;
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("CG_DEBUG_SKIP");
    dbg = e ? atoi(e) : 0;
  }
  kp.dbg = dbg;
  kp.tl = cg_tl_ptr;
  const int sms = cg_device_sms();
  // CTAs along the pixel axis.  Small problems keep >= min_tiles tiles per CTA: the prologue (barriers, TMEM, weight
  // slab) is paid once per CTA, and the SMs left free run the kernels of the other lanes / weight-gradient streams.
  static int min_tiles = -1;
  if (min_tiles < 0) {
    const char* e = getenv("CG_CONV_MIN_TILES");
    min_tiles = e ? atoi(e) : 2;
    if (min_tiles < 1) min_tiles = 1;
  }
  int gx = sms / kp.nN;
  if (gx < 1) gx = 1;
  if (gx > (kp.ntiles + min_tiles - 1) / min_tiles) gx = (kp.ntiles + min_tiles - 1) / min_tiles;
  static int pdl = -1;
  if (pdl < 0) {
    const char* e = getenv("CG_NO_PDL");
    pdl = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  // two MMA issuers (two A rings) as soon as a CTA walks more than one tile; each ring keeps at least 2 stages
  {
    static int one_issuer = -1;
    if (one_issuer < 0) {
      const char* e = getenv("CG_ONE_ISSUER");  // profiling experiment: the round-1 single-issuer pipeline
      one_issuer = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    if (!one_issuer && kp.ntiles > gx && kp.nst >= 4) kp.nst0 = (kp.nst + 1) / 2;
  }
  // two rings = two producers: each owns the operand-ring stages of its tile parity, so that ring's depth must be even
  // (the launch keeps the shared-memory size computed above; an odd depth just leaves its last stage unused)
  if (kp.nst0 < kp.nst && kp.nest > 0 && (kp.nest & 1)) kp.nest -= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, kp.nN);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = cg_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  bool dual = false;
  for (int s = 0; s < a->nseg; ++s) dual = dual || a->seg[s].act_copy != nullptr;
  cudaError_t le;
  if (kp.fold) le = dual ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true>, kp) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true>, kp);
  else le = dual ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false>, kp) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false>, kp);
  if (le != cudaSuccess) {
    cg_set_error("cg_conv2d: launch failed: %s", cudaGetErrorString(le));
    return CG_ERR_CUDA;
  }
  CG_LAUNCH_CHECK("cg_conv2d");
  return CG_OK;
}
