// 7x7 stem convolution (reference src/vae.py:104-110,126): fp32 NCHW image (1 or 3 channels) ->
// bf16 NHWC features, and its weight/bias gradient.  Cin is 1 or 3, far too thin for the tensor
// cores (K = 49..147 with no reuse across Cin), so this is a direct shared-memory-tiled kernel.
#include "cg_common.cuh"

namespace {

constexpr int kT = 16;           // output tile edge
constexpr int kHalo = kT + 6;    // input tile edge
constexpr int kMaxCout = 32;

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, bf16* __restrict__ y, int R,
                                                       long long y_ns) {
  constexpr int Cout = COUT;  // compile-time: the accumulators must stay in registers
  extern __shared__ float sm[];
  float* s_w = sm;                              // [tap][ci][co]
  float* s_x = sm + 49 * CIN * Cout;            // [ci][kHalo][kHalo]
  const int n = blockIdx.z, h0 = blockIdx.y * kT, w0 = blockIdx.x * kT;
  for (int i = threadIdx.x; i < 49 * CIN * Cout; i += 256) {
    int co = i % Cout, r = i / Cout, ci = r % CIN, tap = r / CIN;
    s_w[i] = w[(co * CIN + ci) * 49 + tap];
  }
  for (int i = threadIdx.x; i < CIN * kHalo * kHalo; i += 256) {
    int cc = i % kHalo, rr = (i / kHalo) % kHalo, ci = i / (kHalo * kHalo);
    int hh = h0 - 3 + rr, ww = w0 - 3 + cc;
    s_x[i] = (hh >= 0 && hh < R && ww >= 0 && ww < R) ? x[((long long)(n * CIN + ci) * R + hh) * R + ww] : 0.f;
  }
  __syncthreads();
  const int ty = threadIdx.x / kT, tx = threadIdx.x % kT;
  float acc[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) acc[co] = b[co];
  for (int kh = 0; kh < 7; ++kh)
    for (int kw = 0; kw < 7; ++kw)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float v = s_x[(ci * kHalo + ty + kh) * kHalo + tx + kw];
        const float* wp = s_w + ((kh * 7 + kw) * CIN + ci) * Cout;
#pragma unroll
        for (int co = 0; co < COUT; co += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + co);
          acc[co] += v * w4.x; acc[co + 1] += v * w4.y; acc[co + 2] += v * w4.z; acc[co + 3] += v * w4.w;
        }
      }
  const int h = h0 + ty, wq = w0 + tx;
  if (h < R && wq < R) {
    bf16* o = y + n * y_ns + ((long long)h * R + wq) * 8;  // planar: one 16-byte octet per plane
#pragma unroll
    for (int c8 = 0; c8 < COUT; c8 += 8) *reinterpret_cast<uint4*>(o + (long long)(c8 >> 3) * R * R * 8) = cg_pack8(acc + c8);
  }
}

// dW[co][ci][kh][kw] += sum_pixels dy[p][co] * x[p + (kh-3, kw-3)][ci];  db[co] += sum_pixels dy[p][co].
// Persistent blocks walk 16x16 pixel tiles.  A thread owns a register block of 2 output channels x 4
// consecutive kw taps of one (ci, kh) row and slides a 4-wide window of x along the pixel row, so every pixel
// costs one 8-byte dy load + one x load for 8 FMAs.  Partial sums stay in registers over all tiles of the block;
// one atomicAdd per output per block at the end.
constexpr int kDyPitch = 36;  // floats per pixel row of the dy tile in shared memory (16-byte aligned, conflict-free)

template <int CIN>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ db, int N, int R,
                                                         int Cout, long long dy_ns) {
  __shared__ __align__(16) float s_dy[kT * kT * kDyPitch];
  __shared__ float s_x[CIN * kHalo * kHalo];
  constexpr int kTapBlocks = CIN * 7 * 2;               // (ci, kh, kw half) blocks of 4 kw taps (kw = 7 is padding)
  constexpr int kTasks = (16 * kTapBlocks + 255) / 256;  // (co pair, tap block) tasks per thread
  const int ntask = (Cout / 2) * kTapBlocks;
  float acc[kTasks][2][4];
#pragma unroll
  for (int t = 0; t < kTasks; ++t)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  float bsum = 0.f;  // warp 7: lane = output channel
  const int tiles_1d = (R + kT - 1) / kT;
  const int ntiles = N * tiles_1d * tiles_1d;
  const int C8 = Cout / 8;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_1d * tiles_1d), tr = tile % (tiles_1d * tiles_1d);
    const int h0 = (tr / tiles_1d) * kT, w0 = (tr % tiles_1d) * kT;
    __syncthreads();
    for (int i = threadIdx.x; i < kT * kT * C8; i += 256) {  // one 16-byte octet (8 channels of a pixel) per load
      const int p = i % (kT * kT), c8 = i / (kT * kT);
      const int hh = h0 + p / kT, ww = w0 + p % kT;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (hh < R && ww < R)
        cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + n * dy_ns + ((long long)c8 * R * R + (long long)hh * R + ww) * 8)), f);
      float4* d = reinterpret_cast<float4*>(s_dy + p * kDyPitch + c8 * 8);
      d[0] = make_float4(f[0], f[1], f[2], f[3]);
      d[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = threadIdx.x; i < CIN * kHalo * kHalo; i += 256) {
      int cc = i % kHalo, rr = (i / kHalo) % kHalo, ci = i / (kHalo * kHalo);
      int hh = h0 - 3 + rr, ww = w0 - 3 + cc;
      s_x[i] = (hh >= 0 && hh < R && ww >= 0 && ww < R) ? x[((long long)(n * CIN + ci) * R + hh) * R + ww] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < kTasks; ++t) {
      const int task = threadIdx.x + 256 * t;
      if (task >= ntask) break;
      const int co = (task % (Cout / 2)) * 2, tb = task / (Cout / 2);
      const int kw0 = (tb & 1) * 4, kh = (tb >> 1) % 7, ci = tb / 14;
      // kw0 + 3 = 7 reads one column past the 7-tap window (still inside the 22-wide halo row); its sum is dropped
      const float* xrow = s_x + (ci * kHalo + kh) * kHalo + kw0;
      const float* dyp = s_dy + co;
      for (int py = 0; py < kT; ++py) {
        const float* xr = xrow + py * kHalo;
        float w0v = xr[0], w1v = xr[1], w2v = xr[2], w3v = xr[3];
#pragma unroll
        for (int px = 0; px < kT; ++px) {
          const float2 d = *reinterpret_cast<const float2*>(dyp + (py * kT + px) * kDyPitch);
          acc[t][0][0] += d.x * w0v; acc[t][0][1] += d.x * w1v; acc[t][0][2] += d.x * w2v; acc[t][0][3] += d.x * w3v;
          acc[t][1][0] += d.y * w0v; acc[t][1][1] += d.y * w1v; acc[t][1][2] += d.y * w2v; acc[t][1][3] += d.y * w3v;
          w0v = w1v; w1v = w2v; w2v = w3v;
          if (px + 1 < kT) w3v = (kw0 + px + 4 < kHalo) ? xr[px + 4] : 0.f;
        }
      }
    }
    if (threadIdx.x >= 224 && (int)threadIdx.x - 224 < Cout) {
      const int co = threadIdx.x - 224;
      float a = 0.f;
      for (int p = 0; p < kT * kT; ++p) a += s_dy[p * kDyPitch + co];
      bsum += a;
    }
  }
#pragma unroll
  for (int t = 0; t < kTasks; ++t) {
    const int task = threadIdx.x + 256 * t;
    if (task >= ntask) break;
    const int co = (task % (Cout / 2)) * 2, tb = task / (Cout / 2);
    const int kw0 = (tb & 1) * 4, kh = (tb >> 1) % 7, ci = tb / 14;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (kw0 + j < 7) atomicAdd(dw + ((co + i) * CIN + ci) * 49 + kh * 7 + kw0 + j, acc[t][i][j]);
  }
  if (db != nullptr && threadIdx.x >= 224 && (int)threadIdx.x - 224 < Cout) atomicAdd(db + threadIdx.x - 224, bsum);
}

}  // namespace

extern "C" int cg_stem_fwd(const float* x, const float* w, const float* b, void* y, int32_t N, int32_t Cin, int32_t R,
                           int32_t Cout, int64_t y_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE((Cin == 1 || Cin == 3) && (Cout == 16 || Cout == 32) && y_ld % 8 == 0,
             "cg_stem_fwd: Cin=%d Cout=%d (supported: Cin 1|3, Cout 16|32 = widths[0] of every preset)", Cin, Cout);
  dim3 grid(cg_ceil_div(R, kT), cg_ceil_div(R, kT), N);
  size_t smem = (size_t)(49 * Cin * Cout + Cin * kHalo * kHalo) * sizeof(float);
  bf16* yb = reinterpret_cast<bf16*>(y);
  cudaStream_t st = cg_stream(stream);
  if (Cin == 1 && Cout == 32) stem_fwd_kernel<1, 32><<<grid, 256, smem, st>>>(x, w, b, yb, R, y_ld);
  else if (Cin == 1) stem_fwd_kernel<1, 16><<<grid, 256, smem, st>>>(x, w, b, yb, R, y_ld);
  else if (Cout == 32) stem_fwd_kernel<3, 32><<<grid, 256, smem, st>>>(x, w, b, yb, R, y_ld);
  else stem_fwd_kernel<3, 16><<<grid, 256, smem, st>>>(x, w, b, yb, R, y_ld);
  CG_LAUNCH_CHECK("cg_stem_fwd");
  return CG_OK;
}

extern "C" int cg_stem_wgrad(const float* x, const void* dy, float* dw, float* db, int32_t N, int32_t Cin, int32_t R,
                             int32_t Cout, int64_t dy_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE((Cin == 1 || Cin == 3) && Cout <= kMaxCout && Cout % 8 == 0, "cg_stem_wgrad: Cin=%d Cout=%d", Cin, Cout);
  const int t1 = cg_ceil_div(R, kT);
  int blocks = N * t1 * t1;
  const int cap = 4 * cg_device_sms();
  if (blocks > cap) blocks = cap;
  if (Cin == 1)
    stem_wgrad_kernel<1><<<blocks, 256, 0, cg_stream(stream)>>>(x, reinterpret_cast<const bf16*>(dy), dw, db, N, R, Cout,
                                                                dy_ld);
  else
    stem_wgrad_kernel<3><<<blocks, 256, 0, cg_stream(stream)>>>(x, reinterpret_cast<const bf16*>(dy), dw, db, N, R, Cout,
                                                                dy_ld);
  CG_LAUNCH_CHECK("cg_stem_wgrad");
  return CG_OK;
}
