// 7x7 stem convolution (reference src/vae.py:104-110,126): fp32 NCHW image (1 or 3 channels) ->
// bf16 NHWC features, and its weight/bias gradient.  Cin is 1 or 3, far too thin for the tensor
// cores (K = 49..147 with no reuse across Cin), so this is a direct shared-memory-tiled kernel.
#include "cg_common.cuh"

namespace {

constexpr int kT = 16;           // output tile edge
constexpr int kHalo = kT + 6;    // input tile edge
constexpr int kMaxCout = 32;

template <int CIN>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, bf16* __restrict__ y, int R,
                                                       int Cout, long long y_ns) {
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
  float acc[kMaxCout];
#pragma unroll
  for (int co = 0; co < kMaxCout; ++co) acc[co] = co < Cout ? b[co] : 0.f;
  for (int kh = 0; kh < 7; ++kh)
    for (int kw = 0; kw < 7; ++kw)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float v = s_x[(ci * kHalo + ty + kh) * kHalo + tx + kw];
        const float* wp = s_w + ((kh * 7 + kw) * CIN + ci) * Cout;
#pragma unroll
        for (int co = 0; co < kMaxCout; ++co)
          if (co < Cout) acc[co] += v * wp[co];
      }
  const int h = h0 + ty, wq = w0 + tx;
  if (h < R && wq < R) {
    bf16* o = y + n * y_ns + ((long long)h * R + wq) * 8;  // planar: one 16-byte octet per plane
    for (int c8 = 0; c8 < Cout; c8 += 8) *reinterpret_cast<uint4*>(o + (long long)(c8 >> 3) * R * R * 8) = cg_pack8(acc + c8);
  }
}

// dW[co][ci][tap] += sum_pixels dy[p][co] * x[p+tap][ci];  db[co] += sum dy.  Persistent blocks walk
// tiles and keep their partial sums in registers, one atomicAdd per output per block at the end.
template <int CIN>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ db, int N, int R,
                                                         int Cout, long long dy_ns) {
  __shared__ float s_dy[kT * kT][kMaxCout + 1];
  __shared__ float s_x[CIN * kHalo * kHalo];
  const int nout = Cout * (CIN * 49 + 1);  // + bias column
  constexpr int kPerThread = (kMaxCout * (CIN * 49 + 1) + 255) / 256;
  float acc[kPerThread];
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) acc[j] = 0.f;
  const int tiles_1d = (R + kT - 1) / kT;
  const int ntiles = N * tiles_1d * tiles_1d;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_1d * tiles_1d), tr = tile % (tiles_1d * tiles_1d);
    const int h0 = (tr / tiles_1d) * kT, w0 = (tr % tiles_1d) * kT;
    __syncthreads();
    for (int i = threadIdx.x; i < kT * kT * Cout; i += 256) {
      int co = i % Cout, p = i / Cout, hh = h0 + p / kT, ww = w0 + p % kT;
      s_dy[p][co] = (hh < R && ww < R)
                        ? __bfloat162float(dy[n * dy_ns + ((long long)(co >> 3) * R * R + (long long)hh * R + ww) * 8 + (co & 7)])
                        : 0.f;
    }
    for (int i = threadIdx.x; i < CIN * kHalo * kHalo; i += 256) {
      int cc = i % kHalo, rr = (i / kHalo) % kHalo, ci = i / (kHalo * kHalo);
      int hh = h0 - 3 + rr, ww = w0 - 3 + cc;
      s_x[i] = (hh >= 0 && hh < R && ww >= 0 && ww < R) ? x[((long long)(n * CIN + ci) * R + hh) * R + ww] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int o = threadIdx.x + 256 * j;
      if (o >= nout) break;
      const int co = o % Cout, r = o / Cout;  // r in [0, CIN*49] ; last = bias
      float a = 0.f;
      if (r == CIN * 49) {
        for (int p = 0; p < kT * kT; ++p) a += s_dy[p][co];
      } else {
        const int ci = r / 49, tap = r % 49, kh = tap / 7, kw = tap % 7;
        const float* xs = s_x + (ci * kHalo + kh) * kHalo + kw;
        for (int py = 0; py < kT; ++py)
#pragma unroll
          for (int px = 0; px < kT; ++px) a += s_dy[py * kT + px][co] * xs[py * kHalo + px];
      }
      acc[j] += a;
    }
  }
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int o = threadIdx.x + 256 * j;
    if (o >= nout) break;
    const int co = o % Cout, r = o / Cout;
    if (r == CIN * 49) {
      if (db != nullptr) atomicAdd(db + co, acc[j]);
    } else {
      atomicAdd(dw + (co * CIN + r / 49) * 49 + r % 49, acc[j]);
    }
  }
}

}  // namespace

extern "C" int cg_stem_fwd(const float* x, const float* w, const float* b, void* y, int32_t N, int32_t Cin, int32_t R,
                           int32_t Cout, int64_t y_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE((Cin == 1 || Cin == 3) && Cout % 8 == 0 && Cout <= kMaxCout && y_ld % 8 == 0,
             "cg_stem_fwd: Cin=%d Cout=%d", Cin, Cout);
  dim3 grid(cg_ceil_div(R, kT), cg_ceil_div(R, kT), N);
  size_t smem = (size_t)(49 * Cin * Cout + Cin * kHalo * kHalo) * sizeof(float);
  if (Cin == 1)
    stem_fwd_kernel<1><<<grid, 256, smem, cg_stream(stream)>>>(x, w, b, reinterpret_cast<bf16*>(y), R, Cout, y_ld);
  else
    stem_fwd_kernel<3><<<grid, 256, smem, cg_stream(stream)>>>(x, w, b, reinterpret_cast<bf16*>(y), R, Cout, y_ld);
  CG_LAUNCH_CHECK("cg_stem_fwd");
  return CG_OK;
}

extern "C" int cg_stem_wgrad(const float* x, const void* dy, float* dw, float* db, int32_t N, int32_t Cin, int32_t R,
                             int32_t Cout, int64_t dy_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE((Cin == 1 || Cin == 3) && Cout <= kMaxCout, "cg_stem_wgrad: Cin=%d Cout=%d", Cin, Cout);
  const int t1 = cg_ceil_div(R, kT);
  int blocks = N * t1 * t1;
  const int cap = 2 * cg_device_sms();
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
