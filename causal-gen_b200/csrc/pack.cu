// Weight packing (fp32 OIHW master -> bf16 UMMA core-matrix image) and layout glue kernels.
#include "cg_common.cuh"

namespace {

// Packed image, per GEMM-N chunk y:  [y][kidx][k8(2)][n8(nc/8)][ni(8)][ki(8)]  (bf16)
// = for every K-block of 16 (kidx = c16*taps + tap) a K-major SWIZZLE_NONE B tile of nc rows:
//   core matrix = 8 rows (n) x 8 k  (128 B), SBO (next 8 rows) = 128 B, LBO (next 8 k) = nc*16 B.
__global__ void pack_weights_kernel(const cg_pack_desc* __restrict__ descs) {
  const cg_pack_desc d = descs[blockIdx.y];
  int c16tot = 0;
  for (int s = 0; s < d.nsrc; ++s) c16tot += d.src_c[s] / 16;
  const int ktot16 = c16tot * d.taps;
  const int nc = d.nc;
  const int nN = (d.n_pad + nc - 1) / nc;
  const long long total = (long long)nN * ktot16 * 2 * nc;  // 16-byte groups
  const int kk = d.k * d.k;
  uint4* out = reinterpret_cast<uint4*>(d.out);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int ni = (int)(g & 7);
    const int n8 = (int)((g >> 3) % (nc / 8));
    const int k8 = (int)((g / nc) & 1);
    const int kidx = (int)((g / (2 * nc)) % ktot16);
    const int y = (int)(g / ((long long)2 * nc * ktot16));
    const int n = y * nc + n8 * 8 + ni;
    int c16g = kidx / d.taps;
    const int t = kidx - c16g * d.taps;
    int s = 0;
    while (s < d.nsrc - 1 && c16g >= d.src_c[s] / 16) {
      c16g -= d.src_c[s] / 16;
      ++s;
    }
    int kh = 0, kw = 0;
    if (d.k == 3) {
      if (d.taps == 9) { kh = t / 3; kw = t % 3; } else { kh = 1; kw = 1; }
      if (d.transpose) { kh = 2 - kh; kw = 2 - kw; }
    }
    float f[8];
#pragma unroll
    for (int ki = 0; ki < 8; ++ki) {
      const int c = c16g * 16 + k8 * 8 + ki;  // channel inside source s
      float v = 0.0f;
      if (n < d.n_log && c < d.src_log[s]) {
        const int kch = d.src_off[s] + c;
        const int nch = d.n_off + n;
        const int o = d.transpose ? kch : nch;
        const int i = d.transpose ? nch : kch;
        v = d.w[((long long)o * d.cin_l + i) * kk + kh * d.k + kw];
      }
      f[ki] = v;
    }
    out[g] = cg_pack8(f);
  }
}

__global__ void normalise_u8_kernel(const uint8_t* __restrict__ x8, float* __restrict__ x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = ((float)x8[i] - 127.5f) / 127.5f;
}

__global__ void parents_pack_kernel(const float* __restrict__ pa, long long sstride, long long cstride,
                                    bf16* __restrict__ out, int N, int ctx, int ld, int drop_from, float drop_scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * ld) return;
  int n = i / ld, c = i - n * ld;
  float v = 0.0f;
  if (c < ctx) {
    v = pa[n * sstride + c * cstride];
    if (c >= drop_from) v *= drop_scale;
  }
  out[i] = __float2bfloat16(v);
}

// NCHW fp32 <-> NHWC bf16 through a 32x32 shared-memory transpose
__global__ void nchw2nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int C, int HW, int ld) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, p = p0 + threadIdx.x;
    t[r][threadIdx.x] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < ld) y[((long long)n * HW + p) * ld + c] = __float2bfloat16(c < C ? t[threadIdx.x][r] : 0.0f);
  }
}
__global__ void nhwc2nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y, int C, int HW, int ld) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, c = c0 + threadIdx.x;
    t[r][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(x[((long long)n * HW + p) * ld + c]) : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = t[threadIdx.x][r];
  }
}

__global__ void stats_to_nchw_kernel(const float* __restrict__ src, int ld, int c0, float add, float* __restrict__ dst,
                                     int C, int HW) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, cb = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int p = p0 + r, c = cb + threadIdx.x;
    t[r][threadIdx.x] = (p < HW && c < C) ? src[((long long)n * HW + p) * ld + c0 + c] + add : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = cb + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((long long)n * C + c) * HW + p] = t[threadIdx.x][r];
  }
}

__global__ void fill_rows_kernel(const float* __restrict__ v, bf16* __restrict__ y, long long rows, int C, int ld) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  int c = (int)(i % ld);
  y[i] = __float2bfloat16(c < C ? v[c] : 0.0f);
}

// dv[c] += sum_rows dy[row][c].  Threads walk (row, channel-octet) pairs in memory order (coalesced 16-byte
// loads whatever C is); partial sums are combined per octet through shared memory, one atomicAdd per
// channel per block.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ dv, long long rows,
                                                     int C, int ld) {
  const int C8 = (C + 7) / 8;
  const int rpb = 256 / C8;              // rows per block pass (threads beyond rpb*C8 idle)
  const int oct = threadIdx.x % C8, rsub = threadIdx.x / C8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rsub < rpb) {
    for (long long r = (long long)blockIdx.x * rpb + rsub; r < rows; r += (long long)gridDim.x * rpb) {
      float f[8];
      cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * ld + oct * 8)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
  __shared__ float red[256][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  for (int t = threadIdx.x; t < C8 * 8; t += 256) {  // C may exceed 256 channels
    const int o = t / 8, i = t % 8;
    float s = 0.f;
    for (int j = 0; j < rpb; ++j) s += red[j * C8 + o][i];
    if (o * 8 + i < C) atomicAdd(dv + o * 8 + i, s);
  }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, long long rows,
                           int C8, int a_ld, int b_ld, int y_ld) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C8) return;
  long long r = i / C8;
  int c = (int)(i - r * C8) * 8;
  float fa[8], fb[8];
  cg_unpack8(*reinterpret_cast<const uint4*>(a + r * a_ld + c), fa);
  cg_unpack8(*reinterpret_cast<const uint4*>(b + r * b_ld + c), fb);
#pragma unroll
  for (int k = 0; k < 8; ++k) fa[k] += fb[k];
  *reinterpret_cast<uint4*>(y + r * y_ld + c) = cg_pack8(fa);
}

__global__ void elbo_finalize_kernel(const float* __restrict__ nll, const float* __restrict__ kl, float* __restrict__ out,
                                     int N, int nblk, float kl_scale, float beta) {
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < N; i += 32) {
    a += nll[i];
    for (int j = 0; j < nblk; ++j) b += kl[(long long)j * N + i];
  }
  a = cg_warp_sum(a);
  b = cg_warp_sum(b);
  if (threadIdx.x == 0) {
    float mn = a / N, mk = b * kl_scale / N;
    out[0] = mn + beta * mk;
    out[1] = mn;
    out[2] = mk;
  }
}

}  // namespace

extern "C" int cg_pack_weights(const cg_pack_desc* descs_dev, int32_t n, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(descs_dev != nullptr && n > 0, "cg_pack_weights: empty descriptor table");
  pack_weights_kernel<<<dim3(8, n), 256, 0, cg_stream(stream)>>>(descs_dev);
  CG_LAUNCH_CHECK("cg_pack_weights");
  return CG_OK;
}

extern "C" int cg_normalise_u8(const uint8_t* x8, float* x, int64_t n, void* stream) {
  CG_ARCH_GUARD();
  if (n <= 0) return CG_OK;
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  normalise_u8_kernel<<<blocks, 256, 0, cg_stream(stream)>>>(x8, x, n);
  CG_LAUNCH_CHECK("cg_normalise_u8");
  return CG_OK;
}

extern "C" int cg_parents_pack(const float* pa, int64_t sample_stride, int64_t chan_stride, void* out, int32_t N,
                               int32_t ctx, int32_t ld, int32_t drop_from, float drop_scale, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(ld % 16 == 0 && ld >= ctx, "cg_parents_pack: ld=%d ctx=%d", ld, ctx);
  parents_pack_kernel<<<cg_ceil_div((int64_t)N * ld, 256), 256, 0, cg_stream(stream)>>>(
      pa, sample_stride, chan_stride, reinterpret_cast<bf16*>(out), N, ctx, ld, drop_from, drop_scale);
  CG_LAUNCH_CHECK("cg_parents_pack");
  return CG_OK;
}

extern "C" int cg_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t N, int32_t C, int32_t HW, int32_t ld,
                                        void* stream) {
  CG_ARCH_GUARD();
  dim3 grid(cg_ceil_div(HW, 32), cg_ceil_div(ld, 32), N);
  nchw2nhwc_kernel<<<grid, dim3(32, 8), 0, cg_stream(stream)>>>(x, reinterpret_cast<bf16*>(y), C, HW, ld);
  CG_LAUNCH_CHECK("cg_nchw_f32_to_nhwc_bf16");
  return CG_OK;
}

extern "C" int cg_nhwc_bf16_to_nchw_f32(const void* x, float* y, int32_t N, int32_t C, int32_t HW, int32_t ld,
                                        void* stream) {
  CG_ARCH_GUARD();
  dim3 grid(cg_ceil_div(HW, 32), cg_ceil_div(C, 32), N);
  nhwc2nchw_kernel<<<grid, dim3(32, 8), 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(x), y, C, HW, ld);
  CG_LAUNCH_CHECK("cg_nhwc_bf16_to_nchw_f32");
  return CG_OK;
}

extern "C" int cg_stats_to_nchw(const float* src, int32_t ld, int32_t c0, float add, float* dst, int32_t N, int32_t C,
                                int32_t HW, void* stream) {
  CG_ARCH_GUARD();
  dim3 grid(cg_ceil_div(HW, 32), cg_ceil_div(C, 32), N);
  stats_to_nchw_kernel<<<grid, dim3(32, 8), 0, cg_stream(stream)>>>(src, ld, c0, add, dst, C, HW);
  CG_LAUNCH_CHECK("cg_stats_to_nchw");
  return CG_OK;
}

extern "C" int cg_fill_rows(const float* v, void* y, int64_t rows, int32_t C, int32_t ld, void* stream) {
  CG_ARCH_GUARD();
  fill_rows_kernel<<<cg_ceil_div(rows * ld, 256), 256, 0, cg_stream(stream)>>>(v, reinterpret_cast<bf16*>(y), rows, C,
                                                                               ld);
  CG_LAUNCH_CHECK("cg_fill_rows");
  return CG_OK;
}

extern "C" int cg_colsum(const void* dy, float* dv, int64_t rows, int32_t C, int32_t ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(ld % 8 == 0, "cg_colsum: ld=%d", ld);
  CG_REQUIRE(C <= 2048, "cg_colsum: C=%d too wide", C);
  const int rpb = 256 / ((C + 7) / 8);
  int gx = (int)((rows + (long long)rpb * 8 - 1) / ((long long)rpb * 8));
  if (gx > 148 * 4) gx = 148 * 4;
  if (gx < 1) gx = 1;
  colsum_kernel<<<gx, 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(dy), dv, rows, C, ld);
  CG_LAUNCH_CHECK("cg_colsum");
  return CG_OK;
}

extern "C" int cg_add(const void* a, const void* b, void* y, int64_t rows, int32_t C, int32_t a_ld, int32_t b_ld,
                      int32_t y_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(C % 8 == 0 && a_ld % 8 == 0 && b_ld % 8 == 0 && y_ld % 8 == 0, "cg_add: C/ld must be multiples of 8");
  add_kernel<<<cg_ceil_div(rows * (C / 8), 256), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), reinterpret_cast<bf16*>(y), rows, C / 8, a_ld,
      b_ld, y_ld);
  CG_LAUNCH_CHECK("cg_add");
  return CG_OK;
}

extern "C" int cg_elbo_finalize(const float* nll, const float* kl, float* out, int32_t N, int32_t nblk, float kl_scale,
                                float beta, void* stream) {
  CG_ARCH_GUARD();
  elbo_finalize_kernel<<<1, 32, 0, cg_stream(stream)>>>(nll, kl, out, N, nblk, kl_scale, beta);
  CG_LAUNCH_CHECK("cg_elbo_finalize");
  return CG_OK;
}
