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
  // column-folded image (cg_pack_desc.fold): K-blocks are (channel block, kernel row), the GEMM-N axis carries the three
  // kernel columns side by side: row kx*d.nc + n of chunk y = output channel y*d.nc + n, kernel column kx
  const int fold = d.fold;
  const int taps_k = fold ? 3 : d.taps;      // taps on the K axis
  const int ktot16 = c16tot * taps_k;
  const int nc = fold ? 3 * d.nc : d.nc;     // GEMM-N rows per chunk
  const int nN = (d.n_pad + d.nc - 1) / d.nc;
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
    const int nrow = n8 * 8 + ni;
    const int kx = fold ? nrow / d.nc : 0;
    const int n = y * d.nc + (fold ? nrow - kx * d.nc : nrow);
    int c16g = kidx / taps_k;
    const int t = kidx - c16g * taps_k;
    int s = 0;
    while (s < d.nsrc - 1 && c16g >= d.src_c[s] / 16) {
      c16g -= d.src_c[s] / 16;
      ++s;
    }
    int kh = 0, kw = 0;
    if (d.k == 3) {
      if (fold) { kh = t; kw = kx; }
      else if (d.taps == 9) { kh = t / 3; kw = t % 3; } else { kh = 1; kw = 1; }
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
        if (d.n_scale != nullptr && !d.transpose) v *= d.n_scale[o];  // folded eval-mode BatchNorm
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

// torchvision RandomCrop(size=R, padding=(pad_left, pad_top), fill=0) + RandomHorizontalFlip on uint8 NCHW batches, one
// (top, left, flip) triple per sample (src/datasets.py:107-118 UKBB, :281-286 Morpho-MNIST): out[y, x] = in[y + top -
// pad_top, x' + left - pad_left] with x' = R-1-x when flipped, 0 outside the source image.
__global__ void augment_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ prm,
                                  int C, int Hi, int Wi, int R, int pad_top, int pad_left) {
  const int n = blockIdx.z, c = blockIdx.y;
  const int top = prm[n * 3 + 0], left = prm[n * 3 + 1], flip = prm[n * 3 + 2];
  const uint8_t* src = in + ((long long)n * C + c) * Hi * Wi;
  uint8_t* dst = out + ((long long)n * C + c) * R * R;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < R * R; p += gridDim.x * blockDim.x) {
    const int y = p / R, x = p - y * R;
    const int xs = (flip ? R - 1 - x : x) + left - pad_left, ys = y + top - pad_top;
    dst[p] = (xs >= 0 && xs < Wi && ys >= 0 && ys < Hi) ? src[ys * Wi + xs] : (uint8_t)0;
  }
}

// spatially constant parents -> bf16 planar (N, C/8, HW, 8)
__global__ void parents_plane_kernel(const float* __restrict__ pa, long long sstride, long long cstride,
                                     bf16* __restrict__ out, int N, int ctx, int C8, int HW, long long ns, int drop_from,
                                     float drop_scale, const float* __restrict__ drop_scale_dev) {
  if (drop_scale_dev != nullptr) drop_scale = __ldg(drop_scale_dev);  // device scalar: replayable from a CUDA graph
  const long long total = (long long)N * C8 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW), c8 = (int)((i / HW) % C8), n = (int)(i / ((long long)HW * C8));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      float v = 0.f;
      if (c < ctx) {
        v = __ldg(pa + n * sstride + c * cstride);
        if (c >= drop_from) v *= drop_scale;
      }
      f[k] = v;
    }
    *reinterpret_cast<uint4*>(out + n * ns + ((long long)c8 * HW + hw) * 8) = cg_pack8(f);
  }
}

// fp32 NCHW <-> bf16 planar: per (n, octet, pixel) gather/scatter 8 channels; reads/writes coalesced along pixels
__global__ void nchw_to_planar_kernel(const float* __restrict__ x, bf16* __restrict__ y, int N, int C, int HW, long long ns) {
  const int C8 = (C + 7) / 8;
  const long long total = (long long)N * C8 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW), c8 = (int)((i / HW) % C8), n = (int)(i / ((long long)HW * C8));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      f[k] = c < C ? x[((long long)n * C + c) * HW + hw] : 0.f;
    }
    *reinterpret_cast<uint4*>(y + n * ns + ((long long)c8 * HW + hw) * 8) = cg_pack8(f);
  }
}
__global__ void planar_to_nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y, int N, int C, int HW, long long ns) {
  const int C8 = (C + 7) / 8;
  const long long total = (long long)N * C8 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW), c8 = (int)((i / HW) % C8), n = (int)(i / ((long long)HW * C8));
    float f[8];
    cg_unpack8(*reinterpret_cast<const uint4*>(x + n * ns + ((long long)c8 * HW + hw) * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      if (c < C) y[((long long)n * C + c) * HW + hw] = f[k];
    }
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

__global__ void fill_planar_kernel(const float* __restrict__ v, bf16* __restrict__ y, int N, int HW, int C, long long ns) {
  const int C8 = (C + 7) / 8;
  const long long total = (long long)N * C8 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(i % HW), c8 = (int)((i / HW) % C8), n = (int)(i / ((long long)HW * C8));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (c8 * 8 + k < C) ? v[c8 * 8 + k] : 0.f;
    *reinterpret_cast<uint4*>(y + n * ns + ((long long)c8 * HW + hw) * 8) = cg_pack8(f);
  }
}

// dv[c] += sum over (n, pixel) of dy[n, c, pixel].  grid.y = channel octet; each block strides over the
// (sample, pixel) positions of its octet plane with 16-byte loads, reduces through shared memory and issues
// 8 atomics.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ dv, int N, int HW,
                                                     int C, long long ns) {
  const int c8 = blockIdx.y;
  const long long total = (long long)N * HW;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), hw = (int)(i - (long long)n * HW);
    float f[8];
    cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + n * ns + ((long long)c8 * HW + hw) * 8)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  __shared__ float red[8][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = cg_warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8 && c8 * 8 + threadIdx.x < C) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(dv + c8 * 8 + threadIdx.x, s);
  }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, int N, int HW,
                           int C8, long long a_ns, long long b_ns, long long y_ns) {
  const long long per = (long long)C8 * HW;
  const long long total = (long long)N * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    const long long r = (i - (long long)n * per) * 8;
    float fa[8], fb[8];
    cg_unpack8(*reinterpret_cast<const uint4*>(a + n * a_ns + r), fa);
    cg_unpack8(*reinterpret_cast<const uint4*>(b + n * b_ns + r), fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fb[k];
    *reinterpret_cast<uint4*>(y + n * y_ns + r) = cg_pack8(fa);
  }
}

__global__ void elbo_finalize_kernel(const float* __restrict__ nll, const float* __restrict__ kl, float* __restrict__ out,
                                     int N, int nblk, float kl_scale, float beta, const float* __restrict__ beta_dev) {
  if (beta_dev != nullptr) beta = __ldg(beta_dev);  // device scalar: beta annealing under graph replay
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

static inline int glue_grid(long long work) {
  long long g = (work + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  return g < 1 ? 1 : (int)g;
}

extern "C" int cg_augment_u8(const uint8_t* in, uint8_t* out, const int32_t* params, int32_t N, int32_t C, int32_t Hi,
                             int32_t Wi, int32_t R, int32_t pad_top, int32_t pad_left, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(in != nullptr && out != nullptr && params != nullptr && in != out, "cg_augment_u8: null / aliased buffers");
  CG_REQUIRE(N > 0 && C > 0 && Hi > 0 && Wi > 0 && R > 0 && pad_top >= 0 && pad_left >= 0, "cg_augment_u8: bad shape");
  CG_REQUIRE(Hi + 2 * pad_top >= R && Wi + 2 * pad_left >= R, "cg_augment_u8: padded image smaller than the crop");
  dim3 grid(cg_ceil_div((int64_t)R * R, 256 * 4), C, N);
  augment_u8_kernel<<<grid, 256, 0, cg_stream(stream)>>>(in, out, params, C, Hi, Wi, R, pad_top, pad_left);
  CG_LAUNCH_CHECK("cg_augment_u8");
  return CG_OK;
}

extern "C" int cg_parents_plane(const float* pa, int64_t sample_stride, int64_t chan_stride, void* out, int32_t N,
                                int32_t ctx, int32_t C, int32_t HW, int64_t ns, int32_t drop_from, float drop_scale,
                                const float* drop_scale_dev, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(C % 8 == 0 && C >= ctx && ns % 8 == 0, "cg_parents_plane: C=%d ctx=%d", C, ctx);
  parents_plane_kernel<<<glue_grid((long long)N * (C / 8) * HW), 256, 0, cg_stream(stream)>>>(
      pa, sample_stride, chan_stride, reinterpret_cast<bf16*>(out), N, ctx, C / 8, HW, ns, drop_from, drop_scale, drop_scale_dev);
  CG_LAUNCH_CHECK("cg_parents_plane");
  return CG_OK;
}

extern "C" int cg_nchw_f32_to_planar(const float* x, void* y, int32_t N, int32_t C, int32_t HW, int64_t ns, void* stream) {
  CG_ARCH_GUARD();
  nchw_to_planar_kernel<<<glue_grid((long long)N * ((C + 7) / 8) * HW), 256, 0, cg_stream(stream)>>>(
      x, reinterpret_cast<bf16*>(y), N, C, HW, ns);
  CG_LAUNCH_CHECK("cg_nchw_f32_to_planar");
  return CG_OK;
}

extern "C" int cg_planar_to_nchw_f32(const void* x, float* y, int32_t N, int32_t C, int32_t HW, int64_t ns, void* stream) {
  CG_ARCH_GUARD();
  planar_to_nchw_kernel<<<glue_grid((long long)N * ((C + 7) / 8) * HW), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(x), y, N, C, HW, ns);
  CG_LAUNCH_CHECK("cg_planar_to_nchw_f32");
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

extern "C" int cg_fill_planar(const float* v, void* y, int32_t N, int32_t HW, int32_t C, int64_t ns, void* stream) {
  CG_ARCH_GUARD();
  fill_planar_kernel<<<glue_grid((long long)N * ((C + 7) / 8) * HW), 256, 0, cg_stream(stream)>>>(
      v, reinterpret_cast<bf16*>(y), N, HW, C, ns);
  CG_LAUNCH_CHECK("cg_fill_planar");
  return CG_OK;
}

extern "C" int cg_colsum(const void* dy, float* dv, int32_t N, int32_t HW, int32_t C, int64_t ns, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(ns % 8 == 0, "cg_colsum: ns");
  const int C8 = (C + 7) / 8;
  long long gx = ((long long)N * HW + 2047) / 2048;  // ~8 positions per thread
  const long long cap = (148LL * 4 + C8 - 1) / C8;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  colsum_kernel<<<dim3((int)gx, C8), 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(dy), dv, N, HW, C, ns);
  CG_LAUNCH_CHECK("cg_colsum");
  return CG_OK;
}

extern "C" int cg_add(const void* a, const void* b, void* y, int32_t N, int32_t HW, int32_t C, int64_t a_ns, int64_t b_ns,
                      int64_t y_ns, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(C % 8 == 0, "cg_add: C must be a multiple of 8");
  add_kernel<<<glue_grid((long long)N * (C / 8) * HW), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(a), reinterpret_cast<const bf16*>(b), reinterpret_cast<bf16*>(y), N, HW, C / 8, a_ns,
      b_ns, y_ns);
  CG_LAUNCH_CHECK("cg_add");
  return CG_OK;
}

extern "C" int cg_elbo_finalize(const float* nll, const float* kl, float* out, int32_t N, int32_t nblk, float kl_scale,
                                float beta, const float* beta_dev, void* stream) {
  CG_ARCH_GUARD();
  elbo_finalize_kernel<<<1, 32, 0, cg_stream(stream)>>>(nll, kl, out, N, nblk, kl_scale, beta, beta_dev);
  CG_LAUNCH_CHECK("cg_elbo_finalize");
  return CG_OK;
}
