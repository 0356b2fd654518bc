// Kernels for the anticausal predictors that read the counterfactual image right after the hot path (SURVEY 8 f3;
// reference call site src/pgm/dscm.py:78-83): the BatchNorm `CNN` of src/pgm/layers.py:64-104 and the GroupNorm ResNet-18 of
// src/pgm/resnet.py:9-239, inference (eval mode) only.  Their 3x3 / 1x1 convolutions run on the tcgen05 conv kernel
// (conv_tc.cu; BatchNorm folded into the packed weights + bias, LeakyReLU in the epilogue); this file holds what is left:
// the thin-input stem convolution, max pooling / strided subsampling, GroupNorm (+ residual + ReLU), global average pooling
// and the small fully connected heads.  Activations are bf16 channel-octet planar (N, C/8, H, W, 8) like everywhere else.
#include "cg_common.cuh"

namespace {

__device__ __forceinline__ float act_apply(float v, int act) {
  return act == CG_ACT_LRELU ? (v > 0.f ? v : 0.01f * v) : cg_act(v, act);
}

// eval-mode BatchNorm as a per-channel affine map: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ scale, float* __restrict__ shift,
                               int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] * rsqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = beta[c] - mean[c] * s;
}

// Direct convolution for thin inputs (stem: Cin = 1 or 3, k = 7, stride 1 or 2): fp32 NCHW in, bf16 planar out.
// One thread = one output pixel x one channel octet; the octet's weights sit in shared memory.
// y = act((conv(x, w)) * scale[c] + shift[c])   (scale / shift optional: folded BatchNorm)
constexpr int kDirectMaxW = 8 * 3 * 49;
__global__ void __launch_bounds__(256) conv_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          bf16* __restrict__ y, int Cin, int H, int W, int Cout, int k,
                                                          int stride, int pad, int Ho, int Wo, int act, long long y_ns) {
  __shared__ float s_w[kDirectMaxW];  // [ci][tap][8]
  const int oct = blockIdx.y, n = blockIdx.z;
  const int kk = k * k;
  for (int i = threadIdx.x; i < Cin * kk * 8; i += blockDim.x) {
    const int o = i & 7, r = i >> 3, tap = r % kk, ci = r / kk;
    const int co = oct * 8 + o;
    s_w[i] = co < Cout ? w[((long long)co * Cin + ci) * kk + tap] : 0.f;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo) return;
  const int oh = p / Wo, ow = p - oh * Wo;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int ci = 0; ci < Cin; ++ci) {
    const float* xp = x + ((long long)n * Cin + ci) * H * W;
    for (int kh = 0; kh < k; ++kh) {
      const int ih = oh * stride - pad + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int iw = ow * stride - pad + kw;
        if (iw < 0 || iw >= W) continue;
        const float v = __ldg(xp + (long long)ih * W + iw);
        const float4* wp = reinterpret_cast<const float4*>(s_w + (ci * kk + kh * k + kw) * 8);
        const float4 a = wp[0], b = wp[1];
        acc[0] = fmaf(v, a.x, acc[0]); acc[1] = fmaf(v, a.y, acc[1]); acc[2] = fmaf(v, a.z, acc[2]); acc[3] = fmaf(v, a.w, acc[3]);
        acc[4] = fmaf(v, b.x, acc[4]); acc[5] = fmaf(v, b.y, acc[5]); acc[6] = fmaf(v, b.z, acc[6]); acc[7] = fmaf(v, b.w, acc[7]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int co = oct * 8 + o;
    float v = acc[o];
    if (co < Cout) {
      if (scale != nullptr) v = fmaf(v, scale[co], shift[co]);
      v = act_apply(v, act);
    } else {
      v = 0.f;
    }
    acc[o] = v;
  }
  *reinterpret_cast<uint4*>(y + n * y_ns + ((long long)oct * Ho * Wo + p) * 8) = cg_pack8(acc);
}

// max pooling (k, stride, pad) on planar bf16; k = 1 is plain strided subsampling (the stride-2 convolutions of the
// predictors are evaluated as stride-1 tensor-core convolutions followed by this pick of the even pixels)
__global__ void __launch_bounds__(256) pool_max_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C8, int H, int W,
                                                       int k, int stride, int pad, int Ho, int Wo, long long x_ns,
                                                       long long y_ns) {
  const int n = blockIdx.z, oct = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo) return;
  const int oh = p / Wo, ow = p - oh * Wo;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
  const bf16* xp = x + n * x_ns + (long long)oct * H * W * 8;
  for (int kh = 0; kh < k; ++kh) {
    const int ih = oh * stride - pad + kh;
    if (ih < 0 || ih >= H) continue;
    for (int kw = 0; kw < k; ++kw) {
      const int iw = ow * stride - pad + kw;
      if (iw < 0 || iw >= W) continue;
      float v[8];
      cg_unpack8(*reinterpret_cast<const uint4*>(xp + ((long long)ih * W + iw) * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
    }
  }
  *reinterpret_cast<uint4*>(y + n * y_ns + ((long long)oct * Ho * Wo + p) * 8) = cg_pack8(m);
}

// per-(sample, channel) sum and sum of squares over the pixels: stats[(n*C + c)*2 + {0,1}].  One block per (octet, sample).
__global__ void __launch_bounds__(256) chan_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats, int C, int HW,
                                                         long long x_ns) {
  __shared__ float s_red[8][16];
  const int oct = blockIdx.x, n = blockIdx.y;
  const bf16* xp = x + n * x_ns + (long long)oct * HW * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float v[8];
    cg_unpack8(*reinterpret_cast<const uint4*>(xp + (long long)p * 8), v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = cg_warp_sum(s[i]); q[i] = cg_warp_sum(q[i]); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_red[warp][i] = s[i]; s_red[warp][8 + i] = q[i]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += s_red[wv][threadIdx.x];
    const int c = oct * 8 + (threadIdx.x & 7);
    if (c < C) stats[((long long)n * C + c) * 2 + (threadIdx.x >> 3)] = t;
  }
}

// GroupNorm from the channel statistics + affine + optional residual + activation (nn.GroupNorm(G, C), src/pgm/resnet.py:
// 228; CustomBlock.forward :41-61: out = relu(gn2(conv2(.)) + identity)).  One thread = (pixel, octet).
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int groups, float eps, const bf16* __restrict__ add,
                                                              long long add_ns, int act, bf16* __restrict__ y, int C, int HW,
                                                              long long x_ns, long long y_ns) {
  __shared__ float s_scale[8], s_shift[8];
  const int oct = blockIdx.y, n = blockIdx.z;
  if (threadIdx.x < 8) {
    const int c = oct * 8 + threadIdx.x;
    float sc = 0.f, sh = 0.f;
    if (c < C) {
      const int cpg = C / groups, g0 = (c / cpg) * cpg;
      double su = 0.0, sq = 0.0;
      for (int j = 0; j < cpg; ++j) {
        su += (double)stats[((long long)n * C + g0 + j) * 2];
        sq += (double)stats[((long long)n * C + g0 + j) * 2 + 1];
      }
      const double cnt = (double)cpg * HW, mean = su / cnt;
      double var = sq / cnt - mean * mean;  // biased variance, like torch
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      sc = rstd * gamma[c];
      sh = beta[c] - (float)mean * sc;
    }
    s_scale[threadIdx.x] = sc;
    s_shift[threadIdx.x] = sh;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const long long off = ((long long)oct * HW + p) * 8;
  float v[8], r[8];
  cg_unpack8(*reinterpret_cast<const uint4*>(x + n * x_ns + off), v);
  if (add != nullptr) cg_unpack8(*reinterpret_cast<const uint4*>(add + n * add_ns + off), r);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = fmaf(v[i], s_scale[i], s_shift[i]);
    if (add != nullptr) t += r[i];
    v[i] = act_apply(t, act);
  }
  *reinterpret_cast<uint4*>(y + n * y_ns + off) = cg_pack8(v);
}

// x.mean(dim=(-2,-1)) -> fp32 (N, ld) rows, columns [0, C)
__global__ void __launch_bounds__(256) global_avgpool_kernel(const bf16* __restrict__ x, float* __restrict__ out, int C, int HW,
                                                             long long x_ns, int ld) {
  __shared__ float s_red[8][8];
  const int oct = blockIdx.x, n = blockIdx.y;
  const bf16* xp = x + n * x_ns + (long long)oct * HW * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float v[8];
    cg_unpack8(*reinterpret_cast<const uint4*>(xp + (long long)p * 8), v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = cg_warp_sum(s[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[warp][i] = s[i];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += s_red[wv][threadIdx.x];
    const int c = oct * 8 + threadIdx.x;
    if (c < C) out[(long long)n * ld + c] = t / (float)HW;
  }
}

// out[n][m] = act((sum_k x[n][k] w[m][k] + bias[m]) * scale[m] + shift[m]): one warp per output element
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const float* __restrict__ scale,
                                                     const float* __restrict__ shift, int act, float* __restrict__ out,
                                                     int ldo, int N, int K, int M) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= N * M) return;
  const int n = gw / M, m = gw - n * M;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(x[(long long)n * ldx + k], w[(long long)m * K + k], acc);
  acc = cg_warp_sum(acc);
  if (lane == 0) {
    if (bias != nullptr) acc += bias[m];
    if (scale != nullptr) acc = fmaf(acc, scale[m], shift[m]);
    out[(long long)n * ldo + m] = act_apply(acc, act);
  }
}

}  // namespace

extern "C" int cg_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
                          float* shift, int32_t C, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, "cg_bn_fold: null / empty");
  bn_fold_kernel<<<cg_ceil_div(C, 128), 128, 0, cg_stream(stream)>>>(gamma, beta, mean, var, eps, scale, shift, C);
  CG_LAUNCH_CHECK("cg_bn_fold");
  return CG_OK;
}

extern "C" int cg_conv_direct_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t N,
                                  int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t k, int32_t stride, int32_t pad,
                                  int32_t act, int64_t y_ns, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0 && Cout > 0, "cg_conv_direct_fwd: null / empty");
  CG_REQUIRE(Cin >= 1 && Cin <= 3 && k >= 1 && k <= 7 && stride >= 1 && pad >= 0, "cg_conv_direct_fwd: Cin %d k %d (thin-input stems only)", Cin, k);
  CG_REQUIRE((scale == nullptr) == (shift == nullptr), "cg_conv_direct_fwd: scale and shift go together");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  dim3 grid(cg_ceil_div(Ho * Wo, 256), cg_ceil_div(Cout, 8), N);
  conv_direct_kernel<<<grid, 256, 0, cg_stream(stream)>>>(x, w, scale, shift, reinterpret_cast<bf16*>(y), Cin, H, W, Cout, k,
                                                          stride, pad, Ho, Wo, act, y_ns);
  CG_LAUNCH_CHECK("cg_conv_direct_fwd");
  return CG_OK;
}

extern "C" int cg_pool_max_fwd(const void* x, void* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t k, int32_t stride,
                               int32_t pad, int64_t x_ns, int64_t y_ns, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(x && y && x != y && N > 0 && C > 0 && C % 8 == 0 && k >= 1 && stride >= 1 && pad >= 0 && pad < k + (k == 1),
             "cg_pool_max_fwd: bad arguments");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  CG_REQUIRE(Ho > 0 && Wo > 0, "cg_pool_max_fwd: empty output");
  dim3 grid(cg_ceil_div(Ho * Wo, 256), C / 8, N);
  pool_max_kernel<<<grid, 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), C / 8, H,
                                                       W, k, stride, pad, Ho, Wo, x_ns, y_ns);
  CG_LAUNCH_CHECK("cg_pool_max_fwd");
  return CG_OK;
}

extern "C" int cg_groupnorm_fwd(const void* x, float* stats, const float* gamma, const float* beta, int32_t groups, float eps,
                                const void* add, int64_t add_ns, int32_t act, void* y, int32_t N, int32_t C, int32_t HW,
                                int64_t x_ns, int64_t y_ns, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(x && stats && gamma && beta && y && N > 0 && C > 0 && HW > 0, "cg_groupnorm_fwd: null / empty");
  CG_REQUIRE(groups > 0 && C % groups == 0 && C % 8 == 0, "cg_groupnorm_fwd: C %d groups %d", C, groups);
  chan_stats_kernel<<<dim3(C / 8, N), 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(x), stats, C, HW, x_ns);
  CG_LAUNCH_CHECK("cg_groupnorm_fwd (stats)");
  dim3 grid(cg_ceil_div(HW, 256), C / 8, N);
  groupnorm_apply_kernel<<<grid, 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(x), stats, gamma, beta, groups, eps,
                                                              reinterpret_cast<const bf16*>(add), add_ns, act,
                                                              reinterpret_cast<bf16*>(y), C, HW, x_ns, y_ns);
  CG_LAUNCH_CHECK("cg_groupnorm_fwd (apply)");
  return CG_OK;
}

extern "C" int cg_global_avgpool(const void* x, float* out, int32_t N, int32_t C, int32_t HW, int64_t x_ns, int32_t ld,
                                 void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(x && out && N > 0 && C > 0 && HW > 0 && ld >= C, "cg_global_avgpool: bad arguments");
  global_avgpool_kernel<<<dim3(cg_ceil_div(C, 8), N), 256, 0, cg_stream(stream)>>>(reinterpret_cast<const bf16*>(x), out, C, HW,
                                                                                   x_ns, ld);
  CG_LAUNCH_CHECK("cg_global_avgpool");
  return CG_OK;
}

extern "C" int cg_linear(const float* x, int32_t ldx, const float* w, const float* bias, const float* scale, const float* shift,
                         int32_t act, float* out, int32_t ldo, int32_t N, int32_t K, int32_t M, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(x && w && out && N > 0 && K > 0 && M > 0 && ldx >= K && ldo >= M, "cg_linear: bad arguments");
  CG_REQUIRE((scale == nullptr) == (shift == nullptr), "cg_linear: scale and shift go together");
  linear_kernel<<<cg_ceil_div(N * M * 32, 256), 256, 0, cg_stream(stream)>>>(x, ldx, w, bias, scale, shift, act, out, ldo, N, K,
                                                                             M);
  CG_LAUNCH_CHECK("cg_linear");
  return CG_OK;
}
