// Pixel likelihoods fused with their 1x1 heads:
//   DGaussNet  (reference src/vae.py:322-422)  -- discretised Gaussian, tanh-approx CDF, RGB coefficients
//   DmolNet    (reference src/dmol.py:24-245)  -- discretised mixture of 10 logistics, RGB only
#include "cg_common.cuh"

namespace {

constexpr float kSqrt2OverPi = 0.7978845608028654f;
constexpr float kBin = 1.0f / 255.0f;

__device__ __forceinline__ float approx_cdf(float v) {  // src/vae.py:388-391
  return 0.5f * (1.0f + tanhf(kSqrt2OverPi * (v + 0.044715f * v * v * v)));
}
__device__ __forceinline__ float approx_cdf_grad(float v) {
  float th = tanhf(kSqrt2OverPi * (v + 0.044715f * v * v * v));
  return 0.5f * (1.0f - th * th) * kSqrt2OverPi * (1.0f + 3.0f * 0.044715f * v * v);
}

constexpr int kMaxCw = 64;

struct Heads {  // shared-memory copy of the 1x1 head weights
  float w_loc[3][kMaxCw], w_ls[3][kMaxCw], w_co[3][kMaxCw];
  float b_loc[3], b_ls[3], b_co[3];
};

__device__ void load_heads(Heads& s, const cg_dgauss_args& a) {
  for (int i = threadIdx.x; i < a.C * a.Cw; i += blockDim.x) {
    int c = i / a.Cw, k = i - c * a.Cw;
    s.w_loc[c][k] = a.w_loc[i];
    s.w_ls[c][k] = a.w_ls[i];
    s.w_co[c][k] = a.w_co != nullptr ? a.w_co[i] : 0.f;
  }
  if (threadIdx.x < a.C) {
    s.b_loc[threadIdx.x] = a.b_loc[threadIdx.x];
    s.b_ls[threadIdx.x] = a.b_ls[threadIdx.x];
    s.b_co[threadIdx.x] = a.b_co != nullptr ? a.b_co[threadIdx.x] : 0.f;
  }
  __syncthreads();
}

// per-pixel head evaluation: raw loc / raw logscale / raw coeff for C channels
template <int C>
__device__ __forceinline__ void eval_heads(const Heads& s, const bf16* hrow, long long oct_stride, int Cw, float* loc,
                                           float* ls, float* co, bool rgb, uint4* stash = nullptr) {
#pragma unroll
  for (int c = 0; c < C; ++c) { loc[c] = s.b_loc[c]; ls[c] = s.b_ls[c]; co[c] = s.b_co[c]; }
  for (int k8 = 0; k8 < Cw; k8 += 8) {
    float h[8];
    const uint4 raw = *reinterpret_cast<const uint4*>(hrow + (k8 >> 3) * oct_stride);
    if (stash != nullptr) stash[k8 >> 3] = raw;  // the backward kernel keeps the pixel's features for the weight gradient
    cg_unpack8(raw, h);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        loc[c] += s.w_loc[c][k8 + k] * h[k];
        ls[c] += s.w_ls[c][k8 + k] * h[k];
        if (rgb) co[c] += s.w_co[c][k8 + k] * h[k];
      }
    }
  }
}

// log-prob of one sub-pixel and its derivatives wrt loc and (clamped) logscale  src/vae.py:393-410
__device__ __forceinline__ float dgauss_logprob(float x, float loc, float ls, float* dloc, float* dls) {
  float inv = __expf(-ls);
  float d = x - loc;
  float up = inv * (d + kBin), dn = inv * (d - kBin);
  float cu = approx_cdf(up), cd = approx_cdf(dn);
  float lp, g_up = 0.f, g_dn = 0.f;
  if (x < -0.999f) {
    float v = fmaxf(cu, 1e-12f);
    lp = __logf(v);
    if (cu > 1e-12f) g_up = approx_cdf_grad(up) / v;
  } else if (x > 0.999f) {
    float v = fmaxf(1.0f - cd, 1e-12f);
    lp = __logf(v);
    if (1.0f - cd > 1e-12f) g_dn = -approx_cdf_grad(dn) / v;
  } else {
    float dl = cu - cd;
    float v = fmaxf(dl, 1e-12f);
    lp = __logf(v);
    if (dl > 1e-12f) { g_up = approx_cdf_grad(up) / v; g_dn = -approx_cdf_grad(dn) / v; }
  }
  if (dloc != nullptr) {
    *dloc = -inv * (g_up + g_dn);          // d up/d loc = d dn/d loc = -inv
    *dls = -(g_up * up + g_dn * dn);        // d up/d ls = -up
  }
  return lp;
}

template <int C>
__global__ void __launch_bounds__(256) dgauss_fwd_kernel(const cg_dgauss_args a) {
  __shared__ Heads s;
  __shared__ float red[8];
  load_heads(s, a);
  const int n = blockIdx.y;
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (hw < a.HW) {
    const long long pix = (long long)n * a.HW + hw;
    float loc[C], ls[C], co[C], x[C];
    eval_heads<C>(s, reinterpret_cast<const bf16*>(a.h) + n * a.h_ns + (long long)hw * 8, (long long)a.HW * 8, a.Cw, loc,
                  ls, co, C == 3);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      x[c] = a.x[((long long)n * C + c) * a.HW + hw];
      ls[c] = fmaxf(ls[c], -9.0f);  // src/vae.py:355
    }
    if (C == 3) {  // src/vae.py:370-377 (training branch: conditioned on the true sub-pixels)
      float c0 = tanhf(co[0]), c1 = tanhf(co[1]), c2 = tanhf(co[2]);
      loc[1] += c0 * x[0];
      loc[2] += c1 * x[0] + c2 * x[1];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) acc += dgauss_logprob(x[c], loc[c], ls[c], nullptr, nullptr);
  }
  acc = cg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(a.nll + n, -t / (float)(C * a.HW));
  }
}

// MODE 0: backward of the NLL (dgauss_fwd_kernel).  MODE 1: backward of likelihood.sample(h, return_loc=True)
// (dgauss_sample_kernel, src/vae.py:413-422 + 352-369) given d x_out / d scale_out (fp32 NCHW) -- the counterfactual
// training path back-propagates through cf_loc / cf_scale / rec_loc / rec_scale (src/pgm/dscm.py:53-56,78-88).
template <int C, int MODE>
__global__ void __launch_bounds__(256) dgauss_bwd_kernel(const cg_dgauss_args a, const float* __restrict__ dx_out,
                                                         const float* __restrict__ dscale_out) {
  __shared__ Heads s;
  __shared__ float s_d[256][3 * C + 1];  // per pixel: dloc[C], dls[C], dco[C]
  // the block's features, bf16, one row per pixel with a 16-byte pad (pitch 2*Cw + 16 B: eight lanes of a 16-byte store hit
  // eight different bank groups); dynamic, allocated when Cw <= 32 (every shipped config), else the weight gradient
  // re-reads h from global memory
  extern __shared__ __align__(16) uint8_t s_hraw[];
  const int h_pitch = 2 * a.Cw + 16;
  const bool h_smem = a.Cw <= 32;
  load_heads(s, a);
  const int n = blockIdx.y;
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  const float scale = -a.g / (float)(C * a.HW);  // d loss / d logprob
  float dloc[C], dls[C], dco[C];
#pragma unroll
  for (int c = 0; c < C; ++c) dloc[c] = dls[c] = dco[c] = 0.f;
  if (hw < a.HW) {
    const long long pix = (long long)n * a.HW + hw;
    const bf16* hrow = reinterpret_cast<const bf16*>(a.h) + n * a.h_ns + (long long)hw * 8;
    float loc[C], ls[C], co[C], x[C];
    bool live[C];
    eval_heads<C>(s, hrow, (long long)a.HW * 8, a.Cw, loc, ls, co, C == 3,
                  h_smem ? reinterpret_cast<uint4*>(s_hraw + threadIdx.x * h_pitch) : nullptr);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      x[c] = MODE == 0 ? a.x[((long long)n * C + c) * a.HW + hw] : 0.f;
      live[c] = ls[c] >= -9.0f;
      ls[c] = fmaxf(ls[c], -9.0f);
    }
    float tc[3] = {0.f, 0.f, 0.f};
    if (MODE == 1) {
      float gx[C], gs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const long long o = ((long long)n * C + c) * a.HW + hw;
        gx[c] = dx_out != nullptr ? dx_out[o] : 0.f;
        gs[c] = dscale_out != nullptr ? dscale_out[o] : 0.f;
        dls[c] = live[c] ? gs[c] * __expf(ls[c]) : 0.f;  // scale = exp(max(raw, -9))
      }
      auto pass = [](float v) { return v >= -1.f && v <= 1.f; };  // torch.clamp backward mask (inclusive)
      if (C == 1) {
        dloc[0] = pass(loc[0]) ? gx[0] : 0.f;  // the final clamp(-1, 1) of src/vae.py:421
      } else {
        // r = clamp(l0); g = clamp(l1 + c0 r); b = clamp(l2 + c1 r + c2 g)   (src/vae.py:360-369), then clamp again (no-op)
        tc[0] = tanhf(co[0]); tc[1] = tanhf(co[1]); tc[2] = tanhf(co[2]);
        const float r = fminf(fmaxf(loc[0], -1.f), 1.f);
        const float pg = loc[1] + tc[0] * r, g = fminf(fmaxf(pg, -1.f), 1.f);
        const float pb = loc[2] + tc[1] * r + tc[2] * g;
        const float db_ = pass(pb) ? gx[2] : 0.f;
        const float dg_ = pass(pg) ? gx[1] + db_ * tc[2] : 0.f;
        const float dr_ = pass(loc[0]) ? gx[0] + dg_ * tc[0] + db_ * tc[1] : 0.f;
        dloc[0] = dr_; dloc[1] = dg_; dloc[2] = db_;
        dco[0] = dg_ * r * (1.0f - tc[0] * tc[0]);
        dco[1] = db_ * r * (1.0f - tc[1] * tc[1]);
        dco[2] = db_ * g * (1.0f - tc[2] * tc[2]);
      }
    }
    if (MODE == 0 && C == 3) {
      tc[0] = tanhf(co[0]); tc[1] = tanhf(co[1]); tc[2] = tanhf(co[2]);
      loc[1] += tc[0] * x[0];
      loc[2] += tc[1] * x[0] + tc[2] * x[1];
    }
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float gl, gs;
        dgauss_logprob(x[c], loc[c], ls[c], &gl, &gs);
        dloc[c] = scale * gl;
        dls[c] = live[c] ? scale * gs : 0.f;
      }
    }
    if (MODE == 0 && C == 3) {
      dco[0] = dloc[1] * x[0] * (1.0f - tc[0] * tc[0]);
      dco[1] = dloc[2] * x[0] * (1.0f - tc[1] * tc[1]);
      dco[2] = dloc[2] * x[1] * (1.0f - tc[2] * tc[2]);
    }
    // dh = W_loc^T dloc + W_ls^T dls + W_co^T dco
    bf16* drow = reinterpret_cast<bf16*>(a.dh) + n * a.dh_ns + (long long)hw * 8;
    for (int k8 = 0; k8 < a.Cw; k8 += 8) {
      float g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c)
          v += s.w_loc[c][k8 + k] * dloc[c] + s.w_ls[c][k8 + k] * dls[c] + (C == 3 ? s.w_co[c][k8 + k] * dco[c] : 0.f);
        g[k] = v;
      }
      *reinterpret_cast<uint4*>(drow + (long long)(k8 >> 3) * a.HW * 8) = cg_pack8(g);
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    s_d[threadIdx.x][c] = dloc[c];
    s_d[threadIdx.x][C + c] = dls[c];
    s_d[threadIdx.x][2 * C + c] = dco[c];
  }
  __syncthreads();
  // head weight gradients: dW[head][c][k] = sum_pixels d[head][c](p) * h[p][k].  The block's 256 pixels are split into
  // kParts slices so that all 256 threads work, and h comes from the shared-memory copy made above: the first versions
  // re-read it from global memory with one 2-byte load per (pixel, output) -- 60 % of the kernel's stall samples, 282 /
  // 275 us = 0.32 of HBM (ncu profiles/r2m_ncu_full_summary.txt, r2w_*); slice partials meet in shared memory, one atomic
  // per entry.
  const int nhead = (C == 3) ? 3 : 2;
  const int per = C * a.Cw;
  const int nout = nhead * per;                       // 64 (C=1, Cw=32) ... 144 (C=3, Cw=16)
  const int npx = min(256, a.HW - (int)(blockIdx.x * blockDim.x));
  __shared__ float s_part[4][3 * 3 * kMaxCw];
  const int kParts = nout <= 64 ? 4 : (nout <= 128 ? 2 : 1);
  const int slice = 256 / kParts;
  {
    const int part = threadIdx.x / (256 / kParts), t = threadIdx.x - part * (256 / kParts);
    for (int o = t; o < nout; o += 256 / kParts) {
      const int head = o / per, r = o - head * per, c = r / a.Cw, k = r - c * a.Cw;
      const bf16* hb = reinterpret_cast<const bf16*>(a.h) + n * a.h_ns +
                       ((long long)(k >> 3) * a.HW + blockIdx.x * blockDim.x) * 8 + (k & 7);
      float acc = 0.f;
      const int p1 = min(npx, (part + 1) * slice);
      if (h_smem) {
        const uint8_t* hs = s_hraw + 2 * k;
        for (int p = part * slice; p < p1; ++p)
          acc += s_d[p][head * C + c] * __bfloat162float(*reinterpret_cast<const bf16*>(hs + p * h_pitch));
      } else {
        for (int p = part * slice; p < p1; ++p) acc += s_d[p][head * C + c] * __bfloat162float(hb[(long long)p * 8]);
      }
      s_part[part][o] = acc;
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < nout; o += blockDim.x) {
    const int head = o / per, r = o - head * per;
    float acc = 0.f;
    for (int q = 0; q < kParts; ++q) acc += s_part[q][o];
    float* dst = head == 0 ? a.dw_loc : (head == 1 ? a.dw_ls : a.dw_co);
    if (dst != nullptr) atomicAdd(dst + r, acc);
  }
  for (int t = threadIdx.x; t < nhead * C; t += blockDim.x) {
    int head = t / C, c = t - head * C;
    float acc = 0.f;
    for (int p = 0; p < npx; ++p) acc += s_d[p][head * C + c];
    float* dst = head == 0 ? a.db_loc : (head == 1 ? a.db_ls : a.db_co);
    if (dst != nullptr) atomicAdd(dst + c, acc);
  }
}

template <int C>
__global__ void __launch_bounds__(256) dgauss_sample_kernel(const cg_dgauss_args a, float* __restrict__ x_out,
                                                            float* __restrict__ scale_out, const float* __restrict__ eps,
                                                            float log_t) {
  __shared__ Heads s;
  load_heads(s, a);
  const int n = blockIdx.y;
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  if (hw >= a.HW) return;
  const long long pix = (long long)n * a.HW + hw;
  float loc[C], ls[C], co[C];
  eval_heads<C>(s, reinterpret_cast<const bf16*>(a.h) + n * a.h_ns + (long long)hw * 8, (long long)a.HW * 8, a.Cw, loc, ls,
                co, C == 3);
#pragma unroll
  for (int c = 0; c < C; ++c) ls[c] = fmaxf(ls[c], -9.0f);
  if (C == 3) {  // src/vae.py:360-369 (inference branch: clamped predicted means feed the next sub-pixel)
    float c0 = tanhf(co[0]), c1 = tanhf(co[1]), c2 = tanhf(co[2]);
    float r = fminf(fmaxf(loc[0], -1.f), 1.f);
    float g = fminf(fmaxf(loc[1] + c0 * r, -1.f), 1.f);
    float b = fminf(fmaxf(loc[2] + c1 * r + c2 * g, -1.f), 1.f);
    loc[0] = r; loc[1] = g; loc[2] = b;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    long long o = ((long long)n * C + c) * a.HW + hw;
    float l = ls[c], v = loc[c];
    if (eps != nullptr) { l += log_t; v += __expf(l) * eps[o]; }
    x_out[o] = fminf(fmaxf(v, -1.f), 1.f);  // src/vae.py:421
    scale_out[o] = __expf(l);
  }
}

// ------------------------------------------------------------------------------------- DMoL
// 16 lanes per pixel, lane m < 10 owns mixture m; log-sum-exp over mixtures by half-warp shuffles.
constexpr int kMix = 10;
constexpr int kDmolPix = 16;  // pixels per 256-thread block

__device__ __forceinline__ float hw_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, 16));
  return v;
}
__device__ __forceinline__ float hw_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}
__device__ __forceinline__ float softplusf(float v) {  // F.softplus (threshold 20)
  return v > 20.0f ? v : log1pf(__expf(v));
}
__device__ __forceinline__ float sigmoidf(float v) { return 1.0f / (1.0f + __expf(-v)); }

struct DmolLane {
  float logit, mean[3], ls_raw[3], co_raw[3];
};

// the 10 head outputs this lane (mixture m) needs: logit m and {mean, logscale, coeff}[c][m]
template <int CW>
__device__ __forceinline__ DmolLane dmol_heads(const float* sw, const float* sb, const float* h, int m) {
  DmolLane r;
  auto dot = [&](int o) {
    float v = sb[o];
#pragma unroll
    for (int k = 0; k < CW; ++k) v += sw[o * (CW + 1) + k] * h[k];
    return v;
  };
  r.logit = dot(m);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int base = kMix + c * 3 * kMix;  // src/dmol.py:35-38 channel layout
    r.mean[c] = dot(base + m);
    r.ls_raw[c] = dot(base + kMix + m);
    r.co_raw[c] = dot(base + 2 * kMix + m);
  }
  return r;
}

// log-prob of sub-pixel under one logistic component + derivatives wrt mean and clamped log-scale
__device__ __forceinline__ float logistic_logprob(float x, float mu, float ls, float* dmu, float* dls) {
  float inv = __expf(-ls), d = x - mu;
  float up = inv * (d + kBin), dn = inv * (d - kBin), mid = inv * d;
  float lp;
  float g_up = 0.f, g_dn = 0.f, g_mid = 0.f, g_ls_direct = 0.f;
  if (x < -0.999f) {  // log cdf(up) = up - softplus(up)
    lp = up - softplusf(up);
    g_up = 1.0f - sigmoidf(up);
  } else if (x > 0.999f) {  // log(1 - cdf(dn)) = -softplus(dn)
    lp = -softplusf(dn);
    g_dn = -sigmoidf(dn);
  } else {
    float su = sigmoidf(up), sd = sigmoidf(dn);
    float delta = su - sd;
    if (delta > 1e-5f) {
      lp = __logf(fmaxf(delta, 1e-12f));
      g_up = su * (1.0f - su) / delta;
      g_dn = -sd * (1.0f - sd) / delta;
    } else {  // src/dmol.py:75-77,112
      lp = mid - ls - 2.0f * softplusf(mid) - 4.8481163f;  // log(127.5)
      g_mid = 1.0f - 2.0f * sigmoidf(mid);
      g_ls_direct = -1.0f;
    }
  }
  if (dmu != nullptr) {
    *dmu = -inv * (g_up + g_dn + g_mid);
    *dls = -(g_up * up + g_dn * dn + g_mid * mid) + g_ls_direct;
  }
  return lp;
}

template <int CW>
__device__ __forceinline__ void dmol_stage(const cg_dmol_args& a, float* sw, float* sb) {
  // row pitch CW+1 so the 10 mixture lanes of a pixel hit distinct banks
  for (int i = threadIdx.x; i < 100 * CW; i += blockDim.x) sw[(i / CW) * (CW + 1) + (i % CW)] = a.w[i];
  for (int i = threadIdx.x; i < 100; i += blockDim.x) sb[i] = a.b[i];
  __syncthreads();
}

template <int CW>
__global__ void __launch_bounds__(256) dmol_fwd_kernel(const cg_dmol_args a) {
  __shared__ float sw[100 * (CW + 1)];
  __shared__ float sb[100];
  dmol_stage<CW>(a, sw, sb);
  const int n = blockIdx.y;
  const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int hw = blockIdx.x * kDmolPix + grp;
  const bool live = hw < a.HW;
  const int m = min(lane16, kMix - 1);
  float h[CW];
  float x[3] = {0.f, 0.f, 0.f};
  const long long pix = (long long)n * a.HW + (live ? hw : 0);
#pragma unroll
  for (int k8 = 0; k8 < CW; k8 += 8)
    cg_unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.h) + n * a.h_ns +
                                               ((long long)(k8 >> 3) * a.HW + (live ? hw : 0)) * 8), h + k8);
  for (int c = 0; c < 3; ++c) x[c] = a.x[((long long)n * 3 + c) * a.HW + (live ? hw : 0)];
  DmolLane L = dmol_heads<CW>(sw, sb, h, m);
  float c0 = tanhf(L.co_raw[0]), c1 = tanhf(L.co_raw[1]), c2 = tanhf(L.co_raw[2]);
  float mu[3] = {L.mean[0], L.mean[1] + c0 * x[0], L.mean[2] + c1 * x[0] + c2 * x[1]};  // src/dmol.py:42-51
  float lp = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) lp += logistic_logprob(x[c], mu[c], fmaxf(L.ls_raw[c], -7.0f), nullptr, nullptr);
  const bool act = lane16 < kMix;
  float lg = act ? L.logit : -INFINITY;
  float mx = hw_max(lg);
  float lse_logit = mx + __logf(hw_sum(act ? __expf(lg - mx) : 0.f));
  float t = act ? lp + lg - lse_logit : -INFINITY;  // src/dmol.py:116
  float tm = hw_max(t);
  float logp = tm + __logf(hw_sum(act ? __expf(t - tm) : 0.f));  // src/dmol.py:117
  // block reduce over the 16 pixels
  __shared__ float red[kDmolPix];
  if (lane16 == 0) red[grp] = live ? logp : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kDmolPix; ++i) s += red[i];
    atomicAdd(a.nll + n, -s / (float)(3 * a.HW));  // src/dmol.py:118
  }
}

// One block walks kDmolBwdIters groups of 16 pixels: the head weights are staged once per block and the head weight /
// bias gradients are accumulated in registers across the groups and flushed with ONE set of atomics per block.  (The
// first version staged the weights and issued 1600 + 100 atomics per 16 pixels: 2.9 ms at cmnist batch 1024, ncu
// profiles/r2m_ncu_full_summary.txt, against 0.5 ms for the forward kernel.)
constexpr int kDmolBwdIters = 16;
template <int CW>
__global__ void __launch_bounds__(256) dmol_bwd_kernel(const cg_dmol_args a) {
  __shared__ float sw[100 * (CW + 1)];
  __shared__ float sb[100];
  __shared__ float s_dl[kDmolPix][100];
  __shared__ float s_h[kDmolPix][CW + 1];
  dmol_stage<CW>(a, sw, sb);
  const int n = blockIdx.y;
  const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int m = min(lane16, kMix - 1);
  const bool act = lane16 < kMix;
  constexpr int kOwn = (100 * CW + 255) / 256;  // (output, k) weight-gradient entries owned by a thread
  __shared__ float s_wacc[100 * CW];              // their running sums (each entry is touched by its owner only)
  for (int i = threadIdx.x; i < 100 * CW; i += 256) s_wacc[i] = 0.f;
  float bacc = 0.f;
  auto oidx = [&](int i) {
    if (i == 0) return m;
    int c = (i - 1) / 3, w = (i - 1) % 3;
    return kMix + c * 3 * kMix + w * kMix + m;
  };
#pragma unroll 1
  for (int it = 0; it < kDmolBwdIters; ++it) {
    const int hw0 = (blockIdx.x * kDmolBwdIters + it) * kDmolPix;
    if (hw0 >= a.HW) break;
    const int hw = hw0 + grp;
    const bool live = hw < a.HW;
    float h[CW];
    float x[3];
#pragma unroll
    for (int k8 = 0; k8 < CW; k8 += 8)
      cg_unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.h) + n * a.h_ns +
                                                 ((long long)(k8 >> 3) * a.HW + (live ? hw : 0)) * 8), h + k8);
    for (int c = 0; c < 3; ++c) x[c] = a.x[((long long)n * 3 + c) * a.HW + (live ? hw : 0)];
    DmolLane L = dmol_heads<CW>(sw, sb, h, m);
    float tc[3] = {tanhf(L.co_raw[0]), tanhf(L.co_raw[1]), tanhf(L.co_raw[2])};
    float mu[3] = {L.mean[0], L.mean[1] + tc[0] * x[0], L.mean[2] + tc[1] * x[0] + tc[2] * x[1]};
    float lp = 0.f, dmu[3], dls[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float lsc = fmaxf(L.ls_raw[c], -7.0f);
      lp += logistic_logprob(x[c], mu[c], lsc, &dmu[c], &dls[c]);
      if (L.ls_raw[c] < -7.0f) dls[c] = 0.f;
    }
    float lg = act ? L.logit : -INFINITY;
    float mx = hw_max(lg);
    float se = hw_sum(act ? __expf(lg - mx) : 0.f);
    float lse_logit = mx + __logf(se);
    float prior = act ? __expf(lg - lse_logit) : 0.f;
    float t = act ? lp + lg - lse_logit : -INFINITY;
    float tm = hw_max(t);
    float st = hw_sum(act ? __expf(t - tm) : 0.f);
    float resp = act ? __expf(t - tm) / st : 0.f;  // posterior responsibility of mixture m
    const float gs = live ? -a.g / (float)(3 * a.HW) : 0.f;  // d loss / d logp(pixel)
    // gradients of the 10 head outputs of this lane
    float d_out[10];
    d_out[0] = gs * (resp - prior);
    float gm[3] = {gs * resp * dmu[0], gs * resp * dmu[1], gs * resp * dmu[2]};
    float gco[3] = {gm[1] * x[0] * (1.f - tc[0] * tc[0]), gm[2] * x[0] * (1.f - tc[1] * tc[1]),
                    gm[2] * x[1] * (1.f - tc[2] * tc[2])};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d_out[1 + c * 3] = gm[c];
      d_out[2 + c * 3] = gs * resp * dls[c];
      d_out[3 + c * 3] = gco[c];
    }
    if (!act) {
#pragma unroll
      for (int i = 0; i < 10; ++i) d_out[i] = 0.f;
    }
    // dh[k] = sum over the lane's outputs and over lanes
    bf16* drow = reinterpret_cast<bf16*>(a.dh) + n * a.dh_ns + (long long)(live ? hw : 0) * 8;
#pragma unroll
    for (int k = 0; k < CW; ++k) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 10; ++i) v += d_out[i] * sw[oidx(i) * (CW + 1) + k];
      v = hw_sum(v);
      if (lane16 == 0 && live) drow[(long long)(k >> 3) * a.HW * 8 + (k & 7)] = __float2bfloat16(v);
    }
    __syncthreads();  // the previous group's s_dl / s_h have been consumed
    if (act) {
#pragma unroll
      for (int i = 0; i < 10; ++i) s_dl[grp][oidx(i)] = d_out[i];  // zero for dead pixels (gs = 0)
    }
    if (lane16 == 0) {
#pragma unroll
      for (int k = 0; k < CW; ++k) s_h[grp][k] = h[k];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kOwn; ++i) {
      const int tI = threadIdx.x + i * 256;
      if (tI < 100 * CW) {
        const int o = tI / CW, k = tI - o * CW;
        float acc = 0.f;
#pragma unroll
        for (int p = 0; p < kDmolPix; ++p) acc += s_dl[p][o] * s_h[p][k];
        s_wacc[tI] += acc;
      }
    }
    if (threadIdx.x < 100) {
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < kDmolPix; ++p) acc += s_dl[p][threadIdx.x];
      bacc += acc;
    }
  }
#pragma unroll
  for (int i = 0; i < kOwn; ++i) {
    const int tI = threadIdx.x + i * 256;
    if (tI < 100 * CW) atomicAdd(a.dw + tI, s_wacc[tI]);
  }
  if (threadIdx.x < 100) atomicAdd(a.db + threadIdx.x, bacc);
}

template <int CW>
__global__ void __launch_bounds__(256) dmol_predict_kernel(const cg_dmol_args a, int mode,
                                                           const float* __restrict__ u_gumbel,
                                                           const float* __restrict__ u_logistic, float log_t,
                                                           float* __restrict__ x_out, float* __restrict__ scale_out) {
  __shared__ float sw[100 * (CW + 1)];
  __shared__ float sb[100];
  dmol_stage<CW>(a, sw, sb);
  const int n = blockIdx.y;
  const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int hw = blockIdx.x * kDmolPix + grp;
  const bool live = hw < a.HW;
  const int m = min(lane16, kMix - 1);
  const bool act = lane16 < kMix;
  float h[CW];
  const long long pix = (long long)n * a.HW + (live ? hw : 0);
#pragma unroll
  for (int k8 = 0; k8 < CW; k8 += 8)
    cg_unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.h) + n * a.h_ns +
                                               ((long long)(k8 >> 3) * a.HW + (live ? hw : 0)) * 8), h + k8);
  DmolLane L = dmol_heads<CW>(sw, sb, h, m);
  float sel;
  if (mode == 0) {  // soft: softmax(logits)  src/dmol.py:170-172
    float lg = act ? L.logit : -INFINITY;
    float mx = hw_max(lg);
    float e = act ? __expf(lg - mx) : 0.f;
    sel = e / hw_sum(e);
  } else if (mode > 10) {  // 'top<k>' (k = mode - 10): keep logits >= the k-th largest, renormalise  src/dmol.py:178-189
    const float lg = act ? L.logit : -INFINITY;
    int greater = 0;
#pragma unroll
    for (int j = 0; j < kMix; ++j) greater += __shfl_sync(0xffffffffu, lg, j, 16) > lg ? 1 : 0;
    const bool keep = act && greater < mode - 10;   // "logit >= v[k-1]" of the descending sort, ties kept
    const float lk = keep ? lg : -INFINITY;
    const float mx = hw_max(lk);
    const float e = keep ? __expf(lk - mx) : 0.f;
    sel = e / hw_sum(e);
  } else {  // hard argmax (mode 1) or Gumbel argmax (mode 2)  src/dmol.py:128-131,175-177
    float score = L.logit;
    if (mode == 2) score -= __logf(-__logf(u_gumbel[pix * kMix + m]));
    if (!act) score = -INFINITY;
    float mx = hw_max(score);
    // first index attaining the max (torch.argmax tie rule)
    int cand = (act && score == mx) ? lane16 : 99;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o, 16));
    sel = (lane16 == cand) ? 1.f : 0.f;
  }
  float mean[3], ls[3], co[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mean[c] = hw_sum(sel * L.mean[c]);
    ls[c] = fmaxf(hw_sum(sel * L.ls_raw[c]), -7.0f);
    co[c] = hw_sum(sel * tanhf(L.co_raw[c]));
  }
  if (lane16 == 0 && live) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] = mean[c];
      if (mode == 2) {  // src/dmol.py:138-141
        float u = u_logistic[pix * 3 + c];
        ls[c] += log_t;
        v[c] += __expf(ls[c]) * (__logf(u) - __logf(1.0f - u));
      }
    }
    float x0 = fminf(fmaxf(v[0], -1.f), 1.f);
    float x1 = fminf(fmaxf(v[1] + co[0] * x0, -1.f), 1.f);
    float x2 = fminf(fmaxf(v[2] + co[1] * x0 + co[2] * x1, -1.f), 1.f);
    float xs[3] = {x0, x1, x2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      long long o = ((long long)n * 3 + c) * a.HW + hw;
      x_out[o] = xs[c];
      scale_out[o] = __expf(ls[c]);
    }
  }
}

}  // namespace

#define DGAUSS_CHECK(a, name)                                                                       \
  CG_REQUIRE((a) != nullptr && ((a)->C == 1 || (a)->C == 3), name ": C must be 1 or 3");           \
  CG_REQUIRE((a)->Cw % 8 == 0 && (a)->Cw <= kMaxCw && (a)->h_ns % 8 == 0, name ": Cw=%d", (a)->Cw)

extern "C" int cg_dgauss_nll_fwd(const cg_dgauss_args* a, void* stream) {
  CG_ARCH_GUARD();
  DGAUSS_CHECK(a, "cg_dgauss_nll_fwd");
  dim3 grid(cg_ceil_div(a->HW, 256), a->N);
  if (a->C == 1) dgauss_fwd_kernel<1><<<grid, 256, 0, cg_stream(stream)>>>(*a);
  else dgauss_fwd_kernel<3><<<grid, 256, 0, cg_stream(stream)>>>(*a);
  CG_LAUNCH_CHECK("cg_dgauss_nll_fwd");
  return CG_OK;
}

extern "C" int cg_dgauss_nll_bwd(const cg_dgauss_args* a, void* stream) {
  CG_ARCH_GUARD();
  DGAUSS_CHECK(a, "cg_dgauss_nll_bwd");
  CG_REQUIRE(a->dh != nullptr && a->dh_ns % 8 == 0, "cg_dgauss_nll_bwd: dh");
  dim3 grid(cg_ceil_div(a->HW, 256), a->N);
  const int hs_bytes0 = a->Cw <= 32 ? 256 * (2 * a->Cw + 16) : 0;  // shared-memory copy of the block's features
  if (a->C == 1) dgauss_bwd_kernel<1, 0><<<grid, 256, hs_bytes0, cg_stream(stream)>>>(*a, nullptr, nullptr);
  else dgauss_bwd_kernel<3, 0><<<grid, 256, hs_bytes0, cg_stream(stream)>>>(*a, nullptr, nullptr);
  CG_LAUNCH_CHECK("cg_dgauss_nll_bwd");
  return CG_OK;
}

extern "C" int cg_dgauss_sample_bwd(const cg_dgauss_args* a, const float* dx_out, const float* dscale_out, void* stream) {
  CG_ARCH_GUARD();
  DGAUSS_CHECK(a, "cg_dgauss_sample_bwd");
  CG_REQUIRE(a->dh != nullptr && a->dh_ns % 8 == 0, "cg_dgauss_sample_bwd: dh");
  CG_REQUIRE(dx_out != nullptr || dscale_out != nullptr, "cg_dgauss_sample_bwd: no upstream gradient");
  dim3 grid(cg_ceil_div(a->HW, 256), a->N);
  const int hs_bytes1 = a->Cw <= 32 ? 256 * (2 * a->Cw + 16) : 0;  // shared-memory copy of the block's features
  if (a->C == 1) dgauss_bwd_kernel<1, 1><<<grid, 256, hs_bytes1, cg_stream(stream)>>>(*a, dx_out, dscale_out);
  else dgauss_bwd_kernel<3, 1><<<grid, 256, hs_bytes1, cg_stream(stream)>>>(*a, dx_out, dscale_out);
  CG_LAUNCH_CHECK("cg_dgauss_sample_bwd");
  return CG_OK;
}

extern "C" int cg_dgauss_sample(const cg_dgauss_args* a, float* x_out, float* scale_out, const float* eps, float log_t,
                                void* stream) {
  CG_ARCH_GUARD();
  DGAUSS_CHECK(a, "cg_dgauss_sample");
  dim3 grid(cg_ceil_div(a->HW, 256), a->N);
  if (a->C == 1) dgauss_sample_kernel<1><<<grid, 256, 0, cg_stream(stream)>>>(*a, x_out, scale_out, eps, log_t);
  else dgauss_sample_kernel<3><<<grid, 256, 0, cg_stream(stream)>>>(*a, x_out, scale_out, eps, log_t);
  CG_LAUNCH_CHECK("cg_dgauss_sample");
  return CG_OK;
}

#define DMOL_CHECK(a, name)                                                                                 \
  CG_REQUIRE((a) != nullptr && ((a)->Cw == 16 || (a)->Cw == 32) && (a)->h_ns % 8 == 0, name ": Cw=%d must be 16 or 32", \
             (a) ? (a)->Cw : -1)
#define DMOL_LAUNCH(kern, a, ...)                                                            \
  do {                                                                                       \
    dim3 grid(cg_ceil_div((a)->HW, kDmolPix), (a)->N);                                       \
    if ((a)->Cw == 16) kern<16><<<grid, 256, 0, cg_stream(stream)>>>(__VA_ARGS__);           \
    else kern<32><<<grid, 256, 0, cg_stream(stream)>>>(__VA_ARGS__);                         \
  } while (0)

extern "C" int cg_dmol_loss_fwd(const cg_dmol_args* a, void* stream) {
  CG_ARCH_GUARD();
  DMOL_CHECK(a, "cg_dmol_loss_fwd");
  DMOL_LAUNCH(dmol_fwd_kernel, a, *a);
  CG_LAUNCH_CHECK("cg_dmol_loss_fwd");
  return CG_OK;
}

extern "C" int cg_dmol_loss_bwd(const cg_dmol_args* a, void* stream) {
  CG_ARCH_GUARD();
  DMOL_CHECK(a, "cg_dmol_loss_bwd");
  CG_REQUIRE(a->dh != nullptr && a->dw != nullptr && a->db != nullptr, "cg_dmol_loss_bwd: null gradient buffers");
  {
    dim3 grid(cg_ceil_div((a)->HW, kDmolPix * kDmolBwdIters), (a)->N);
    if ((a)->Cw == 16) dmol_bwd_kernel<16><<<grid, 256, 0, cg_stream(stream)>>>(*a);
    else dmol_bwd_kernel<32><<<grid, 256, 0, cg_stream(stream)>>>(*a);
  }
  CG_LAUNCH_CHECK("cg_dmol_loss_bwd");
  return CG_OK;
}

extern "C" int cg_dmol_predict(const cg_dmol_args* a, int32_t mode, const float* u_gumbel, const float* u_logistic,
                               float log_t, float* x_out, float* scale_out, void* stream) {
  CG_ARCH_GUARD();
  DMOL_CHECK(a, "cg_dmol_predict");
  CG_REQUIRE(((mode >= 0 && mode <= 2) || (mode > 10 && mode < 20)) && (mode != 2 || (u_gumbel && u_logistic)),
             "cg_dmol_predict: mode %d", mode);
  DMOL_LAUNCH(dmol_predict_kernel, a, *a, mode, u_gumbel, u_logistic, log_t, x_out, scale_out);
  CG_LAUNCH_CHECK("cg_dmol_predict");
  return CG_OK;
}
