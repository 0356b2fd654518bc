// Resampling, latent sample+KL (forward/backward), mediator mixture and counterfactual combine.
#include "cg_common.cuh"

namespace {

// All bf16 tensors here are channel-octet planar (N, C/8, H, W, 8): one thread handles one 16-byte octet of one
// pixel, consecutive threads walk consecutive pixels of a plane -> fully coalesced.
// ------------------------------------------------------------------ avg-pool / nearest upsample
// grid = (pixels of a plane / 256, channel octets, samples): no 64-bit index arithmetic per element
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C8, int d,
                                   long long x_ns, long long y_ns, int Po) {
  const int Ho = H / d, Wo = W / d;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Po * Po) return;
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int ho = p / Po, wo = p - ho * Po;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ho < Ho && wo < Wo) {
    const bf16* xp = x + n * x_ns + (long long)c8 * H * W * 8;
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        float f[8];
        cg_unpack8(__ldg(reinterpret_cast<const uint4*>(xp + ((ho * d + a) * W + wo * d + b) * 8)), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    const float inv = 1.0f / (d * d);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
  }
  *reinterpret_cast<uint4*>(y + n * y_ns + ((long long)c8 * Po * Po + p) * 8) = cg_pack8(acc);
}

__global__ void avgpool_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W, int C8,
                                   int d, long long dy_ns, long long dx_ns, int Po, int accumulate) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int h = p / W, w = p - h * W;
  const float inv = 1.0f / (d * d);
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // F.avg_pool2d floors: rows/columns beyond (H/d)*d belong to no window (8x8 map pooled by 7, src/vae.py:79-83)
  if (h / d < H / d && w / d < W / d)
    cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + n * dy_ns + ((long long)c8 * Po * Po + (h / d) * Po + w / d) * 8)), f);
  bf16* o = dx + n * dx_ns + ((long long)c8 * H * W + p) * 8;
  float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (accumulate) cg_unpack8(*reinterpret_cast<const uint4*>(o), g);
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] += f[k] * inv;
  *reinterpret_cast<uint4*>(o) = cg_pack8(g);
}

// the common case (H, W multiples of d, plain write): one thread per POOLED pixel octet scales its gradient once and
// writes the d x d window -- a quarter of the threads, no per-output division, d consecutive 16-byte stores per row
__global__ void avgpool_bwd_exact_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int H, int W, int d,
                                         long long dy_ns, long long dx_ns, int Po) {
  const int Ho = H / d, Wo = W / d;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Wo) return;
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int ho = p / Wo, wo = p - ho * Wo;
  float f[8];
  cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + n * dy_ns + ((long long)c8 * Po * Po + ho * Po + wo) * 8)), f);
  const float inv = 1.0f / (d * d);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] *= inv;
  const uint4 v = cg_pack8(f);
  bf16* o = dx + n * dx_ns + ((long long)c8 * H * W + (long long)(ho * d) * W + wo * d) * 8;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) *reinterpret_cast<uint4*>(o + ((long long)a * W + b) * 8) = v;
}

// blockIdx.z = chunk of kUpN samples: the learned bias (fp32, 32 bytes per output octet -- twice the output itself) is
// read once per chunk instead of once per sample
constexpr int kUpN = 8;
__global__ void upsample_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ bias, bf16* __restrict__ y,
                                    int N, int Hi, int Ho, int C, long long x_ns, long long y_ns) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Ho) return;
  const int c8 = blockIdx.y;
  const int n0 = blockIdx.z * kUpN, n1 = min(N, n0 + kUpN);
  const int h = p / Ho, w = p - h * Ho;
  const int hs = (h * Hi) / Ho, ws = (w * Hi) / Ho;
  float b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (bias != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c8 * 8 + k;
      if (c < C) b[k] = __ldg(bias + (long long)c * Ho * Ho + p);
    }
  }
  const bf16* xp = x + ((long long)c8 * Hi * Hi + hs * Hi + ws) * 8;
  bf16* yp = y + ((long long)c8 * Ho * Ho + p) * 8;
#pragma unroll 4
  for (int n = n0; n < n1; ++n) {
    float f[8];
    cg_unpack8(__ldg(reinterpret_cast<const uint4*>(xp + n * x_ns)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += b[k];
    *reinterpret_cast<uint4*>(yp + n * y_ns) = cg_pack8(f);
  }
}

__global__ void upsample_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int Hi, int Ho, int C8,
                                    long long dy_ns, long long dx_ns, int accumulate) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Hi * Hi) return;
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int h = p / Hi, w = p - h * Hi;
  // destination rows/cols whose nearest source is (h, w): floor(ho*Hi/Ho) == h
  const int h0 = (h * Ho + Hi - 1) / Hi, h1 = ((h + 1) * Ho + Hi - 1) / Hi;
  const int w0 = (w * Ho + Hi - 1) / Hi, w1 = ((w + 1) * Ho + Hi - 1) / Hi;
  float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bf16* o = dx + n * dx_ns + ((long long)c8 * Hi * Hi + p) * 8;
  if (accumulate) cg_unpack8(*reinterpret_cast<const uint4*>(o), g);
  const bf16* dp = dy + n * dy_ns + (long long)c8 * Ho * Ho * 8;
  for (int a = h0; a < h1; ++a)
    for (int b = w0; b < w1; ++b) {
      float f[8];
      cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dp + (a * Ho + b) * 8)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += f[k];
    }
  *reinterpret_cast<uint4*>(o) = cg_pack8(g);
}

// dbias[c,h,w] += sum_n dy[n,c,h,w]; blockIdx.z = chunk of 16 samples, partial sums meet through atomics
__global__ void upsample_dbias_kernel(const bf16* __restrict__ dy, float* __restrict__ dbias, int N, int Ho, int C,
                                      long long dy_ns) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ho * Ho) return;
  const int c8 = blockIdx.y;
  const int n0 = blockIdx.z * 16, n1 = min(N, n0 + 16);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = n0; n < n1; ++n) {
    float f[8];
    cg_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + n * dy_ns + ((long long)c8 * Ho * Ho + p) * 8)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (c8 * 8 + k < C) atomicAdd(dbias + (long long)(c8 * 8 + k) * Ho * Ho + p, s[k]);
}

// ------------------------------------------------------------------ Philox4x32-10 -> N(0,1)
// One call = four standard normals (two Box-Muller pairs).  Noise element (n, c, pixel) of a latent block is
// output (c & 3) of counter  offset + 2 * ((n*HW + pixel)*2 + c/8) + ((c & 7) >> 2)  -- the same in every
// kernel below, so backward passes regenerate exactly the forward noise without storing it.
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t idx, float* out) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float s = 2.3283064365386963e-10f;
  const float r0 = sqrtf(-2.0f * __logf(((float)c0 + 0.5f) * s)), r1 = sqrtf(-2.0f * __logf(((float)c2 + 0.5f) * s));
  float sn0, cs0, sn1, cs1;
  __sincosf(6.283185307179586f * ((float)c1 + 0.5f) * s, &sn0, &cs0);
  __sincosf(6.283185307179586f * ((float)c3 + 0.5f) * s, &sn1, &cs1);
  out[0] = r0 * cs0; out[1] = r0 * sn0; out[2] = r1 * cs1; out[3] = r1 * sn1;
}
// eight normals of work item (pixel, channel octet): channels [8*oc, 8*oc + 8)
__device__ __forceinline__ void philox_normal8(uint64_t seed, uint64_t offset, long long item, float* out) {
  philox_normal4(seed, offset + 2ull * (uint64_t)item, out);
  philox_normal4(seed, offset + 2ull * (uint64_t)item + 1ull, out + 4);
}
// single element (slow path: explicit-layout kernels that need one channel at a time)
__device__ __forceinline__ float philox_normal_at(uint64_t seed, uint64_t offset, long long pix, int c) {
  float v[4];
  philox_normal4(seed, offset + 2ull * (uint64_t)(pix * 2 + (c >> 3)) + (uint64_t)((c & 7) >> 2), v);
  return v[c & 3];
}

// ------------------------------------------------------------------ latent block forward
// One block = up to 64 pixels of one image x zdim(16) channels.  eps / z_f32 are NCHW (reference
// layout), q / p / z_bf16 are NHWC, so the tile goes through shared memory to keep both coalesced.
constexpr int kLatPix = 64;
__global__ void __launch_bounds__(256) latent_fwd_kernel(const cg_latent_args a) {
  __shared__ float s_eps[16][kLatPix + 1];
  __shared__ float s_z[16][kLatPix + 1];
  __shared__ float s_red[8];
  __shared__ float s_ch[16];
  __shared__ float s_kl[16][kLatPix + 1];
  if (threadIdx.x < 16) s_ch[threadIdx.x] = 0.f;
  const int n = blockIdx.y;
  const int hw0 = blockIdx.x * kLatPix;
  const int npx = min(kLatPix, a.HW - hw0);
  const int zd = a.zdim;  // 16
  const int tid = threadIdx.x;
  const uint64_t seed = a.seed + (a.seed_dev != nullptr ? *a.seed_dev : 0ull);
  if (a.mode != 2) {
    for (int e = tid; e < zd * kLatPix; e += 256) {
      int c = e / kLatPix, px = e - c * kLatPix;
      float v = 0.f;
      if (px < npx) {
        long long gi = ((long long)n * zd + c) * a.HW + hw0 + px;
        v = a.eps != nullptr ? a.eps[gi] : philox_normal_at(seed, a.offset, (long long)n * a.HW + hw0 + px, c);
        if (a.eps_out != nullptr) a.eps_out[gi] = v;
      }
      s_eps[c][px] = v;
    }
  }
  __syncthreads();
  float kl_acc = 0.f;
  bf16* zb = reinterpret_cast<bf16*>(a.z_bf16);
  for (int e = tid; e < 2 * kLatPix; e += 256) {  // work item = (pixel, channel octet)
    const int px = e >> 1, oc = e & 1;
    if (px >= npx) continue;
    const long long pix = (long long)n * a.HW + hw0 + px;
    float zv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = oc * 8 + k;
      const float p_loc = a.p[pix * a.p_ld + c], p_ls = a.p[pix * a.p_ld + zd + c] + a.log_t;
      float z;
      if (a.mode == 0) {
        const float q_loc = a.q[pix * a.q_ld + c], q_ls = a.q[pix * a.q_ld + zd + c] + a.log_t;
        z = q_loc + __expf(q_ls) * s_eps[c][px];
        // src/vae.py:14-25 (same term order)
        const float eq = __expf(q_ls), ep = __expf(p_ls), dm = q_loc - p_loc;
        const float kl = -0.5f + p_ls - q_ls + 0.5f * (eq * eq + dm * dm) / (ep * ep);
        kl_acc += kl;
        s_kl[c][px] = kl;
        if (a.kl_ch != nullptr) atomicAdd(&s_ch[c], kl);  // free-bits statistics (rare path): shared-memory atomics
      } else if (a.mode == 1) {
        z = p_loc + __expf(p_ls) * s_eps[c][px];
      } else {
        z = p_loc;
      }
      zv[k] = z;
      s_z[c][px] = z;
    }
    *reinterpret_cast<uint4*>(zb + n * a.z_ns + ((long long)oc * a.HW + hw0 + px) * 8) = cg_pack8(zv);
  }
  if (a.kl_out != nullptr && a.mode == 0) {
    kl_acc = cg_warp_sum(kl_acc);
    if ((tid & 31) == 0) s_red[tid >> 5] = kl_acc;
  }
  __syncthreads();
  if (a.kl_out != nullptr && a.mode == 0 && tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    atomicAdd(a.kl_out + n, s);
  }
  if (a.kl_ch != nullptr && a.mode == 0 && tid < 16) atomicAdd(a.kl_ch + tid, s_ch[tid]);
  if (a.kl_elem != nullptr && a.mode == 0) {
    for (int e = tid; e < zd * kLatPix; e += 256) {
      int c = e / kLatPix, px = e - c * kLatPix;
      if (px < npx) a.kl_elem[((long long)n * zd + c) * a.HW + hw0 + px] = s_kl[c][px];
    }
  }
  if (a.z_f32 != nullptr) {
    for (int e = tid; e < zd * kLatPix; e += 256) {
      int c = e / kLatPix, px = e - c * kLatPix;
      if (px < npx) a.z_f32[((long long)n * zd + c) * a.HW + hw0 + px] = s_z[c][px];
    }
  }
}

// gradients of g_kl*KL + <dz,z> wrt (q_loc,q_ls) and (p_loc,p_ls), written bf16 as channels [0,2*zd)
__global__ void __launch_bounds__(256) latent_bwd_kernel(const cg_latent_bwd_args a) {
  __shared__ float s_eps[16][kLatPix + 1];
  const int n = blockIdx.y;
  const int hw0 = blockIdx.x * kLatPix;
  const int npx = min(kLatPix, a.HW - hw0);
  const int zd = a.zdim;
  const int tid = threadIdx.x;
  const uint64_t seed = a.seed + (a.seed_dev != nullptr ? *a.seed_dev : 0ull);
  const float g_kl = a.g_kl * (a.g_kl_dev != nullptr ? __ldg(a.g_kl_dev) : 1.0f);
  if (a.mode != 2) {
    for (int e = tid; e < zd * kLatPix; e += 256) {
      int c = e / kLatPix, px = e - c * kLatPix;
      float v = 0.f;
      if (px < npx) {
        long long gi = ((long long)n * zd + c) * a.HW + hw0 + px;
        v = a.eps != nullptr ? a.eps[gi] : philox_normal_at(seed, a.offset, (long long)n * a.HW + hw0 + px, c);
      }
      s_eps[c][px] = v;
    }
  }
  __syncthreads();
  bf16* dq = reinterpret_cast<bf16*>(a.dq);
  bf16* dp = reinterpret_cast<bf16*>(a.dp);
  const bf16* dzp = reinterpret_cast<const bf16*>(a.dz);
  for (int e = tid; e < 2 * kLatPix; e += 256) {  // work item = (pixel, channel octet)
    const int px = e >> 1, oc = e & 1;
    if (px >= npx) continue;
    const long long hw = hw0 + px;
    const long long pix = (long long)n * a.HW + hw;
    float dz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (dzp != nullptr) cg_unpack8(*reinterpret_cast<const uint4*>(dzp + n * a.dz_ns + ((long long)oc * a.HW + hw) * 8), dz);
    float g_qloc[8], g_qls[8], g_ploc[8], g_pls[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = oc * 8 + k;
      const float p_loc = a.p[pix * a.p_ld + c], p_ls = a.p[pix * a.p_ld + zd + c] + a.log_t;
      if (a.mode == 0) {
        const float q_loc = a.q[pix * a.q_ld + c], q_ls = a.q[pix * a.q_ld + zd + c] + a.log_t;
        const float eq = __expf(q_ls), ivp = __expf(-2.0f * p_ls), dm = q_loc - p_loc;
        const float gk = a.kl_gate != nullptr ? g_kl * __ldg(a.kl_gate + c) : g_kl;
        g_qloc[k] = gk * dm * ivp + dz[k];
        g_qls[k] = gk * (eq * eq * ivp - 1.0f) + dz[k] * eq * s_eps[c][px];
        g_ploc[k] = -gk * dm * ivp;
        g_pls[k] = gk * (1.0f - (eq * eq + dm * dm) * ivp);
      } else if (a.mode == 1) {
        g_ploc[k] = dz[k];
        g_pls[k] = dz[k] * __expf(p_ls) * s_eps[c][px];
      } else {
        g_ploc[k] = dz[k];
        g_pls[k] = 0.f;
      }
    }
    // channels [0,16) = loc -> octets 0,1 ; channels [16,32) = logscale -> octets 2,3
    if (a.mode == 0) {
      bf16* q0 = dq + n * a.dq_ns + hw * 8;
      *reinterpret_cast<uint4*>(q0 + (long long)oc * a.HW * 8) = cg_pack8(g_qloc);
      *reinterpret_cast<uint4*>(q0 + (long long)(2 + oc) * a.HW * 8) = cg_pack8(g_qls);
    }
    bf16* p0 = dp + n * a.dp_ns + hw * 8;
    *reinterpret_cast<uint4*>(p0 + (long long)oc * a.HW * 8) = cg_pack8(g_ploc);
    *reinterpret_cast<uint4*>(p0 + (long long)(2 + oc) * a.HW * 8) = cg_pack8(g_pls);
  }
}

// ------------------------------------------------------------------ streaming variants (in-kernel noise)
// Training / sampling with Philox noise needs no NCHW staging: one thread = (pixel, channel octet), 128-bit loads of
// the fp32 statistics rows, noise generated in registers, one 16-byte bf16 store, block-reduced KL.
__global__ void __launch_bounds__(256) latent_fwd_stream_kernel(const cg_latent_args a) {
  __shared__ float s_red[8];
  __shared__ float s_ch[16];
  if (threadIdx.x < 16) s_ch[threadIdx.x] = 0.f;
  float klc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // this thread's KL per channel of its octet (free-bits statistics)
  const int n = blockIdx.y;
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;  // (pixel, octet) inside the sample
  const bool live = item < 2ll * a.HW;
  float kl_acc = 0.f;
  if (live) {
    const long long hw = item >> 1;
    const int oc = (int)(item & 1);
    const long long pix = (long long)n * a.HW + hw;
    const uint64_t seed = a.seed + (a.seed_dev != nullptr ? *a.seed_dev : 0ull);
    float pl[8], ps[8], zv[8], e[8];
    const float4* pp = reinterpret_cast<const float4*>(a.p + pix * a.p_ld + oc * 8);
    const float4* pq = reinterpret_cast<const float4*>(a.p + pix * a.p_ld + 16 + oc * 8);
    *reinterpret_cast<float4*>(pl) = pp[0]; *reinterpret_cast<float4*>(pl + 4) = pp[1];
    *reinterpret_cast<float4*>(ps) = pq[0]; *reinterpret_cast<float4*>(ps + 4) = pq[1];
    if (a.mode != 2) philox_normal8(seed, a.offset, pix * 2 + oc, e);
    if (a.mode == 0) {
      float ql[8], qs[8];
      const float4* qp = reinterpret_cast<const float4*>(a.q + pix * a.q_ld + oc * 8);
      const float4* qq = reinterpret_cast<const float4*>(a.q + pix * a.q_ld + 16 + oc * 8);
      *reinterpret_cast<float4*>(ql) = qp[0]; *reinterpret_cast<float4*>(ql + 4) = qp[1];
      *reinterpret_cast<float4*>(qs) = qq[0]; *reinterpret_cast<float4*>(qs + 4) = qq[1];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float q_ls = qs[k] + a.log_t, p_ls = ps[k] + a.log_t;
        const float eq = __expf(q_ls), ep = __expf(p_ls), dm = ql[k] - pl[k];
        zv[k] = ql[k] + eq * e[k];
        klc[k] = -0.5f + p_ls - q_ls + 0.5f * (eq * eq + dm * dm) / (ep * ep);  // src/vae.py:14-25
        kl_acc += klc[k];
      }
    } else if (a.mode == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) zv[k] = pl[k] + __expf(ps[k] + a.log_t) * e[k];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) zv[k] = pl[k];
    }
    bf16* zb = reinterpret_cast<bf16*>(a.z_bf16);
    *reinterpret_cast<uint4*>(zb + n * a.z_ns + ((long long)oc * a.HW + hw) * 8) = cg_pack8(zv);
    if (a.eps_out != nullptr && a.mode != 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a.eps_out[((long long)n * 16 + oc * 8 + k) * a.HW + hw] = e[k];
    }
  }
  if (a.kl_out != nullptr && a.mode == 0) {
    kl_acc = cg_warp_sum(kl_acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = kl_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += s_red[i];
      atomicAdd(a.kl_out + n, s);
    }
  }
  if (a.kl_ch != nullptr && a.mode == 0) {  // kl_free_bits statistics: lanes of equal parity hold the same octet
    __syncthreads();                         // s_ch zeroed
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = klc[k];
#pragma unroll
      for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) < 2) atomicAdd(&s_ch[(threadIdx.x & 1) * 8 + k], v);
    }
    __syncthreads();
    if (threadIdx.x < 16) atomicAdd(a.kl_ch + threadIdx.x, s_ch[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(256) latent_bwd_stream_kernel(const cg_latent_bwd_args a) {
  const int n = blockIdx.y;
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  if (item >= 2ll * a.HW) return;
  const long long hw = item >> 1;
  const int oc = (int)(item & 1);
  const long long pix = (long long)n * a.HW + hw;
  const uint64_t seed = a.seed + (a.seed_dev != nullptr ? *a.seed_dev : 0ull);
  const float g_kl = a.g_kl * (a.g_kl_dev != nullptr ? __ldg(a.g_kl_dev) : 1.0f);
  float pl[8], ps[8], e[8], dz[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float4* pp = reinterpret_cast<const float4*>(a.p + pix * a.p_ld + oc * 8);
  const float4* pq = reinterpret_cast<const float4*>(a.p + pix * a.p_ld + 16 + oc * 8);
  *reinterpret_cast<float4*>(pl) = pp[0]; *reinterpret_cast<float4*>(pl + 4) = pp[1];
  *reinterpret_cast<float4*>(ps) = pq[0]; *reinterpret_cast<float4*>(ps + 4) = pq[1];
  if (a.dz != nullptr)
    cg_unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.dz) + n * a.dz_ns + ((long long)oc * a.HW + hw) * 8), dz);
  if (a.mode != 2) philox_normal8(seed, a.offset, pix * 2 + oc, e);
  float g_ploc[8], g_pls[8];
  if (a.mode == 0) {
    float ql[8], qs[8], g_qloc[8], g_qls[8];
    const float4* qp = reinterpret_cast<const float4*>(a.q + pix * a.q_ld + oc * 8);
    const float4* qq = reinterpret_cast<const float4*>(a.q + pix * a.q_ld + 16 + oc * 8);
    *reinterpret_cast<float4*>(ql) = qp[0]; *reinterpret_cast<float4*>(ql + 4) = qp[1];
    *reinterpret_cast<float4*>(qs) = qq[0]; *reinterpret_cast<float4*>(qs + 4) = qq[1];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float eq = __expf(qs[k] + a.log_t), ivp = __expf(-2.0f * (ps[k] + a.log_t)), dm = ql[k] - pl[k];
      const float gk = a.kl_gate != nullptr ? g_kl * __ldg(a.kl_gate + oc * 8 + k) : g_kl;
      g_qloc[k] = gk * dm * ivp + dz[k];
      g_qls[k] = gk * (eq * eq * ivp - 1.0f) + dz[k] * eq * e[k];
      g_ploc[k] = -gk * dm * ivp;
      g_pls[k] = gk * (1.0f - (eq * eq + dm * dm) * ivp);
    }
    bf16* q0 = reinterpret_cast<bf16*>(a.dq) + n * a.dq_ns + hw * 8;
    *reinterpret_cast<uint4*>(q0 + (long long)oc * a.HW * 8) = cg_pack8(g_qloc);
    *reinterpret_cast<uint4*>(q0 + (long long)(2 + oc) * a.HW * 8) = cg_pack8(g_qls);
  } else if (a.mode == 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { g_ploc[k] = dz[k]; g_pls[k] = dz[k] * __expf(ps[k] + a.log_t) * e[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) { g_ploc[k] = dz[k]; g_pls[k] = 0.f; }
  }
  bf16* p0 = reinterpret_cast<bf16*>(a.dp) + n * a.dp_ns + hw * 8;
  *reinterpret_cast<uint4*>(p0 + (long long)oc * a.HW * 8) = cg_pack8(g_ploc);
  *reinterpret_cast<uint4*>(p0 + (long long)(2 + oc) * a.HW * 8) = cg_pack8(g_pls);
}

// kl_free_bits (src/vae.py:443-449): one block; thread i = (stochastic block, latent channel)
__global__ void free_bits_kernel(const float* __restrict__ kl_ch, int n, float fb, float inv_batch, float* __restrict__ gate,
                                 float* __restrict__ kl_row, int N) {
  __shared__ float s_part[32];
  __shared__ float s_tot;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = kl_ch[i] * inv_batch;
    gate[i] = v > fb ? 1.0f : 0.0f;
    acc += fmaxf(fb, v);
  }
  acc = cg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_part[i];
    s_tot = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) kl_row[i] = s_tot;
}

__global__ void latent_mix_kernel(const float* __restrict__ z, const float* __restrict__ q_loc,
                                  const float* __restrict__ q_ls, const float* __restrict__ p_loc,
                                  const float* __restrict__ p_ls, float* __restrict__ out, long long n, float alpha,
                                  float t, int has_t) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // src/vae.py:487-513
    float qs = __expf(q_ls[i]);
    float u = (z[i] - q_loc[i]) / qs;
    float pv = __expf(p_ls[i]);
    pv *= pv;
    float r_loc = alpha * q_loc[i] + (1.0f - alpha) * p_loc[i];
    float r_sc = sqrtf(alpha * alpha * qs * qs + (1.0f - alpha) * (1.0f - alpha) * pv);
    if (has_t) r_sc *= t;
    out[i] = r_loc + r_sc * u;
  }
}

__global__ void cf_combine_kernel(const float* __restrict__ x, const float* __restrict__ rec_loc,
                                  const float* __restrict__ rec_scale, const float* __restrict__ cf_loc,
                                  const float* __restrict__ cf_scale, float* __restrict__ cf_x, float* __restrict__ sum,
                                  float* __restrict__ sum2, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // src/pgm/dscm.py:55-56
    float u = (x[i] - rec_loc[i]) / fmaxf(rec_scale[i], 1e-12f);
    float v = fminf(fmaxf(cf_loc[i] + cf_scale[i] * u, -1.0f), 1.0f);
    cf_x[i] = v;
    if (sum != nullptr) sum[i] += v;      // src/pgm/dscm.py:59
    if (sum2 != nullptr) sum2[i] += v * v;  // src/pgm/dscm.py:61
  }
}

// backward of cf_combine_kernel: d cf_x -> d cf_loc, d cf_scale, d rec_loc, d rec_scale (torch.clamp masks inclusive)
__global__ void cf_combine_bwd_kernel(const float* __restrict__ x, const float* __restrict__ rec_loc,
                                      const float* __restrict__ rec_scale, const float* __restrict__ cf_loc,
                                      const float* __restrict__ cf_scale, const float* __restrict__ dcf,
                                      float* __restrict__ d_rec_loc, float* __restrict__ d_rec_scale,
                                      float* __restrict__ d_cf_loc, float* __restrict__ d_cf_scale, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = fmaxf(rec_scale[i], 1e-12f);
    const float u = (x[i] - rec_loc[i]) / s;
    const float pre = cf_loc[i] + cf_scale[i] * u;
    const float g = (pre >= -1.0f && pre <= 1.0f) ? dcf[i] : 0.f;
    const float du = g * cf_scale[i];
    d_cf_loc[i] = g;
    d_cf_scale[i] = g * u;
    d_rec_loc[i] = -du / s;
    d_rec_scale[i] = rec_scale[i] >= 1e-12f ? -du * u / s : 0.f;
  }
}

inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int cg_avgpool_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t d,
                              int64_t x_ld, int64_t y_ld, int32_t pad_to, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(d >= 1 && d <= H && H == W, "cg_avgpool_fwd: H=%d W=%d d=%d", H, W, d);
  CG_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0, "cg_avgpool_fwd: C/ld multiples of 8");
  const int Po = pad_to > 0 ? pad_to : H / d;
  CG_REQUIRE(Po >= H / d, "cg_avgpool_fwd: pad_to %d < %d", Po, H / d);
  avgpool_fwd_kernel<<<dim3(cg_ceil_div(Po * Po, 256), C / 8, N), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(x), reinterpret_cast<bf16*>(y), N, H, W, C / 8, d, x_ld, y_ld, Po);
  CG_LAUNCH_CHECK("cg_avgpool_fwd");
  return CG_OK;
}

extern "C" int cg_avgpool_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t d,
                              int64_t dy_ld, int64_t dx_ld, int32_t pad_to, int32_t accumulate, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(d >= 1 && d <= H && H == W, "cg_avgpool_bwd: H=%d W=%d d=%d", H, W, d);
  CG_REQUIRE(C % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0, "cg_avgpool_bwd: C/ld multiples of 8");
  const int Po = pad_to > 0 ? pad_to : H / d;
  if (!accumulate && H % d == 0 && W % d == 0 && Po == H / d) {  // (a padded gradient plane keeps the general kernel)
    avgpool_bwd_exact_kernel<<<dim3(cg_ceil_div((H / d) * (W / d), 256), C / 8, N), 256, 0, cg_stream(stream)>>>(
        reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dx), H, W, d, dy_ld, dx_ld, Po);
    CG_LAUNCH_CHECK("cg_avgpool_bwd");
    return CG_OK;
  }
  avgpool_bwd_kernel<<<dim3(cg_ceil_div(H * W, 256), C / 8, N), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dx), N, H, W, C / 8, d, dy_ld, dx_ld, Po, accumulate);
  CG_LAUNCH_CHECK("cg_avgpool_bwd");
  return CG_OK;
}

extern "C" int cg_upsample_fwd(const void* x, const float* bias, void* y, int32_t N, int32_t Hi, int32_t Ho, int32_t C,
                               int64_t x_ld, int64_t y_ld, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(Ho >= Hi && x_ld % 8 == 0 && y_ld % 8 == 0, "cg_upsample_fwd: Hi=%d Ho=%d", Hi, Ho);
  upsample_fwd_kernel<<<dim3(cg_ceil_div(Ho * Ho, 256), (C + 7) / 8, cg_ceil_div(N, kUpN)), 256, 0, cg_stream(stream)>>>(
      reinterpret_cast<const bf16*>(x), bias, reinterpret_cast<bf16*>(y), N, Hi, Ho, C, x_ld, y_ld);
  CG_LAUNCH_CHECK("cg_upsample_fwd");
  return CG_OK;
}

extern "C" int cg_upsample_bwd(const void* dy, void* dx, float* dbias, int32_t N, int32_t Hi, int32_t Ho, int32_t C,
                               int64_t dy_ld, int64_t dx_ld, int32_t accumulate, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(Ho >= Hi && dy_ld % 8 == 0 && dx_ld % 8 == 0, "cg_upsample_bwd: Hi=%d Ho=%d", Hi, Ho);
  if (dx != nullptr) {
    upsample_bwd_kernel<<<dim3(cg_ceil_div(Hi * Hi, 256), (C + 7) / 8, N), 256, 0, cg_stream(stream)>>>(
        reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dx), N, Hi, Ho, (C + 7) / 8, dy_ld, dx_ld, accumulate);
    CG_LAUNCH_CHECK("cg_upsample_bwd");
  }
  if (dbias != nullptr) {
    upsample_dbias_kernel<<<dim3(cg_ceil_div(Ho * Ho, 256), (C + 7) / 8, cg_ceil_div(N, 16)), 256, 0, cg_stream(stream)>>>(
        reinterpret_cast<const bf16*>(dy), dbias, N, Ho, C, dy_ld);
    CG_LAUNCH_CHECK("cg_upsample_bwd(dbias)");
  }
  return CG_OK;
}

extern "C" int cg_latent_fwd(const cg_latent_args* a, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(a != nullptr && a->zdim == 16, "cg_latent_fwd: zdim must be 16");
  CG_REQUIRE(a->p != nullptr && a->z_bf16 != nullptr && (a->mode != 0 || a->q != nullptr), "cg_latent_fwd: null operand");
  const bool rows16 = (((uintptr_t)a->p & 15) == 0) && a->p_ld % 4 == 0 &&
                      (a->q == nullptr || ((((uintptr_t)a->q & 15) == 0) && a->q_ld % 4 == 0));
  if (a->eps == nullptr && a->z_f32 == nullptr && a->kl_elem == nullptr && rows16) {  // in-kernel noise: streaming kernel
    dim3 g2(cg_ceil_div(2ll * a->HW, 256), a->N);
    latent_fwd_stream_kernel<<<g2, 256, 0, cg_stream(stream)>>>(*a);
    CG_LAUNCH_CHECK("cg_latent_fwd");
    return CG_OK;
  }
  dim3 grid(cg_ceil_div(a->HW, kLatPix), a->N);
  latent_fwd_kernel<<<grid, 256, 0, cg_stream(stream)>>>(*a);
  CG_LAUNCH_CHECK("cg_latent_fwd");
  return CG_OK;
}

extern "C" int cg_latent_bwd(const cg_latent_bwd_args* a, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(a != nullptr && a->zdim == 16, "cg_latent_bwd: zdim must be 16");
  CG_REQUIRE(a->p != nullptr && a->dp != nullptr && (a->mode != 0 || (a->q != nullptr && a->dq != nullptr)),
             "cg_latent_bwd: null operand");
  const bool rows16 = (((uintptr_t)a->p & 15) == 0) && a->p_ld % 4 == 0 &&
                      (a->q == nullptr || ((((uintptr_t)a->q & 15) == 0) && a->q_ld % 4 == 0));
  if (a->eps == nullptr && rows16) {
    dim3 g2(cg_ceil_div(2ll * a->HW, 256), a->N);
    latent_bwd_stream_kernel<<<g2, 256, 0, cg_stream(stream)>>>(*a);
    CG_LAUNCH_CHECK("cg_latent_bwd");
    return CG_OK;
  }
  dim3 grid(cg_ceil_div(a->HW, kLatPix), a->N);
  latent_bwd_kernel<<<grid, 256, 0, cg_stream(stream)>>>(*a);
  CG_LAUNCH_CHECK("cg_latent_bwd");
  return CG_OK;
}

extern "C" int cg_free_bits(const float* kl_ch, int32_t nblk, float free_bits, float inv_batch, float* gate, float* kl_row,
                            int32_t N, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(kl_ch != nullptr && gate != nullptr && kl_row != nullptr && nblk > 0 && N > 0, "cg_free_bits: null / empty");
  free_bits_kernel<<<1, 256, 0, cg_stream(stream)>>>(kl_ch, nblk * 16, free_bits, inv_batch, gate, kl_row, N);
  CG_LAUNCH_CHECK("cg_free_bits");
  return CG_OK;
}

extern "C" int cg_latent_mix(const float* z, const float* q_loc, const float* q_ls, const float* p_loc,
                             const float* p_ls, float* out, int64_t n, float alpha, float t, int32_t has_t,
                             void* stream) {
  CG_ARCH_GUARD();
  latent_mix_kernel<<<grid_for(n, 256), 256, 0, cg_stream(stream)>>>(z, q_loc, q_ls, p_loc, p_ls, out, n, alpha, t, has_t);
  CG_LAUNCH_CHECK("cg_latent_mix");
  return CG_OK;
}

extern "C" int cg_cf_combine(const float* x, const float* rec_loc, const float* rec_scale, const float* cf_loc,
                             const float* cf_scale, float* cf_x, float* sum, float* sum2, int64_t n, void* stream) {
  CG_ARCH_GUARD();
  cf_combine_kernel<<<grid_for(n, 256), 256, 0, cg_stream(stream)>>>(x, rec_loc, rec_scale, cf_loc, cf_scale, cf_x, sum,
                                                                     sum2, n);
  CG_LAUNCH_CHECK("cg_cf_combine");
  return CG_OK;
}

extern "C" int cg_cf_combine_bwd(const float* x, const float* rec_loc, const float* rec_scale, const float* cf_loc,
                                 const float* cf_scale, const float* dcf, float* d_rec_loc, float* d_rec_scale,
                                 float* d_cf_loc, float* d_cf_scale, int64_t n, void* stream) {
  CG_ARCH_GUARD();
  cf_combine_bwd_kernel<<<grid_for(n, 256), 256, 0, cg_stream(stream)>>>(x, rec_loc, rec_scale, cf_loc, cf_scale, dcf,
                                                                         d_rec_loc, d_rec_scale, d_cf_loc, d_cf_scale, n);
  CG_LAUNCH_CHECK("cg_cf_combine_bwd");
  return CG_OK;
}
