// Optimiser tail on flat fp32 buffers, no host synchronisation:
//   global-norm clip (nn.utils.clip_grad_norm_, reference src/trainer.py:67), NaN / grad_skip test
//   (src/trainer.py:71-85), AdamW with linear LR warm-up (src/train_setup.py:42-53,
//   src/utils.py:32-36) and the inverse-decay EMA of src/utils.py:169-220.
#include "cg_common.cuh"

namespace {

// partials != nullptr: block b writes its partial sum to partials[b] (no atomics) and sumsq_final_kernel adds them in a
// fixed order -> bit-identical result on every data-parallel replica (their clip coefficients must agree exactly)
__global__ void sumsq_kernel(const float* __restrict__ g, float* __restrict__ out, long long n,
                             float* __restrict__ partials) {
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    float v = g[n4 * 4 + threadIdx.x];
    acc += v * v;
  }
  __shared__ float red[8];
  acc = cg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    if (partials != nullptr) partials[blockIdx.x] = s;
    else atomicAdd(out, s);
  }
}

__global__ void sumsq_final_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];  // fixed assignment, fixed tree below
  acc = cg_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    out[0] = s;
  }
}

// state[0]=adam step, [1]=ema calls, [2]=skipped count, [3]=skip flag of the current step
// dyn[0]=lr, [1]=1-b1^t, [2]=1-b2^t, [3]=ema decay, [4]=clip coefficient, [5]=grad norm
__global__ void optim_advance_kernel(int* __restrict__ state, float* __restrict__ dyn, const float* __restrict__ gsumsq,
                                     const float* __restrict__ loss_terms, float base_lr, int warmup, float beta1,
                                     float beta2, float grad_clip, float grad_skip, float grad_scale, float ema_beta,
                                     int ema_after) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float norm = sqrtf(gsumsq[0]) * grad_scale;
  bool bad = !(norm < grad_skip);  // also true for NaN
  if (loss_terms != nullptr) bad = bad || isnan(loss_terms[1]) || isnan(loss_terms[2]);
  dyn[5] = norm;
  dyn[4] = fminf(1.0f, grad_clip / (norm + 1e-6f)) * grad_scale;
  state[3] = bad ? 1 : 0;
  if (bad) {
    state[2] += 1;
    return;
  }
  // LambdaLR: the lr used by optimizer.step() number t (1-based) is base_lr * f(t-1), f = linear_warmup
  const int t = state[0] + 1;
  state[0] = t;
  const int it = t - 1;
  dyn[0] = base_lr * ((it > warmup || warmup <= 0) ? 1.0f : (float)it / (float)warmup);
  dyn[1] = 1.0f - powf(beta1, (float)t);
  dyn[2] = 1.0f - powf(beta2, (float)t);
  // EMA.update() (src/utils.py:169-193), s = calls so far: s <= update_after_step copies (decay 0), the first call
  // after that copies once more (`initted`), then decay = 1 - 1/(1 + k) with k = s - update_after_step --
  // get_current_decay() reads the step AFTER update() incremented it -- clamped to [0, beta].
  // Pinned against the reference's own trajectory: tests/golden/ema_schedule.npz, oracle ema_decay().
  const int s = state[1];
  state[1] = s + 1;
  float decay = 0.f;
  if (s > ema_after + 1) {
    const float k = (float)(s - ema_after);
    decay = fminf(fmaxf(1.0f - 1.0f / (1.0f + k), 0.f), ema_beta);
  }
  dyn[3] = decay;
}

__global__ void adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, long long n,
                                 const int* __restrict__ state, const float* __restrict__ dyn, float beta1, float beta2,
                                 float eps, float wd) {
  if (state[3] != 0) return;  // update skipped (src/trainer.py:71-85)
  const float lr = dyn[0], bc1 = dyn[1], bc2s = sqrtf(dyn[2]), decay = dyn[3], clip = dyn[4];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    pi -= (lr / bc1) * mi / (sqrtf(vi) / bc2s + eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (ema != nullptr) ema[i] = decay * ema[i] + (1.0f - decay) * pi;
  }
}

}  // namespace

extern "C" int cg_sumsq(const float* g, float* out, int64_t n, float* scratch, int32_t scratch_n, void* stream) {
  CG_ARCH_GUARD();
  CG_REQUIRE(((uintptr_t)g & 15) == 0, "cg_sumsq: unaligned");
  int blocks = (int)((n + 2047) / 2048);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  if (scratch != nullptr) {
    CG_REQUIRE(scratch_n >= 148 * 4, "cg_sumsq: scratch needs %d floats", 148 * 4);
    sumsq_kernel<<<blocks, 256, 0, cg_stream(stream)>>>(g, out, n, scratch);
    sumsq_final_kernel<<<1, 256, 0, cg_stream(stream)>>>(scratch, blocks, out);
  } else {
    sumsq_kernel<<<blocks, 256, 0, cg_stream(stream)>>>(g, out, n, nullptr);
  }
  CG_LAUNCH_CHECK("cg_sumsq");
  return CG_OK;
}

extern "C" int cg_optim_advance(int32_t* state, float* dyn, const float* gsumsq, const float* loss_terms, float base_lr,
                                int32_t warmup, float beta1, float beta2, float grad_clip, float grad_skip,
                                float grad_scale, float ema_beta, int32_t ema_after, void* stream) {
  CG_ARCH_GUARD();
  optim_advance_kernel<<<1, 32, 0, cg_stream(stream)>>>(state, dyn, gsumsq, loss_terms, base_lr, warmup, beta1, beta2,
                                                        grad_clip, grad_skip, grad_scale, ema_beta, ema_after);
  CG_LAUNCH_CHECK("cg_optim_advance");
  return CG_OK;
}

extern "C" int cg_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                                 const int32_t* state, const float* dyn, float beta1, float beta2, float eps, float wd,
                                 void* stream) {
  CG_ARCH_GUARD();
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adamw_ema_kernel<<<blocks, 256, 0, cg_stream(stream)>>>(p, g, m, v, ema, n, state, dyn, beta1, beta2, eps, wd);
  CG_LAUNCH_CHECK("cg_adamw_ema_step");
  return CG_OK;
}
