// Shared device/host helpers for the causalgen_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "causalgen_b200.h"

// ---- host-side error plumbing (api.cu) ------------------------------------------------
void cg_set_error(const char* fmt, ...);
int cg_require_sm100();  // CG_OK or CG_ERR_ARCH

#define CG_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      cg_set_error(__VA_ARGS__);     \
      return CG_ERR_ARG;             \
    }                                \
  } while (0)

#define CG_ARCH_GUARD()                        \
  do {                                         \
    int _s = cg_require_sm100();               \
    if (_s != CG_OK) return _s;                \
  } while (0)

#define CG_LAUNCH_CHECK(name)                                                   \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      cg_set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));      \
      return CG_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

// TMA tensor map (CUtensorMap, 128 B) over a planar bf16 view; see conv_tc.cu
int cg_make_planar_map(void* map_out, const void* ptr, long long ns, int N, int H, int W, int C8, int flat, int box_w8,
                       int box_h, int box_c8);

// mma.sync weight-gradient path for small-channel 3x3 problems (wgrad_mma.cu); *handled = 0 -> use tcgen05
int cg_wgrad_mma_try(const cg_wgrad_args* a, void* stream, int* handled);

static inline cudaStream_t cg_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cg_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- optional in-kernel timeline (debug builds only: make EXTRA=-DCG_TIMELINE) ---------
extern unsigned long long* cg_tl_ptr;  // device buffer set by cg_debug_timeline(), NULL otherwise (api.cu)

// ---- device helpers ------------------------------------------------------------------
#ifdef __CUDACC__
#ifdef CG_TIMELINE
__device__ __forceinline__ void cg_tl_mark(unsigned long long* tl, int k) {
  if (tl != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[k] = t;
  }
}
#define CG_TL(tl, k) cg_tl_mark(tl, k)
#else
#define CG_TL(tl, k)
#endif
typedef __nv_bfloat16 bf16;

// nn.GELU() (exact erf form, src/vae.py:58) for bf16 activations.  erf through Abramowitz-Stegun 7.1.25
// (|error| <= 2.5e-5, one reciprocal + one exp2 + three FMAs): Phi is off by <= 1.3e-5, far below the 2^-9 relative
// step of the bf16 value that is stored, at about a third of the instructions of erff/expf.  One shared
// exp(-x^2/2) serves both Phi(x) and the density of the derivative.
__device__ __forceinline__ float cg_ex2_approx(float v) {  // MUFU.EX2
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float cg_rcp_approx(float v) {  // MUFU.RCP
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void cg_phi_pdf(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x);
  const float e = cg_ex2_approx(-0.72134752044448170f * x * x);          // exp(-x^2 / 2)
  const float t = cg_rcp_approx(fmaf(0.33267074000000000f, ax, 1.0f));   // 1 / (1 + 0.47047 |x| / sqrt(2))
  const float poly = t * fmaf(t, fmaf(t, 0.7478556f, -0.0958798f), 0.3480242f);
  const float half_erfc = 0.5f * poly * e;                          // 0.5 * (1 - erf(|x| / sqrt(2)))
  cdf = x >= 0.0f ? 1.0f - half_erfc : half_erfc;
  pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float cg_gelu(float x) {
  float cdf, pdf;
  cg_phi_pdf(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float cg_dgelu(float x) {
  // d/dx [x Phi(x)] = Phi(x) + x phi(x)
  float cdf, pdf;
  cg_phi_pdf(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}
__device__ __forceinline__ float cg_act(float x, int act) {
  return act == CG_ACT_RELU ? fmaxf(x, 0.0f) : (act == CG_ACT_GELU ? cg_gelu(x) : (act == CG_ACT_LRELU ? (x > 0.0f ? x : 0.01f * x) : x));
}
__device__ __forceinline__ float cg_dact(float x, int act) {
  return act == CG_ACT_RELU ? (x > 0.0f ? 1.0f : 0.0f) : (act == CG_ACT_GELU ? cg_dgelu(x) : 1.0f);
}

// two floats -> packed bf16x2 (a in the low half), round to nearest even: ONE cvt; bf16x2 -> two floats: a shift and a mask
// (exact).  These sit on the conv epilogue's per-tile instruction chain, where every instruction counts.
__device__ __forceinline__ uint32_t cg_pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 cg_unpack2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// 8 bf16 <-> 8 floats
__device__ __forceinline__ void cg_unpack8(const uint4& u, float* f) {
  float2 a = cg_unpack2(u.x), b = cg_unpack2(u.y), c = cg_unpack2(u.z), d = cg_unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 cg_pack8(const float* f) {
  return make_uint4(cg_pack2(f[0], f[1]), cg_pack2(f[2], f[3]), cg_pack2(f[4], f[5]), cg_pack2(f[6], f[7]));
}

__device__ __forceinline__ float cg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- PTX wrappers: mbarrier / bulk copy / tcgen05 -----------------------------------
__device__ __forceinline__ uint32_t cg_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// one thread of a fully converged warp.  ptxas knows a region guarded by elect.sync has a single active
// thread, so uniform-datapath instructions inside it (tcgen05.mma, TMA, commit) are emitted straight instead
// of inside a per-active-thread ELECT/BRA.U.ANY loop (what `if (lane == 0)` costs: ~80 clk per tcgen05.mma).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (A try_wait with an explicit suspend-time hint -- ptxas lowers it to TRYWAIT + NANOSLEEP.SYNCS -- was measured neutral
// for the conv pipeline, profiles/r2t_conv_microbench_trywait_hint.txt, so the plain polling form stays.)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
#ifdef CG_MBAR_HINT_NS  // A/B builds only: try_wait with a suspend-time hint (TRYWAIT + NANOSLEEP.SYNCS)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"((uint32_t)CG_MBAR_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 16-byte global->shared async copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, one CTA
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of the warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// same load without the wait: issue several, then tmem_ld_wait() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, version 1 (sm_100): 16-byte units.
//   K-major : lbo = byte step between the two 8-element K halves, sbo = step between 8-row groups
//   MN-major: lbo = byte step between 8-element K groups,        sbo = step between 8-element MN groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: bf16 x bf16 -> fp32, M=128; a_mn/b_mn select MN-major operands
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#endif  // __CUDACC__
