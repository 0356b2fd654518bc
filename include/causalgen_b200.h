/* causalgen_b200 -- C ABI of the B200 (sm_100a) HVAE causal-mechanism engine.
 *
 * The reference (biomedia-mira/causal-gen) has no FFI: its boundary for this path is the
 * Python nn.Module surface of src/vae.py / src/dmol.py / src/pgm/dscm.py.  This header is
 * the C boundary UNDER that surface: every arithmetic op the path executes is one of the
 * entry points below.  Each entry cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller (PyTorch) owns every buffer, the library
 *     never allocates, frees or synchronises device memory (CUDA-graph capturable);
 *   - every launch goes on the `stream` argument (a cudaStream_t passed as void*);
 *   - return 0 on success, a negative cg_status otherwise; cg_last_error() gives the text;
 *   - activations are bf16 "channel-octet planar" (N, C/8, H, W, 8) with a sample stride `ns` (elements), so
 *     channel slices that start at a multiple of 8 are valid operands of a wider buffer; channel counts handed
 *     to the tensor-core kernels are multiples of 16 (zero padded); latent statistics / KL / likelihood math
 *     is fp32 (layout notes at cg_src below);
 *   - sm_100a only: any other device returns CG_ERR_ARCH.  There is no CPU fallback.
 */
#ifndef CAUSALGEN_B200_H
#define CAUSALGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CG_OK = 0,
  CG_ERR_ARG = -1,     /* bad shape / alignment / enum */
  CG_ERR_ARCH = -2,    /* device is not sm_100 */
  CG_ERR_CUDA = -3,    /* CUDA runtime error (text in cg_last_error) */
  CG_ERR_UNSUPPORTED = -4
} cg_status;

enum { CG_ACT_NONE = 0, CG_ACT_RELU = 1, CG_ACT_GELU = 2,    /* nn.ReLU / nn.GELU(erf), src/vae.py:50,58 */
       CG_ACT_LRELU = 3 };  /* nn.LeakyReLU(0.01) of the predictor CNN (src/pgm/layers.py:70): cg_seg.out_act and the predictor
                               kernels only (no derivative / input pre-activation) */
enum { CG_BF16 = 0, CG_F32 = 1 };

int cg_version(void);
const char* cg_last_error(void);
/* number of SMs of the current device, 0 if it is not sm_100 */
int cg_device_sms(void);

/* ---------------------------------------------------------------------------------------
 * Convolution (implicit GEMM on tcgen05, fp32 accumulation in TMEM; weight gradients of small-channel
 * problems on warp-level mma.sync, see csrc/wgrad_mma.cu)
 *   replaces nn.Conv2d forward/backward in Block / DecoderBlock: src/vae.py:49-84,165-170
 * ------------------------------------------------------------------------------------- */
#define CG_MAX_SRC 3
#define CG_MAX_SEG 4

/* Activation layout: bf16 "channel-octet planar" (N, C/8, H, W, 8) -- element (n,c,h,w) lives at
 *   ptr + n*ns + (c/8)*H*W*8 + (h*W+w)*8 + c%8        (ns = sample stride in elements)
 * so a channel slice [c0, c0+C) with c0 % 8 == 0 of a wider tensor is again a valid operand (same ns,
 * ptr advanced by (c0/8)*H*W*8).  A TMA box of this layout is directly the tcgen05 shared-memory operand. */
typedef struct {
  const void* ptr;  /* first channel-octet plane of the view */
  int64_t ns;       /* sample stride in elements (multiple of 8) */
  int32_t C;        /* channels taken from this source, multiple of 16: the K-blocks of the GEMM */
  int32_t c8;       /* channel octets PHYSICALLY stored (0 = C/8).  A tensor with 8, 24, 40 ... channels keeps 1, 3, 5 ... octet
                       planes in HBM; the TMA box of the last 16-channel K-block runs past that extent and is zero-filled, so
                       the zero half of the block costs no HBM bytes */
} cg_src;

typedef struct {
  void* ptr;        /* destination of output channels [c0, c0+cn): bf16 planar view, or fp32 rows */
  const void* add;  /* optional bf16 planar tensor added after everything else (residual / accumulate) */
  const void* add2; /* optional second addend (h + p_feat + z_proj(...) in one pass, src/vae.py:292-294) */
  const void* mul;  /* optional bf16 planar pre-activation tensor x: result *= act'(x)  (backward) */
  int64_t ns;       /* bf16: sample stride; CG_F32: row pitch of a (N*H*W, ns) fp32 row tensor */
  int64_t add_ns, add2_ns, mul_ns;
  int32_t c0, cn;   /* c0 multiple of 16, cn multiple of 8 */
  int32_t dtype;    /* CG_BF16 or CG_F32 */
  int32_t mul_act;  /* activation whose derivative is applied with `mul` */
  int32_t out_act;  /* activation applied to the finished value before the store: lets a producer write
                       act(y) when every consumer of y applies the same pre-activation (Block, src/vae.py:49-56) */
  int32_t _pad;
  void* act_copy;   /* optional second bf16 planar destination: when set, `ptr` receives the raw value and
                       act_copy receives out_act(value) (GELU blocks need both: act(y) feeds the next conv and
                       the weight gradient, y feeds act'(y) in the data gradient, src/vae.py:57-68) */
  int64_t act_copy_ns;
} cg_seg;

typedef struct {
  int32_t N, H, W;     /* stride 1, "same" padding: output spatial == input spatial */
  int32_t ksize;       /* 1 or 3 */
  int32_t act;         /* activation applied to the (concatenated) input on load */
  int32_t nsrc, nseg;
  int32_t cout;        /* output channels incl. zero padding, multiple of 16 */
  cg_src src[CG_MAX_SRC];   /* K-concatenated inputs: torch.cat([...], dim=1) without the copy */
  cg_seg seg[CG_MAX_SEG];   /* channel-split outputs (loc | logscale | features ...), disjoint ranges */
  const void* wpack;   /* weights packed by cg_pack_weights for exactly this (src list, cout) */
  const float* bias;   /* fp32 [bias_n] or NULL */
  int32_t bias_n;      /* valid bias entries (logical output channels); the rest is 0 */
  int32_t nc;          /* GEMM-N chunk the weights were packed for (cg_conv_nchunk_ex); 0 = cg_conv_nchunk(ktot16, cout) */
  int32_t fold;        /* 1: the weights were packed with cg_pack_desc.fold = 1 ("column-folded" 3x3 conv, wide inputs and
                          cout <= 32): the three kernel columns sit side by side on the GEMM-N axis (N = 3*cout), each input
                          tile is multiplied once per kernel ROW (3 MMAs per K-block instead of 9) and the epilogue adds the
                          three column partials of the left / same / right pixel.  Same result, a third of the shared-
                          memory operand reads.  Requires ksize == 3 and nc == cout <= 32 */
  int32_t _pad2;
} cg_conv_args;

/* y = conv(act(cat(src))) + bias, split/added per segment.  Also the data-gradient pass when
 * given transposed+flipped packed weights and seg.mul = saved pre-activation. */
int cg_conv2d(const cg_conv_args* a, void* stream);

/* GEMM-N chunk (output channels per CTA) the conv kernel uses for this problem size; the packed
 * weight image is laid out per chunk, so cg_pack_desc.nc must carry this value */
int32_t cg_conv_nchunk(int32_t ktot16, int32_t cout);
/* same with a hint: want_e = 0 when no launch of this pack has fused epilogue operands (add / add2 / mul), so the chunk
 * need not leave shared memory for the operand ring (first convs of a Block: wide K, narrow N) */
int32_t cg_conv_nchunk_ex(int32_t ktot16, int32_t cout, int32_t want_e);
/* 1 when the 3x3 conv (ktot16 = 9 * sum(C)/16 K-blocks, cout padded output channels) can run column-folded
 * (cg_conv_args.fold / cg_pack_desc.fold, nc = cout); want_e as above */
int32_t cg_conv_fold_ok(int32_t ktot16, int32_t cout, int32_t want_e);
/* bytes of the packed-weight image for a conv with `ktot16` K-blocks of 16 (= taps * sum(C)/16) */
int64_t cg_packed_weight_bytes(int32_t ktot16, int32_t cout);
int64_t cg_packed_weight_bytes_nc(int32_t ktot16, int32_t cout, int32_t nc);

typedef struct {
  const float* w;      /* fp32 OIHW master weight (Cout_l, Cin_l, k, k) -- reference layout */
  void* out;           /* bf16 packed image */
  int32_t cout_l, cin_l, k;   /* logical dims of w */
  int32_t transpose;   /* 0: forward pack; 1: data-gradient pack (swap O/I, flip taps) */
  int32_t taps;        /* taps packed: k*k, or 1 = centre tap only (3x3 conv on a 1x1 image) */
  int32_t n_pad;       /* GEMM-N (padded output channels of this pass), multiple of 16 */
  int32_t nc;          /* cg_conv_nchunk(ktot16, n_pad) */
  int32_t n_off;       /* forward: 0.  transpose: first logical input channel of the source */
  int32_t n_log;       /* logical channels valid on the N side (others are zero) */
  int32_t nsrc;        /* K side: sources in concat order */
  int32_t src_c[CG_MAX_SRC];      /* padded channels per source (multiple of 16) */
  int32_t src_log[CG_MAX_SRC];    /* logical channels per source */
  int32_t src_off[CG_MAX_SRC];    /* first logical channel of the source on the K side */
  int32_t fold;        /* 1: column-folded image for cg_conv_args.fold (k == 3, taps == 9, nc == n_pad <= 32):
                          K-blocks are (channel block, kernel row), GEMM-N row kx*nc + n holds kernel column kx */
  const float* n_scale; /* optional fp32 [cout_l] multiplier per OUTPUT channel of w, applied while packing (forward packs
                           only): eval-mode BatchNorm folded into the convolution, src/pgm/layers.py:72-92 */
} cg_pack_desc;

/* pack `n` weight tensors in one launch; `descs_dev` is a device copy of the descriptors */
int cg_pack_weights(const cg_pack_desc* descs_dev, int32_t n, void* stream);

typedef struct {
  int32_t N, H, W, ksize, act;
  int32_t nsrc;
  cg_src src[CG_MAX_SRC];  /* forward inputs (activation re-applied on load) */
  const void* dy;          /* bf16 planar gradient of the conv output */
  int64_t dy_ns;           /* its sample stride */
  int32_t dy_c, dy_c8;     /* padded channels (multiple of 16); octets physically stored (0 = dy_c/8), see cg_src.c8 */
  float* dw;               /* fp32 OIHW gradient, ACCUMULATED into (atomics) */
  float* dbias;            /* fp32 [cout_l] accumulated, or NULL */
  int32_t cout_l, cin_l;   /* logical dims of dw */
  int32_t src_log[CG_MAX_SRC], src_off[CG_MAX_SRC];
  int32_t taps;            /* k*k or 1 (centre tap only) */
  int32_t min_tiles;       /* scheduling hint: at least this many 128-pixel tiles per CTA (0 = 24).  Weight gradients are
                              filler next to the dependent conv chain; few CTAs with many tiles leave the SMs to that chain
                              and flush fewer accumulator tiles (the caller knows how much time the step has to hide them) */
} cg_wgrad_args;

/* dW += sum_pixels dy (x) act(cat(src)) shifted per tap; dbias += sum_pixels dy */
int cg_conv2d_wgrad(const cg_wgrad_args* a, void* stream);
/* number of kernels the call above launches for these arguments (1 or 2): launch accounting only */
int32_t cg_conv2d_wgrad_launches(const cg_wgrad_args* a);

/* 7x7 stem, fp32 NCHW image in -> bf16 planar out (src/vae.py:104-110,126) and its weight grad.  Cin 1 | 3, Cout 16 | 32.
 * Single-channel images run on warp-level tensor-core instructions with the image and the weights split into two bf16
 * halves (fp32-accurate); three-channel images on direct fp32 kernels */
int cg_stem_fwd(const float* x, const float* w, const float* b, void* y, int32_t N, int32_t Cin,
                int32_t R, int32_t Cout, int64_t y_ns, void* stream);
int cg_stem_wgrad(const float* x, const void* dy, float* dw, float* db, int32_t N, int32_t Cin,
                  int32_t R, int32_t Cout, int64_t dy_ns, void* stream);

/* ---------------------------------------------------------------------------------------
 * Resampling: F.avg_pool2d (src/vae.py:79-83), F.interpolate nearest + learned bias
 * (src/vae.py:251-262), F.pad odd resolutions (src/vae.py:130-132); bf16 planar, *_ns = sample strides
 * ------------------------------------------------------------------------------------- */
int cg_avgpool_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t d,
                   int64_t x_ns, int64_t y_ns, int32_t pad_to, void* stream);
/* dx (+)= avgpool^T(dy); accumulate!=0 adds into dx */
int cg_avgpool_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t d,
                   int64_t dy_ns, int64_t dx_ns, int32_t pad_to, int32_t accumulate, void* stream);
/* y[n,h,w,c] = bias[c,h,w] + x[n,h/s,w/s,c]; bias fp32 (C,Ho,Wo) reference layout or NULL.
 * Ho need not be a multiple of Hi (7 -> 8 style): source index floor(h*Hi/Ho). */
int cg_upsample_fwd(const void* x, const float* bias, void* y, int32_t N, int32_t Hi, int32_t Ho,
                    int32_t C, int64_t x_ns, int64_t y_ns, void* stream);
/* dx (+)= sum over replicas of dy; dbias += sum over batch of dy (fp32, may be NULL) */
int cg_upsample_bwd(const void* dy, void* dx, float* dbias, int32_t N, int32_t Hi, int32_t Ho,
                    int32_t C, int64_t dy_ns, int64_t dx_ns, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------------------------------
 * Latent blocks: sample_gaussian + gaussian_kl fused (src/vae.py:14-30,268-269)
 *   q,p: fp32 NHWC (npix, 32) = [loc(16) | logscale(16)]; eps fp32 NCHW (N,16,H,W) (reference
 *   layout, drawn by the caller) or NULL -> in-kernel Philox(seed, block offset).
 *   z_bf16: bf16 planar (N,2,H,W,8) conv operand with sample stride z_ns; z_f32: NCHW (N,16,H,W) fp32 or NULL (abduct output);
 *   kl_out: fp32 [N] += sum over (16,H,W) of the block's KL.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const float* q; const float* p; int32_t q_ld, p_ld;
  const float* eps; uint64_t seed; uint64_t offset;
  const uint64_t* seed_dev; /* optional device counter added to seed (graph-replayable noise) */
  float log_t;            /* log temperature added to both logscales (0 when t is None) */
  void* z_bf16; int64_t z_ns;
  float* z_f32;
  float* eps_out;         /* optional: NCHW eps actually used (needed by backward in Philox mode) */
  float* kl_out;          /* [N] accumulated, or NULL */
  int32_t N, HW, zdim;
  int32_t mode;           /* 0: z~q, KL(q||p)   1: z~p (prior sample)   2: z=p_loc (deterministic) */
  float* kl_ch;           /* optional [zdim] accumulated: sum over (N,H,W) of the block's KL per latent CHANNEL -- the
                             quantity kl_free_bits thresholds (src/vae.py:443-449), see cg_free_bits */
  float* kl_elem;         /* optional fp32 NCHW (N,zdim,H,W): the element-wise KL, i.e. stats[i]["kl"] of Decoder.forward
                             (src/vae.py:268); only the stand-alone decoder call materialises it */
} cg_latent_args;
int cg_latent_fwd(const cg_latent_args* a, void* stream);

/* gradients of  sum_n g_kl * kl[n]  +  <dz, z>  w.r.t. q and p statistics, written as bf16
 * into channels [0,32) of the bf16 planar conv-output-gradient buffers dq / dp */
typedef struct {
  const float* q; const float* p; int32_t q_ld, p_ld;
  const float* eps;           /* NCHW eps used in forward, or NULL -> regenerate Philox(seed, offset) */
  uint64_t seed; uint64_t offset; const uint64_t* seed_dev;
  const void* dz; int64_t dz_ns;   /* bf16 planar gradient wrt z or NULL */
  float g_kl;                 /* d loss / d kl[n]  (= beta / (B * C*H*W)) */
  void* dq; int64_t dq_ns; void* dp; int64_t dp_ns;
  int32_t N, HW, zdim, mode;
  const float* g_kl_dev;      /* optional device scalar multiplied into g_kl (beta annealing under CUDA-graph replay,
                                 src/trainer.py:52-57): g_kl then carries 1 / (B * C*H*W) and *g_kl_dev the live beta */
  float log_t;                /* log temperature the forward added to both logscales (abduction with t, src/vae.py:176-190) */
  int32_t _pad;
  const float* kl_gate;       /* optional [zdim] in {0,1}: kl_free_bits gate of this block's channels (cg_free_bits); the KL
                                 gradient of a channel whose batch-mean KL sits below the free-bits floor is zero */
} cg_latent_bwd_args;
int cg_latent_bwd(const cg_latent_bwd_args* a, void* stream);

/* kl_free_bits > 0 (src/vae.py:443-449): kl = sum_blocks sum_c max(free_bits, mean_batch sum_hw KL[b,c,h,w]).
 *   kl_ch: [nblk*16] per-channel sums written by cg_latent_fwd (summed over the GLOBAL batch: all-reduce it first under
 *   data parallelism); inv_batch = 1 / global batch.  Writes gate[nblk*16] (1 where the mean exceeds the
 *   floor: those channels get a KL gradient) and fills kl_row[0..N) with the total, so that cg_elbo_finalize(kl = kl_row,
 *   nblk = 1) yields the reference's kl / elbo. */
int cg_free_bits(const float* kl_ch, int32_t nblk, float free_bits, float inv_batch, float* gate, float* kl_row, int32_t N,
                 void* stream);

/* mediator mixture r = alpha q + (1-alpha) p of HVAE.abduct (src/vae.py:485-513); all fp32 NCHW */
int cg_latent_mix(const float* z, const float* q_loc, const float* q_ls, const float* p_loc,
                  const float* p_ls, float* out, int64_t n, float alpha, float t, int32_t has_t,
                  void* stream);

/* ---------------------------------------------------------------------------------------
 * Likelihoods
 * ------------------------------------------------------------------------------------- */
/* DGaussNet (src/vae.py:322-422): 1x1 heads fused with the discretised-Gaussian NLL.
 * h bf16 planar with Cw channels, sample stride h_ns; x fp32 NCHW (N,C,H,W); w_* fp32 (C,Cw); C in {1,3}. */
typedef struct {
  const void* h; int64_t h_ns; int32_t Cw, _pad0;
  const float* x;
  const float *w_loc, *b_loc, *w_ls, *b_ls, *w_co, *b_co;   /* w_co/b_co NULL when C==1 */
  int32_t N, HW, C;
  float* nll;        /* fwd: [N] accumulated: -mean over (C,H,W) of log-prob */
  float g;           /* bwd: d loss / d nll[n]  (= 1/B) */
  void* dh; int64_t dh_ns;                  /* bwd: bf16 planar, written */
  float *dw_loc, *db_loc, *dw_ls, *db_ls, *dw_co, *db_co;   /* bwd: accumulated */
} cg_dgauss_args;
int cg_dgauss_nll_fwd(const cg_dgauss_args* a, void* stream);
int cg_dgauss_nll_bwd(const cg_dgauss_args* a, void* stream);
/* likelihood.sample(h, return_loc=True): x = clamp(loc), scale = exp(logscale) as fp32 NCHW;
 * eps!=NULL -> x = clamp(loc + exp(logscale + log_t) * eps) (intended return_loc=False semantics) */
int cg_dgauss_sample(const cg_dgauss_args* a, float* x_out, float* scale_out, const float* eps,
                     float log_t, void* stream);

/* backward of cg_dgauss_sample (eps == NULL): given d x_out and/or d scale_out (fp32 NCHW, either may be NULL) writes
 * a->dh (bf16 planar) and accumulates a->dw_* / a->db_*.  Gradients reach the likelihood this way when the counterfactual
 * is trained through (src/pgm/dscm.py:53-56,78-88; src/pgm/train_cf.py:159-180). */
int cg_dgauss_sample_bwd(const cg_dgauss_args* a, const float* dx_out, const float* dscale_out, void* stream);

/* DmolNet (src/dmol.py:24-245): 1x1 conv to 100 channels fused with the mixture loss.
 * w (100,Cw) b (100) fp32; x fp32 NCHW (N,3,H,W). */
typedef struct {
  const void* h; int64_t h_ns; int32_t Cw, _pad0;
  const float* x; const float* w; const float* b;
  int32_t N, HW;
  float* nll;   /* fwd [N] accumulated */
  float g; void* dh; int64_t dh_ns; float* dw; float* db;   /* bwd */
} cg_dmol_args;
int cg_dmol_loss_fwd(const cg_dmol_args* a, void* stream);
int cg_dmol_loss_bwd(const cg_dmol_args* a, void* stream);
/* mode 0 soft mean, 1 hard (argmax) mean, 10+k 'top<k>' mean (the k most probable components, renormalised; src/dmol.py:
 * 178-189), 2 sample (u_gumbel (N,H,W,10), u_logistic (N,H,W,3)
 * uniform(1e-5,1-1e-5) drawn by caller); outputs fp32 NCHW x (clamped) and scale */
int cg_dmol_predict(const cg_dmol_args* a, int32_t mode, const float* u_gumbel,
                    const float* u_logistic, float log_t, float* x_out, float* scale_out,
                    void* stream);

/* ---------------------------------------------------------------------------------------
 * DSCM counterfactual combine (src/pgm/dscm.py:55-72), fp32 NCHW
 *   u = (x-rec_loc)/max(rec_scale,1e-12); cf = clamp(cf_loc + cf_scale*u, -1, 1)
 *   sum/sum2 (optional) accumulate particles.
 * ------------------------------------------------------------------------------------- */
int cg_cf_combine(const float* x, const float* rec_loc, const float* rec_scale, const float* cf_loc,
                  const float* cf_scale, float* cf_x, float* sum, float* sum2, int64_t n,
                  void* stream);

/* backward of the combine: d cf_x -> gradients of the four likelihood outputs (src/pgm/dscm.py:55-56) */
int cg_cf_combine_bwd(const float* x, const float* rec_loc, const float* rec_scale, const float* cf_loc,
                      const float* cf_scale, const float* dcf, float* d_rec_loc, float* d_rec_scale,
                      float* d_cf_loc, float* d_cf_scale, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Layout glue (src/trainer.py:16-21, src/pgm/dscm.py:121-132)
 * ------------------------------------------------------------------------------------- */
/* uint8 (N,C,H,W) -> fp32 NCHW (x-127.5)/127.5 */
int cg_normalise_u8(const uint8_t* x8, float* x, int64_t n, void* stream);
/* on-device train-time augmentation of a uint8 NCHW batch (src/datasets.py:107-118,281-286): torchvision
 * RandomCrop(size=R, padding=[pad_left, pad_top], fill=0) followed by RandomHorizontalFlip; params = int32 (N,3) device
 * array of (top, left, flip) per sample, top in [0, Hi + 2*pad_top - R], left in [0, Wi + 2*pad_left - R] */
int cg_augment_u8(const uint8_t* in, uint8_t* out, const int32_t* params, int32_t N, int32_t C, int32_t Hi, int32_t Wi,
                  int32_t R, int32_t pad_top, int32_t pad_left, void* stream);
/* parents (N,ctx,R,R) fp32 [sampled at pixel (0,0)] or (N,ctx) -> bf16 planar (N, ctx16/8, H, W, 8), spatially
 * constant, zero padded channels; channels >= drop_from multiplied by drop_scale (conditioning dropout,
 * src/vae.py:244-247).  This is the materialised parents[..., :res, :res] of src/vae.py:241 at bf16. */
int cg_parents_plane(const float* pa, int64_t sample_stride, int64_t chan_stride, void* out, int32_t N,
                     int32_t ctx, int32_t C, int32_t HW, int64_t ns, int32_t drop_from, float drop_scale,
                     const float* drop_scale_dev /* optional device scalar overriding drop_scale (graph replay) */,
                     void* stream);
/* fp32 NCHW (N,C,H,W) <-> bf16 planar (latents in / out) */
int cg_nchw_f32_to_planar(const float* x, void* y, int32_t N, int32_t C, int32_t HW, int64_t ns,
                          void* stream);
int cg_planar_to_nchw_f32(const void* x, float* y, int32_t N, int32_t C, int32_t HW, int64_t ns,
                          void* stream);
/* fp32 row statistics slice [c0,c0+C) (+add) -> fp32 NCHW (abduct's q_loc/q_logscale dict entries) */
int cg_stats_to_nchw(const float* src, int32_t ld, int32_t c0, float add, float* dst, int32_t N, int32_t C,
                     int32_t HW, void* stream);
/* y[n, c, :] = v[c] for every pixel (decoder initial state bias[1].repeat, src/vae.py:232) */
int cg_fill_planar(const float* v, void* y, int32_t N, int32_t HW, int32_t C, int64_t ns, void* stream);
/* dv[c] += sum_{n,pixels} dy[n,c,pixel] (fp32 accumulate): bias gradients */
int cg_colsum(const void* dy, float* dv, int32_t N, int32_t HW, int32_t C, int64_t ns, void* stream);
/* y = a + b (bf16 planar) */
int cg_add(const void* a, const void* b, void* y, int32_t N, int32_t HW, int32_t C, int64_t a_ns, int64_t b_ns,
           int64_t y_ns, void* stream);

/* kl rows (nblk, N): per-block per-sample KL sums.  kl_pp[n] = kl_scale * sum_blk kl[blk][n];
 * elbo = mean(nll) + beta*mean(kl_pp)  (src/vae.py:451-458); out[3] = {elbo,nll,kl} */
int cg_elbo_finalize(const float* nll, const float* kl, float* out, int32_t N, int32_t nblk,
                     float kl_scale, float beta, const float* beta_dev /* optional device scalar overriding beta */,
                     void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimiser tail on flat fp32 buffers (src/trainer.py:66-87, src/train_setup.py:42-53,
 * src/utils.py:169-220): global-norm clip, NaN/skip test, AdamW, EMA -- no host sync.
 * ------------------------------------------------------------------------------------- */
/* scratch == NULL: out[0] += sum g^2 (atomics).  scratch (>= 592 floats) given: out[0] = sum g^2 through per-block
 * partials added in a fixed order -- bit-identical on every data-parallel replica, so their clip coefficients agree */
int cg_sumsq(const float* g, float* out, int64_t n, float* scratch, int32_t scratch_n, void* stream);
/* Device-side step bookkeeping so the whole training step is CUDA-graph replayable:
 *   state[4] (int32): {adam step t, ema update calls, skipped updates, skip flag of this step}
 *   dyn[6]   (fp32) : {lr, 1-b1^t, 1-b2^t, ema decay, grad scale*clip coefficient, grad norm}
 * gsumsq = sum of squares of the (all-reduced, still unscaled) flat gradient; grad_scale = 1/world.
 * Skips (flag=1, counters untouched) when the norm >= grad_skip or nll/kl in loss_terms[3] is NaN. */
int cg_optim_advance(int32_t* state, float* dyn, const float* gsumsq, const float* loss_terms,
                     float base_lr, int32_t warmup, float beta1, float beta2, float grad_clip,
                     float grad_skip, float grad_scale, float ema_beta, int32_t ema_after, void* stream);
/* p,m,v,ema updated in place from g using state/dyn written by cg_optim_advance; ema may be NULL */
int cg_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                      const int32_t* state, const float* dyn, float beta1, float beta2, float eps,
                      float wd, void* stream);

/* ---------------------------------------------------------------------------------------
 * Anticausal predictors on counterfactual images (call site src/pgm/dscm.py:78-83), inference only:
 * `CNN` (BatchNorm, src/pgm/layers.py:64-104) and GroupNorm ResNet-18 (src/pgm/resnet.py:9-239).  Their 3x3 / 1x1
 * convolutions are cg_conv2d launches (BatchNorm folded through cg_pack_desc.n_scale + bias, LeakyReLU as out_act);
 * stride-2 convolutions = stride-1 cg_conv2d + cg_pool_max_fwd(k = 1, stride = 2).
 * ------------------------------------------------------------------------------------- */
/* eval-mode nn.BatchNorm{1,2}d as an affine map: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale */
int cg_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
               float* shift, int32_t C, void* stream);
/* thin-input stem convolution (Cin <= 3, k <= 7, any stride / padding; src/pgm/layers.py:72, src/pgm/resnet.py:229-231):
 * x fp32 NCHW, w fp32 OIHW, y bf16 planar (N, ceil(Cout/8), Ho, Wo, 8) = act(conv(x) * scale[c] + shift[c]); scale/shift
 * optional (folded BatchNorm) */
int cg_conv_direct_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int32_t N,
                       int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act,
                       int64_t y_ns, void* stream);
/* nn.MaxPool2d(k, stride, pad) on bf16 planar tensors (src/pgm/layers.py:75, src/pgm/resnet.py:98); k = 1: strided pick */
int cg_pool_max_fwd(const void* x, void* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t k, int32_t stride, int32_t pad,
                    int64_t x_ns, int64_t y_ns, void* stream);
/* y = act(GroupNorm(groups, C)(x) [+ add]) on bf16 planar tensors (src/pgm/resnet.py:228, CustomBlock.forward :41-61);
 * stats: fp32 scratch [N*C*2] (per-channel sum / sum of squares, written by the call) */
int cg_groupnorm_fwd(const void* x, float* stats, const float* gamma, const float* beta, int32_t groups, float eps,
                     const void* add, int64_t add_ns, int32_t act, void* y, int32_t N, int32_t C, int32_t HW, int64_t x_ns,
                     int64_t y_ns, void* stream);
/* x.mean(dim=(-2,-1)) of a bf16 planar tensor -> fp32 rows out[n*ld + c]  (src/pgm/layers.py:101, nn.AdaptiveAvgPool2d(1)) */
int cg_global_avgpool(const void* x, float* out, int32_t N, int32_t C, int32_t HW, int64_t x_ns, int32_t ld, void* stream);
/* out[n][m] = act((x[n] . w[m] + bias[m]) * scale[m] + shift[m]), fp32; bias / (scale, shift) optional
 * (nn.Linear [+ BatchNorm1d + LeakyReLU], src/pgm/layers.py:94-99, src/pgm/resnet.py:233) */
int cg_linear(const float* x, int32_t ldx, const float* w, const float* bias, const float* scale, const float* shift,
              int32_t act, float* out, int32_t ldo, int32_t N, int32_t K, int32_t M, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAUSALGEN_B200_H */
