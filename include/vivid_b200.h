/*
 * vivid_b200 — C ABI of libvividb200.so
 *
 * B200-native (sm_100a) kernels for VIVID's guided EDM2 denoising hot path.
 * The reference (danielcodelavin/vivid) has no FFI of its own: its boundary is
 * Python duck-typing.  This header is the boundary a maintainer
 * would bind instead of the PyTorch library calls listed per entry point below;
 * INTEGRATION.md shows the ctypes stub.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     stated otherwise; tensors are contiguous in the layout stated;
 *   - every call returns 0 on success, a negative vb_status otherwise, never
 *     throws or aborts; vb_last_error() returns a thread-local message;
 *   - all work is enqueued on the caller's CUDA stream (cudaStream_t passed as
 *     void*), no implicit synchronisation, no allocation on the hot path;
 *   - the library is GPU-only: there is no CPU fallback.
 *
 * Activation layout inside the library is NHWC ("pixels x channels") in ONE
 * 16-bit format for GEMM operands and the residual stream alike: IEEE fp16, the
 * reference's own reduced precision (training/models.py:632) — a build with
 * -DVB_OP_BF16 switches it to bf16.  Accumulation (TMEM), normalisation
 * statistics and the sampler state are fp32.  Channel counts of GEMM operands
 * are padded to multiples of 64 (inputs) / 16 (outputs).
 */
#ifndef VIVID_B200_H_
#define VIVID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: vb_conv_desc grew ks_ws; the whole-sampler (vb_sample), library-side plan recording (vb_net_plan_*) and feature / trace
 * entry points were added.  vb_struct_size() lets a binding verify every descriptor layout. */
#define VB_ABI_VERSION 2

typedef enum vb_status {
  VB_OK = 0,
  VB_ERR_INVALID = -1,  /* bad argument / unsupported shape */
  VB_ERR_CUDA = -2,     /* CUDA runtime / driver error (message has details) */
  VB_ERR_NO_DEVICE = -3 /* no sm_100 device is current */
} vb_status;

typedef enum vb_dtype { VB_F32 = 0, VB_F16 = 1, VB_BF16 = 2, VB_F64 = 3, VB_U8 = 4 } vb_dtype;

const char* vb_last_error(void);
int vb_abi_version(void);
/* 0 when the current CUDA device is compute capability 10.x, VB_ERR_NO_DEVICE otherwise. */
int vb_device_check(void);
/* Measurement plumbing: occupies `stream` for the given time (one spinning thread) so that work enqueued behind it
 * executes back to back, independent of the host's launch rate. */
int vb_spin(int microseconds, void* stream);
/* Programmatic dependent launch of the library's kernels (default on; VB_PDL=0 in the environment turns it off).  Returns
 * the previous setting.  Affects launches and graph captures made AFTER the call (measurement aid: with it off, kernel
 * activity records of consecutive kernels do not overlap). */
int vb_set_pdl(int on);
/* sizeof() of the descriptor structs below, in declaration order (0 = vb_weight_prep_desc ... 7 = vb_heun_desc, 8 = vb_stats_desc, 9 = vb_f32_conv_desc, 10 = vb_f32_op_desc, 11 = vb_io_desc, 12 = vb_sample_desc, 13 = vb_unet_desc, 14 = vb_net_desc, 15 = vb_param);
 * lets a foreign-language binding verify its mirror of the layout.  -1 for an unknown index. */
int vb_struct_size(int which);
/* vb_dtype of GEMM operands / stream in this build (VB_F16 unless built with -DVB_OP_BF16). */
int vb_operand_dtype(void);

/* ------------------------------------------------------------------------
 * Weight preparation — replaces the per-call prologue of MPConv.forward
 * (reference training/models.py:115-121 + normalize :37-42):
 *   w_eff = gain * w / (eps*sqrt(K) + ||w||_2 per out-channel),  K = cin*taps, eps = 1e-4
 * computed in fp32 from fp32/fp16 parameters, rounded once to 16 bits and
 * repacked OIHW -> [cout_pad][tap][cin_pad] (K-major GEMM B operand).
 * Optional: fold the two mp_cat scale factors (models.py:78-84) into the
 * input-channel segments [0,split) and [split,cin); de-interleave the qkv /
 * kv output-channel order c = h*P*D + d*P + j  ->  h*P*D + j*D + d
 * (models.py:192,285) so one (head, q|k|v) vector is contiguous.
 * dst_dtype VB_F32 writes the unpadded fp32 matrix [cout][cin*taps] instead
 * (used by the small embedding linears, models.py:123).
 * ------------------------------------------------------------------------ */
typedef struct vb_weight_prep_desc {
  const void* src; /* [cout][cin][taps] contiguous */
  void* dst;
  int32_t src_dtype; /* vb_dtype */
  int32_t dst_dtype; /* VB_F16 / VB_BF16 = the library's operand format (padded, repacked) or VB_F32 (plain) */
  int32_t cout, cin, taps;
  int32_t cout_pad; /* rows of dst (>= cout, extra rows zero) */
  int32_t split;    /* = cin when there is a single input segment */
  int32_t seg_a_pad, seg_b_pad; /* padded channel extents of the two segments in dst */
  int32_t perm_parts, perm_dim; /* 0,0 = keep order; (3,D) qkv; (2,D) kv */
  float gain;
  float scale_a, scale_b;
} vb_weight_prep_desc;
int vb_weight_prep(const vb_weight_prep_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Implicit-GEMM convolution (3x3 same-pad or 1x1) on tcgen05/TMEM fed by TMA —
 * replaces F.conv2d in MPConv.forward (models.py:126) together with the
 * pointwise ops the reference runs around it (Block.forward, models.py:165-206):
 *   flags VB_F_MODSILU : v = mp_silu(v * mod[b][c])                 (:175-176)
 *   res_mode           : v = mp_sum(res, v, res_t)                  (:184,202)
 *   flags VB_F_CLIP    : v = clamp(v, -clip, clip)                  (:204-205)
 *   flags VB_F_RESB_FOLDED : the prepared weights already carry mp_sum's coefficient of v (t / sqrt((1-t)^2 + t^2), e.g.
 *                        through vb_weight_prep_desc.gain): the epilogue adds the residual term only
 *   epi  VB_EPI_QKVNORM: per-(token, head, q|k|v) normalize over D and scatter
 *                        to [B][heads][seq][D] 16-bit tensors       (:192-193,283-297)
 * GEMM view: M = B*H*W pixels, N = cout, K = taps*(cin_pad+cin2_pad).
 * ------------------------------------------------------------------------ */
enum { VB_EPI_PLAIN = 0, VB_EPI_QKVNORM = 1 };
enum { VB_F_MODSILU = 1, VB_F_CLIP = 4, VB_F_RESB_FOLDED = 8 };
/* residual input: none | mp_sum(res, v) | mp_sum(pixel_norm(res), v)  (the enc-flavour block normalises its input
 * before using it as the residual base, models.py:171,184; recomputed in the epilogue instead of stored) |
 * mp_sum(res * res_rnorm[pixel], v): the same, with the per-pixel 1/(eps + rms) taken from the fp32 side channel the
 * producer of `res` wrote through out_rnorm (one pass over the residual tile instead of two) */
enum { VB_RES_NONE = 0, VB_RES_PLAIN = 1, VB_RES_PIXNORM = 2, VB_RES_SCALED = 3 };
/* output slots: v | mp_silu(v*scale) | pixel_norm(v) | mp_silu(pixel_norm(v)).  The NORM kinds need the whole
 * channel extent in one tile (cout_pad == block_n <= 256): this is the next block's pixel-norm fused here. */
enum { VB_OUT_NONE = 0, VB_OUT_RAW = 1, VB_OUT_SILU = 2, VB_OUT_NORM = 3, VB_OUT_NORM_SILU = 4 };

typedef struct vb_conv_desc {
  const void* x;  /* 16-bit NHWC [B][H][W][cin_pad] */
  const void* x2; /* optional second channel segment (mp_cat folded into K), 16-bit NHWC [B][H][W][cin2_pad] */
  const void* w;  /* 16-bit [cout_pad][taps*(cin_pad+cin2_pad)] from vb_weight_prep */
  const float* mod; /* fp32 [B][mod_stride], pre-offset to this layer's first channel */
  const void* res;  /* 16-bit NHWC [B*H*W][cout_pad] residual stream (TMA-staged through shared memory) */
  void* out[3];     /* 16-bit NHWC [B*H*W][cout_pad] outputs (shared-memory staged, TMA stores) */
  float* out_f32;   /* optional fp32 [B*H*W][ld_f32] copy of v, direct stores (the 3-channel out_conv) */
  float* out_rnorm; /* optional fp32 [B*H*W]: 1/(1e-4 + rms_c(v)) per pixel, written when a NORM kind is emitted */
  const float* res_rnorm; /* VB_RES_SCALED: fp32 [B*H*W] per-pixel scale of the residual */
  void* part_out[3]; /* QKVNORM: q,k,v (or k,v) 16-bit [B/seg_div][heads][part_seq[j]][part_ld or head_dim] */
  int32_t B, H, W;
  int32_t cin_pad, cin2_pad;
  int32_t cout_pad; /* multiple of block_n */
  int32_t taps;     /* 1 or 9 */
  int32_t block_n;  /* 16..256, multiple of 16 (multiple of 64 when res/out[] are used) */
  int32_t epi_mode, flags;
  int32_t mod_stride, ld_f32;
  int32_t res_mode;
  int32_t out_kind[3];
  int32_t head_dim, parts, seg_div;
  int32_t part_seq[3]; /* total sequence length of each destination */
  int32_t part_off[3]; /* first sequence slot written by image segment 0 */
  float out_scale[3];  /* VB_OUT_SILU: mp_silu(v * scale) (mp_cat weight folded in).  VB_EPI_QKVNORM: out_scale[j] != 0 multiplies
                          the normalised rows of part j (the plans fold the softmax's log2(e)/sqrt(D) into q this way) */
  float res_t, clip;
  int32_t tune;        /* plan-time autotuning; 0 = the library decides.  bits 0-1: 1 single CTA, 2 CTA pair;
                        * bits 2-3: 1 per-tap operand boxes, 2 shared haloed boxes wherever they fit;
                        * bits 4-5: 1 row-rolling input-stationary layout (3x3, 64 -> 64 channels, rows >= 128 px);
                        * bit 6: ping-pong epilogue (two groups of four warps on alternate tiles; block_n == 64);
                        * bit 7: 1x1 layers keep their N tile's weights resident in shared memory;
                        * bit 8: K-split — the two CTAs of a cluster share a tile, each summing half of the 64-channel blocks of
                        *        every tap; the second CTA's fp32 partial goes through ks_ws and is added as (first + second) by
                        *        the first.  For layers with fewer tiles than SMs and long K loops (the 8x8 level).  Unlike
                        *        bits 0-7 this changes the summation order: decide it from the layer's geometry, never from the
                        *        batch size, or results stop being identical across batch splits. */
  int32_t part_ld;     /* QKVNORM: elements per destination row; 0 = head_dim.  64 with head_dim 32: the rows are written into
                        * zero-initialised 64-element rows (the upper half is never touched), see vb_attn_desc.ld */
  void* ks_ws;         /* tune bit 8: fp32 workspace of vb_conv_ksplit_ws_bytes(B, H, W, cout_pad) bytes; may be shared by all
                        * K-split convs launched in order on one stream */
} vb_conv_desc;
int vb_conv(const vb_conv_desc* d, void* stream);
/* Bytes of K-split workspace a conv over [B][H][W] pixels with cout_pad output channels needs (whole 128-pixel tiles). */
int64_t vb_conv_ksplit_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t cout_pad);
/* Diagnostics (micro-benchmarks only): with the environment variable VB_DBG & 16 set when an op is prepared, the conv
 * kernel records the longest CTA lifetime in SM cycles; this call synchronises, returns the maximum since the last
 * call (HOST pointer) and resets it. */
int vb_debug_conv_cycles(unsigned long long* out);
/* VB_DBG & 32: clock64 stamps of CTA 0 of the last conv launch: [0] start, [1] set-up done, [2] first operand stage
 * landed, [3] last MMA committed, [4] first accumulator ready, [5] last store drained, [6] exit (HOST pointer, 8 values). */
int vb_debug_conv_stamps(long long* out8);

/* ------------------------------------------------------------------------
 * Fused cosine attention — replaces einsum/softmax/einsum (snapshot
 * experiments/code/training/models.py:190-191,274-280) and
 * F.scaled_dot_product_attention (current tree training/models.py:198,305):
 *   y[b][s][h*D+d] = sum_k softmax_k(q.k/sqrt(D)) v ; q,k,v already normalised.
 * zero_keys extra keys with k = v = 0 are accounted for analytically (the
 * unconditional gnet's all-zero source features, snapshot models.py:616-625).
 * ------------------------------------------------------------------------ */
typedef struct vb_attn_desc {
  const void* q; /* 16-bit [B][heads][sq][D] */
  const void* k; /* 16-bit [B][heads][sk][D] */
  const void* v; /* 16-bit [B][heads][sk][D] */
  void* y;       /* 16-bit [B][sq][heads*D] (NHWC) */
  int32_t B, heads, sq, sk, head_dim;
  int32_t zero_keys;
  int32_t ld;          /* elements per q/k/v row; 0 = head_dim.  64 with head_dim 32: rows zero-padded to 64 elements (what the
                          plans emit for D = 32 so that the tcgen05 kernel serves it; y stays dense [.., heads*32]) */
  int32_t q_prescaled; /* != 0: q already carries log2(e)/sqrt(D) (the QKV GEMM epilogue folded it in: vb_conv_desc.out_scale[0]
                          in VB_EPI_QKVNORM mode), the kernel computes p = 2^(q.k) without a per-logit multiply */
} vb_attn_desc;
int vb_attn(const vb_attn_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Fused elementwise passes over the 16-bit NHWC stream (vectorised, coalesced;
 * warp-shuffle reductions; fp32 math).  Only the passes that cannot ride in a
 * GEMM epilogue remain:
 *   VB_EW_PIXNORM      out = x/(eps+||x||_C/sqrt(C)); out_silu = mp_silu(out)    (models.py:171, :37-42)
 *                      (levels whose channel count exceeds one GEMM tile: C > 256)
 *   VB_EW_DOWN_PIXNORM 2x2 mean pool (resample 'down', :48-61) then PIXNORM
 *   VB_EW_UP           nearest x2 (resample 'up'): out = up(x); out_silu = mp_silu(up(x))
 *   VB_EW_CAT          mp_cat(a,b,t) (:78-84): out = cat; out_silu = mp_silu(cat)  (kept for odd channel counts;
 *                      the plans fold mp_cat into the consumer GEMM's K loop instead)
 *   VB_EW_SILU         out = x*wa (optional), out_silu = mp_silu(x*wa)
 * ------------------------------------------------------------------------ */
enum { VB_EW_PIXNORM = 0, VB_EW_DOWN_PIXNORM = 1, VB_EW_UP = 2, VB_EW_CAT = 3, VB_EW_SILU = 4 };
typedef struct vb_ew_desc {
  const void* a; /* 16-bit [pixels_in][ca] */
  const void* b; /* 16-bit [pixels][cb] (CAT) */
  void* out;      /* optional 16-bit result */
  void* out_silu; /* optional 16-bit mp_silu(result) */
  int32_t kind;
  int32_t B, H, W; /* OUTPUT spatial extent */
  int32_t ca, cb;
  float wa, wb; /* CAT scale factors */
} vb_ew_desc;
int vb_eltwise(const vb_ew_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Embedding path — MPFourier (models.py:96-101), emb_noise / emb_label linears
 * (:123), mp_sum(.., label_balance) and mp_silu (UNet.forward :388-391), then
 * every block's  c = emb_linear(emb, gain=emb_gain) + 1  (:175) in one pass:
 *   emb[b] = mp_silu(mp_sum(Wn . fourier(c_noise[b]), Wl . geom[b], t))
 *   mod[b][j] = Wmod[j] . emb[b] + 1          (Wmod = all blocks' emb_linear, stacked)
 * ------------------------------------------------------------------------ */
typedef struct vb_emb_desc {
  const float* sigma;  /* [B] or [1] (sigma_n) */
  const float* geom;   /* [B][label_dim] or NULL */
  const float* freqs;  /* [cnoise] */
  const float* phases; /* [cnoise] */
  const float* w_noise; /* fp32 [cemb][cnoise] prepared */
  const float* w_label; /* fp32 [cemb][label_dim] prepared, or NULL */
  const float* w_mod;   /* fp32 [mod_total][cemb] prepared */
  float* emb;           /* scratch [B][cemb] */
  float* mod;           /* out [B][mod_total] */
  int32_t B, sigma_n, sigma_stride, cnoise, cemb, label_dim, mod_total;
  int32_t geom_rows;    /* rows available in geom (1 => broadcast) */
  float label_balance;
  float noise_scale;    /* 1 normally; 0 for no_time_enc encoders (c_noise*0) */
  float geom_scale;     /* 1 normally; 0 for uncond (geometry*0) */
} vb_emb_desc;
int vb_embed(const vb_emb_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * EDM preconditioning (NVPrecond.forward, snapshot models.py:588-595,608-611,632)
 *   vb_precond_in : x_in = c_in(sigma)*x  -> NHWC 16-bit, + ones channel, + SR
 *                   conditioning channels (cond + noisy_sr*noise), zero padded to cpad
 *   vb_precond_out: D = c_skip*x + c_out*F   (fp32 NCHW)
 * ------------------------------------------------------------------------ */
typedef struct vb_precond_in_desc {
  const float* x;     /* fp32 NCHW [B][3][R][R] (image stride img_stride floats) */
  const float* cond;  /* optional fp32 NCHW [B][3][R][R] */
  const float* noise; /* optional fp32 NCHW [B][3][R][R], added as noisy_sr*noise */
  const float* sigma; /* [B] or [1]; NULL => no c_in scaling (encoder input) */
  void* out;          /* 16-bit NHWC [B][R][R][cpad] */
  int32_t B, R, cpad, sigma_n, sigma_stride;
  int32_t im2col;     /* 1: write the 3x3 neighbourhood (channel index ci*9+tap, zero outside the image) so that the first
                         3x3 conv (models.py:351,394) becomes a K=64 1x1 GEMM instead of 9 taps of 64 padded channels */
  int64_t img_stride; /* elements between consecutive images of x (allows x[::2]) */
  float sigma_data, noisy_sr;
} vb_precond_in_desc;
int vb_precond_in(const vb_precond_in_desc* d, void* stream);

typedef struct vb_precond_out_desc {
  const float* x;     /* fp32 NCHW noisy input */
  const float* f;     /* fp32 NHWC [B][R][R][ldf] raw network output (3 valid channels) */
  const float* sigma; /* [B] or [1] */
  float* d_out;       /* fp32 NCHW [B][3][R][R] */
  int32_t B, R, ldf, sigma_n, sigma_stride;
  int64_t img_stride;
  float sigma_data;
} vb_precond_out_desc;
int vb_precond_out(const vb_precond_out_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Heun step + autoguidance (edm_sampler, generate_images.py:62,93-114):
 *   D  = lerp(D_g, D_n, guidance)              (D_g may be NULL)
 *   phase 0 (Euler):  d_cur = (x_hat - D)/t_hat;  x_next = x_hat + (t_next-t_hat)*d_cur
 *   phase 1 (2nd order): d' = (x_next - D)/t_next; x_next = x_hat + (t_next-t_hat)*(d_cur+d')/2
 * all fp32, elementwise over n values.
 * Zero-copy hand-off to the next denoiser call (optional, NULL = off): the new x_next is ALSO stored to x_out[0..1]
 * (the input buffers of the plans that run next: net and gnet), and sigma_next is broadcast to sigma_out[k][0..sigma_n)
 * (their per-sample noise-level inputs), so no host-side copy or fill sits between two calls.
 * ------------------------------------------------------------------------ */
typedef struct vb_heun_desc {
  const float* d_net;
  const float* d_gnet; /* or NULL */
  const float* x_hat;
  float* d_cur;  /* phase 0: written; phase 1: read */
  float* x_next; /* phase 0: written; phase 1: read then overwritten */
  float* x_out[2];     /* or NULL: extra copies of x_next */
  float* sigma_out[2]; /* or NULL: receive sigma_next, sigma_n values each */
  int64_t n;
  int32_t phase;
  int32_t sigma_n;
  float guidance, t_hat, t_next, sigma_next;
} vb_heun_desc;
int vb_heun(const vb_heun_desc* d, void* stream);

/* Uncertainty head u(sigma) = logvar_linear(logvar_fourier(ln(sigma)/4)) (training/models.py:746-747; dual-source mode
 * :686-688 passes sigma_stride = 2 for c_noise[::2]).  weight is the RAW [channels] fp32 parameter of logvar_linear (the
 * forced weight normalisation is applied inside), freqs/phases the MPFourier buffers; out is [n] fp32. */
int vb_logvar(const float* sigma, int32_t n, int32_t sigma_stride, const float* weight, const float* freqs,
              const float* phases, int32_t channels, float* out, void* stream);

/* ------------------------------------------------------------------------
 * Metric statistics of `calculate_metrics.py gen` (calculate_stats_for_iterable_nvs, :158-172, :148):
 *   vb_stats_update: cum_mu[F] += sum_n f[n,:], cum_sigma[F][F] += f^T f, fp64, for a batch of detector features
 *     f = [feat | feat2] (F = f1 + f2; feat2/f2 = 0 for the plain statistics, the src-view features for the joint_* ones,
 *     replacing torch.cat + .to(float64) + .sum(0) + features.T @ features).  Deterministic (no atomics).
 *   vb_psnr_u8: psnr_out[i] = 10 log10(255^2 / mean((images[i] - tgt[i])^2)) per image (fp64 statistics); when cum_sum is
 *     not NULL, cum_sum[0] += sum_i psnr_out[i] (image order).  tgt is uint8 or fp32 in [0, 255], images uint8.
 * ------------------------------------------------------------------------ */
typedef struct vb_stats_desc {
  const void* feat;  /* [n][ld1] detector features */
  const void* feat2; /* [n][ld2] or NULL */
  double* cum_mu;    /* [f1 + f2] */
  double* cum_sigma; /* [f1 + f2][f1 + f2] */
  int64_t ld1, ld2;
  int32_t dtype; /* vb_dtype of feat and feat2: VB_F32 / VB_F16 / VB_BF16 / VB_F64 */
  int32_t n, f1, f2;
} vb_stats_desc;
int vb_stats_update(const vb_stats_desc* d, void* stream);
int vb_psnr_u8(const uint8_t* images, const void* tgt, int32_t tgt_dtype, int32_t n, int64_t per_image,
               int64_t tgt_image_stride, double* psnr_out, double* cum_sum, void* stream);

/* Image resize, fp32 NCHW planes [planes][h_in][w_in] -> [planes][h_out][w_out]: the arithmetic of
 * torch.nn.functional.interpolate(mode="bilinear", align_corners=False, antialias=<0|1>) — the x4 inter-stage upscale
 * (generate_images.py:322) and the anti-aliased x1/4 low-res conditioning of the SR-only path (:282-283). */
int vb_resize(const float* src, float* dst, int32_t planes, int32_t h_in, int32_t w_in, int32_t h_out, int32_t w_out,
              int32_t antialias, void* stream);

/* ------------------------------------------------------------------------
 * fp32 validation path — NVPrecond(use_fp16=False) / forward(force_fp32=True) (training/models.py:632,697).
 * CUDA-core kernels on NHWC fp32 activations [B*H*W][C], one per reference op, unfused; meets the north star's
 * "fp32 mode" parity bound (rel-L2 <= 1e-4), which the tensor cores cannot (tf32 keeps 10 mantissa bits).
 * For validation, not speed.  Weights come from vb_weight_prep with dst_dtype VB_F32 ([cout][cin*taps], OIHW order).
 * ------------------------------------------------------------------------ */
typedef struct vb_f32_conv_desc {
  const float* x; /* [B*H*W][cin] */
  const float* w; /* [cout][cin*taps], k = ci*taps + tap (MPConv's F.conv2d, models.py:126) */
  float* out;     /* [B*H*W][ldo], cout columns written */
  int32_t B, H, W, cin, cout, taps, ldo;
} vb_f32_conv_desc;
int vb_f32_conv(const vb_f32_conv_desc* d, void* stream);

enum { VB_F32_ACT = 0, VB_F32_SUM = 1, VB_F32_CAT = 2, VB_F32_DOWN = 3, VB_F32_UP = 4, VB_F32_QKV = 5, VB_F32_PRECOND_IN = 6 };
enum { VB_F32_NORM = 1, VB_F32_MOD = 2, VB_F32_SILU = 4 };
typedef struct vb_f32_op_desc {
  const float* a;
  const float* b;   /* SUM: second operand or NULL; CAT: second source; PRECOND_IN: conditioning image (NCHW) or NULL */
  const float* b2;  /* PRECOND_IN: noise added to the conditioning image as wb * noise, or NULL */
  const float* mod; /* ACT+MOD: [B][mod_stride] per-channel gains; PRECOND_IN: sigma [B*mod_stride] or NULL */
  float* out;
  float* out2;      /* QKV: destination of part 1 / 2 */
  float* out3;
  int64_t img_stride; /* PRECOND_IN: floats between consecutive images of a */
  int32_t kind, flags; /* VB_F32_* ; ACT: VB_F32_NORM | VB_F32_MOD | VB_F32_SILU */
  int32_t B, H, W;     /* extents of OUT */
  int32_t ca, cb;      /* channels of a (and of out, except CAT: ca + cb) and of b */
  int32_t mod_stride;
  int32_t heads, parts, head_dim, seg_div; /* QKV */
  int32_t part_seq[3], part_off[3];
  float wa, wb, clip;  /* SUM/CAT weights; clip <= 0: none; PRECOND_IN: wa = sigma_data, wb = noisy_sr */
} vb_f32_op_desc;
int vb_f32_op(const vb_f32_op_desc* d, void* stream);
/* y[b][s][h*D + d] = softmax(q k^T / sqrt(D)) v, q [B*heads][sq][D], k/v [B*heads][sk][D], plus zero_keys all-zero keys. */
int vb_f32_attn(const float* q, const float* k, const float* v, float* y, int32_t B, int32_t heads, int32_t sq, int32_t sk,
                int32_t head_dim, int32_t zero_keys, void* stream);

/* Pixel codec (training/encoders.py:58-62). */
int vb_encode_u8(const uint8_t* src, float* dst, int64_t n, void* stream); /* x/127.5 - 1 */
int vb_decode_u8(const float* src, uint8_t* dst, int64_t n, void* stream); /* clip(x*127.5+128) */

/* ------------------------------------------------------------------------
 * Plans: a recorded sequence of the ops above over fixed buffers, replayed per
 * denoiser call (optionally as one CUDA graph).  TMA descriptors are encoded
 * once when an op is added.
 * ------------------------------------------------------------------------ */
typedef struct vb_plan vb_plan;
int vb_plan_create(vb_plan** out);
void vb_plan_destroy(vb_plan* p);
int vb_plan_add_conv(vb_plan* p, const vb_conv_desc* d);
int vb_plan_add_attn(vb_plan* p, const vb_attn_desc* d);
int vb_plan_add_eltwise(vb_plan* p, const vb_ew_desc* d);
int vb_plan_add_embed(vb_plan* p, const vb_emb_desc* d);
int vb_plan_add_precond_in(vb_plan* p, const vb_precond_in_desc* d);
int vb_plan_add_precond_out(vb_plan* p, const vb_precond_out_desc* d);
int vb_plan_add_heun(vb_plan* p, const vb_heun_desc* d);
int vb_plan_num_ops(const vb_plan* p);
/* Run ops [first, last) eagerly on stream. last < 0 means "to the end". */
int vb_plan_run(vb_plan* p, int first, int last, void* stream);
/* Capture the whole plan into a CUDA graph (once), then launch it. */
int vb_plan_launch_graph(vb_plan* p, void* stream);
/* Same for ops [first, last) only; one graph is kept per distinct range.  Used for no_time_enc feature caching
 * (generate_images.py:52-57): the source-view encoder runs once per batch, the denoising UNet once per sampler step. */
int vb_plan_launch_graph_range(vb_plan* p, int first, int last, void* stream);
/* Algorithmic work of one plan run: kind 0 = GEMM+attention FLOPs, 1 = kernel launches. */
double vb_plan_query(const vb_plan* p, int kind);

/* ------------------------------------------------------------------------
 * Whole-call entry point: one NVPrecond.forward — reference training/models.py:589-689 (current tree),
 * experiments/code/training/models.py:547-638 (snapshot) — on a recorded plan, with plain device pointers.
 * The host that BUILT the plan (vivid_b200/engine.py walks the module tree) registers the plan's persistent input /
 * output buffers once; any host can then run denoiser calls without knowing the op sequence.
 * ------------------------------------------------------------------------ */
typedef struct vb_io_desc {
  float* in_x;     /* [n_x, 3, R, R] noisy images (dual-source: 2B interleaved)                     required */
  float* in_src;   /* [n_x, 3, R, R] source views; NULL for nets without source-view encoder                 */
  float* in_sigma; /* [n_x] noise levels                                                            required */
  float* in_geom;  /* [n_x, geom_dim] pose vectors (zeros when the call passes none)                required */
  float* in_cond;  /* [B, 3, R, R] SR conditioning image; NULL unless super_res                              */
  float* in_noise; /* [B, 3, R, R] SR conditioning noise (the kernel forms cond + noisy_sr * noise)          */
  float* out_d;    /* [B, 3, R, R] D_x                                                              required */
  int64_t n_x, n_out, img_elems /* 3*R*R */, geom_dim;
  int64_t workspace_bytes; /* device bytes the plan holds (weights, activations, I/O): reported by vb_workspace_bytes */
} vb_io_desc;
int vb_plan_bind_io(vb_plan* p, const vb_io_desc* io);
/* D_out = D(x; sigma | src, geometry[, cond]) for the plan's batch.  All pointers are DEVICE pointers to contiguous fp32;
 * sigma_n is 1 (broadcast) or n_x; geometry may be NULL (zeros) and geometry_rows is 1 (broadcast) or n_x; src is ignored by
 * nets without encoder; cond and noise are required for super_res plans (noise: N(0,1) drawn by the caller, the reference
 * draws it with torch.randn_like on every call, experiments/code/training/models.py:608-611).  Copies the inputs into the
 * plan's buffers, replays the plan as one CUDA graph and copies D_x out, all on `stream`, no host synchronisation. */
int vb_denoise(vb_plan* p, const float* src, const float* x, const float* sigma, int32_t sigma_n, const float* geometry,
               int32_t geometry_rows, const float* cond, const float* noise, float* D_out, void* stream);
int64_t vb_workspace_bytes(const vb_plan* p);

/* ------------------------------------------------------------------------
 * Whole-sampler entry point — the Heun loop of edm_sampler (generate_images.py:72-118, deterministic branch: S_churn = 0;
 * snapshot experiments/code/generate_images.py:58-91) over plans with bound I/O, for any host:
 *   x = noise * t_0;  per step:  d = (x - D(x; t_i)) / t_i,  x' = x + (t_{i+1} - t_i) d,  and unless t_{i+1} = 0 the 2nd-order
 *   correction with D(x'; t_{i+1});  D = lerp(D_gnet, D_net, guidance) when guidance != 1 (autoguidance, :57-62).
 * vb_plan_set_inputs uploads what stays constant over a sampler call (source views, pose vectors, SR conditioning image) into
 * the plan's buffers once; vb_sample then enqueues 2*num_steps - 1 graph replays per net with a vb_heun pass after each, whose
 * outputs are the next replay's x / sigma inputs.  Nothing synchronises with the host.
 * ------------------------------------------------------------------------ */
int vb_plan_set_inputs(vb_plan* p, const float* src, const float* geometry /* or NULL: zeros */, int32_t geometry_rows /* 1 or n_x */,
                       const float* cond /* super_res only */, void* stream);
/* super_res plans: fills dst[0..n) with N(0,1) samples on `stream` before every call of the net (the reference draws them
 * with torch.randn_like from the global generator, experiments/code/training/models.py:608-611); returns 0 on success. */
typedef int (*vb_noise_fn)(void* user, float* dst, int64_t n, void* stream);
typedef struct vb_sample_desc {
  vb_plan* net;           /* plan of the main net */
  vb_plan* gnet;          /* plan of the guiding net (same batch / resolution, not super_res); ignored when guidance == 1 */
  const float* noise;     /* DEVICE [B,3,R,R]: N(0,1) latents (generate_images.py:299-300) */
  const float* t_steps;   /* HOST [num_steps + 1]: the noise levels, t_steps[num_steps] = 0 (generate_images.py:68-70) */
  float* workspace;       /* DEVICE, vb_sample_workspace_bytes(net) bytes */
  float* x_out;           /* DEVICE [B,3,R,R]: the sample */
  vb_noise_fn sr_noise;   /* required for super_res plans */
  void* sr_noise_user;
  void* side_stream;      /* optional second stream for the guiding net's replays (joined before every update) */
  int32_t num_steps;
  int32_t net_first_op;   /* 0, or the first op after the source-view encoder when the caller has run it once (no_time_enc nets,
                           * generate_images.py:52-57) */
  float guidance;
} vb_sample_desc;
int64_t vb_sample_workspace_bytes(const vb_plan* p);
int vb_sample(const vb_sample_desc* d, void* stream);

/* ------------------------------------------------------------------------
 * Plan recording inside the library (a plan created from a net description plus a name -> pointer table of
 * the weights).
 * vb_net_plan_create walks the reference's network topology itself — UNet / XAttnUNet / SRXAttnUNet / UNetEncoder
 * constructors, training/models.py:340-383, 438-480, 523-534, 575-582; NVPrecond.forward :628-689 (snapshot
 * experiments/code/training/models.py:581-638) — from a description of the constructor arguments and a table of the
 * parameters under their reference state_dict names, prepares the weights (vb_weight_prep), allocates every buffer,
 * tunes and records the ops and binds the I/O, so that a host without Python can go from a checkpoint to vb_denoise /
 * vb_sample.  It records the same op sequence as vivid_b200/engine.py (tests/test_netplan.py compares the two op for op).
 * ------------------------------------------------------------------------ */
typedef struct vb_unet_desc {
  int32_t img_resolution;
  int32_t in_channels;    /* img_channels + 1 (the ones channel), + img_channels more for the SR UNet's conditioning image */
  int32_t out_channels;   /* 3; 0 = UNetEncoder (no out_conv, trailing decoder blocks without attention dropped, :523-534) */
  int32_t model_channels;
  int32_t num_levels;
  int32_t channel_mult[8];
  int32_t num_blocks;
  int32_t num_attn_res;
  int32_t attn_resolutions[8];
  int32_t extra_attn;     /* block slot that gets attention on every level but the first; -1 = none */
  int32_t channels_per_head;
  int32_t xattn;          /* attention blocks also attend to the source-view features (XAttnBlock) */
  int32_t label_dim;
  int32_t cnoise, cemb;   /* widths of the noise embedding and of the block modulation vector */
  double label_balance, concat_balance, res_balance, attn_balance;
  double clip_act;        /* < 0: no clipping */
} vb_unet_desc;
typedef struct vb_net_desc {
  vb_unet_desc unet;      /* parameters under "unet." */
  vb_unet_desc encoder;   /* parameters under "encoder."; ignored unless has_encoder */
  int32_t has_encoder;    /* 0 for uncond nets */
  int32_t img_resolution;
  int32_t uncond, super_res, dual_source, no_time_enc;
  double sigma_data, noisy_sr;
} vb_net_desc;
typedef struct vb_param {
  const char* name;       /* reference state_dict key, e.g. "unet.enc.64x64_block0.conv_res0.weight" */
  const void* data;       /* DEVICE pointer, contiguous */
  int32_t dtype;          /* VB_F32 or VB_F16 (persisted EMA snapshots are fp16) */
  int32_t ndim;
  int64_t shape[4];
} vb_param;
/* Records the plan of one denoiser call for `batch` target images.  The plan owns its device buffers (freed by
 * vb_plan_destroy); the parameter tensors are read once (weights are normalised, scaled, rounded and repacked into the plan) and
 * need not outlive the call.  Synchronises `stream` (layout tuning times candidate launches). */
int vb_net_plan_create(const vb_net_desc* net, const vb_param* params, int32_t n_params, int32_t batch, void* stream, vb_plan** out);
/* Re-prepares every weight of a plan recorded by vb_net_plan_create from another parameter table of the SAME architecture (the
 * next checkpoint, updated EMA weights): normalise / scale / round / repack into the plan's existing buffers, on `stream`.
 * Nothing is re-recorded or re-tuned and captured graphs stay valid.  All or nothing: the table is validated before the first
 * launch. */
int vb_net_plan_set_weights(vb_plan* p, const vb_param* params, int32_t n_params, void* stream);
/* The plan's persistent I/O buffers (vb_plan_bind_io's descriptor) and the index of the first op after the source-view encoder. */
int vb_plan_get_io(const vb_plan* p, vb_io_desc* out, int32_t* enc_ops);
/* The source-view encoder's cross-attention feature maps of a library-recorded plan, in consumption order: 16-bit NHWC
 * [B][R][R][C] device buffers that ops [0, enc_ops) write and the denoising UNet's K/V GEMMs read — what the reference's
 * return_features / inject_features hand around (training/models.py:664-672; edm_sampler's no_time_enc caching,
 * generate_images.py:52-57: run vb_plan_launch_graph_range(p, 0, enc_ops) once, then (enc_ops, -1) per step). */
int vb_plan_num_features(const vb_plan* p);
int vb_plan_get_feature(const vb_plan* p, int32_t i, void** ptr, int32_t* B, int32_t* R, int32_t* C);
/* Dry run of the recording, with no device: writes one text line per buffer allocation, weight preparation and recorded op —
 * pointers canonicalised as buffer-id/offset and parameter-index/offset — into buf (NUL-terminated, truncated to cap) and returns
 * the full length, or a negative status.  vb_trace_desc formats one descriptor the same way (kind: 0 weight_prep, 1 conv,
 * 2 attn, 3 eltwise, 4 embed, 5 precond_in, 6 precond_out, 7 io): the op-for-op comparison with another recorder. */
int64_t vb_net_plan_trace(const vb_net_desc* net, const vb_param* params, int32_t n_params, int32_t batch, char* buf, int64_t cap);
int64_t vb_trace_desc(int32_t kind, const void* desc, char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* VIVID_B200_H_ */
