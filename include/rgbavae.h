/*
 * rgbavae.h -- C ABI of librgbavae.so, the sm_100a kernel library behind the RGBA-VAE hot path.
 *
 * The reference (jaejung-dev/ragb-vae) has no FFI of its own: its boundary is the diffusers
 * ModelMixin duck-type (vae.encode(x).latent_dist / vae.decode(z).sample).  The
 * Python package ragb_vae_b200 implements that surface and is the ONLY caller of this library
 * (through ctypes).  Each entry point below names the reference call site whose arithmetic it
 * replaces.  All pointers are DEVICE pointers unless stated otherwise; `stream` is a
 * cudaStream_t passed as void*.  Every function returns 0 on success and a non-zero code on
 * failure (1 = CUDA error, 2 = bad argument), in which case rv_last_error() returns a
 * thread-local message.  There is no CPU path.
 *
 * dtype codes: RV_F32 = 0, RV_BF16 = 1.
 * Activations inside the network are NHWC ("pixels x channels"); image tensors at the
 * boundary are NCHW exactly as the reference passes them.
 */
#ifndef RGBAVAE_H_
#define RGBAVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RV_F32 0
#define RV_BF16 1

#define RV_ABI_VERSION 25
#define RV_PROF_CATEGORIES 10

int rv_abi_version(void);
const char* rv_last_error(void);
/* Binds the driver entry point used for TMA descriptors and raises the dynamic shared memory
 * limit of the tcgen05 kernel on the current device.  Call once per process and device. */
int rv_init(void);
/* Number of kernels this library has launched since load (all threads, all streams). */
int64_t rv_launch_count(void);
/* Per-category device timing with CUDA events recorded on the launching stream, for the
 * roofline figures in bench.py.  Categories: 0 tcgen05 conv/GEMM, 1 direct conv, 2 norm+SiLU,
 * 3 softmax, 4 layout, 5 reparam, 6 recon loss, 7 composite+PSNR, 8 fused attention, 9 the nearest-x2
 * up-sampling convs (tcgen05, phase-folded: booked at the algorithmic 3x3-on-the-upsampled-grid FLOPs as the
 * reference computes them, executing 4/9 of that).
 * rv_prof_end synchronises the device and fills ms[c] (summed kernel time), launches[c] and
 * work[c] (algorithmic FLOPs for 0/1/8/9, algorithmic bytes for the others). */
int rv_prof_begin(void);
int rv_prof_end(double* ms, int64_t* launches, double* work);

/* ---- convolution / GEMM ---------------------------------------------------------------- */
/* One descriptor drives both convolution paths.  It covers every conv of diffusers'
 * Encoder/Decoder (vae.encode / vae.decode: src/models/rgba_vae.py:277,279;
 * src/training/rgba_vae_stage.py:449,452; src/models/flux_kontext_textalpha.py:331,497):
 * 3x3 pad 1, 3x3 stride 2 with (0,1,0,1) padding, nearest-x2 upsample + 3x3, and 1x1.
 * The same kernels run plain GEMMs (attention projections, QK^T, PV) as 1x1 convs:
 * y[pixel][co] = alpha * sum_k x[pixel][k] * w[co][k] + bias. */
typedef struct rv_conv_desc {
  int32_t n, h, w;           /* input batch and spatial size (of the tensor actually stored) */
  int32_t cin, cout;         /* logical channels */
  int32_t ksize;             /* 1 or 3 */
  int32_t stride;            /* 1 or 2 */
  int32_t pad_lo;            /* zero padding on top/left (1 for 3x3 s1, 0 for the s2 downsampler) */
  int32_t upsample;          /* 1: nearest x2 upsample of the input fused in front of the conv */
  int32_t oh, ow;            /* output spatial size */
  int32_t x_dtype, y_dtype;  /* RV_F32 / RV_BF16 */
  int32_t x_nchw, y_nchw;    /* 1: tensor is NCHW, else NHWC (x_nchw: direct path only) */
  int32_t x_cstride;         /* NHWC pixel pitch of x in elements (>= cin) */
  int32_t y_cstride;         /* NHWC pixel pitch of y / residual in elements (>= cout) */
  int32_t bias_mode;         /* 0 none, 1 per output channel, 2 per output pixel (GEMM row) */
  float   in_scale, in_shift;   /* x*in_scale+in_shift applied when loading (direct path only) */
  float   out_scale, out_shift; /* y*out_scale+out_shift applied after bias and residual */
  int32_t clamp;             /* 1: clamp to [clamp_lo, clamp_hi] after scale/shift */
  float   clamp_lo, clamp_hi;
  float   alpha;             /* accumulator scale (1/sqrt(d) for QK^T), applied before bias */
  int32_t taps_1d;           /* 1 (tensor-core path, ksize 3, stride 1): a 3x1 kernel -- vertical taps only, weights
                              * [cout][3][cin].  With an rv_nchw_to_nhwc_hpack input this is the 3x3 few-channel stem conv. */
} rv_conv_desc;

/* CUDA-core implicit GEMM, fp32 accumulate, any shape / layout / dtype.  Weights are fp32
 * [cout][ksize*ksize][cin] (K-major).  Used for the 4-channel edge layers, for the fp32
 * parity mode (config c1) and as the in-library cross-check of the tensor-core path.
 * residual (optional) has the layout and dtype of y. */
int rv_conv2d_direct(const rv_conv_desc* d, const void* x, const float* w, const float* bias,
                     const void* residual, void* y, void* stream);

/* tcgen05/TMEM implicit GEMM fed by TMA (bf16 in, fp32 accumulate).  x is NHWC bf16 with
 * x_cstride % 8 == 0 and cin % 16 == 0.  Weights are bf16 K-major rows of pitch `w_ld`
 * elements: [cout][taps][cin] for ksize 1/3 (rv_pack_conv_weights), or for upsample=1 the
 * four folded 2x2 phase kernels [cout][4 phases][4 taps][cin].  residual (optional) is NHWC
 * bf16 with pitch y_cstride. */
int rv_conv2d_tc(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld,
                 const float* bias, const void* residual, void* y, void* stream);

/* conv_out of the decoder (the last call inside vae.decode: src/models/rgba_vae.py:279, rgba_vae_stage.py:452,
 * flux_kontext_textalpha.py:497): 3x3 / stride 1 / pad 1, cin = 64 | 96 | 128 NHWC bf16 -> cout <= 5 NCHW (bf16 or fp32) with
 * bias, out_scale / out_shift and clamp of the descriptor (the (y+1)/2 + clamp(0,1) of RgbaVAE.forward).  HBM-bound: the
 * kernel-row index rides in the MMA's N (3 * cout <= 16) so that every input row is multiplied once.  w_taps: bf16
 * [3 (dx)][16 rows: dy * cout + co, zero padded][cin]. */
int rv_conv_out(const rv_conv_desc* d, const void* x, const void* w_taps, const float* bias, void* y, void* stream);

/* rv_conv2d_tc plus the consumer's QwenImageRMS_norm (+SiLU) fused into the epilogue: besides (or, with
 * y == NULL, instead of) the raw output it writes y_act = act(v / max(||v||_2 over cout, 1e-12) * gamma_scaled)
 * as a second NHWC bf16 tensor of the same pitch, gamma_scaled = gamma * sqrt(cout) (fp32 [cout]).  Needs
 * cout <= 256 (one accumulator tile holds the pixel's whole channel vector) and an NHWC bf16 output.
 * Saves the standalone norm kernel's read+write of the tensor (norm1/norm2 of QwenImageResidualBlock). */
int rv_conv2d_tc_norm(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld,
                      const float* bias, const void* residual, void* y, void* y_act,
                      const float* gamma_scaled, int apply_silu, void* stream);

/* rv_conv2d_tc that also leaves the GroupNorm(groups = 32) statistics of its OUTPUT -- the [n][groups][2] fp64 (sum, sum of
 * squares) rv_groupnorm_stats would compute from y, in the layout rv_groupnorm_silu reads -- so that the consumer's GroupNorm
 * (norm1 / norm2 of diffusers' ResnetBlock2D inside vae.encode / vae.decode) needs no statistics pass over the tensor: the
 * epilogue sums the values it stores per (tile, channel group), a finishing kernel adds a sample's tiles in a fixed order.
 * Covers the CTA-pair implicit-GEMM layers with 256 / 512 output channels (NHWC bf16, per-channel bias, optional
 * residual); rv_conv2d_tc_gnstats_scratch_bytes returns the device scratch it needs, or 0 where the layer is not covered
 * (the caller then runs rv_conv2d_tc + rv_groupnorm_stats). */
int64_t rv_conv2d_tc_gnstats_scratch_bytes(const rv_conv_desc* d, int groups);
int rv_conv2d_tc_gnstats(const rv_conv_desc* d, const void* x, const void* w_packed, int64_t w_ld, const float* bias,
                         const void* residual, void* y, int groups, double* stats, void* scratch, int64_t scratch_bytes,
                         void* stream);

/* Packs fp32 weights [cout][cin][ksize][ksize] (PyTorch layout; for the Qwen causal-conv3d
 * the caller passes the live temporal slice w[:, :, kt-1]) into the bf16 K-major matrix
 * rv_conv2d_tc reads: [cout][taps][cin].  With upsample=1 the 3x3 kernel is folded into four
 * 2x2 phase kernels, [cout][16][cin].  The row length is returned through *w_ld. */
int rv_pack_conv_weights(const float* w, int cout, int cin, int ksize, int upsample,
                         void* out_bf16, int64_t* w_ld, void* stream);
/* Same source layout -> fp32 [cout][taps][cin] for rv_conv2d_direct. */
int rv_pack_conv_weights_direct(const float* w, int cout, int cin, int ksize, float* out, void* stream);

/* ---- normalisation + SiLU (residual blocks: diffusers ResnetBlock2D / QwenImageResidualBlock) */
/* y = act(x / max(||x||_2 over C, 1e-12) * sqrt(C) * gamma), per pixel (QwenImageRMS_norm);
 * x, y NHWC dense [pixels][c]; act = SiLU when apply_silu. */
int rv_rmsnorm_silu(const void* x, const float* gamma, void* y, int64_t pixels, int c, int dtype,
                    int apply_silu, void* stream);
/* GroupNorm(groups, eps): rv_groupnorm_stats accumulates sum and sum of squares per (sample,
 * group) into stats [n][groups][2] fp64 (it zeroes the buffer first); rv_groupnorm_silu applies
 * y = act((x-mean)*rstd*gamma+beta).  x, y NHWC dense [n][hw][c]. */
int rv_groupnorm_stats(const void* x, double* stats, int n, int64_t hw, int c, int groups, int dtype,
                       void* stream);
int rv_groupnorm_silu(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                      int n, int64_t hw, int c, int groups, float eps, int dtype, int apply_silu,
                      void* stream);

/* ---- mid-block attention (single head, d = C) ------------------------------------------- */
/* Row softmax of fp32 scores [rows][cols] (already scaled) into probabilities of `dtype`. */
int rv_softmax_rows(const float* s, void* p, int64_t rows, int64_t cols, int64_t ld_s, int64_t ld_p,
                    int dtype, void* stream);

/* Fused single-head attention O = softmax(Q K^T / sqrt(d)) V for d = 384 (QwenImageAttentionBlock) and d = 512 (the
 * diffusers Attention block of the Flux AutoencoderKL; two output-column passes inside the kernel): scores and
 * probabilities stay in TMEM / shared memory.  q, k: bf16 [n_img*tokens][ld_qk] (row pitch in elements; q and k may
 * be column slices of one tensor); vt: bf16 V transposed, [n_img][d][tokens]; out: bf16 [n_img*tokens][ld_out].
 * tokens % 128 == 0. */
int rv_attention(const void* q, const void* k, int64_t ld_qk, const void* vt, void* out, int64_t ld_out,
                 int n_img, int tokens, int d, void* stream);
/* rv_attention that also leaves, per query row, the base-2 log-sum-exp of its scaled scores in lse (fp32
 * [n_img*tokens]): softmax(QK^T/sqrt(d))[i][j] = exp2(q_i.k_j * log2(e)/sqrt(d) - lse[i]).  The training step keeps it so
 * that the backward of the attention block (loss.backward() through scaled_dot_product_attention,
 * src/training/rgba_vae_stage.py:516) recomputes probabilities without a second softmax pass. */
int rv_attention_lse(const void* q, const void* k, int64_t ld_qk, const void* vt, void* out, int64_t ld_out, float* lse,
                     int n_img, int tokens, int d, void* stream);
/* rv_attention_lse with a device workspace of rv_attention_workspace_bytes(tokens, d) bytes (0: this shape uses none; d = 512
 * does).  With it the d = 512 kernel keeps pass 1's probability tiles there and its second output pass streams them back
 * instead of recomputing S and the softmax (64 instead of 96 MMAs per key block).  The first 1024 bytes (slot ownership
 * flags) must be zero before the first call and are zero again whenever no call is in flight; the rest needs no
 * initialisation.  One workspace serves one stream at a time.  workspace == NULL: same as rv_attention_lse.
 * V^T is addressed through two pitches in elements: ld_vt between consecutive d-rows and vt_img_pitch between images --
 * (tokens, d * tokens) for the dense [n_img][d][tokens] of rv_attention, (n_img * tokens, tokens) for a [d][n_img * tokens]
 * matrix that ONE projection GEMM over all images writes. */
int64_t rv_attention_workspace_bytes(int tokens, int d);
int rv_attention_ws(const void* q, const void* k, int64_t ld_qk, const void* vt, int64_t ld_vt, int64_t vt_img_pitch, void* out,
                    int64_t ld_out, float* lse, void* workspace, int64_t workspace_bytes, int n_img, int tokens, int d,
                    void* stream);

/* ---- layout plumbing at the NCHW boundary ----------------------------------------------- */
/* y[n][hw][c_pad] = x[n][c][hw]*scale+shift (extra channels zero). */
int rv_nchw_to_nhwc(const void* x, void* y, int n, int c, int64_t hw, int c_pad, int x_dtype,
                    int y_dtype, float scale, float shift, void* stream);
/* Few-channel stem loader (3*c <= 16): NCHW -> NHWC bf16 with 16 channels per pixel holding the pixel's three
 * HORIZONTAL neighbours, y[n][h][w][dx*c + ch] = x[n][ch][h][w+dx-1]*scale+shift (zero outside the image and past 3c).
 * A 3x3 conv over x is then a 3x1 conv (taps_1d) over y with K = 16 per tap: 3 MMAs and 3 TMA boxes per tile instead
 * of 9 each through 16-channel zero padding. */
int rv_nchw_to_nhwc_hpack(const void* x, void* y, int n, int c, int h, int w, int x_dtype, float scale, float shift,
                          void* stream);
int rv_nhwc_to_nchw(const void* x, void* y, int n, int c, int64_t hw, int x_cstride, int x_dtype,
                    int y_dtype, void* stream);

/* ---- data formats either side of the VAE ----------------------------------- */
/* build_detail_augmented_triplet (src/training/rgba_vae_stage.py:606-625): target NCHW [b][4][hw] in [-1,1] ->
 * out [3b][4][hw] = [target | on black, alpha 1 | on white, alpha 1]. */
int rv_triplet_augment(const void* target, void* out, int b, int64_t hw, int dtype, void* stream);
/* RandomBackgroundBlend._blend_tensor (src/training/rgba_vae_stage.py:85-130) for a whole batch: x NCHW [n][4][hw] in
 * [0,1]; samples with mask[n] != 0 are composited over the opaque colour colors[n][3] (fp32) and get alpha 1, the others
 * are copied.  colors and mask are device arrays. */
int rv_background_blend(const void* x, const float* colors, const unsigned char* mask, void* y, int n, int64_t hw,
                        int dtype, void* stream);
/* FluxPipeline._pack_latents / _unpack_latents (src/models/flux_kontext_textalpha.py:334-349): (n,c,h,w) <->
 * (n, (h/2)(w/2), 4c) with feature order (c, dy, dx).  unpack = 0: y = (pack(x) - shift) * scale;
 * unpack = 1: y = unpack(x) * scale + shift. */
int rv_pack_latents(const void* x, void* y, int n, int c, int h, int w, int dtype, float shift, float scale,
                    int unpack, void* stream);
/* blend_v / blend_h of diffusers' tiled encode / decode (enable_tiling, src/training/rgba_vae_stage.py:296-299):
 * the first `extent` rows (vertical = 1) or columns of every plane of b become a linear ramp from the last `extent`
 * rows / columns of a to b.  a: [planes][ah][aw], b: [planes][bh][bw], modified in place. */
int rv_blend_tiles(const void* a, void* b, int planes, int ah, int aw, int bh, int bw, int extent, int vertical,
                   int dtype, void* stream);
/* uint8 RGBA, HWC per image (inference_rgba_flux.py:15-26) <-> NCHW float: y = x/255*scale+shift; and back
 * (clamp to [0,1], *255, truncate). */
int rv_rgba_u8_to_nchw(const void* x_u8, void* y, int n, int64_t hw, int y_dtype, float scale, float shift,
                       void* stream);
int rv_nchw_to_rgba_u8(const void* x, void* y_u8, int n, int64_t hw, int x_dtype, void* stream);

/* ---- posterior (diffusers DiagonalGaussianDistribution; src/models/rgba_vae.py:278) ------ */
/* moments NCHW [n][2*zc][hw]; noise / z NCHW [n][zc][hw].
 * z = (mean + exp(0.5*clamp(logvar,-30,20))*noise - z_shift) * z_scale   (z_shift 0 / z_scale 1
 * for the plain sample(); the Flux latent normalisation of flux_kontext_textalpha.py:330-332
 * otherwise).  kl_out (optional, [n] fp32) = 0.5*sum(mean^2 + var - 1 - logvar). */
int rv_reparam(const void* moments, const void* noise, void* z, float* kl_out, int n, int zc,
               int64_t hw, int dtype, float z_shift, float z_scale, void* stream);

/* ---- losses and validation metrics -------------------------------------------------------*/
/* The three reductions below read both images once and finish in the same launch (the last block of a sample sums
 * the per-block fp64 partials in a fixed order: deterministic, and a sample's bits do not depend on its batch).
 * `partial` is caller-provided scratch of n * rv_reduce_blocks(hw) * K doubles (K per function below).  Calls on
 * one stream are ordered; calls on different streams are independent. */
/* AlphaVaeLoss.reconstruction_loss (src/models/losses.py:67-83): pred/target NCHW [n][4][hw] in
 * [-1,1].  per_sample[n] receives the per-sample SUM of the loss map (naive: over 4 channels).  K = 1. */
int rv_recon_loss(const void* pred, const void* target, const float* eb_host, const float* eb2_host,
                  int naive_mse, float* per_sample, double* partial, int n, int64_t hw, int dtype,
                  void* stream);
/* composite_over_background + compute_psnr + alpha MAE (src/models/rgba_vae.py:75-84,
 * src/training/rgba_vae_stage.py:712-715,742-753) in one pass over recon/target NCHW [n][4][hw]
 * in [0,1].  bgs_host: nbg (<= 4) RGB triples (HOST pointer).  out: [n][nbg+1] fp32: PSNR per
 * background, then alpha MAE.  K = 5. */
int rv_composite_psnr(const void* recon, const void* target, const float* bgs_host, int nbg,
                      float* out, double* partial, int n, int64_t hw, int dtype, void* stream);
/* Every term RgbaVAE.loss weighs (src/models/rgba_vae.py:283-316) in one pass over recon/target NCHW [n][4][hw] in
 * [0,1].  out [n][6] fp32 per-sample SUMS: [0] the AlphaVAE map (:324-336) on the pair rescaled to [-1,1],
 * [1] (recon_rgb - target_rgb)^2 (the use_naive_mse branch, :291-293), [2] / [3] squared error of the composites over
 * white / black (:298-306), [4] (alpha_r - alpha_t)^2 (:308-309), [5] |alpha_r - alpha_t| (:311-312).  K = 6. */
int rv_rgba_loss_terms(const void* recon, const void* target, const float* eb_host, const float* eb2_host, float* out,
                       double* partial, int n, int64_t hw, int dtype, void* stream);
/* number of partial blocks per sample the two reductions above use for `hw` pixels */
int rv_reduce_blocks(int64_t hw);

/* ---- training-step building blocks (src/training/rgba_vae_stage.py:433-523) ------------------ */
/* d loss / d pred of AlphaVaeLoss.reconstruction_loss (losses.py:67-83); grad_scale = upstream gradient times the
 * reduction factor (1/(B*3*HW) for reduce_mean -- 1/(B*4*HW) for the naive MSE -- else 1/B).  NCHW [n][4][hw].
 * clamp_lo < clamp_hi: pred is the decoder's clamped output, the gradient is zero where it sits on either bound. */
int rv_recon_loss_bwd(const void* pred, const void* target, const float* eb_host, const float* eb2_host,
                      int naive_mse, float grad_scale, float clamp_lo, float clamp_hi, void* dpred, int n, int64_t hw,
                      int dtype, void* stream);
/* Backward of posterior.sample() (+ kl_weight * posterior.kl()): dmoments NCHW [n][2*zc][hw] from dz [n][zc][hw]
 * (dz / noise may both be NULL for the KL term alone); the logvar gradient is zero outside the clamp range. */
int rv_reparam_bwd(const void* moments, const void* noise, const void* dz, void* dmoments, int n, int zc,
                   int64_t hw, int dtype, float kl_weight, void* stream);
/* KL(posterior || frozen reference posterior), the reference-KL term of the train step (rgba_vae_stage.py:489-508;
 * DiagonalGaussianDistribution.kl(other)): kl_out[n] (fp32) += per-sample sum; dmoments (optional, same layout and
 * dtype as moments) = weight * d kl / d moments.  moments / ref_moments NCHW [n][2*zc][hw]. */
int rv_kl_ref(const void* moments, const void* ref_moments, float* kl_out, void* dmoments, int n, int zc, int64_t hw,
              int dtype, float weight, void* stream);
/* Backward of rv_rmsnorm_silu: dx, and dgamma[c] += dgamma_scale * d loss / d (gamma*sqrt(C)) (fp32, ACCUMULATED;
 * dgamma_scale = sqrt(C) gives d loss / d gamma, so dgamma may point straight into a gradient buffer).
 * gamma_scaled = gamma * sqrt(C).  c = 3 * 2^k 16-byte chunks (96, 192, 384 in bf16).  add (optional, same layout as dx) is
 * added to dx: the gradient of the skip branch that meets the normalised one at this point. */
int rv_rmsnorm_silu_bwd(const void* x, const float* gamma_scaled, const void* dy, const void* add, void* dx,
                        float* dgamma, float dgamma_scale, int64_t pixels, int c, int dtype, int apply_silu, void* stream);
/* Backward of rv_groupnorm_stats + rv_groupnorm_silu (GroupNorm(32) + SiLU of the Flux AutoencoderKL blocks; the VAE that
 * configs/flux_vae.yaml:73 trains).  x / dy / dx / add NHWC [n][hw][c]; stats = the forward's [n][groups][2] raw sums;
 * dgamma / dbeta [c] fp32 are ACCUMULATED (they may point into a gradient buffer); add (optional) is added to dx (the skip
 * branch's gradient); scratch: n * c * 2 doubles. */
int rv_groupnorm_silu_bwd(const void* x, const double* stats, const float* gamma, const float* beta, const void* dy,
                          const void* add, void* dx, float* dgamma, float* dbeta, double* scratch, int n, int64_t hw, int c,
                          int groups, float eps, int dtype, int apply_silu, void* stream);
/* Weight (and bias) gradient of a stride-1 3x3 or 1x1 convolution on the tensor cores: x NHWC bf16 [n][h][w][cin],
 * dy NHWC bf16 [n][h][w][cout].  Element (co, ci, tap) is ACCUMULATED (fp32 atomics) at
 * dw[co*dw_co_stride + ci*dw_ci_stride + tap*dw_tap_stride], so the gradient can land directly in the parameter's own
 * layout ([cout][cin][k][k]: strides cin*k*k, k*k, 1; the last temporal slice of a [cout][cin][3][3][3] causal kernel:
 * strides 27*cin, 27, 1 from dw + 18).  dbias [cout_valid] optional, accumulated.  Only co < cout_valid and
 * ci < cin_valid are written (channel-padded stems).  cin % 16 == 0, cout % 8 == 0.  pad = ksize/2 for a 'same' conv;
 * pad = 0 with a zero-inserted dy gives the weight gradient of the stride-2 (0,1,0,1)-padded down-sampling conv. */
int rv_conv2d_wgrad(const void* x, const void* dy, float* dw, int64_t dw_co_stride, int64_t dw_ci_stride,
                    int64_t dw_tap_stride, float* dbias, int n, int h, int w, int cin, int cout, int cin_valid,
                    int cout_valid, int ksize, int pad, void* stream);
/* Packed bf16 weights of the convolution that maps dY to dX (data gradient of a stride-1 conv): reads the parameter in
 * place (bf16, element (co, ci, tap) at w[co*w_co_stride + ci*w_ci_stride + tap]) and writes
 * out[ci][tap'*cout_pad + co] = W[co][ci][taps-1-tap'] (zero for co >= cout): taps flipped, channels transposed. */
int rv_pack_dgrad_weights(const void* w, int64_t w_co_stride, int64_t w_ci_stride, void* out, int cout, int cin,
                          int cout_pad, int ksize, void* stream);
/* Spatial helpers of the backward pass, NHWC bf16, c % 8 == 0.  mode 0: zero-insert x2 (dY of a stride-2 conv onto
 * the input grid); mode 1: nearest x2 upsample; mode 2: 2x2 sum pool (backward of the nearest upsample). */
int rv_resample2x(const void* x, void* y, int n, int h, int w, int c, int mode, void* stream);
/* y = a + b over n bf16 elements (n % 8 == 0): gradient accumulation where two branches meet. */
int rv_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream);
/* Softmax backward for a block of `rows` query rows: dS = P * (dP - rowsum(dP*P)) * scale as bf16 [rows][cols] and as
 * (optionally, ds_t != NULL) its transpose written into ds_t [cols][ld_t] at column offset row0.  p bf16 [rows][cols],
 * dp fp32 [rows][cols]. */
int rv_softmax_bwd(const void* p, const float* dp, void* ds, void* ds_t, int64_t rows, int64_t cols, int64_t ld_t,
                   int64_t row0, float scale, void* stream);
/* The two score-matrix GEMMs of the attention backward with their elementwise step fused into the tensor-core
 * epilogue (rv_conv2d_tc run as a plain GEMM y[row][col] = sum_k x[row][k] w[col][k], d = a 1x1 descriptor with
 * n = h = 1, w = rows; bf16 output), so that fp32 scores never reach HBM:
 *   mode 1: y = exp2(alpha * acc - rowstat[row])     P from Q K^T with alpha = log2(e)/sqrt(d), rowstat = rv_attention_lse's lse
 *   mode 2: y = mul_in * (alpha * acc - rowstat[row])  dS from dO V^T with alpha = 1/sqrt(d), mul_in = P (bf16, pitch
 *           y_cstride), rowstat = rv_rowdot(dO, O) / sqrt(d)
 * mul_in must be NULL in mode 1. */
int rv_gemm_rowstat(const rv_conv_desc* d, const void* x, const void* w, int64_t w_ld, const float* rowstat, int mode,
                    const void* mul_in, void* y, void* stream);
/* out[row] = scale * sum_c a[row][c] * b[row][c] for bf16 matrices of `cols` (% 8 == 0) columns and row pitches ld_a, ld_b:
 * the delta = rowsum(dO * O) = rowsum(dP * P) term of the softmax backward. */
int rv_rowdot(const void* a, const void* b, int64_t rows, int cols, int64_t ld_a, int64_t ld_b, float scale, float* out,
              void* stream);
/* *out += sum(g^2) over a flat fp32 gradient buffer (accelerator.clip_grad_norm_, rgba_vae_stage.py:520-521).
 * Deterministic (fixed partition and summation order, fp64 partials in `scratch`, a device buffer of
 * rv_grad_sqnorm_scratch_bytes() bytes): data-parallel replicas get bit-identical clip factors. */
int rv_grad_sqnorm_scratch_bytes(void);
int rv_grad_sqnorm(const float* g, int64_t n, float* out, void* scratch, void* stream);
/* torch.optim.AdamW step (rgba_vae_stage.py:321-331, 522) over flat fp32 buffers, fused with the gradient scaling
 * of a data-parallel SUM all-reduce (grad_scale = 1/world) and with clip_grad_norm_ (sqnorm = device pointer to the
 * squared norm of the UNscaled gradient, max_norm <= 0 disables).  p_bf16 (optional) receives the bf16 copy.
 * Bias corrections come from the host step count `step` (>= 1), or from `state` (see rv_adamw_advance) when non-NULL. */
int rv_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, const float* state, float grad_scale,
                  const float* sqnorm, float max_norm, void* stream);
/* Device-resident step counter for CUDA-graph replay of the optimizer: state = [t, 1-beta1^t, 1-beta2^t] (fp32 [3],
 * zero-initialised); call once per step ahead of rv_adamw_step(..., step = 0, state, ...). */
int rv_adamw_advance(float* state, float beta1, float beta2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGBAVAE_H_ */
