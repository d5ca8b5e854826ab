/* libl2i -- C ABI of the B200-native layout2img hot path.
 *
 * This is the drop-in seam the reference declares in setup.py:46-54 (the CUDAExtension
 * `model.roi_layers._C`, whose sources are absent upstream) widened to every operator on the
 * G+D train-step path.  Plain C: borrowed raw device pointers, explicit sizes, an explicit
 * cudaStream_t (passed as void*), int return codes.  No function allocates, synchronises or
 * keeps a pointer past its return.  All tensors are fp32 NHWC unless stated; "pair" tensors
 * are two bf16 tensors (hi, lo) with x ~= hi + lo and a channel count padded to a multiple
 * of 8.  Return 0 on success, <0 on error (l2i_last_error() holds the message).
 *
 * There is no CPU fallback: every entry point launches sm_100a kernels.
 */
#ifndef L2I_H_
#define L2I_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define L2I_OK 0
#define L2I_ERR_BAD_ARG (-1)
#define L2I_ERR_UNSUPPORTED (-2)
#define L2I_ERR_LAUNCH (-3)
#define L2I_ERR_DRIVER (-4)

int l2i_version(void);
const char* l2i_last_error(void);
/* number of libl2i kernels launched by this process since the last reset (bench.py's gpu_launches). */
int l2i_launch_count(int reset);

/* ---- convolution (replaces nn.Conv2d -> cuDNN; reference model/resnet_generator_app_v2.py:633-639,
 *      681-686 and model/rcnn_discriminator_app.py:10-15,297-326) ------------------------------- */

/* W [cout][cin][kh][kw] fp32 (torch layout), taps = kh*kw in {1, 9}; optional device scalar sigma
 * divides W (spectral norm).  Writes the forward pair [cout][taps][cin_pad] and, when dg_hi != NULL,
 * the data-gradient pair [cin][taps][cout_pad] (flipped + transposed filter). */
int l2i_conv_weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* fwd_hi, void* fwd_lo,
                         int cin_pad, void* dg_hi, void* dg_lo, int cout_pad, void* stream);

/* x [N,H,W,C] fp32 -> pair [N,H<<up2,W<<up2,cpad]; optional ReLU, optional nearest x2 up-sampling. */
int l2i_act_split(const float* x, int N, int H, int W, int C, int relu, int up2, void* hi, void* lo, int cpad,
                  void* stream);

/* Residual-block operand preparation: one read of x [N,H,W,C] fp32 gives pair a = relu_a ? relu(x) : x at
 * full resolution and, for b_mode 1 / 2, pair b = b_scale * x / b_scale * (2x2 sum of x) (never ReLU'd):
 * b_scale 0.25 = avg_pool2d (input of the pooled 1x1 shortcut of rcnn_discriminator_app.py:294-344), 1.0 = the
 * backward of nearest x2 up-sampling (resnet_generator_app_v2.py:665-670).  a_hi / a_lo may be NULL (pair b only). */
int l2i_act_split2(const float* x, int N, int H, int W, int C, int relu_a, void* a_hi, void* a_lo, int b_mode,
                   float b_scale, void* b_hi, void* b_lo, int cpad, void* stream);
/* Gradient arriving at a block output, g [N,H,W,C] fp32: pair (lo_hi, lo_lo) = g (nullable), pair (up_hi, up_lo)
 * = up_scale * nearest_x2(g) at [N,2H,2W,cpad] (nullable; up_scale 0.25 = backward of avg_pool2d), and
 * colsum [C] = sum over pixels of g (the bias gradients). */
int l2i_grad_split(const float* g, int N, int H, int W, int C, void* lo_hi, void* lo_lo, float up_scale, void* up_hi,
                   void* up_lo, float* colsum, int cpad, void* stream);
/* colsum [C] = sum over pixels of (hi + lo) [pixels, cpad]. */
int l2i_pair_colsum(const void* hi, const void* lo, long long pixels, int C, int cpad, float* colsum, void* stream);

/* v = (conv(x, w) + bias) * out_scale, stride 1, "same" padding; H, W powers of two.
 * x pair [N,H,W,cin_pad]; w pair [cout][taps][cin_pad]; bias [cout] or NULL.
 * mask_hi (nullable): bf16 [N,H,W,mask_cpad]; v is zeroed where mask_hi <= 0 -- the ReLU derivative taken
 *   from the saved (ReLU'd) activation pair when this call computes a data gradient.
 * pool: 0 none; 1 = 2x2 average, 2 = 2x2 sum of v, stored at (H/2, W/2) (F.avg_pool2d(.,2) of the D blocks;
 *   the backward of the G blocks' nearest x2 up-sampling).
 * y = pool(v) + res_scale * residual; residual (nullable) has the stored resolution, or half of it, read
 *   with nearest x2 up-sampling, when res_up2 = 1 (pool must be 0 then).
 * Outputs (either may be NULL, not both): out fp32 [N,Ho,Wo,cout]; pair [N,Ho,Wo,cout_pad] of
 * relu_split ? relu(y) : y.  The data gradient is the same call with the dgrad weight pair. */
int l2i_conv2d_fwd(int N, int H, int W, int cin_pad, int cout, int taps, const void* x_hi, const void* x_lo,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, int res_up2,
                   float res_scale, float out_scale, const void* mask_hi, int mask_cpad, int pool, float* out, void* out_hi,
                   void* out_lo, int cout_pad, int relu_split, void* stream);

/* 3x3 convolutions with <= 4 channels on one side (the discriminator's first convolution 3 -> 64, the generator's RGB
 * head 64 -> 3) as 1x1 convolutions over a 9C-channel im2col tensor (channel order c*9 + tap = torch's weight order, so
 * W.view(cout, 9C) is the 1x1 weight); d(tap) = (tap/3 - 1, tap%3 - 1), sign = +1 / -1:
 *   l2i_im2col3_pair: pair[p, c*9+tap] = x[p + sign d(tap), c] (zero outside), x fp32 [N,H,W,C], pair [N,H,W,cpad] with
 *                     cpad >= 9C (channels >= 9C zero); colsum [C] (nullable) = sum over pixels of x (a bias gradient)
 *   l2i_col2im3:      out[q, c] = sum_tap col[q - sign d(tap), c*9+tap] + bias[c] + res_scale * residual[q or q/2, c]
 *                     (the adjoint of im2col(sign)); col fp32 [N,H,W,ldc], out / residual fp32 [N,H,W,C] ([N,H/2,W/2,C]
 *                     read with nearest x2 up-sampling when res_up2). */
int l2i_im2col3_pair(const float* x, int N, int H, int W, int C, int sign, void* hi, void* lo, int cpad, float* colsum,
                     void* stream);
int l2i_col2im3(const float* col, int ldc, int N, int H, int W, int C, int sign, const float* bias, const float* residual,
                int res_up2, float res_scale, float* out, void* stream);

/* dw [cout][taps][cin] fp32 = sum over pixels of dy (x) shifted x.  dy pair [N,H,W,cout_pad],
 * x pair [N,H,W,cin_pad]. */
int l2i_conv2d_wgrad(int N, int H, int W, int cin, int cin_pad, int cout, int cout_pad, int taps, const void* dy_hi,
                     const void* dy_lo, const void* x_hi, const void* x_lo, float* dw, void* stream);

/* ---- batch-norm statistics + ISLA modulation (reference model/norm_module.py:152-189,
 *      model/sync_batchnorm/batchnorm.py:48-53,113-125) ------------------------------------------- */

/* sums[2c] = sum x, sums[2c+1] = sum x^2 over `pixels` rows of x [pixels, C] (fp64 accumulators). */
int l2i_bn_stats(const float* x, long long pixels, int C, double* sums, void* stream);
/* mean_invstd [2C] = (mean, 1/sqrt(biased var + eps)); running stats (nullable) updated with
 * momentum and the unbiased variance, as F.batch_norm does in training mode. */
int l2i_bn_finalize(const double* sums, double count, int C, float eps, float momentum, float* running_mean,
                    float* running_var, float* mean_invstd, void* stream);
/* eval mode: mean_invstd from the running statistics. */
int l2i_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean_invstd,
                      void* stream);
/* out = (sum_o m_o gamma_o/(sum_o m_o + 1e-6) + 1) * xhat + sum_o m_o beta_o/(sum_o m_o + 1e-6).
 * x [B,H,W,C]; mask [B,H,W,O] (pixel-major); gamma, beta [B,O,C]; O <= 32.  O == 0: plain batch norm with the optional
 * affine (aff_w, aff_b) and the optional per-(image, channel) factor chan_scale [B,C] (the scaled Dropout2d keep-mask of
 * the PSP head, resnet_generator_app_v2.py:736: relu(y) * k == relu(y * k) for k >= 0):  y = (xhat aff_w + aff_b) chan_scale.
 * Outputs: out fp32 [B,H,W,C] (nullable; ReLU'd when relu & 2) and/or the pair [B,H<<up2,W<<up2,cpad] of
 * (relu & 1) ? relu(y) : y, nearest-upsampled when up2. */
int l2i_isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
                 const float* aff_w, const float* aff_b, const float* chan_scale, int B, int H, int W, int C, int O,
                 float* out, void* hi, void* lo, int cpad, int relu, int up2, void* stream);
/* Backward of l2i_isla_fwd followed by (relu != 0: ReLU) and (nearest x2): dout [B,H<<up2,W<<up2,C].
 * Writes dx [B,H,W,C]; for O > 0 dmask [B,H,W,O], dgamma, dbeta [B,O,C]; csum [2C] fp64 receives
 * (sum dxhat, sum dxhat*xhat) for O > 0 or (dbias, dweight) of the affine form for O == 0.
 * gbuf is unused (may be NULL; the backward writes no intermediate tensor: two passes over x and dout, 20 B per
 * element).  train = 0 skips the batch-statistics terms (eval-mode BN).
 * phase 0 runs everything; for a batch norm whose statistics span several ranks (the reference's multi-GPU
 * SynchronizedBatchNorm2d, sync_batchnorm/batchnorm.py:90-111) call phase 1 (all reductions; dx untouched),
 * all-reduce csum, then phase 2 (dx) with count = the global pixel count (count <= 0: B*H*W). */
int l2i_isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
                 const float* aff_w, const float* aff_b, const float* chan_scale, const float* dout, int B, int H, int W,
                 int C, int O, int relu, int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta,
                 double* csum, float* dx, int phase, double count, void* stream);

/* ---- layout maps (reference model/resnet_generator_app_v2.py:466-470,697-721, utils/bilinear.py:137-192)
 *      bbox [B*O,4] xywh in [0,1]; maps are [B,O,S,S] unless stated ------------------------------- */
int l2i_bbox_mask(const float* bbox, int BO, int H, int W, float* out, void* stream);
int l2i_masks_to_layout_fwd(const float* bbox, const float* masks, int BO, int M, int S, float* out, void* stream);
int l2i_masks_to_layout_bwd(const float* bbox, const float* dout, int BO, int M, int S, float* dmasks, void* stream);
/* bilinear (align_corners=False) resize of [B,O,hi,wi] to [B,O,h,w] (pixel_major = 0) or [B,h,w,O]
 * (pixel_major = 1); identity copy/transpose when sizes match.  bwd is a deterministic gather. */
int l2i_mask_resize_fwd(const float* in, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* out,
                        void* stream);
int l2i_mask_resize_bwd(const float* dout, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* din,
                        void* stream);
/* out[b,o] = bilinear(bmask -> h) * (1 - a) + sigmoid(stage[b,:,:,y]) * nearest(hard -> h) * a,
 * a = sigmoid(alpha[y]); stage [B,h,w,NC] NHWC, y [B,O] int64, alpha [NC]. */
int l2i_stage_mix_fwd(const float* stage, const int64_t* y, const float* alpha, const float* bmask, const float* hard,
                      int B, int O, int h, int w, int NC, int S, float* out, void* stream);
/* dstage [B,h,w,NC] and dalpha [NC] must be zero-initialised; dsoft [B,O,h,w] = dout * (1 - a). */
int l2i_stage_mix_bwd(const float* stage, const int64_t* y, const float* alpha, const float* bmask, const float* hard,
                      const float* dout, int B, int O, int h, int w, int NC, int S, float* dstage, float* dalpha,
                      float* dsoft, void* stream);

/* Gathered mask head + stage-mask mixing in one kernel (reference resnet_generator_app_v2.py:646/651 and :466-470): only
 * the O class channels an image's objects select are formed, never the 184-channel stage mask.
 *   sel[b,o,p] = bc[y[b,o]] + sum_c Wc[y[b,o], c] * t[b,p,c];   out = bilinear(bmask) (1 - a) + sigmoid(sel) nearest(hard) a,
 *   a = sigmoid(alpha[y[b,o]]).   t [B,h,w,C] (the head's features after BatchNorm / ReLU / Dropout2d), Wc [NC,C] (the 1x1
 *   convolution's weight), bc [NC] (nullable), out / sel [B,O,h,w] (sel is kept for the backward).  O <= 32.
 * bwd: dt [B,h,w,C]; dW [NC,C], db [NC] (nullable), dalpha [NC] (zeroed by the call, class-indexed atomics);
 *   dsoft [B,O,h,w] = dout (1 - a) (feed it to l2i_mask_resize_bwd for d bmask). */
int l2i_class_mix_fwd(const float* t, const float* Wc, const float* bc, const int64_t* y, const float* alpha,
                      const float* bmask, const float* hard, int B, int O, int h, int w, int C, int NC, int S, float* sel,
                      float* out, void* stream);
int l2i_class_mix_bwd(const float* t, const float* Wc, const int64_t* y, const float* alpha, const float* bmask,
                      const float* hard, const float* sel, const float* dout, int B, int O, int h, int w, int C, int NC, int S,
                      float* dt, float* dW, float* db, float* dalpha, float* dsoft, void* stream);

/* ---- mask-regression trunk (reference model/mask_regression.py:66-99): InstanceNorm2d (affine=False, eps) -> ReLU
 *      -> optional bilinear x2 (align_corners=False) of x [N,H,W,C], written as the next convolution's operand
 *      pair [N,H<<up2,W<<up2,cpad]; stats [N,C,2] = (mean, 1/sqrt(var+eps)) is kept for the backward. ---------- */
int l2i_inorm_relu_fwd(const float* x, int N, int H, int W, int C, int up2, float eps, float* stats, void* hi, void* lo,
                       int cpad, void* stream);
/* dx [N,H,W,C] from da [N,H<<up2,W<<up2,C], the gradient w.r.t. the (up-sampled) output of l2i_inorm_relu_fwd. */
int l2i_inorm_relu_bwd(const float* x, const float* stats, const float* da, int N, int H, int W, int C, int up2, float* dx,
                       void* stream);

/* ---- ROIAlign (the operator of the reference's setup.py:46-54 extension `model.roi_layers._C`;
 *      call sites model/rcnn_discriminator_app.py:98-99,139,143; torchvision semantics, aligned=False,
 *      sampling_ratio=0).  feat [N,H,W,C]; rois [K,5] = (image, x0, y0, x1, y1) px; out [K,P,P,C]. */
int l2i_roi_align_fwd(const float* feat, const float* rois, int K, int N, int H, int W, int C, int P, float scale,
                      float* out, void* stream);
int l2i_roi_align_bwd(const float* dout, const float* rois, int K, int N, int H, int W, int C, int P, float scale,
                      float* dfeat, void* stream);
/* Device-side ROI preparation (reference rcnn_discriminator_app.py:402-417,131-146), no host round trip: bbox [B,O,4]
 * xywh in [0,1], label [B*O].  Writes ALL B*O rows of rois [B*O,5] = (image, x0, y0, x1, y1) px, y_sorted, level and
 * perm (source row): first the large ROIs (level 0), then the small ones (level 1: width < small_thresh and height <
 * small_thresh), each in (b, o) order -- the reference's order -- then the dropped rows (label == 0, level 2).
 * counts [2] = (n_large, n_small).  Bit-exact with the reference's float arithmetic. */
int l2i_roi_prepare(const float* bbox, const int64_t* label, int B, int O, float img_size, float small_thresh, float* rois,
                    int64_t* y_sorted, int32_t* level, int32_t* perm, int32_t* counts, void* stream);
/* ROIAlign of the two-level object path in one launch: row k samples feat_l [N,Hl,Wl,C] at scale_l (level 0), feat_s
 * [N,Hs,Ws,C] at scale_s (level 1), or is zero (level 2).  out [K,P,P,C].  bwd zero-fills both feature gradients. */
int l2i_roi_align2_fwd(const float* feat_l, int Hl, int Wl, float scale_l, const float* feat_s, int Hs, int Ws, float scale_s,
                       const float* rois, const int32_t* level, int K, int N, int C, int P, float* out, void* stream);
int l2i_roi_align2_bwd(const float* dout, const float* rois, const int32_t* level, int K, int N, int C, int P, int Hl, int Wl,
                       float scale_l, float* dfeat_l, int Hs, int Ws, float scale_s, float* dfeat_s, void* stream);
/* 2x2 average pooling, NHWC (F.avg_pool2d(x, 2) in the discriminator blocks). */
int l2i_avgpool2_fwd(const float* x, int N, int H, int W, int C, float* out, void* stream);
int l2i_avgpool2_bwd(const float* dout, int N, int H, int W, int C, float* dx, void* stream);

/* 2x2 / stride 2 max pooling, NHWC (the VGG19 feature extractor of the perceptual loss, reference utils/util.py:49-94);
 * the backward sends the gradient to the first maximum of each window, as torch does. */
int l2i_maxpool2_fwd(const float* x, int N, int H, int W, int C, float* out, void* stream);
int l2i_maxpool2_bwd(const float* x, const float* dout, int N, int H, int W, int C, float* dx, void* stream);

/* ---- object-context attention (reference model/resnet_generator_app_v2.py:17-120,172-192), one head.
 *      q, k, v, out [B,O,D]; bbox [B,O,4]; y [B,O] int64 (0 = padding key); wg [64], bg [1];
 *      p_save, glin_save [B,O,O] are kept for the backward. -------------------------------------- */
int l2i_box_attention_fwd(const float* q, const float* k, const float* v, const float* bbox, const int64_t* y,
                          const float* wg, const float* bg, int B, int O, int D, float* out, float* p_save,
                          float* glin_save, void* stream);
int l2i_box_attention_bwd(const float* q, const float* k, const float* v, const float* bbox, const int64_t* y,
                          const float* p_save, const float* glin_save, const float* dout, int B, int O, int D,
                          float* dq, float* dk, float* dv, float* dwg, float* dbg, void* stream);

/* ---- pyramid pooling of the res4 mask head (reference model/resnet_generator_app_v2.py:724-752, PSPModule;
 *      sizes (1, 2, 3, 6) -> 50 cells per image in that order, row-major inside a size) -------------------- */
/* pooled [B,50,C] = the four AdaptiveAvgPool2d of x [B,H,W,C]. */
int l2i_psp_pool_fwd(const float* x, int B, int H, int W, int C, float* pooled, void* stream);
/* dx [B,H,W,C] = (base ? base[b,h,w,base_off + c] with channel stride base_stride : 0) + pooling backward of
 * dpooled [B,50,C]. */
int l2i_psp_pool_bwd(const float* dpooled, const float* base, int base_stride, int base_off, int B, int H, int W, int C,
                     float* dx, void* stream);
/* pair [B,H,W,cpad] of cat[ up(priors_1), up(priors_2), up(priors_3), up(priors_6), feats ]: priors [B,50,CP]
 * bilinearly up-sampled with align_corners=True, feats [B,H,W,C]; channels >= 4*CP + C are zero. */
int l2i_psp_concat_fwd(const float* feats, const float* priors, int B, int H, int W, int C, int CP, void* hi, void* lo,
                       int cpad, void* stream);
/* dpriors [B,50,CP] = backward of the four up-samplings from dcat [B,H,W,cstride] (channels [0, 4*CP)). */
int l2i_psp_concat_bwd(const float* dcat, int B, int H, int W, int CP, int cstride, float* dpriors, void* stream);

/* ---- spectral normalisation (torch.nn.utils.spectral_norm as used at resnet_generator_app_v2.py:681-686 and
 *      rcnn_discriminator_app.py:10-15).  W [R, Cc] fp32 (weight_orig viewed 2-D), u [R], v [Cc]. ------------ */
/* training != 0: one power iteration, u and v updated in place (v <- normalize(W^T u), u <- normalize(W v), eps
 * as in F.normalize).  Always: sigma[0] = u . (W v); u_used / v_used receive the vectors sigma was computed with
 * (for the backward).  work: Cc + R floats of scratch.  W / sigma itself is never materialised:
 * l2i_conv_weight_prep takes sigma. */
int l2i_sn_sigma(const float* W, int R, int Cc, float* u, float* v, int training, float eps, float* u_used, float* v_used,
                 float* sigma, float* work, void* stream);
/* dW [R][cin][taps] (torch layout of weight_orig) = (G - <G, W>/sigma * u v^T) / sigma from the tensor-core
 * weight gradient G [R][taps][cin] (l2i_conv2d_wgrad's layout).  scratch: 1 float. */
int l2i_sn_weight_grad(const float* G, const float* W, const float* u, const float* v, const float* sigma, int R, int cin,
                       int taps, float* dW, float* scratch, void* stream);

/* Grouped form of the two calls above plus l2i_conv_weight_prep: ONE call runs the power iteration / sigma of every
 * spectrally normalised module of a network forward and writes the tensor-core operand pairs of every convolution
 * weight (6 launches instead of ~4 per module).  Power iterations depend on the weights only, so running them all
 * before the first layer is the same arithmetic as torch's per-module pre-forward hooks.
 *   table: device array of n_modules entries { const float* W; float* u; float* v; int64 f32_off; int64 bf_off;
 *     int R, Cc; float eps; int training, has_sn; int cin, taps; int pad; } (72 bytes).  has_sn = 0: plain conv weight
 *     (operand pairs only).  bf_off < 0: no operand pairs (linear / embedding weights).
 *   f32 (f32_floats floats, zeroed by the call): per module at f32_off: sigma [4], u_used [pad4(R)], v_used [pad4(Cc)]
 *     and 2 scratch vectors (pad4(Cc), pad4(R)).
 *   bf16: per module at bf_off: forward pair hi, lo ([R][taps][pad(cin)] each, rounded up to 64 elements), then, when
 *     want_dgrad, the data-gradient pair hi, lo ([cin][taps][pad(R)] each); pad() as in l2i_conv_weight_prep's callers.
 *   work lists (device, static per network): wt_items int4 (module, column block of 256, r0, r1), wv_items int2 (module,
 *     block of 8 rows), prep9_items / prep1_items int4 (module, tile_x, tile_y, 0) for 3x3 / 1x1 weights in 32 x 32 tiles;
 *     wt_smem_floats = max (r1 - r0), max_cc = max Cc over the modules. */
int l2i_sn_prepare_group(const void* table, int n_modules, const int* wt_items, int n_wt, int wt_smem_floats,
                         const int* wv_items, int n_wv, int max_cc, const int* prep9_items, int n9, const int* prep1_items,
                         int n1, float* f32, long long f32_floats, void* bf16, int want_dgrad, void* stream);

/* ---- LayerNorm of the attention block (replaces nn.LayerNorm -> ATen at resnet_generator_app_v2.py:201-212; the
 *      nn.Linear layers of the path run as 1x1 convolutions through l2i_conv2d_fwd / l2i_conv2d_wgrad). ------------------ */
/* out [N] = column sums of x [M,N] (row-major): the bias gradient of a linear layer wider than l2i_grad_split's 2048. */
int l2i_colsum(const float* x, int M, int N, float* out, void* stream);
/* y = LayerNorm(a + b) * w + bias over rows of D elements (b nullable); stats [rows,2] = (mean, 1/std) for the backward. */
int l2i_add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, int rows, int D, float eps,
                          float* y, float* stats, void* stream);
/* ds [rows,D] = gradient of (a + b);  dw, dbias [D] (zeroed by the call). */
int l2i_add_layernorm_bwd(const float* a, const float* b, const float* w, const float* stats, const float* dy, int rows, int D,
                          float* ds, float* dw, float* dbias, void* stream);

/* ---- discriminator output heads (reference model/rcnn_discriminator_app.py:125-127 image head, :160-166 object head,
 *      :148-157 appearance head).  feat / x: [N, P, C] fp32 (NHWC feature maps, P pixels).  w, emb are the ORIGINAL
 *      (un-normalised) spectrally normalised weights; sigma_w / sigma_e device scalars from l2i_sn_sigma. ----------- */
/* s[n,:] = sum_p relu(feat[n,p,:]) (kept for the backward);  out[n] = s[n,:] . w/sigma_w + bias[0]
 *          (+ s[n,:] . emb[y[n],:]/sigma_e when emb != NULL: the class projection of the object head). */
int l2i_head_fwd(const float* feat, int N, int P, int C, const float* w, const float* sigma_w, const float* bias,
                 const float* emb, const float* sigma_e, const int64_t* y, float* s, float* out, void* stream);
/* dfeat [N,P,C] (nullable); gw [C] = dL/d(w/sigma_w), gemb [num_emb,C] = dL/d(emb/sigma_e) (both zeroed by the call,
 * nullable; feed them to l2i_sn_weight_grad with taps = 1); dbias [1]. */
int l2i_head_bwd(const float* feat, const float* s, const float* dout, int N, int P, int C, const float* w,
                 const float* sigma_w, const float* emb, const float* sigma_e, const int64_t* y, int num_emb, float* dfeat,
                 float* gw, float* gemb, float* dbias, void* stream);
/* Appearance head without the (K,C,C) Gram matrix or the (K,C,2C) concatenation: F = relu(x);
 *   out[k] = (1/C^2) sum_p (sum_c F[k,p,c]) (sum_c F[k,p,c] w1[c]) + e[y[k],:] . w2 + bias[0],
 * w = [w1 | w2] (2C, divided by sigma_w), e = emb / sigma_e.  colsum, proj [K,P] are kept for the backward. */
int l2i_gram_proj_fwd(const float* x, int K, int P, int C, const float* w, const float* sigma_w, const float* bias,
                      const float* emb, const float* sigma_e, const int64_t* y, float* colsum, float* proj, float* out,
                      void* stream);
/* dx [K,P,C]; gw [2C] = dL/d(w/sigma_w); gemb [num_emb,C] = dL/d(emb/sigma_e); dbias [1] (all zeroed by the call). */
int l2i_gram_proj_bwd(const float* x, const float* colsum, const float* proj, const float* dout, int K, int P, int C,
                      const float* w, const float* sigma_w, const float* emb, const float* sigma_e, const int64_t* y,
                      int num_emb, float* dx, float* gw, float* gemb, float* dbias, void* stream);

/* ---- optimizer (reference train_context_app_v2.py:113-127,174,189: torch.optim.Adam(betas=(0, 0.999)),
 *      one parameter group per tensor).  One launch updates every tensor of a network.
 *      tensors: device array of { float* p; const float* g; float* m; float* v; int64 n; float step_size;
 *      float bc2_sqrt; } (48 bytes each) with step_size = lr / (1 - beta1^step), bc2_sqrt = sqrt(1 - beta2^step) for
 *      the tensor's OWN step count (torch keeps one per parameter) and n = 0 for a tensor without a gradient
 *      (skipped); chunks: device array of n_chunks (tensor index, chunk index) int pairs, each covering
 *      chunk_elems (multiple of 4) consecutive elements.  Arithmetic = torch's Adam step.
 *      step_dev != NULL (CUDA-graph capturable form): *step_dev is the common step count on the device, the entries'
 *      step_size holds the plain lr and both bias corrections are formed in the kernel. ---- */
int l2i_adam_step(const void* tensors, const int* chunks, int n_chunks, int chunk_elems, double beta1, double beta2,
                  double eps, const int64_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2I_H_ */
