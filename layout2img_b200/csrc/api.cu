// extern "C" entry points of libl2i.so (declared in include/l2i.h) + error plumbing.
#include <stdarg.h>
#include <stdio.h>
#include "../../include/l2i.h"
#include "kernels.h"

#include <atomic>
namespace l2i {
static thread_local char g_err[512] = "";
static std::atomic<int> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return L2I_ERR_LAUNCH;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return L2I_OK;
}
}  // namespace l2i

using namespace l2i;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int l2i_version(void) { return 100; }
const char* l2i_last_error(void) { return g_err; }
int l2i_launch_count(int reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

int l2i_conv_weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* fwd_hi, void* fwd_lo,
                         int cin_pad, void* dg_hi, void* dg_lo, int cout_pad, void* stream) {
  return weight_prep(w, sigma, cout, cin, taps, fwd_hi, fwd_lo, cin_pad, dg_hi, dg_lo, cout_pad, ST(stream));
}

int l2i_act_split(const float* x, int N, int H, int W, int C, int relu, int up2, void* hi, void* lo, int cpad,
                  void* stream) {
  return act_split(x, N, H, W, C, relu, up2, hi, lo, cpad, ST(stream));
}

int l2i_act_split2(const float* x, int N, int H, int W, int C, int relu_a, void* a_hi, void* a_lo, int b_mode,
                   float b_scale, void* b_hi, void* b_lo, int cpad, void* stream) {
  return act_split2(x, N, H, W, C, relu_a, a_hi, a_lo, b_mode, b_scale, b_hi, b_lo, cpad, ST(stream));
}
int l2i_grad_split(const float* g, int N, int H, int W, int C, void* lo_hi, void* lo_lo, float up_scale, void* up_hi,
                   void* up_lo, float* colsum, int cpad, void* stream) {
  return grad_split(g, N, H, W, C, lo_hi, lo_lo, up_scale, up_hi, up_lo, colsum, cpad, ST(stream));
}
int l2i_pair_colsum(const void* hi, const void* lo, long long pixels, int C, int cpad, float* colsum, void* stream) {
  return pair_colsum(hi, lo, pixels, C, cpad, colsum, ST(stream));
}

int l2i_conv2d_fwd(int N, int H, int W, int cin_pad, int cout, int taps, const void* x_hi, const void* x_lo,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, int res_up2,
                   float res_scale, float out_scale, const void* mask_hi, int mask_cpad, int pool, float* out, void* out_hi,
                   void* out_lo, int cout_pad, int relu_split, void* stream) {
  ConvFwdArgs a;
  a.N = N; a.H = H; a.W = W; a.cin_pad = cin_pad; a.cout = cout; a.taps = taps;
  a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias; a.residual = residual;
  a.res_shift = res_up2 ? 1 : 0; a.out = out; a.out_hi = out_hi; a.out_lo = out_lo; a.cout_pad = cout_pad;
  a.relu_split = relu_split; a.out_scale = out_scale; a.res_scale = res_scale;
  a.mask_hi = mask_hi; a.mask_cpad = mask_cpad; a.pool = pool;
  return conv_fwd_tc(a, ST(stream));
}

int l2i_conv2d_wgrad(int N, int H, int W, int cin, int cin_pad, int cout, int cout_pad, int taps, const void* dy_hi,
                     const void* dy_lo, const void* x_hi, const void* x_lo, float* dw, void* stream) {
  ConvWgradArgs a;
  a.N = N; a.H = H; a.W = W; a.cin = cin; a.cin_pad = cin_pad; a.cout = cout; a.cout_pad = cout_pad; a.taps = taps;
  a.dy_hi = dy_hi; a.dy_lo = dy_lo; a.x_hi = x_hi; a.x_lo = x_lo; a.dw = dw;
  return conv_wgrad_tc(a, ST(stream));
}

int l2i_bn_stats(const float* x, long long pixels, int C, double* sums, void* stream) {
  return bn_stats(x, pixels, C, sums, ST(stream));
}
int l2i_bn_finalize(const double* sums, double count, int C, float eps, float momentum, float* running_mean,
                    float* running_var, float* mean_invstd, void* stream) {
  return bn_finalize(sums, count, C, eps, momentum, running_mean, running_var, mean_invstd, ST(stream));
}
int l2i_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean_invstd,
                      void* stream) {
  return bn_eval_stats(running_mean, running_var, C, eps, mean_invstd, ST(stream));
}
int l2i_isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
                 const float* aff_w, const float* aff_b, const float* chan_scale, int B, int H, int W, int C, int O,
                 float* out, void* hi, void* lo, int cpad, int relu, int up2, void* stream) {
  return isla_fwd(x, mean_invstd, mask, gamma, beta, aff_w, aff_b, chan_scale, B, H, W, C, O, out, hi, lo, cpad, relu, up2,
                  ST(stream));
}
int l2i_isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
                 const float* aff_w, const float* aff_b, const float* chan_scale, const float* dout, int B, int H, int W,
                 int C, int O, int relu, int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta,
                 double* csum, float* dx, int phase, double count, void* stream) {
  return isla_bwd(x, mean_invstd, mask, gamma, beta, aff_w, aff_b, chan_scale, dout, B, H, W, C, O, relu, up2, train, gbuf,
                  dmask, dgamma, dbeta, csum, dx, phase, count, ST(stream));
}
int l2i_bbox_mask(const float* bbox, int BO, int H, int W, float* out, void* stream) {
  return bbox_mask(bbox, BO, H, W, out, ST(stream));
}
int l2i_masks_to_layout_fwd(const float* bbox, const float* masks, int BO, int M, int S, float* out, void* stream) {
  return masks_to_layout_fwd(bbox, masks, BO, M, S, out, ST(stream));
}
int l2i_masks_to_layout_bwd(const float* bbox, const float* dout, int BO, int M, int S, float* dmasks, void* stream) {
  return masks_to_layout_bwd(bbox, dout, BO, M, S, dmasks, ST(stream));
}
int l2i_mask_resize_fwd(const float* in, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* out,
                        void* stream) {
  return mask_resize_fwd(in, B, O, hi, wi, h, w, pixel_major, out, ST(stream));
}
int l2i_mask_resize_bwd(const float* dout, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* din,
                        void* stream) {
  return mask_resize_bwd(dout, B, O, hi, wi, h, w, pixel_major, din, ST(stream));
}
int l2i_stage_mix_fwd(const float* stage, const int64_t* y, const float* alpha, const float* bmask, const float* hard,
                      int B, int O, int h, int w, int NC, int S, float* out, void* stream) {
  return stage_mix_fwd(stage, reinterpret_cast<const long long*>(y), alpha, bmask, hard, B, O, h, w, NC, S, out, ST(stream));
}
int l2i_stage_mix_bwd(const float* stage, const int64_t* y, const float* alpha, const float* bmask, const float* hard,
                      const float* dout, int B, int O, int h, int w, int NC, int S, float* dstage, float* dalpha,
                      float* dsoft, void* stream) {
  return stage_mix_bwd(stage, reinterpret_cast<const long long*>(y), alpha, bmask, hard, dout, B, O, h, w, NC, S, dstage,
                       dalpha, dsoft, ST(stream));
}
int l2i_class_mix_fwd(const float* t, const float* Wc, const float* bc, const int64_t* y, const float* alpha,
                      const float* bmask, const float* hard, int B, int O, int h, int w, int C, int NC, int S, float* sel,
                      float* out, void* stream) {
  return class_mix_fwd(t, Wc, bc, reinterpret_cast<const long long*>(y), alpha, bmask, hard, B, O, h, w, C, NC, S, sel, out,
                       ST(stream));
}
int l2i_class_mix_bwd(const float* t, const float* Wc, const int64_t* y, const float* alpha, const float* bmask,
                      const float* hard, const float* sel, const float* dout, int B, int O, int h, int w, int C, int NC, int S,
                      float* dt, float* dW, float* db, float* dalpha, float* dsoft, void* stream) {
  return class_mix_bwd(t, Wc, reinterpret_cast<const long long*>(y), alpha, bmask, hard, sel, dout, B, O, h, w, C, NC, S, dt,
                       dW, db, dalpha, dsoft, ST(stream));
}
int l2i_inorm_relu_fwd(const float* x, int N, int H, int W, int C, int up2, float eps, float* stats, void* hi, void* lo,
                       int cpad, void* stream) {
  return inorm_relu_fwd(x, N, H, W, C, up2, eps, stats, hi, lo, cpad, ST(stream));
}
int l2i_inorm_relu_bwd(const float* x, const float* stats, const float* da, int N, int H, int W, int C, int up2, float* dx,
                       void* stream) {
  return inorm_relu_bwd(x, stats, da, N, H, W, C, up2, dx, ST(stream));
}
int l2i_roi_align_fwd(const float* feat, const float* rois, int K, int N, int H, int W, int C, int P, float scale,
                      float* out, void* stream) {
  return roi_align_fwd(feat, rois, K, N, H, W, C, P, scale, out, ST(stream));
}
int l2i_roi_align_bwd(const float* dout, const float* rois, int K, int N, int H, int W, int C, int P, float scale,
                      float* dfeat, void* stream) {
  return roi_align_bwd(dout, rois, K, N, H, W, C, P, scale, dfeat, ST(stream));
}
int l2i_avgpool2_fwd(const float* x, int N, int H, int W, int C, float* out, void* stream) {
  return avgpool2_fwd(x, N, H, W, C, out, ST(stream));
}
int l2i_avgpool2_bwd(const float* dout, int N, int H, int W, int C, float* dx, void* stream) {
  return avgpool2_bwd(dout, N, H, W, C, dx, ST(stream));
}
int l2i_box_attention_fwd(const float* q, const float* k, const float* v, const float* bbox, const int64_t* y,
                          const float* wg, const float* bg, int B, int O, int D, float* out, float* p_save,
                          float* glin_save, void* stream) {
  return box_attention_fwd(q, k, v, bbox, reinterpret_cast<const long long*>(y), wg, bg, B, O, D, out, p_save, glin_save,
                           ST(stream));
}
int l2i_box_attention_bwd(const float* q, const float* k, const float* v, const float* bbox, const int64_t* y,
                          const float* p_save, const float* glin_save, const float* dout, int B, int O, int D,
                          float* dq, float* dk, float* dv, float* dwg, float* dbg, void* stream) {
  return box_attention_bwd(q, k, v, bbox, reinterpret_cast<const long long*>(y), p_save, glin_save, dout, B, O, D, dq, dk,
                           dv, dwg, dbg, ST(stream));
}

int l2i_psp_pool_fwd(const float* x, int B, int H, int W, int C, float* pooled, void* stream) {
  return psp_pool_fwd(x, B, H, W, C, pooled, ST(stream));
}
int l2i_psp_pool_bwd(const float* dpooled, const float* base, int base_stride, int base_off, int B, int H, int W, int C,
                     float* dx, void* stream) {
  return psp_pool_bwd(dpooled, base, base_stride, base_off, B, H, W, C, dx, ST(stream));
}
int l2i_psp_concat_fwd(const float* feats, const float* priors, int B, int H, int W, int C, int CP, void* hi, void* lo,
                       int cpad, void* stream) {
  return psp_concat_fwd(feats, priors, B, H, W, C, CP, hi, lo, cpad, ST(stream));
}
int l2i_psp_concat_bwd(const float* dcat, int B, int H, int W, int CP, int cstride, float* dpriors, void* stream) {
  return psp_concat_bwd(dcat, B, H, W, CP, cstride, dpriors, ST(stream));
}
int l2i_sn_sigma(const float* W, int R, int Cc, float* u, float* v, int training, float eps, float* u_used, float* v_used,
                 float* sigma, float* work, void* stream) {
  return sn_sigma(W, R, Cc, u, v, training, eps, u_used, v_used, sigma, work, ST(stream));
}
int l2i_sn_weight_grad(const float* G, const float* W, const float* u, const float* v, const float* sigma, int R, int cin,
                       int taps, float* dW, float* scratch, void* stream) {
  return sn_weight_grad(G, W, u, v, sigma, R, cin, taps, dW, scratch, ST(stream));
}
int l2i_sn_prepare_group(const void* table, int n_modules, const int* wt_items, int n_wt, int wt_smem_floats,
                         const int* wv_items, int n_wv, int max_cc, const int* prep9_items, int n9, const int* prep1_items,
                         int n1, float* f32, long long f32_floats, void* bf16, int want_dgrad, void* stream) {
  int rc = sn_group_sigma(table, n_modules, wt_items, n_wt, wt_smem_floats, wv_items, n_wv, max_cc, f32, f32_floats, ST(stream));
  if (rc) return rc;
  if (n9 + n1 > 0) rc = weight_prep_group(table, prep9_items, n9, prep1_items, n1, f32, bf16, want_dgrad, ST(stream));
  return rc;
}
int l2i_roi_prepare(const float* bbox, const int64_t* label, int B, int O, float img_size, float small_thresh, float* rois,
                    int64_t* y_sorted, int32_t* level, int32_t* perm, int32_t* counts, void* stream) {
  return roi_prepare(bbox, reinterpret_cast<const long long*>(label), B, O, img_size, small_thresh, rois,
                     reinterpret_cast<long long*>(y_sorted), level, perm, counts, ST(stream));
}
int l2i_roi_align2_fwd(const float* feat_l, int Hl, int Wl, float scale_l, const float* feat_s, int Hs, int Ws, float scale_s,
                       const float* rois, const int32_t* level, int K, int N, int C, int P, float* out, void* stream) {
  return roi_align2_fwd(feat_l, Hl, Wl, scale_l, feat_s, Hs, Ws, scale_s, rois, level, K, N, C, P, out, ST(stream));
}
int l2i_roi_align2_bwd(const float* dout, const float* rois, const int32_t* level, int K, int N, int C, int P, int Hl, int Wl,
                       float scale_l, float* dfeat_l, int Hs, int Ws, float scale_s, float* dfeat_s, void* stream) {
  return roi_align2_bwd(dout, rois, level, K, N, C, P, Hl, Wl, scale_l, dfeat_l, Hs, Ws, scale_s, dfeat_s, ST(stream));
}
int l2i_im2col3_pair(const float* x, int N, int H, int W, int C, int sign, void* hi, void* lo, int cpad, float* colsum,
                     void* stream) {
  return im2col3_pair(x, N, H, W, C, sign, hi, lo, cpad, colsum, ST(stream));
}
int l2i_col2im3(const float* col, int ldc, int N, int H, int W, int C, int sign, const float* bias, const float* residual,
                int res_up2, float res_scale, float* out, void* stream) {
  return col2im3(col, ldc, N, H, W, C, sign, bias, residual, res_up2, res_scale, out, ST(stream));
}
int l2i_colsum(const float* x, int M, int N, float* out, void* stream) { return colsum(x, M, N, out, ST(stream)); }
int l2i_add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, int rows, int D, float eps,
                          float* y, float* stats, void* stream) {
  return add_layernorm_fwd(a, b, w, bias, rows, D, eps, y, stats, ST(stream));
}
int l2i_add_layernorm_bwd(const float* a, const float* b, const float* w, const float* stats, const float* dy, int rows, int D,
                          float* ds, float* dw, float* dbias, void* stream) {
  return add_layernorm_bwd(a, b, w, stats, dy, rows, D, ds, dw, dbias, ST(stream));
}
int l2i_maxpool2_fwd(const float* x, int N, int H, int W, int C, float* out, void* stream) {
  return maxpool2_fwd(x, N, H, W, C, out, ST(stream));
}
int l2i_maxpool2_bwd(const float* x, const float* dout, int N, int H, int W, int C, float* dx, void* stream) {
  return maxpool2_bwd(x, dout, N, H, W, C, dx, ST(stream));
}
int l2i_head_fwd(const float* feat, int N, int P, int C, const float* w, const float* sigma_w, const float* bias,
                 const float* emb, const float* sigma_e, const int64_t* y, float* s, float* out, void* stream) {
  return head_fwd(feat, N, P, C, w, sigma_w, bias, emb, sigma_e, reinterpret_cast<const long long*>(y), s, out, ST(stream));
}
int l2i_head_bwd(const float* feat, const float* s, const float* dout, int N, int P, int C, const float* w,
                 const float* sigma_w, const float* emb, const float* sigma_e, const int64_t* y, int num_emb, float* dfeat,
                 float* gw, float* gemb, float* dbias, void* stream) {
  return head_bwd(feat, s, dout, N, P, C, w, sigma_w, emb, sigma_e, reinterpret_cast<const long long*>(y), num_emb, dfeat, gw,
                  gemb, dbias, ST(stream));
}
int l2i_gram_proj_fwd(const float* x, int K, int P, int C, const float* w, const float* sigma_w, const float* bias,
                      const float* emb, const float* sigma_e, const int64_t* y, float* colsum, float* proj, float* out,
                      void* stream) {
  return gram_proj_fwd(x, K, P, C, w, sigma_w, bias, emb, sigma_e, reinterpret_cast<const long long*>(y), colsum, proj, out,
                       ST(stream));
}
int l2i_gram_proj_bwd(const float* x, const float* colsum, const float* proj, const float* dout, int K, int P, int C,
                      const float* w, const float* sigma_w, const float* emb, const float* sigma_e, const int64_t* y,
                      int num_emb, float* dx, float* gw, float* gemb, float* dbias, void* stream) {
  return gram_proj_bwd(x, colsum, proj, dout, K, P, C, w, sigma_w, emb, sigma_e, reinterpret_cast<const long long*>(y), num_emb,
                       dx, gw, gemb, dbias, ST(stream));
}
int l2i_adam_step(const void* tensors, const int* chunks, int n_chunks, int chunk_elems, double beta1, double beta2,
                  double eps, const int64_t* step_dev, void* stream) {
  return adam_step(tensors, chunks, n_chunks, chunk_elems, beta1, beta2, eps, reinterpret_cast<const long long*>(step_dev),
                   ST(stream));
}

}  // extern "C"
