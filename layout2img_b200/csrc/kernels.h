// Internal launcher declarations shared between the .cu files and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "conv_tc.h"

namespace l2i {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// prep.cu
int weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* f_hi, void* f_lo, int cin_pad,
                void* d_hi, void* d_lo, int cout_pad, cudaStream_t stream);
int act_split(const float* x, int N, int H, int W, int C, int relu, int up2, void* hi, void* lo, int cpad,
              cudaStream_t stream);

int act_split2(const float* x, int N, int H, int W, int C, int relu_a, void* a_hi, void* a_lo, int b_mode, float b_scale,
               void* b_hi, void* b_lo, int cpad, cudaStream_t stream);
int grad_split(const float* g, int N, int H, int W, int C, void* lo_hi, void* lo_lo, float up_scale, void* up_hi,
               void* up_lo, float* colsum, int cpad, cudaStream_t stream);
int pair_colsum(const void* hi, const void* lo, long long pixels, int C, int cpad, float* colsum, cudaStream_t stream);

// isla.cu
int bn_stats(const float* x, long long pixels, int C, double* sums, cudaStream_t stream);
int bn_finalize(const double* sums, double count, int C, float eps, float momentum, float* running_mean,
                float* running_var, float* mean_invstd, cudaStream_t stream);
int bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* mean_invstd, cudaStream_t stream);
int isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* chan_scale, int B, int H, int W, int C, int O, float* out,
             void* hi, void* lo, int cpad, int relu, int up2, cudaStream_t stream);
int isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* chan_scale, const float* dout, int B, int H, int W, int C,
             int O, int relu, int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta, double* csum,
             float* dx, int phase, double count, cudaStream_t stream);

// layout_ops.cu
int bbox_mask(const float* bbox, int BO, int H, int W, float* out, cudaStream_t stream);
int masks_to_layout_fwd(const float* bbox, const float* masks, int BO, int M, int S, float* out, cudaStream_t stream);
int masks_to_layout_bwd(const float* bbox, const float* dout, int BO, int M, int S, float* dmasks, cudaStream_t stream);
int mask_resize_fwd(const float* in, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* out,
                    cudaStream_t stream);
int mask_resize_bwd(const float* dout, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* din,
                    cudaStream_t stream);
int stage_mix_fwd(const float* stage, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  int B, int O, int h, int w, int NC, int S, float* out, cudaStream_t stream);
int stage_mix_bwd(const float* stage, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  const float* dout, int B, int O, int h, int w, int NC, int S, float* dstage, float* dalpha,
                  float* dsoft, cudaStream_t stream);
int class_mix_fwd(const float* t, const float* Wc, const float* bc, const long long* y, const float* alpha, const float* bmask,
                  const float* hard, int B, int O, int h, int w, int C, int NC, int S, float* sel, float* out,
                  cudaStream_t stream);
int class_mix_bwd(const float* t, const float* Wc, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  const float* sel, const float* dout, int B, int O, int h, int w, int C, int NC, int S, float* dt, float* dW,
                  float* db, float* dalpha, float* dsoft, cudaStream_t stream);
int inorm_relu_fwd(const float* x, int N, int H, int W, int C, int up2, float eps, float* stats, void* hi, void* lo,
                   int cpad, cudaStream_t stream);
int inorm_relu_bwd(const float* x, const float* stats, const float* da, int N, int H, int W, int C, int up2, float* dx,
                   cudaStream_t stream);
// roi_align.cu
int roi_align_fwd(const float* feat, const float* rois, int K, int N, int H, int W, int C, int P, float scale, float* out,
                  cudaStream_t stream);
int roi_align_bwd(const float* dout, const float* rois, int K, int N, int H, int W, int C, int P, float scale, float* dfeat,
                  cudaStream_t stream);
int roi_prepare(const float* bbox, const long long* label, int B, int O, float img, float thresh, float* rois,
                long long* y_sorted, int* level, int* perm, int* counts, cudaStream_t stream);
int roi_align2_fwd(const float* feat_l, int Hl, int Wl, float scale_l, const float* feat_s, int Hs, int Ws, float scale_s,
                   const float* rois, const int* level, int K, int N, int C, int P, float* out, cudaStream_t stream);
int roi_align2_bwd(const float* dout, const float* rois, const int* level, int K, int N, int C, int P, int Hl, int Wl,
                   float scale_l, float* dfeat_l, int Hs, int Ws, float scale_s, float* dfeat_s, cudaStream_t stream);
int maxpool2_fwd(const float* x, int N, int H, int W, int C, float* out, cudaStream_t stream);
int maxpool2_bwd(const float* x, const float* dout, int N, int H, int W, int C, float* dx, cudaStream_t stream);
int avgpool2_fwd(const float* x, int N, int H, int W, int C, float* out, cudaStream_t stream);
int avgpool2_bwd(const float* dout, int N, int H, int W, int C, float* dx, cudaStream_t stream);
// attention.cu
int box_attention_fwd(const float* q, const float* k, const float* v, const float* bbox, const long long* y,
                      const float* wg, const float* bg, int B, int O, int D, float* out, float* p_save, float* glin_save,
                      cudaStream_t stream);
int box_attention_bwd(const float* q, const float* k, const float* v, const float* bbox, const long long* y,
                      const float* p_save, const float* glin_save, const float* dout, int B, int O, int D, float* dq,
                      float* dk, float* dv, float* dwg, float* dbg, cudaStream_t stream);

// psp.cu
int psp_pool_fwd(const float* x, int B, int H, int W, int C, float* pooled, cudaStream_t stream);
int psp_pool_bwd(const float* dpooled, const float* base, int base_stride, int base_off, int B, int H, int W, int C,
                 float* dx, cudaStream_t stream);
int psp_concat_fwd(const float* feats, const float* priors, int B, int H, int W, int C, int CP, void* hi, void* lo, int cpad,
                   cudaStream_t stream);
int psp_concat_bwd(const float* dcat, int B, int H, int W, int CP, int cstride, float* dpriors, cudaStream_t stream);

// specnorm.cu
int sn_sigma(const float* W, int R, int Cc, float* u, float* v, int training, float eps, float* u_used, float* v_used,
             float* sigma, float* work, cudaStream_t stream);
int sn_weight_grad(const float* G, const float* W, const float* u, const float* v, const float* sigma, int R, int cin,
                   int taps, float* dW, float* scratch, cudaStream_t stream);

// Grouped spectral norm + weight preparation: one table entry per module of a network (72 bytes; include/l2i.h).
struct SnEntry {
  const float* W;        // weight_orig (or the plain weight when has_sn = 0), [R, Cc] / torch layout [cout][cin][kh][kw]
  float* u;              // [R]  (updated in place when training)
  float* v;              // [Cc]
  long long f32_off;     // this module's slice of the per-call fp32 buffer (floats): sigma, u_used, v_used, t, s
  long long bf_off;      // this module's slice of the per-call bf16 buffer (elements), < 0: no operand pairs wanted
  int R, Cc;
  float eps;
  int training, has_sn;
  int cin, taps;         // conv geometry for the operand pairs (cout = R)
  int pad_;
};
__host__ __device__ inline long long sn_pad4(long long x) { return (x + 3) & ~3LL; }
__host__ __device__ inline long long sn_off_u() { return 4; }                                    // sigma at 0
__host__ __device__ inline long long sn_off_v(int R, int Cc) { return 4 + sn_pad4(R); }
__host__ __device__ inline long long sn_off_t(int R, int Cc) { return 4 + sn_pad4(R) + sn_pad4(Cc); }
__host__ __device__ inline long long sn_off_s(int R, int Cc) { return 4 + sn_pad4(R) + 2 * sn_pad4(Cc); }
int sn_group_sigma(const void* table, int n_modules, const int* wt_items, int n_wt, int wt_smem_floats, const int* wv_items,
                   int n_wv, int max_cc, float* f32, long long f32_floats, cudaStream_t stream);
int weight_prep_group(const void* table, const int* items9, int n9, const int* items1, int n1, const float* f32, void* bf16,
                      int want_dgrad, cudaStream_t stream);

// im2col.cu
int im2col3_pair(const float* x, int N, int H, int W, int C, int sign, void* hi, void* lo, int cpad, float* colsum,
                 cudaStream_t stream);
int col2im3(const float* col, int ldc, int N, int H, int W, int C, int sign, const float* bias, const float* residual,
            int res_up2, float res_scale, float* out, cudaStream_t stream);

// linear.cu
int colsum(const float* X, int M, int N, float* out, cudaStream_t stream);
int add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, int rows, int D, float eps, float* y,
                      float* stats, cudaStream_t stream);
int add_layernorm_bwd(const float* a, const float* b, const float* w, const float* stats, const float* dy, int rows, int D,
                      float* ds, float* dw, float* dbias, cudaStream_t stream);

// heads.cu
int head_fwd(const float* feat, int N, int P, int C, const float* w, const float* sigma_w, const float* bias, const float* emb,
             const float* sigma_e, const long long* y, float* s, float* out, cudaStream_t stream);
int head_bwd(const float* feat, const float* s, const float* dout, int N, int P, int C, const float* w, const float* sigma_w,
             const float* emb, const float* sigma_e, const long long* y, int num_emb, float* dfeat, float* gw, float* gemb,
             float* dbias, cudaStream_t stream);
int gram_proj_fwd(const float* x, int K, int P, int C, const float* w, const float* sigma_w, const float* bias, const float* emb,
                  const float* sigma_e, const long long* y, float* colsum, float* proj, float* out, cudaStream_t stream);
int gram_proj_bwd(const float* x, const float* colsum, const float* proj, const float* dout, int K, int P, int C, const float* w,
                  const float* sigma_w, const float* emb, const float* sigma_e, const long long* y, int num_emb, float* dx,
                  float* gw, float* gemb, float* dbias, cudaStream_t stream);

// optim.cu
struct AdamTensor {       // one entry of the device-resident tensor table (48 bytes, see include/l2i.h)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;            // 0: no gradient this step, the tensor is skipped
  float step_size;        // lr / (1 - beta1^step), step = this tensor's own update count
  float bc2_sqrt;         // sqrt(1 - beta2^step)
};
int adam_step(const void* tensors, const int* chunks, int n_chunks, int chunk_elems, double beta1, double beta2, double eps,
              const long long* step_dev, cudaStream_t stream);

}  // namespace l2i
