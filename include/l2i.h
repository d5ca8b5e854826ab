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

/* y = (conv(x, w) + bias + residual) * out_scale, stride 1, "same" padding; H, W powers of two.
 * x pair [N,H,W,cin_pad]; w pair [cout][taps][cin_pad]; bias [cout] or NULL; residual [N,H,W,cout]
 * (res_up2 = 0) or [N,H/2,W/2,cout] read with nearest x2 up-sampling (res_up2 = 1), or NULL.
 * Outputs (either may be NULL, not both): out fp32 [N,H,W,cout]; pair [N,H,W,cout_pad] of
 * relu_split ? relu(y) : y.  The data gradient is the same call with the dgrad weight pair. */
int l2i_conv2d_fwd(int N, int H, int W, int cin_pad, int cout, int taps, const void* x_hi, const void* x_lo,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, int res_up2,
                   float out_scale, float* out, void* out_hi, void* out_lo, int cout_pad, int relu_split,
                   void* stream);

/* dw [cout][taps][cin] fp32 = sum over pixels of dy (x) shifted x.  dy pair [N,H,W,cout_pad],
 * x pair [N,H,W,cin_pad]. */
int l2i_conv2d_wgrad(int N, int H, int W, int cin, int cin_pad, int cout, int cout_pad, int taps, const void* dy_hi,
                     const void* dy_lo, const void* x_hi, const void* x_lo, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2I_H_ */
