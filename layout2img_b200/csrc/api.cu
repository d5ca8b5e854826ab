// extern "C" entry points of libl2i.so (declared in include/l2i.h) + error plumbing.
#include <stdarg.h>
#include <stdio.h>
#include "../../include/l2i.h"
#include "kernels.h"

namespace l2i {
static thread_local char g_err[512] = "";
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
  return L2I_OK;
}
}  // namespace l2i

using namespace l2i;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int l2i_version(void) { return 100; }
const char* l2i_last_error(void) { return g_err; }

int l2i_conv_weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* fwd_hi, void* fwd_lo,
                         int cin_pad, void* dg_hi, void* dg_lo, int cout_pad, void* stream) {
  return weight_prep(w, sigma, cout, cin, taps, fwd_hi, fwd_lo, cin_pad, dg_hi, dg_lo, cout_pad, ST(stream));
}

int l2i_act_split(const float* x, int N, int H, int W, int C, int relu, int up2, void* hi, void* lo, int cpad,
                  void* stream) {
  return act_split(x, N, H, W, C, relu, up2, hi, lo, cpad, ST(stream));
}

int l2i_conv2d_fwd(int N, int H, int W, int cin_pad, int cout, int taps, const void* x_hi, const void* x_lo,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, int res_up2,
                   float out_scale, float* out, void* out_hi, void* out_lo, int cout_pad, int relu_split,
                   void* stream) {
  ConvFwdArgs a;
  a.N = N; a.H = H; a.W = W; a.cin_pad = cin_pad; a.cout = cout; a.taps = taps;
  a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias; a.residual = residual;
  a.res_shift = res_up2 ? 1 : 0; a.out = out; a.out_hi = out_hi; a.out_lo = out_lo; a.cout_pad = cout_pad;
  a.relu_split = relu_split; a.out_scale = out_scale;
  return conv_fwd_tc(a, ST(stream));
}

int l2i_conv2d_wgrad(int N, int H, int W, int cin, int cin_pad, int cout, int cout_pad, int taps, const void* dy_hi,
                     const void* dy_lo, const void* x_hi, const void* x_lo, float* dw, void* stream) {
  ConvWgradArgs a;
  a.N = N; a.H = H; a.W = W; a.cin = cin; a.cin_pad = cin_pad; a.cout = cout; a.cout_pad = cout_pad; a.taps = taps;
  a.dy_hi = dy_hi; a.dy_lo = dy_lo; a.x_hi = x_hi; a.x_lo = x_lo; a.dw = dw;
  return conv_wgrad_tc(a, ST(stream));
}

}  // extern "C"
