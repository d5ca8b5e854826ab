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

// isla.cu
int bn_stats(const float* x, long long pixels, int C, double* sums, cudaStream_t stream);
int bn_finalize(const double* sums, double count, int C, float eps, float momentum, float* running_mean,
                float* running_var, float* mean_invstd, cudaStream_t stream);
int bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* mean_invstd, cudaStream_t stream);
int isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, int B, int H, int W, int C, int O, float* out, void* hi, void* lo,
             int cpad, int relu, int up2, cudaStream_t stream);
int isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* dout, int B, int H, int W, int C, int O, int relu,
             int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta, double* csum, float* dx,
             cudaStream_t stream);

}  // namespace l2i
