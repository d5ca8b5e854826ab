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

}  // namespace l2i
