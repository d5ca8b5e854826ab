// 3x3 convolutions with a tiny channel count on one side -- the discriminator's first convolution (3 -> 64,
// rcnn_discriminator_app.py:297) and the generator's RGB head (64 -> 3, resnet_generator_app_v2.py:418) -- rewritten as
// 1x1 convolutions over a 9C-channel (C <= 4) im2col tensor, so that the tensor-core kernel runs ONE K chunk per tile
// instead of nine taps of a 64-wide chunk that is 95 % padding (the padded forms ran at 8-12 TFLOP/s):
//
//   small input :  y[p, co] = sum_{c,tap} W[co][c][tap] x[p + d(tap), c]  =  (im2col(+1)(x) [p, c*9+tap]) . W.view(co, 9C)
//   small output:  y[p, s]  = sum_tap P[p + d(tap), s*9+tap],  P = x . W2^T,  W2[s*9+tap][l] = W[s][l][tap]  =  col2im(-1)(P)
//
//   im2col(sign):  col[p, c*9+tap] = x[p + sign * d(tap), c]            (zero outside the image; written as a bf16 pair)
//   col2im(sign):  out[q, c] = sum_tap col[q - sign * d(tap), c*9+tap]  (the adjoint of im2col(sign))
// with d(tap) = (tap / 3 - 1, tap % 3 - 1).  The channel order c*9+tap is torch's weight order, so W.view(co, 9C) IS the
// 1x1 weight and its gradient needs no permutation.  HBM-bound: 4 B per im2col channel written.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

// A block takes 256 consecutive pixels.  Phase 1: thread t gathers the 9C values of pixel t (its 3 x 3 neighbourhood,
// L1-resident) and parks their (hi, lo) halves in shared memory; phase 2: the block streams the [256][cpad] tile of
// each half out in consecutive 16-byte pieces (channels >= 9C as zeros).  Optional per-channel sums of x.
static constexpr int kColCh = 40;          // 9 * 4 = 36 im2col channels, rounded to whole 8-channel groups
__global__ void __launch_bounds__(256)
im2col3_pair_kernel(const float* __restrict__ x, int N, int H, int W, int C, int sign, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo, int cpad, float* __restrict__ colsum) {
  __shared__ __align__(16) __nv_bfloat16 s_hi[256][kColCh], s_lo[256][kColCh];
  __shared__ float s_sum[4];
  if (threadIdx.x < 4) s_sum[threadIdx.x] = 0.f;
  const long long P = 1LL * N * H * W;
  const int groups = cpad >> 3;
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long p0 = blockIdx.x * 256LL; p0 < P; p0 += 256LL * gridDim.x) {
    const long long p = p0 + threadIdx.x;
    __syncthreads();                                   // previous tile fully streamed out (and s_sum initialised)
    if (p < P) {
      const int w = static_cast<int>(p % W), h = static_cast<int>((p / W) % H);
      const long long n = p / (1LL * W * H);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int hh = h + sign * (tap / 3 - 1), ww = w + sign * (tap % 3 - 1);
          float v = 0.f;
          if (c < C && hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + ((n * H + hh) * W + ww) * C + c);
          if (tap == 4) csum[c] += v;
          __nv_bfloat16 a, b;
          split_bf16(v, a, b);
          s_hi[threadIdx.x][c * 9 + tap] = a;
          s_lo[threadIdx.x][c * 9 + tap] = b;
        }
      }
#pragma unroll
      for (int k = 36; k < kColCh; ++k) { s_hi[threadIdx.x][k] = __float2bfloat16(0.f); s_lo[threadIdx.x][k] = __float2bfloat16(0.f); }
    }
    __syncthreads();
    const int npix = static_cast<int>(min(256LL, P - p0));
    for (int i = threadIdx.x; i < npix * groups; i += 256) {
      const int px = i / groups, g = i - px * groups;
      uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
      if (g * 8 < kColCh) {
        a = *reinterpret_cast<const uint4*>(&s_hi[px][g * 8]);
        b = *reinterpret_cast<const uint4*>(&s_lo[px][g * 8]);
      }
      *reinterpret_cast<uint4*>(hi + (p0 + px) * cpad + g * 8) = a;
      *reinterpret_cast<uint4*>(lo + (p0 + px) * cpad + g * 8) = b;
    }
  }
  if (colsum) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float t = warp_sum(csum[c]);
      if ((threadIdx.x & 31) == 0 && c < C) atomicAdd(&s_sum[c], t);
    }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(colsum + threadIdx.x, s_sum[threadIdx.x]);
  }
}

// one thread per output element: out[q, c] = sum_tap col[q - sign d(tap), c*9+tap] + bias[c] + res_scale * residual
__global__ void __launch_bounds__(256)
col2im3_kernel(const float* __restrict__ col, int ldc, int N, int H, int W, int C, int sign, const float* __restrict__ bias,
               const float* __restrict__ residual, int res_up2, float res_scale, float* __restrict__ out) {
  const long long total = 1LL * N * H * W * C;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long q = i / C;
    const int w = static_cast<int>(q % W), h = static_cast<int>((q / W) % H);
    const long long n = q / (1LL * W * H);
    float acc = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - sign * (tap / 3 - 1), ww = w - sign * (tap % 3 - 1);
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      acc += __ldg(col + ((n * H + hh) * W + ww) * ldc + c * 9 + tap);
    }
    if (residual) {
      const long long r = res_up2 ? ((n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) : q;
      acc = fmaf(res_scale, __ldg(residual + r * C + c), acc);
    }
    out[i] = acc;
  }
}

int im2col3_pair(const float* x, int N, int H, int W, int C, int sign, void* hi, void* lo, int cpad, float* colsum,
                 cudaStream_t stream) {
  if (!x || !hi || !lo || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 4 || cpad % 8 || cpad < 9 * C || (sign != 1 && sign != -1)) {
    set_error("im2col3: bad arguments (C=%d must be <= 4, cpad=%d)", C, cpad);
    return L2I_ERR_BAD_ARG;
  }
  if (colsum) {
    cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * C, stream);
    if (e != cudaSuccess) { set_error("im2col3: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  }
  long long blocks = (1LL * N * H * W + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  im2col3_pair_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, N, H, W, C, sign, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                   reinterpret_cast<__nv_bfloat16*>(lo), cpad, colsum);
  return check_launch("im2col3_pair_kernel");
}

int col2im3(const float* col, int ldc, int N, int H, int W, int C, int sign, const float* bias, const float* residual,
            int res_up2, float res_scale, float* out, cudaStream_t stream) {
  if (!col || !out || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 4 || ldc < 9 * C || (sign != 1 && sign != -1) ||
      (residual && res_up2 && ((H | W) & 1))) {
    set_error("col2im3: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  const long long total = 1LL * N * H * W * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  col2im3_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(col, ldc, N, H, W, C, sign, bias, residual, res_up2, res_scale, out);
  return check_launch("col2im3_kernel");
}

}  // namespace l2i
