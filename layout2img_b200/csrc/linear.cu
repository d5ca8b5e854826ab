// LayerNorm(x + residual) of the object-context attention block (reference resnet_generator_app_v2.py:201-212), forward
// and backward, one warp per row.  (The nn.Linear layers of the path run on the tensor-core convolution kernel as 1x1
// convolutions over one-pixel images: functional.LinearFn.)
#include "common.cuh"
#include "kernels.h"

namespace l2i {

// column sums of a [M, N] row-major matrix (the bias gradient of a wide linear layer): one block per 32 columns
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int M, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (n < N)
    for (int m = warp; m < M; m += 8) acc += __ldg(X + static_cast<size_t>(m) * N + n);
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    out[n] = t;
  }
}

int colsum(const float* X, int M, int N, float* out, cudaStream_t stream) {
  if (!X || !out || M <= 0 || N <= 0) { set_error("colsum: bad arguments"); return L2I_ERR_BAD_ARG; }
  colsum_kernel<<<(N + 31) / 32, 256, 0, stream>>>(X, M, N, out);
  return check_launch("colsum_kernel");
}

// ------------------------------------------------------------------------------------------ LayerNorm(a + b)
// y = (s - mean) * rstd * w + bias with s = a + b (b nullable), per row of D elements; one warp per row.
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                         const float* __restrict__ bias, int rows, int D, float eps, float* __restrict__ y,
                         float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pa = a + static_cast<size_t>(row) * D;
  const float* pb = b ? b + static_cast<size_t>(row) * D : nullptr;
  float s1 = 0.f;
  for (int i = lane; i < D; i += 32) s1 += __ldg(pa + i) + (pb ? __ldg(pb + i) : 0.f);
  const float mean = warp_sum(s1) / static_cast<float>(D);
  float s2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float d = __ldg(pa + i) + (pb ? __ldg(pb + i) : 0.f) - mean;
    s2 = fmaf(d, d, s2);
  }
  const float rstd = rsqrtf(warp_sum(s2) / static_cast<float>(D) + eps);
  for (int i = lane; i < D; i += 32) {
    const float d = __ldg(pa + i) + (pb ? __ldg(pb + i) : 0.f) - mean;
    y[static_cast<size_t>(row) * D + i] = fmaf(d * rstd, __ldg(w + i), __ldg(bias + i));
  }
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// ds = rstd * (dy w - mean_D(dy w) - xh mean_D(dy w xh))  (gradient of both a and b);  dw += dy xh;  dbias += dy
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                         const float* __restrict__ stats, const float* __restrict__ dy, int rows, int D,
                         float* __restrict__ ds, float* __restrict__ dw, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pa = a + static_cast<size_t>(row) * D;
  const float* pb = b ? b + static_cast<size_t>(row) * D : nullptr;
  const float* pd = dy + static_cast<size_t>(row) * D;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float c1 = 0.f, c2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float xh = (__ldg(pa + i) + (pb ? __ldg(pb + i) : 0.f) - mean) * rstd;
    const float g = __ldg(pd + i) * __ldg(w + i);
    c1 += g;
    c2 = fmaf(g, xh, c2);
  }
  c1 = warp_sum(c1) / static_cast<float>(D);
  c2 = warp_sum(c2) / static_cast<float>(D);
  for (int i = lane; i < D; i += 32) {
    const float xh = (__ldg(pa + i) + (pb ? __ldg(pb + i) : 0.f) - mean) * rstd;
    const float d = __ldg(pd + i);
    ds[static_cast<size_t>(row) * D + i] = rstd * (d * __ldg(w + i) - c1 - xh * c2);
    atomicAdd(dw + i, d * xh);
    atomicAdd(dbias + i, d);
  }
}

int add_layernorm_fwd(const float* a, const float* b, const float* w, const float* bias, int rows, int D, float eps, float* y,
                      float* stats, cudaStream_t stream) {
  if (!a || !w || !bias || !y || !stats || rows <= 0 || D <= 0) { set_error("add_layernorm_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  add_layernorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(a, b, w, bias, rows, D, eps, y, stats);
  return check_launch("add_layernorm_fwd_kernel");
}

int add_layernorm_bwd(const float* a, const float* b, const float* w, const float* stats, const float* dy, int rows, int D,
                      float* ds, float* dw, float* dbias, cudaStream_t stream) {
  if (!a || !w || !stats || !dy || !ds || !dw || !dbias || rows <= 0 || D <= 0) { set_error("add_layernorm_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * D, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbias, 0, sizeof(float) * D, stream);
  if (e != cudaSuccess) { set_error("add_layernorm_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  add_layernorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(a, b, w, stats, dy, rows, D, ds, dw, dbias);
  return check_launch("add_layernorm_bwd_kernel");
}

}  // namespace l2i
