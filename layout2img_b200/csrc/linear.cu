// The small dense layers of the path as hand-written fp32 kernels (no cuBLAS / ATen on the hot path):
//   * nn.Linear layers -- the 308-wide attention projections (resnet_generator_app_v2.py:148-151,208-212), the generator's
//     fc (:409, 128 -> 16384), the mask-regression fc (mask_regression.py:64), the 20 ISLA gamma / beta projections
//     (norm_module.py:158-159), the 1x1 convolutions of the PSP stages on pooled cells (:741-746) -- are one generic
//     strided SIMT GEMM:  C[m,n] = (sum_k A(m,k) B(k,n)) / sigma + bias[n]   with arbitrary element strides, which covers
//     y = x W^T (+ b), dx = dy W and dW = dy^T x without transposing anything.  The GEMMs are tiny (<= 0.003 GMAC per
//     image): a 64 x 64 x 16 register-tiled fp32 kernel is ample; the forward form accumulates in a fixed order.
//   * LayerNorm(x + residual) of the attention block (:201-212), forward and backward, one warp per row.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

static constexpr int kGT = 64, kGK = 16;      // tile M = N = 64, K step 16; 256 threads, 4 x 4 outputs each

__global__ void __launch_bounds__(256)
gemm_strided_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
                    long long sbn, int M, int N, int K, const float* __restrict__ sigma, const float* __restrict__ bias,
                    float* __restrict__ C, long long scm, int k_per_split) {
  __shared__ __align__(16) float sA[kGK][kGT + 4], sB[kGK][kGT + 4];      // rows 16-byte aligned: float4 reads below
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kGT, n0 = blockIdx.x * kGT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // split-K (gridDim.z > 1): this block reduces k in [kb, ke) and adds its partial with atomics into the zeroed C
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  // each thread stages 4 elements of each operand per K step; the faster-varying index follows the unit stride.  The
  // global loads of step i + 1 are issued before the FMAs of step i (register prefetch), so their latency is hidden.
  int am[4], ak[4], bn[4], bk[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = threadIdx.x + r * 256;
    if (sak == 1) { ak[r] = e % kGK; am[r] = e / kGK; } else { am[r] = e % kGT; ak[r] = e / kGT; }
    if (sbk == 1) { bk[r] = e % kGK; bn[r] = e / kGK; } else { bn[r] = e % kGT; bk[r] = e / kGT; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + am[r], k = k0 + ak[r];
      ra[r] = (m < M && k < ke) ? __ldg(A + m * sam + k * sak) : 0.f;
      const int n = n0 + bn[r], k2 = k0 + bk[r];
      rb[r] = (n < N && k2 < ke) ? __ldg(B + k2 * sbk + n * sbn) : 0.f;
    }
  };
  if (kb < ke) fetch(kb);
  for (int k0 = kb; k0 < ke; k0 += kGK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) { sA[ak[r]][am[r]] = ra[r]; sB[bk[r]][bn[r]] = rb[r]; }
    __syncthreads();
    if (k0 + kGK < ke) fetch(k0 + kGK);
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float sg = sigma ? __ldg(sigma) : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = sigma ? acc[i][j] / sg : acc[i][j];
      if (bias && blockIdx.z == 0) v += __ldg(bias + n);
      float* o = C + m * scm + n;
      if (gridDim.z > 1) atomicAdd(o, v); else *o = v;
    }
  }
}

int gemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, int M, int N, int K,
                 const float* sigma, const float* bias, float* C, long long scm, int allow_split, cudaStream_t stream) {
  if (!A || !B || !C || M < 0 || N <= 0 || K <= 0) { set_error("gemm: bad arguments (M=%d N=%d K=%d)", M, N, K); return L2I_ERR_BAD_ARG; }
  if (M == 0) return L2I_OK;
  const int gx = (N + kGT - 1) / kGT, gy = (M + kGT - 1) / kGT;
  // few output tiles and a long reduction (dx of the 16384-wide fc: 2 tiles, K = 16384): split K over ~2 waves of CTAs
  int splits = 1;
  // (backward GEMMs only: the forward stays bit-reproducible from run to run)
  if (allow_split && gx * gy < 148 && K >= 256) {
    splits = (296 + gx * gy - 1) / (gx * gy);
    if (splits > K / 64) splits = K / 64;
    if (splits < 1) splits = 1;
  }
  int kps = (K + splits - 1) / splits;
  kps = (kps + kGK - 1) / kGK * kGK;
  splits = (K + kps - 1) / kps;
  if (splits > 1) {
    if (scm != N) { set_error("gemm: split-K needs a dense output"); return L2I_ERR_UNSUPPORTED; }
    cudaError_t e = cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, stream);
    if (e != cudaSuccess) { set_error("gemm: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  }
  dim3 grid(gx, gy, splits);
  gemm_strided_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, M, N, K, sigma, bias, C, scm, kps);
  return check_launch("gemm_strided_kernel");
}

// column sums of a [M, N] row-major matrix (bias gradients): one block per 32 columns
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
