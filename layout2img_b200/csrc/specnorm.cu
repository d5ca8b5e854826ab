// Spectral normalisation as the reference applies it to every conv / linear weight
// (torch.nn.utils.spectral_norm, call sites resnet_generator_app_v2.py:681-686, rcnn_discriminator_app.py:10-15):
//   training:  v <- normalize(W^T u, eps);  u <- normalize(W v, eps)      (one power iteration, in place)
//   always:    sigma = u . (W v);   W_sn = W / sigma
//   backward:  dW = (G - <G, W>/sigma * u v^T) / sigma                    (u, v constants)
// W is viewed as [R, Cc] (R = out channels).  torch runs this as ~15 small launches per module and
// materialises W / sigma; here: 3 launches (two passes over W) produce sigma, which the operand-preparation
// kernel folds into the bf16 split, and 2 launches map the tensor-core weight gradient G (layout
// [R][taps][cin]) to the gradient of weight_orig (torch layout [R][cin][taps]).  HBM-bound: 2 reads of W
// forward, 1 read of W + 1 read of G + 1 write backward.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

__device__ __forceinline__ float block_sum(float v, float* red) {   // blockDim.x <= 1024; all threads call
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// t[c] (+)= sum_{r in [r0, r1)} W[r,c] * u[r] for the 256 columns of column block `cb`
__device__ __forceinline__ void sn_wt_u_body(const float* __restrict__ W, const float* __restrict__ u, int Cc, int r0, int r1,
                                             int cb, float* __restrict__ t, float* s_u) {
  for (int i = threadIdx.x; i < r1 - r0; i += blockDim.x) s_u[i] = __ldg(u + r0 + i);
  __syncthreads();
  const int c = cb * blockDim.x + threadIdx.x;
  if (c >= Cc) return;
  float acc = 0.f;
  const float* wp = W + static_cast<size_t>(r0) * Cc + c;
#pragma unroll 4
  for (int r = r0; r < r1; ++r, wp += Cc) acc = fmaf(__ldg(wp), s_u[r - r0], acc);
  atomicAdd(t + c, acc);
}
__global__ void __launch_bounds__(256) sn_wt_u_kernel(const float* __restrict__ W, const float* __restrict__ u, int R, int Cc,
                                                      int rows_per_split, float* __restrict__ t) {
  extern __shared__ float s_u[];
  const int r0 = blockIdx.y * rows_per_split, r1 = min(R, r0 + rows_per_split);
  sn_wt_u_body(W, u, Cc, r0, r1, blockIdx.x, t, s_u);
}

// v = normalize_v ? t / max(|t|, eps) : t;  s[r] = W[r,:] . v  (one warp per row);  block 0 also stores v
__device__ __forceinline__ void sn_w_v_body(const float* __restrict__ W, const float* __restrict__ t, int R, int Cc,
                                            int normalize_v, float eps, float* __restrict__ v_out,
                                            float* __restrict__ v_out2, float* __restrict__ s, int row_block, float* s_v) {
  float* red = s_v + Cc;
  float inv = 1.f;
  if (normalize_v) {
    float q = 0.f;
    for (int i = threadIdx.x; i < Cc; i += blockDim.x) { const float x = __ldg(t + i); q = fmaf(x, x, q); }
    const float n2 = block_sum(q, red);
    inv = 1.0f / fmaxf(sqrtf(n2), eps);
  }
  for (int i = threadIdx.x; i < Cc; i += blockDim.x) {
    const float x = __ldg(t + i) * inv;
    s_v[i] = x;
    if (row_block == 0) {
      if (v_out) v_out[i] = x;
      if (v_out2) v_out2[i] = x;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = row_block * (blockDim.x >> 5) + warp;
  if (r >= R) return;
  const float* wp = W + static_cast<size_t>(r) * Cc;
  float acc = 0.f;
  for (int c = lane; c < Cc; c += 32) acc = fmaf(__ldg(wp + c), s_v[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) s[r] = acc;
}
__global__ void __launch_bounds__(256) sn_w_v_kernel(const float* __restrict__ W, const float* __restrict__ t, int R, int Cc,
                                                     int normalize_v, float eps, float* __restrict__ v_out,
                                                     float* __restrict__ v_out2, float* __restrict__ s) {
  extern __shared__ float s_v[];          // [Cc] + 32
  sn_w_v_body(W, t, R, Cc, normalize_v, eps, v_out, v_out2, s, blockIdx.x, s_v);
}

// update_u: u = s / max(|s|, eps) (stored to u_out and u_out2);  sigma = u . s
__device__ __forceinline__ void sn_finish_body(const float* __restrict__ s, const float* __restrict__ u_in, int R,
                                               int update_u, float eps, float* __restrict__ u_out,
                                               float* __restrict__ u_out2, float* __restrict__ sigma, float* red) {
  float inv = 1.f;
  if (update_u) {
    float q = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) { const float x = __ldg(s + i); q = fmaf(x, x, q); }
    inv = 1.0f / fmaxf(sqrtf(block_sum(q, red)), eps);
  }
  float d = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const float sv = __ldg(s + i);
    const float uv = update_u ? sv * inv : __ldg(u_in + i);
    if (update_u && u_out) u_out[i] = uv;
    if (u_out2) u_out2[i] = uv;
    d = fmaf(uv, sv, d);
  }
  d = block_sum(d, red);
  if (threadIdx.x == 0) *sigma = d;
}
__global__ void __launch_bounds__(1024) sn_finish_kernel(const float* __restrict__ s, const float* __restrict__ u_in, int R,
                                                         int update_u, float eps, float* __restrict__ u_out,
                                                         float* __restrict__ u_out2, float* __restrict__ sigma) {
  __shared__ float red[32];
  sn_finish_body(s, u_in, R, update_u, eps, u_out, u_out2, sigma, red);
}

// ---- grouped forms: one launch per phase for ALL spectrally normalised modules of a network call.  `tab` is the
// device table of SnEntry (kernels.h), `items` the static work list of the phase, `f32` the per-call fp32 buffer that
// receives every module's (sigma, u_used, v_used) and holds its scratch (t, s) at the entry's offsets.
__global__ void __launch_bounds__(256) sn_wt_u_group_kernel(const SnEntry* __restrict__ tab, const int4* __restrict__ items,
                                                            float* __restrict__ f32) {
  extern __shared__ float s_u[];
  const int4 it = items[blockIdx.x];                  // (module, column block, r0, r1)
  const SnEntry e = tab[it.x];
  if (!e.has_sn || !e.training) return;
  sn_wt_u_body(e.W, e.u, e.Cc, it.z, it.w, it.y, f32 + e.f32_off + sn_off_t(e.R, e.Cc), s_u);
}
__global__ void __launch_bounds__(256) sn_w_v_group_kernel(const SnEntry* __restrict__ tab, const int2* __restrict__ items,
                                                           float* __restrict__ f32) {
  extern __shared__ float s_v[];
  const int2 it = items[blockIdx.x];                  // (module, row block)
  const SnEntry e = tab[it.x];
  if (!e.has_sn) return;
  float* base = f32 + e.f32_off;
  sn_w_v_body(e.W, e.training ? base + sn_off_t(e.R, e.Cc) : e.v, e.R, e.Cc, e.training, e.eps, e.training ? e.v : nullptr,
              base + sn_off_v(e.R, e.Cc), base + sn_off_s(e.R, e.Cc), it.y, s_v);
}
__global__ void __launch_bounds__(1024) sn_finish_group_kernel(const SnEntry* __restrict__ tab, float* __restrict__ f32) {
  __shared__ float red[32];
  const SnEntry e = tab[blockIdx.x];
  if (!e.has_sn) return;
  float* base = f32 + e.f32_off;
  sn_finish_body(base + sn_off_s(e.R, e.Cc), e.u, e.R, e.training, e.eps, e.u, base + sn_off_u(), base, red);
}

int sn_group_sigma(const void* table, int n_modules, const int* wt_items, int n_wt, int wt_smem_floats, const int* wv_items,
                   int n_wv, int max_cc, float* f32, long long f32_floats, cudaStream_t stream) {
  if (!table || n_modules <= 0 || !f32 || f32_floats <= 0 || max_cc > 40000) { set_error("sn_group_sigma: bad arguments"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(f32, 0, sizeof(float) * f32_floats, stream);      // zeroes every module's t accumulator
  if (e != cudaSuccess) { set_error("sn_group_sigma: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(sn_w_v_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40032 * 4);
    configured.done();
  }
  const SnEntry* tab = reinterpret_cast<const SnEntry*>(table);
  int rc;
  if (n_wt > 0) {
    sn_wt_u_group_kernel<<<n_wt, 256, sizeof(float) * wt_smem_floats, stream>>>(tab, reinterpret_cast<const int4*>(wt_items), f32);
    if ((rc = check_launch("sn_wt_u_group_kernel"))) return rc;
  }
  if (n_wv > 0) {
    sn_w_v_group_kernel<<<n_wv, 256, sizeof(float) * (max_cc + 32), stream>>>(tab, reinterpret_cast<const int2*>(wv_items), f32);
    if ((rc = check_launch("sn_w_v_group_kernel"))) return rc;
    sn_finish_group_kernel<<<n_modules, 1024, 0, stream>>>(tab, f32);
    if ((rc = check_launch("sn_finish_group_kernel"))) return rc;
  }
  return L2I_OK;
}

int sn_sigma(const float* W, int R, int Cc, float* u, float* v, int training, float eps, float* u_used, float* v_used,
             float* sigma, float* work, cudaStream_t stream) {
  // work: [Cc + R] floats of scratch (t, s)
  if (!W || !u || !v || !sigma || !work || !u_used || !v_used || R <= 0 || Cc <= 0 || Cc > 40000) {
    set_error("sn_sigma: bad arguments (R=%d Cc=%d)", R, Cc);
    return L2I_ERR_BAD_ARG;
  }
  float* t = work;
  float* s = work + Cc;
  const size_t smem_v = sizeof(float) * (Cc + 32);
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(sn_w_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40032 * 4);
    configured.done();
  }
  if (training) {
    cudaError_t e = cudaMemsetAsync(t, 0, sizeof(float) * Cc, stream);
    if (e != cudaSuccess) { set_error("sn_sigma: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
    const int col_blocks = (Cc + 255) / 256;
    int splits = (296 + col_blocks - 1) / col_blocks;
    if (splits > (R + 15) / 16) splits = (R + 15) / 16;
    if (splits < 1) splits = 1;
    const int rps = (R + splits - 1) / splits;
    splits = (R + rps - 1) / rps;
    sn_wt_u_kernel<<<dim3(col_blocks, splits), 256, sizeof(float) * rps, stream>>>(W, u, R, Cc, rps, t);
    int rc = check_launch("sn_wt_u_kernel");
    if (rc) return rc;
    sn_w_v_kernel<<<(R + 7) / 8, 256, smem_v, stream>>>(W, t, R, Cc, 1, eps, v, v_used, s);
  } else {
    sn_w_v_kernel<<<(R + 7) / 8, 256, smem_v, stream>>>(W, v, R, Cc, 0, eps, nullptr, v_used, s);
  }
  int rc = check_launch("sn_w_v_kernel");
  if (rc) return rc;
  sn_finish_kernel<<<1, 1024, 0, stream>>>(s, u, R, training, eps, u, u_used, sigma);
  return check_launch("sn_finish_kernel");
}

// d += sum G[r][tap][ci] * W[r][ci][tap].  One block per output channel r: both rows are contiguous
// (taps * cin floats), G is read coalesced and the tap-strided reads of W stay inside the row's L1 lines.
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ G, const float* __restrict__ W, int R, int cin,
                                                         int taps, float* __restrict__ d) {
  __shared__ float red[32];
  const int n = cin * taps;
  float acc = 0.f;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* g = G + static_cast<size_t>(r) * n;
    const float* w = W + static_cast<size_t>(r) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {     // i indexes G: (tap, ci)
      const int tap = i / cin, ci = i - tap * cin;
      acc = fmaf(__ldg(g + i), __ldg(w + ci * taps + tap), acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(d, acc);
}

// dW[r][ci][tap] = (G[r][tap][ci] - d / sigma * u[r] * v[ci * taps + tap]) / sigma   (block per r, coalesced writes)
__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* __restrict__ G, const float* __restrict__ u,
                                                           const float* __restrict__ v, const float* __restrict__ sigma,
                                                           const float* __restrict__ d, int R, int cin, int taps,
                                                           float* __restrict__ dW) {
  const float inv = 1.0f / __ldg(sigma);
  const float coef = __ldg(d) * inv;
  const int n = cin * taps;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* g = G + static_cast<size_t>(r) * n;
    float* o = dW + static_cast<size_t>(r) * n;
    const float cu = coef * __ldg(u + r);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {     // i indexes dW: (ci, tap)
      const int ci = i / taps, tap = i - ci * taps;
      o[i] = (__ldg(g + tap * cin + ci) - cu * __ldg(v + i)) * inv;
    }
  }
}

int sn_weight_grad(const float* G, const float* W, const float* u, const float* v, const float* sigma, int R, int cin,
                   int taps, float* dW, float* scratch, cudaStream_t stream) {
  if (!G || !W || !u || !v || !sigma || !dW || !scratch || R <= 0 || cin <= 0 || taps <= 0) {
    set_error("sn_weight_grad: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float), stream);
  if (e != cudaSuccess) { set_error("sn_weight_grad: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  const int blocks = R < 148 * 8 ? R : 148 * 8;
  sn_bwd_dot_kernel<<<blocks, 256, 0, stream>>>(G, W, R, cin, taps, scratch);
  int rc = check_launch("sn_bwd_dot_kernel");
  if (rc) return rc;
  sn_bwd_apply_kernel<<<blocks, 256, 0, stream>>>(G, u, v, sigma, scratch, R, cin, taps, dW);
  return check_launch("sn_bwd_apply_kernel");
}

}  // namespace l2i

static_assert(sizeof(l2i::SnEntry) == 72, "SnEntry layout is part of the C ABI (include/l2i.h)");
