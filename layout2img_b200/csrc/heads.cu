// The discriminator's output heads (reference model/rcnn_discriminator_app.py):
//
//  * projection heads (:125-127 image head, :160-166 object head):
//        s[n,:] = sum over pixels of relu(feat[n,p,:]);   out[n] = s[n,:] . w + bias  (+ s[n,:] . e[y[n],:])
//    with w = l7 / l_obj weight (1, C) and e = l_y embedding (num_classes, C), both spectrally normalised
//    (the 1/sigma factors are device scalars produced by the spectral-norm kernels).
//  * appearance head (:148-157):  F = relu(app_conv(x)) as (K, C, P);  Gram = F F^T / C;
//        out[k] = (1/C) sum_i Linear([Gram[i,:], e_app[y[k],:]])
//               = (1/C^2) sum_p (sum_c F[k,p,c]) (sum_c F[k,p,c] w1[c]) + e_app[y[k],:] . w2 + bias
//    -- evaluated without forming the (K, C, C) Gram matrix or the (K, C, 2C) concatenation (1 GB at config A).
//
// One block per image / object; the feature map of that image ([P, C] fp32, NHWC) is read once in the forward and once
// in the backward.  HBM-bound: 4 B per element forward, 8 B backward.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

__device__ __forceinline__ float block_sum_256(float v, float* red) {   // blockDim.x == 256, all threads call
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ------------------------------------------------------------------------------------------ projection heads
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ feat, int P, int C, const float* __restrict__ w, const float* __restrict__ sigma_w,
                const float* __restrict__ bias, const float* __restrict__ emb, const float* __restrict__ sigma_e,
                const long long* __restrict__ y, float* __restrict__ s, float* __restrict__ out) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float iw = 1.0f / __ldg(sigma_w);
  const float ie = emb ? 1.0f / __ldg(sigma_e) : 0.f;
  const float* e = emb ? emb + static_cast<size_t>(y[n]) * C : nullptr;
  const float* f = feat + static_cast<size_t>(n) * P * C;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += fmaxf(__ldg(f + static_cast<size_t>(p) * C + c), 0.f);
    s[static_cast<size_t>(n) * C + c] = acc;
    float wv = __ldg(w + c) * iw;
    if (e) wv = fmaf(__ldg(e + c), ie, wv);
    dot = fmaf(acc, wv, dot);
  }
  dot = block_sum_256(dot, red);
  if (threadIdx.x == 0) out[n] = dot + (bias ? __ldg(bias) : 0.f);
}

// dfeat[n,p,c] = relu'(feat) * dout[n] * (w[c]/sigma_w + e[y[n],c]/sigma_e);  gw[c] += dout[n] s[n,c] (= dL/d(w/sigma_w));
// gemb[y[n],c] += dout[n] s[n,c];  dbias += dout[n]
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ s, const float* __restrict__ dout, int P, int C,
                const float* __restrict__ w, const float* __restrict__ sigma_w, const float* __restrict__ emb,
                const float* __restrict__ sigma_e, const long long* __restrict__ y, float* __restrict__ dfeat,
                float* __restrict__ gw, float* __restrict__ gemb, float* __restrict__ dbias) {
  const int n = blockIdx.x;
  const float d = __ldg(dout + n);
  const float iw = 1.0f / __ldg(sigma_w);
  const float ie = emb ? 1.0f / __ldg(sigma_e) : 0.f;
  const long long yn = emb ? y[n] : 0;
  const float* e = emb ? emb + static_cast<size_t>(yn) * C : nullptr;
  const float* f = feat + static_cast<size_t>(n) * P * C;
  float* df = dfeat ? dfeat + static_cast<size_t>(n) * P * C : nullptr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float wv = __ldg(w + c) * iw;
    if (e) wv = fmaf(__ldg(e + c), ie, wv);
    const float g = d * wv;
    if (df)
      for (int p = 0; p < P; ++p) df[static_cast<size_t>(p) * C + c] = (__ldg(f + static_cast<size_t>(p) * C + c) > 0.f) ? g : 0.f;
    const float ds = d * __ldg(s + static_cast<size_t>(n) * C + c);
    if (gw) atomicAdd(gw + c, ds);
    if (gemb) atomicAdd(gemb + static_cast<size_t>(yn) * C + c, ds);
  }
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, d);
}

int head_fwd(const float* feat, int N, int P, int C, const float* w, const float* sigma_w, const float* bias, const float* emb,
             const float* sigma_e, const long long* y, float* s, float* out, cudaStream_t stream) {
  if (!feat || !w || !sigma_w || !s || !out || N < 0 || P <= 0 || C <= 0 || (emb && (!sigma_e || !y))) {
    set_error("head_fwd: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  if (N == 0) return L2I_OK;
  head_fwd_kernel<<<N, 256, 0, stream>>>(feat, P, C, w, sigma_w, bias, emb, sigma_e, y, s, out);
  return check_launch("head_fwd_kernel");
}

int head_bwd(const float* feat, const float* s, const float* dout, int N, int P, int C, const float* w, const float* sigma_w,
             const float* emb, const float* sigma_e, const long long* y, int num_emb, float* dfeat, float* gw, float* gemb,
             float* dbias, cudaStream_t stream) {
  if (!feat || !s || !dout || !w || !sigma_w || N < 0 || P <= 0 || C <= 0 || (emb && (!sigma_e || !y)) || (gemb && (!emb || num_emb <= 0))) {
    set_error("head_bwd: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaSuccess;
  if (gw) e = cudaMemsetAsync(gw, 0, sizeof(float) * C, stream);
  if (e == cudaSuccess && gemb) e = cudaMemsetAsync(gemb, 0, sizeof(float) * static_cast<size_t>(num_emb) * C, stream);
  if (e == cudaSuccess && dbias) e = cudaMemsetAsync(dbias, 0, sizeof(float), stream);
  if (e != cudaSuccess) { set_error("head_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  if (N == 0) return L2I_OK;
  head_bwd_kernel<<<N, 256, 0, stream>>>(feat, s, dout, P, C, w, sigma_w, emb, sigma_e, y, dfeat, gw, gemb, dbias);
  return check_launch("head_bwd_kernel");
}

// ------------------------------------------------------------------------------------------ appearance head
// block = one object k; warp <-> pixel rows p = warp, warp + 8, ...; lanes stride the channels (float4)
__global__ void __launch_bounds__(256)
gram_proj_fwd_kernel(const float* __restrict__ x, int P, int C, const float* __restrict__ w, const float* __restrict__ sigma_w,
                     const float* __restrict__ bias, const float* __restrict__ emb, const float* __restrict__ sigma_e,
                     const long long* __restrict__ y, float* __restrict__ colsum, float* __restrict__ proj,
                     float* __restrict__ out) {
  __shared__ float red[32];
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float iw = 1.0f / __ldg(sigma_w);
  const float* xk = x + static_cast<size_t>(k) * P * C;
  float acc = 0.f;
  for (int p = warp; p < P; p += 8) {
    float cs = 0.f, pr = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xk + static_cast<size_t>(p) * C + c));
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
      const float a0 = fmaxf(v.x, 0.f), a1 = fmaxf(v.y, 0.f), a2 = fmaxf(v.z, 0.f), a3 = fmaxf(v.w, 0.f);
      cs += (a0 + a1) + (a2 + a3);
      pr = fmaf(a0, w4.x, fmaf(a1, w4.y, fmaf(a2, w4.z, fmaf(a3, w4.w, pr))));
    }
    cs = warp_sum(cs);
    pr = warp_sum(pr) * iw;
    if (lane == 0) {
      colsum[static_cast<size_t>(k) * P + p] = cs;
      proj[static_cast<size_t>(k) * P + p] = pr;
    }
    acc = fmaf(cs, pr, acc);                    // identical in every lane
  }
  float tot = (lane == 0) ? acc : 0.f;
  // class term: e_app[y[k],:] . w2
  const float ie = 1.0f / __ldg(sigma_e);
  const float* e = emb + static_cast<size_t>(y[k]) * C;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dot = fmaf(__ldg(e + c) * ie, __ldg(w + C + c) * iw, dot);
  const float inv_c2 = 1.0f / (static_cast<float>(C) * static_cast<float>(C));
  tot = block_sum_256(tot * inv_c2 + dot, red);
  if (threadIdx.x == 0) out[k] = tot + (bias ? __ldg(bias) : 0.f);
}

// dx[k,p,c] = relu'(x) dout[k]/C^2 (proj[k,p] + colsum[k,p] w1[c]);  gw[c] += dout[k]/C^2 sum_p colsum[k,p] F[k,p,c]  (c < C),
// gw[C + c] += dout[k] e[y[k],c];  gemb[y[k],c] += dout[k] w2[c];  dbias += dout[k]     (w1, w2, e already divided by sigma)
__global__ void __launch_bounds__(256)
gram_proj_bwd_kernel(const float* __restrict__ x, const float* __restrict__ colsum, const float* __restrict__ proj,
                     const float* __restrict__ dout, int P, int C, const float* __restrict__ w, const float* __restrict__ sigma_w,
                     const float* __restrict__ emb, const float* __restrict__ sigma_e, const long long* __restrict__ y,
                     float* __restrict__ dx, float* __restrict__ gw, float* __restrict__ gemb, float* __restrict__ dbias) {
  extern __shared__ float s_gw[];               // [C] cross-warp reduction of the w1 gradient
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float iw = 1.0f / __ldg(sigma_w), ie = 1.0f / __ldg(sigma_e);
  const float d = __ldg(dout + k);
  const float coef = d / (static_cast<float>(C) * static_cast<float>(C));
  const float* xk = x + static_cast<size_t>(k) * P * C;
  float* dk = dx + static_cast<size_t>(k) * P * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_gw[c] = 0.f;
  __syncthreads();
  for (int c = lane * 4; c < C; c += 128) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
    const float w1[4] = {w4.x * iw, w4.y * iw, w4.z * iw, w4.w * iw};
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = warp; p < P; p += 8) {
      const float cs = __ldg(colsum + static_cast<size_t>(k) * P + p), pr = __ldg(proj + static_cast<size_t>(k) * P + p);
      const float4 v = __ldg(reinterpret_cast<const float4*>(xk + static_cast<size_t>(p) * C + c));
      const float xv[4] = {v.x, v.y, v.z, v.w};
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool on = xv[j] > 0.f;
        r[j] = on ? coef * fmaf(cs, w1[j], pr) : 0.f;
        g[j] = fmaf(cs, on ? xv[j] : 0.f, g[j]);
      }
      *reinterpret_cast<float4*>(dk + static_cast<size_t>(p) * C + c) = make_float4(r[0], r[1], r[2], r[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&s_gw[c + j], g[j]);
  }
  __syncthreads();
  const long long yk = y[k];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(gw + c, coef * s_gw[c]);
    atomicAdd(gw + C + c, d * __ldg(emb + static_cast<size_t>(yk) * C + c) * ie);
    atomicAdd(gemb + static_cast<size_t>(yk) * C + c, d * __ldg(w + C + c) * iw);
  }
  if (threadIdx.x == 0) atomicAdd(dbias, d);
}

int gram_proj_fwd(const float* x, int K, int P, int C, const float* w, const float* sigma_w, const float* bias, const float* emb,
                  const float* sigma_e, const long long* y, float* colsum, float* proj, float* out, cudaStream_t stream) {
  if (!x || !w || !sigma_w || !emb || !sigma_e || !y || !colsum || !proj || !out || K < 0 || P <= 0 || C <= 0 || (C & 3)) {
    set_error("gram_proj_fwd: bad arguments (C must be a multiple of 4)");
    return L2I_ERR_BAD_ARG;
  }
  if (K == 0) return L2I_OK;
  gram_proj_fwd_kernel<<<K, 256, 0, stream>>>(x, P, C, w, sigma_w, bias, emb, sigma_e, y, colsum, proj, out);
  return check_launch("gram_proj_fwd_kernel");
}

int gram_proj_bwd(const float* x, const float* colsum, const float* proj, const float* dout, int K, int P, int C, const float* w,
                  const float* sigma_w, const float* emb, const float* sigma_e, const long long* y, int num_emb, float* dx,
                  float* gw, float* gemb, float* dbias, cudaStream_t stream) {
  if (!x || !colsum || !proj || !dout || !w || !sigma_w || !emb || !sigma_e || !y || !dx || !gw || !gemb || !dbias || K < 0 ||
      P <= 0 || C <= 0 || (C & 3) || num_emb <= 0 || C > 8192) {
    set_error("gram_proj_bwd: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaMemsetAsync(gw, 0, sizeof(float) * 2 * C, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(gemb, 0, sizeof(float) * static_cast<size_t>(num_emb) * C, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbias, 0, sizeof(float), stream);
  if (e != cudaSuccess) { set_error("gram_proj_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  if (K == 0) return L2I_OK;
  gram_proj_bwd_kernel<<<K, 256, sizeof(float) * C, stream>>>(x, colsum, proj, dout, P, C, w, sigma_w, emb, sigma_e, y, dx, gw,
                                                              gemb, dbias);
  return check_launch("gram_proj_bwd_kernel");
}

}  // namespace l2i
