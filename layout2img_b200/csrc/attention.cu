// Object-context attention: box_attention + BoxRelationalEmbedding + the WGs geometry gate of
// BoxMultiHeadedAttention (reference model/resnet_generator_app_v2.py:17-120,172-192), one head.
//
//   emb_ij  = [sin(100*pos_ij*f), cos(100*pos_ij*f)]  (64-d), pos = log-ratios of box centres / sizes
//   geo_ij  = relu(emb_ij . wg + bg)
//   w_ij    = log(max(geo_ij, 1e-6)) + masked_fill(q_i.k_j / sqrt(D), y_j == 0, -1e9)
//   out_i   = sum_j softmax_j(w_ij) v_j
//
// The sequence is the object list (O <= 48): one CTA per image, one warp per query row, the O x O
// score matrix lives in shared memory; q/k/v rows are read with coalesced lane-strided loads and
// reduced with warp shuffles.  Backward in the same launch shape.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

static constexpr int kMaxO = 48;

// bbox columns are read as (x0, y0, x1, y1) although callers pass xywh -- reference behaviour (:30-35)
__device__ void rel_embedding(const float* __restrict__ bb, int i, int j, float* emb /*64*/) {
  const float xi0 = bb[i * 4], yi0 = bb[i * 4 + 1], xi1 = bb[i * 4 + 2], yi1 = bb[i * 4 + 3];
  const float xj0 = bb[j * 4], yj0 = bb[j * 4 + 1], xj1 = bb[j * 4 + 2], yj1 = bb[j * 4 + 3];
  const float cxi = (xi0 + xi1) * 0.5f, cyi = (yi0 + yi1) * 0.5f, wi = (xi1 - xi0) + 1.0f, hi = (yi1 - yi0) + 1.0f;
  const float cxj = (xj0 + xj1) * 0.5f, cyj = (yj0 + yj1) * 0.5f, wj = (xj1 - xj0) + 1.0f, hj = (yj1 - yj0) + 1.0f;
  float pos[4];
  pos[0] = logf(fmaxf(fabsf((cxi - cxj) / wi), 1e-3f));
  pos[1] = logf(fmaxf(fabsf((cyi - cyj) / hi), 1e-3f));
  pos[2] = logf(wi / wj);
  pos[3] = logf(hi / hj);
#pragma unroll
  for (int d = 0; d < 4; ++d)
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      const float freq = 1.0f / powf(1000.0f, static_cast<float>(f) / 8.0f);
      const float ang = 100.0f * pos[d] * freq;
      emb[d * 8 + f] = sinf(ang);
      emb[32 + d * 8 + f] = cosf(ang);
    }
}

// q,k,v,out (B,O,D); bbox (B,O,4); y (B,O) int64; wg (64), bg (1); saves p (B,O,O), glin (B,O,O)
__global__ void __launch_bounds__(256) box_attention_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, const float* __restrict__ bbox,
                                                                const long long* __restrict__ y, const float* __restrict__ wg,
                                                                const float* __restrict__ bg, int O, int D,
                                                                float* __restrict__ out, float* __restrict__ p_save,
                                                                float* __restrict__ glin_save) {
  __shared__ float s_w[kMaxO * kMaxO];
  __shared__ float s_bb[kMaxO * 4];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < O * 4; i += blockDim.x) s_bb[i] = __ldg(bbox + 1LL * b * O * 4 + i);
  __syncthreads();
  // geometry gate (pre-ReLU linear value saved for the backward)
  for (int ij = threadIdx.x; ij < O * O; ij += blockDim.x) {
    float emb[64];
    rel_embedding(s_bb, ij / O, ij % O, emb);
    float acc = __ldg(bg);
#pragma unroll
    for (int t = 0; t < 64; ++t) acc = fmaf(emb[t], __ldg(wg + t), acc);
    glin_save[1LL * b * O * O + ij] = acc;
    s_w[ij] = logf(fmaxf(fmaxf(acc, 0.f), 1e-6f));
  }
  __syncthreads();
  const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(D));
  for (int i = warp; i < O; i += nwarp) {
    const float* qi = q + (1LL * b * O + i) * D;
    for (int j = 0; j < O; ++j) {
      const float* kj = k + (1LL * b * O + j) * D;
      float acc = 0.f;
      for (int d = lane; d < D; d += 32) acc = fmaf(__ldg(qi + d), __ldg(kj + d), acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        float sc = acc * inv_sqrt_d;
        if (__ldg(y + 1LL * b * O + j) == 0) sc = -1e9f;
        s_w[i * O + j] += sc;
      }
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < O; j += 32) mx = fmaxf(mx, s_w[i * O + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < O; j += 32) {
      const float e = expf(s_w[i * O + j] - mx);
      s_w[i * O + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int j = lane; j < O; j += 32) {
      const float pv = s_w[i * O + j] / sum;
      s_w[i * O + j] = pv;
      p_save[(1LL * b * O + i) * O + j] = pv;
    }
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < O; ++j) acc = fmaf(s_w[i * O + j], __ldg(v + (1LL * b * O + j) * D + d), acc);
      out[(1LL * b * O + i) * D + d] = acc;
    }
  }
}

// dq, dk, dv (B,O,D); dwg (64) and dbg (1) accumulate over images with atomics (zero-initialised)
__global__ void __launch_bounds__(256) box_attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                const float* __restrict__ v, const float* __restrict__ bbox,
                                                                const long long* __restrict__ y, const float* __restrict__ p_save,
                                                                const float* __restrict__ glin_save, const float* __restrict__ dout,
                                                                int O, int D, float* __restrict__ dq, float* __restrict__ dk,
                                                                float* __restrict__ dv, float* __restrict__ dwg,
                                                                float* __restrict__ dbg) {
  __shared__ float s_p[kMaxO * kMaxO];
  __shared__ float s_ds[kMaxO * kMaxO];   // d(score + log geo)
  __shared__ float s_bb[kMaxO * 4];
  __shared__ float s_dwg[65];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < O * 4; i += blockDim.x) s_bb[i] = __ldg(bbox + 1LL * b * O * 4 + i);
  for (int i = threadIdx.x; i < O * O; i += blockDim.x) s_p[i] = __ldg(p_save + 1LL * b * O * O + i);
  for (int i = threadIdx.x; i < 65; i += blockDim.x) s_dwg[i] = 0.f;
  __syncthreads();
  // dp_ij = dout_i . v_j ; ds_ij = p_ij (dp_ij - sum_j p_ij dp_ij)
  for (int i = warp; i < O; i += nwarp) {
    const float* go = dout + (1LL * b * O + i) * D;
    for (int j = 0; j < O; ++j) {
      const float* vj = v + (1LL * b * O + j) * D;
      float acc = 0.f;
      for (int d = lane; d < D; d += 32) acc = fmaf(__ldg(go + d), __ldg(vj + d), acc);
      acc = warp_sum(acc);
      if (lane == 0) s_ds[i * O + j] = acc;
    }
    __syncwarp();
    float dot = 0.f;
    for (int j = lane; j < O; j += 32) dot += s_p[i * O + j] * s_ds[i * O + j];
    dot = warp_sum(dot);
    for (int j = lane; j < O; j += 32) s_ds[i * O + j] = s_p[i * O + j] * (s_ds[i * O + j] - dot);
  }
  __syncthreads();
  const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(D));
  // dq_i = sum_j ds_ij k_j / sqrt(D) (unmasked keys) ; dk_j = sum_i ds_ij q_i / sqrt(D) ; dv_j = sum_i p_ij dout_i
  for (int r = warp; r < O; r += nwarp) {
    const bool key_ok = __ldg(y + 1LL * b * O + r) != 0;
    for (int d = lane; d < D; d += 32) {
      float aq = 0.f, ak = 0.f, av = 0.f;
      for (int t = 0; t < O; ++t) {
        if (__ldg(y + 1LL * b * O + t) != 0) aq = fmaf(s_ds[r * O + t], __ldg(k + (1LL * b * O + t) * D + d), aq);
        if (key_ok) ak = fmaf(s_ds[t * O + r], __ldg(q + (1LL * b * O + t) * D + d), ak);
        av = fmaf(s_p[t * O + r], __ldg(dout + (1LL * b * O + t) * D + d), av);
      }
      dq[(1LL * b * O + r) * D + d] = aq * inv_sqrt_d;
      dk[(1LL * b * O + r) * D + d] = ak * inv_sqrt_d;
      dv[(1LL * b * O + r) * D + d] = av;
    }
  }
  // geometry gate: d glin = ds * [geo >= 1e-6]/max(geo,1e-6) * [glin > 0]
  for (int ij = threadIdx.x; ij < O * O; ij += blockDim.x) {
    const float gl = __ldg(glin_save + 1LL * b * O * O + ij);
    const float geo = fmaxf(gl, 0.f);
    if (gl > 0.f && geo >= 1e-6f) {
      const float dg = s_ds[ij] / geo;
      float emb[64];
      rel_embedding(s_bb, ij / O, ij % O, emb);
#pragma unroll
      for (int t = 0; t < 64; ++t) atomicAdd(&s_dwg[t], dg * emb[t]);
      atomicAdd(&s_dwg[64], dg);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 64; t += blockDim.x) atomicAdd(dwg + t, s_dwg[t]);
  if (threadIdx.x == 0) atomicAdd(dbg, s_dwg[64]);
}

int box_attention_fwd(const float* q, const float* k, const float* v, const float* bbox, const long long* y,
                      const float* wg, const float* bg, int B, int O, int D, float* out, float* p_save, float* glin_save,
                      cudaStream_t stream) {
  if (!q || !k || !v || !bbox || !y || !wg || !bg || !out || !p_save || !glin_save || B <= 0 || O <= 0 || D <= 0) { set_error("box_attention_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  if (O > kMaxO) { set_error("box_attention: at most %d objects per image supported (got %d)", kMaxO, O); return L2I_ERR_UNSUPPORTED; }
  box_attention_fwd_kernel<<<B, 256, 0, stream>>>(q, k, v, bbox, y, wg, bg, O, D, out, p_save, glin_save);
  return check_launch("box_attention_fwd_kernel");
}

int box_attention_bwd(const float* q, const float* k, const float* v, const float* bbox, const long long* y,
                      const float* p_save, const float* glin_save, const float* dout, int B, int O, int D, float* dq,
                      float* dk, float* dv, float* dwg, float* dbg, cudaStream_t stream) {
  if (!q || !k || !v || !bbox || !y || !p_save || !glin_save || !dout || !dq || !dk || !dv || !dwg || !dbg || B <= 0 || O <= 0 || D <= 0) { set_error("box_attention_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  if (O > kMaxO) { set_error("box_attention: at most %d objects per image supported (got %d)", kMaxO, O); return L2I_ERR_UNSUPPORTED; }
  cudaError_t e = cudaMemsetAsync(dwg, 0, sizeof(float) * 64, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbg, 0, sizeof(float), stream);
  if (e != cudaSuccess) { set_error("box_attention_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  box_attention_bwd_kernel<<<B, 256, 0, stream>>>(q, k, v, bbox, y, p_save, glin_save, dout, O, D, dq, dk, dv, dwg, dbg);
  return check_launch("box_attention_bwd_kernel");
}

}  // namespace l2i
