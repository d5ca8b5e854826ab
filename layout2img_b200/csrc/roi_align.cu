// ROIAlign forward / backward on NHWC features -- the operator the reference's setup.py:46-54
// extension seam (`model.roi_layers._C`) was meant to provide and that the code now takes from
// torchvision.ops.RoIAlign((8,8), scale, sampling_ratio=0), aligned=False
// (model/rcnn_discriminator_app.py:98-99,139,143).  Semantics per SURVEY.md Appendix C.
// Gather-bound: thread <-> (roi, bin, 4 consecutive channels), float4 taps, float4 atomics in bwd.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

struct RoiGeom {
  int n;
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float inv_count;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int P) {
  RoiGeom g;
  g.n = static_cast<int>(__ldg(roi));
  g.start_w = __ldg(roi + 1) * scale;
  g.start_h = __ldg(roi + 2) * scale;
  const float end_w = __ldg(roi + 3) * scale, end_h = __ldg(roi + 4) * scale;
  const float rw = fmaxf(end_w - g.start_w, 1.0f), rh = fmaxf(end_h - g.start_h, 1.0f);
  g.bin_w = rw / static_cast<float>(P);
  g.bin_h = rh / static_cast<float>(P);
  g.grid_w = static_cast<int>(ceilf(rw / static_cast<float>(P)));
  g.grid_h = static_cast<int>(ceilf(rh / static_cast<float>(P)));
  g.inv_count = 1.0f / fmaxf(static_cast<float>(g.grid_w * g.grid_h), 1.0f);
  return g;
}

struct Taps { int y0, y1, x0, x1; float w1, w2, w3, w4; bool valid; };
__device__ __forceinline__ Taps bilinear_taps(float y, float x, int H, int W) {
  Taps t;
  t.valid = !(y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W));
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  t.y0 = static_cast<int>(y);
  t.x0 = static_cast<int>(x);
  if (t.y0 >= H - 1) { t.y1 = t.y0 = H - 1; y = static_cast<float>(t.y0); } else { t.y1 = t.y0 + 1; }
  if (t.x0 >= W - 1) { t.x1 = t.x0 = W - 1; x = static_cast<float>(t.x0); } else { t.x1 = t.x0 + 1; }
  const float ly = y - t.y0, lx = x - t.x0, hy = 1.0f - ly, hx = 1.0f - lx;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
  return t;
}

__global__ void roi_align_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois, int K, int H, int W,
                                     int C, int P, float scale, float* __restrict__ out) {
  const int c4n = C >> 2;
  const long long total = 1LL * K * P * P * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int pw = static_cast<int>((i / c4n) % P), ph = static_cast<int>((i / (1LL * c4n * P)) % P);
    const int k = static_cast<int>(i / (1LL * c4n * P * P));
    const RoiGeom g = roi_geom(rois + k * 5, scale, P);
    const float* f = feat + 1LL * g.n * H * W * C + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
        const Taps t = bilinear_taps(y, x, H, W);
        if (!t.valid) continue;
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y0 * W + t.x0) * C));
        const float4 v2 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y0 * W + t.x1) * C));
        const float4 v3 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y1 * W + t.x0) * C));
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y1 * W + t.x1) * C));
        acc.x += t.w1 * v1.x + t.w2 * v2.x + t.w3 * v3.x + t.w4 * v4.x;
        acc.y += t.w1 * v1.y + t.w2 * v2.y + t.w3 * v3.y + t.w4 * v4.y;
        acc.z += t.w1 * v1.z + t.w2 * v2.z + t.w3 * v3.z + t.w4 * v4.z;
        acc.w += t.w1 * v1.w + t.w2 * v2.w + t.w3 * v3.w + t.w4 * v4.w;
      }
    }
    acc.x *= g.inv_count; acc.y *= g.inv_count; acc.z *= g.inv_count; acc.w *= g.inv_count;
    *reinterpret_cast<float4*>(out + ((1LL * k * P + ph) * P + pw) * C + c) = acc;
  }
}

__device__ __forceinline__ void atomic_add4(float* p, float w, const float4& g) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
}

__global__ void roi_align_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ rois, int K, int H, int W,
                                     int C, int P, float scale, float* __restrict__ dfeat) {
  const int c4n = C >> 2;
  const long long total = 1LL * K * P * P * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int pw = static_cast<int>((i / c4n) % P), ph = static_cast<int>((i / (1LL * c4n * P)) % P);
    const int k = static_cast<int>(i / (1LL * c4n * P * P));
    const RoiGeom g = roi_geom(rois + k * 5, scale, P);
    float* f = dfeat + 1LL * g.n * H * W * C + c;
    float4 go = __ldg(reinterpret_cast<const float4*>(dout + ((1LL * k * P + ph) * P + pw) * C + c));
    go.x *= g.inv_count; go.y *= g.inv_count; go.z *= g.inv_count; go.w *= g.inv_count;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
        const Taps t = bilinear_taps(y, x, H, W);
        if (!t.valid) continue;
        atomic_add4(f + (1LL * t.y0 * W + t.x0) * C, t.w1, go);
        atomic_add4(f + (1LL * t.y0 * W + t.x1) * C, t.w2, go);
        atomic_add4(f + (1LL * t.y1 * W + t.x0) * C, t.w3, go);
        atomic_add4(f + (1LL * t.y1 * W + t.x1) * C, t.w4, go);
      }
    }
  }
}

int roi_align_fwd(const float* feat, const float* rois, int K, int N, int H, int W, int C, int P, float scale, float* out,
                  cudaStream_t stream) {
  if (K == 0) return L2I_OK;
  if (!feat || !rois || !out || K < 0 || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || P <= 0) { set_error("roi_align_fwd: bad arguments (C must be a multiple of 4)"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * K * P * P * (C >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_align_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(feat, rois, K, H, W, C, P, scale, out);
  return check_launch("roi_align_fwd_kernel");
}

int roi_align_bwd(const float* dout, const float* rois, int K, int N, int H, int W, int C, int P, float scale, float* dfeat,
                  cudaStream_t stream) {
  if (!dfeat || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || P <= 0 || K < 0) { set_error("roi_align_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(dfeat, 0, sizeof(float) * N * H * W * C, stream);
  if (e != cudaSuccess) { set_error("roi_align_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  if (K == 0) return L2I_OK;
  if (!dout || !rois) { set_error("roi_align_bwd: null pointer"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * K * P * P * (C >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_align_bwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dout, rois, K, H, W, C, P, scale, dfeat);
  return check_launch("roi_align_bwd_kernel");
}

// ------------------------------------------------------------------------------------------
// Device-side ROI preparation (reference rcnn_discriminator_app.py:402-417 and :131-146) -- no host round trip:
//   xywh in [0,1] -> (image, x0, y0, x1, y1) in pixels; rows with label == 0 dropped; the survivors ordered as
//   [all "large" ROIs, all "small" ROIs] (small: width < 64 AND height < 64), each group in (b, o) row-major order.
// Writes all B*O rows: the valid ones first in that order, the dropped ones last (level 2).  counts = (n_large, n_small).
// One block; a stable three-way partition by a block-wide exclusive scan.  The arithmetic is the reference's, rounding
// by rounding: x1 = (w + x0) * S, y1 = (h + y0) * S, small = (x1 - x0 < T) * (y1 - y0 < T) in fp32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) roi_prepare_kernel(const float* __restrict__ bbox, const long long* __restrict__ label,
                                                           int B, int O, float img, float thresh, float* __restrict__ rois,
                                                           long long* __restrict__ y_sorted, int* __restrict__ level,
                                                           int* __restrict__ perm, int* __restrict__ counts) {
  __shared__ int s_cnt[3][1024];
  __shared__ int s_tot[3];
  const int K = B * O;
  const int per = (K + blockDim.x - 1) / blockDim.x;         // consecutive rows per thread: keeps the (b, o) order
  const int r0 = threadIdx.x * per, r1 = min(K, r0 + per);
  int c[3] = {0, 0, 0};
  for (int r = r0; r < r1; ++r) {
    const float x0 = __fmul_rn(__ldg(bbox + 4 * r), img), y0 = __fmul_rn(__ldg(bbox + 4 * r + 1), img);
    const float x1 = __fmul_rn(__fadd_rn(__ldg(bbox + 4 * r + 2), __ldg(bbox + 4 * r)), img);
    const float y1 = __fmul_rn(__fadd_rn(__ldg(bbox + 4 * r + 3), __ldg(bbox + 4 * r + 1)), img);
    const bool small = (__fsub_rn(x1, x0) < thresh) && (__fsub_rn(y1, y0) < thresh);
    const int lv = (label[r] == 0) ? 2 : (small ? 1 : 0);
    ++c[lv];
  }
  for (int j = 0; j < 3; ++j) s_cnt[j][threadIdx.x] = c[j];
  __syncthreads();
  // exclusive scan over the threads (Hillis-Steele on the three counters)
  for (int off = 1; off < static_cast<int>(blockDim.x); off <<= 1) {
    int v[3];
    for (int j = 0; j < 3; ++j) v[j] = (static_cast<int>(threadIdx.x) >= off) ? s_cnt[j][threadIdx.x - off] : 0;
    __syncthreads();
    for (int j = 0; j < 3; ++j) s_cnt[j][threadIdx.x] += v[j];
    __syncthreads();
  }
  if (threadIdx.x == blockDim.x - 1) {
    for (int j = 0; j < 3; ++j) s_tot[j] = s_cnt[j][threadIdx.x];
    counts[0] = s_cnt[0][threadIdx.x];
    counts[1] = s_cnt[1][threadIdx.x];
  }
  __syncthreads();
  int pos[3];
  pos[0] = s_cnt[0][threadIdx.x] - c[0];
  pos[1] = s_tot[0] + s_cnt[1][threadIdx.x] - c[1];
  pos[2] = s_tot[0] + s_tot[1] + s_cnt[2][threadIdx.x] - c[2];
  for (int r = r0; r < r1; ++r) {
    const float x0 = __fmul_rn(__ldg(bbox + 4 * r), img), y0 = __fmul_rn(__ldg(bbox + 4 * r + 1), img);
    const float x1 = __fmul_rn(__fadd_rn(__ldg(bbox + 4 * r + 2), __ldg(bbox + 4 * r)), img);
    const float y1 = __fmul_rn(__fadd_rn(__ldg(bbox + 4 * r + 3), __ldg(bbox + 4 * r + 1)), img);
    const bool small = (__fsub_rn(x1, x0) < thresh) && (__fsub_rn(y1, y0) < thresh);
    const long long lab = label[r];
    const int lv = (lab == 0) ? 2 : (small ? 1 : 0);
    const int d = pos[lv]++;
    rois[5 * d] = static_cast<float>(r / O);
    rois[5 * d + 1] = x0; rois[5 * d + 2] = y0; rois[5 * d + 3] = x1; rois[5 * d + 4] = y1;
    y_sorted[d] = lab;
    level[d] = lv;
    perm[d] = r;
  }
}

int roi_prepare(const float* bbox, const long long* label, int B, int O, float img, float thresh, float* rois,
                long long* y_sorted, int* level, int* perm, int* counts, cudaStream_t stream) {
  if (!bbox || !label || !rois || !y_sorted || !level || !perm || !counts || B <= 0 || O <= 0) { set_error("roi_prepare: bad arguments"); return L2I_ERR_BAD_ARG; }
  roi_prepare_kernel<<<1, 1024, 0, stream>>>(bbox, label, B, O, img, thresh, rois, y_sorted, level, perm, counts);
  return check_launch("roi_prepare_kernel");
}

// Two-level ROIAlign: row k takes feat_l at scale_l (level 0), feat_s at scale_s (level 1) or is zero-filled (level 2:
// a dropped object in the fixed-size, graph-capturable form).  One launch writes the [large ..., small ...] stack that
// the reference builds with two RoIAlign calls and a concatenation (rcnn_discriminator_app.py:137-146).
__global__ void roi_align2_fwd_kernel(const float* __restrict__ feat_l, int Hl, int Wl, float scale_l,
                                      const float* __restrict__ feat_s, int Hs, int Ws, float scale_s,
                                      const float* __restrict__ rois, const int* __restrict__ level, int K, int C, int P,
                                      float* __restrict__ out) {
  const int c4n = C >> 2;
  const long long total = 1LL * K * P * P * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int pw = static_cast<int>((i / c4n) % P), ph = static_cast<int>((i / (1LL * c4n * P)) % P);
    const int k = static_cast<int>(i / (1LL * c4n * P * P));
    const int lv = __ldg(level + k);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lv < 2) {
      const int H = lv ? Hs : Hl, W = lv ? Ws : Wl;
      const RoiGeom g = roi_geom(rois + k * 5, lv ? scale_s : scale_l, P);
      const float* f = (lv ? feat_s : feat_l) + 1LL * g.n * H * W * C + c;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
          const Taps t = bilinear_taps(y, x, H, W);
          if (!t.valid) continue;
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y0 * W + t.x0) * C));
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y0 * W + t.x1) * C));
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y1 * W + t.x0) * C));
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(f + (1LL * t.y1 * W + t.x1) * C));
          acc.x += t.w1 * v1.x + t.w2 * v2.x + t.w3 * v3.x + t.w4 * v4.x;
          acc.y += t.w1 * v1.y + t.w2 * v2.y + t.w3 * v3.y + t.w4 * v4.y;
          acc.z += t.w1 * v1.z + t.w2 * v2.z + t.w3 * v3.z + t.w4 * v4.z;
          acc.w += t.w1 * v1.w + t.w2 * v2.w + t.w3 * v3.w + t.w4 * v4.w;
        }
      }
      acc.x *= g.inv_count; acc.y *= g.inv_count; acc.z *= g.inv_count; acc.w *= g.inv_count;
    }
    *reinterpret_cast<float4*>(out + ((1LL * k * P + ph) * P + pw) * C + c) = acc;
  }
}

__global__ void roi_align2_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ rois,
                                      const int* __restrict__ level, int K, int C, int P, int Hl, int Wl, float scale_l,
                                      float* __restrict__ dfeat_l, int Hs, int Ws, float scale_s, float* __restrict__ dfeat_s) {
  const int c4n = C >> 2;
  const long long total = 1LL * K * P * P * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int pw = static_cast<int>((i / c4n) % P), ph = static_cast<int>((i / (1LL * c4n * P)) % P);
    const int k = static_cast<int>(i / (1LL * c4n * P * P));
    const int lv = __ldg(level + k);
    if (lv >= 2) continue;
    const int H = lv ? Hs : Hl, W = lv ? Ws : Wl;
    const RoiGeom g = roi_geom(rois + k * 5, lv ? scale_s : scale_l, P);
    float* f = (lv ? dfeat_s : dfeat_l) + 1LL * g.n * H * W * C + c;
    float4 go = __ldg(reinterpret_cast<const float4*>(dout + ((1LL * k * P + ph) * P + pw) * C + c));
    go.x *= g.inv_count; go.y *= g.inv_count; go.z *= g.inv_count; go.w *= g.inv_count;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
        const Taps t = bilinear_taps(y, x, H, W);
        if (!t.valid) continue;
        atomic_add4(f + (1LL * t.y0 * W + t.x0) * C, t.w1, go);
        atomic_add4(f + (1LL * t.y0 * W + t.x1) * C, t.w2, go);
        atomic_add4(f + (1LL * t.y1 * W + t.x0) * C, t.w3, go);
        atomic_add4(f + (1LL * t.y1 * W + t.x1) * C, t.w4, go);
      }
    }
  }
}

int roi_align2_fwd(const float* feat_l, int Hl, int Wl, float scale_l, const float* feat_s, int Hs, int Ws, float scale_s,
                   const float* rois, const int* level, int K, int N, int C, int P, float* out, cudaStream_t stream) {
  if (K == 0) return L2I_OK;
  if (!feat_l || !feat_s || !rois || !level || !out || K < 0 || N <= 0 || C <= 0 || (C & 3) || P <= 0 || Hl <= 0 || Wl <= 0 ||
      Hs <= 0 || Ws <= 0) { set_error("roi_align2_fwd: bad arguments (C must be a multiple of 4)"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * K * P * P * (C >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_align2_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(feat_l, Hl, Wl, scale_l, feat_s, Hs, Ws, scale_s, rois, level,
                                                                     K, C, P, out);
  return check_launch("roi_align2_fwd_kernel");
}

int roi_align2_bwd(const float* dout, const float* rois, const int* level, int K, int N, int C, int P, int Hl, int Wl,
                   float scale_l, float* dfeat_l, int Hs, int Ws, float scale_s, float* dfeat_s, cudaStream_t stream) {
  if (!dfeat_l || !dfeat_s || N <= 0 || C <= 0 || (C & 3) || P <= 0 || K < 0) { set_error("roi_align2_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(dfeat_l, 0, sizeof(float) * N * Hl * Wl * C, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dfeat_s, 0, sizeof(float) * N * Hs * Ws * C, stream);
  if (e != cudaSuccess) { set_error("roi_align2_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  if (K == 0) return L2I_OK;
  if (!dout || !rois || !level) { set_error("roi_align2_bwd: null pointer"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * K * P * P * (C >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  roi_align2_bwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dout, rois, level, K, C, P, Hl, Wl, scale_l, dfeat_l, Hs, Ws,
                                                                     scale_s, dfeat_s);
  return check_launch("roi_align2_bwd_kernel");
}

// ------------------------------------------------------------------------------------------
// 2x2 average pooling, NHWC (F.avg_pool2d(x, 2), rcnn_discriminator_app.py:304-312,333-344)
// ------------------------------------------------------------------------------------------
__global__ void avgpool2_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ out) {
  const int c4n = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = 1LL * N * Ho * Wo * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int wo = static_cast<int>((i / c4n) % Wo), ho = static_cast<int>((i / (1LL * c4n * Wo)) % Ho);
    const int n = static_cast<int>(i / (1LL * c4n * Wo * Ho));
    const float* p = x + ((1LL * n * H + 2 * ho) * W + 2 * wo) * C + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + C));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + 1LL * W * C)), e = __ldg(reinterpret_cast<const float4*>(p + 1LL * W * C + C));
    *reinterpret_cast<float4*>(out + ((1LL * n * Ho + ho) * Wo + wo) * C + c) =
        make_float4((a.x + b.x + d.x + e.x) * 0.25f, (a.y + b.y + d.y + e.y) * 0.25f, (a.z + b.z + d.z + e.z) * 0.25f,
                    (a.w + b.w + d.w + e.w) * 0.25f);
  }
}
// dx[n,h,w,c] = 0.25 * dout[n,h/2,w/2,c]
__global__ void avgpool2_bwd_kernel(const float* __restrict__ dout, int N, int H, int W, int C, float* __restrict__ dx) {
  const int c4n = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = 1LL * N * H * W * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int w = static_cast<int>((i / c4n) % W), h = static_cast<int>((i / (1LL * c4n * W)) % H);
    const int n = static_cast<int>(i / (1LL * c4n * W * H));
    const float4 g = __ldg(reinterpret_cast<const float4*>(dout + ((1LL * n * Ho + (h >> 1)) * Wo + (w >> 1)) * C + c));
    *reinterpret_cast<float4*>(dx + ((1LL * n * H + h) * W + w) * C + c) = make_float4(g.x * 0.25f, g.y * 0.25f, g.z * 0.25f, g.w * 0.25f);
  }
}
// 2x2 / stride 2 max pooling, NHWC (torchvision VGG19 features 4, 9, 18, 27 inside the perceptual loss, reference
// utils/util.py:49-94).  Backward routes the gradient to the FIRST maximum of the window in (row, column) order, as
// torch's max_pool2d does.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, int N, int H, int W, int C, float* __restrict__ out) {
  const int c4n = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const long long total = 1LL * N * Ho * Wo * c4n;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % c4n) * 4;
    const int wo = static_cast<int>((i / c4n) % Wo), ho = static_cast<int>((i / (1LL * c4n * Wo)) % Ho);
    const int n = static_cast<int>(i / (1LL * c4n * Wo * Ho));
    const float* p = x + ((1LL * n * H + 2 * ho) * W + 2 * wo) * C + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + C));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + 1LL * W * C)), e = __ldg(reinterpret_cast<const float4*>(p + 1LL * W * C + C));
    *reinterpret_cast<float4*>(out + ((1LL * n * Ho + ho) * Wo + wo) * C + c) =
        make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                    fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
  }
}
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, int N, int H, int W, int C,
                                    float* __restrict__ dx) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = 1LL * N * Ho * Wo * C;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int wo = static_cast<int>((i / C) % Wo), ho = static_cast<int>((i / (1LL * C * Wo)) % Ho);
    const int n = static_cast<int>(i / (1LL * C * Wo * Ho));
    const long long base = ((1LL * n * H + 2 * ho) * W + 2 * wo) * C + c;
    const long long off[4] = {0, C, 1LL * W * C, 1LL * W * C + C};
    float best = __ldg(x + base);
    int arg = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = __ldg(x + base + off[k]);
      if (v > best) { best = v; arg = k; }
    }
    const float g = __ldg(dout + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[base + off[k]] = (k == arg) ? g : 0.f;
  }
}
int maxpool2_fwd(const float* x, int N, int H, int W, int C, float* out, cudaStream_t stream) {
  if (!x || !out || N <= 0 || (H & 1) || (W & 1) || (C & 3) || C <= 0) { set_error("maxpool2_fwd: need even H, W and C %% 4 == 0"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * N * (H / 2) * (W / 2) * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  maxpool2_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, N, H, W, C, out);
  return check_launch("maxpool2_fwd_kernel");
}
int maxpool2_bwd(const float* x, const float* dout, int N, int H, int W, int C, float* dx, cudaStream_t stream) {
  if (!x || !dout || !dx || N <= 0 || (H & 1) || (W & 1) || C <= 0) { set_error("maxpool2_bwd: need even H, W"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * N * (H / 2) * (W / 2) * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  maxpool2_bwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, dout, N, H, W, C, dx);
  return check_launch("maxpool2_bwd_kernel");
}

int avgpool2_fwd(const float* x, int N, int H, int W, int C, float* out, cudaStream_t stream) {
  if (!x || !out || N <= 0 || (H & 1) || (W & 1) || (C & 3) || C <= 0) { set_error("avgpool2_fwd: need even H, W and C %% 4 == 0"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * N * (H / 2) * (W / 2) * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  avgpool2_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, N, H, W, C, out);
  return check_launch("avgpool2_fwd_kernel");
}
int avgpool2_bwd(const float* dout, int N, int H, int W, int C, float* dx, cudaStream_t stream) {
  if (!dout || !dx || N <= 0 || (H & 1) || (W & 1) || (C & 3) || C <= 0) { set_error("avgpool2_bwd: need even H, W and C %% 4 == 0"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * N * H * W * (C / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  avgpool2_bwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dout, N, H, W, C, dx);
  return check_launch("avgpool2_bwd_kernel");
}

}  // namespace l2i
