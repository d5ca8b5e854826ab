// Layout-map kernels of the generator (HBM-trivial, index-exact):
//   bbox_mask         reference model/resnet_generator_app_v2.py:697-721   (bit-exact {0,1} map)
//   masks_to_layout   reference utils/bilinear.py:137-192                  (grid_sample paste, fwd + bwd)
//   mask_resize       F.interpolate(mode='bilinear', align_corners=False)   norm_module.py:176,
//                     resnet_generator_app_v2.py:470 (fwd + deterministic gather-form bwd)
//   stage_mask_mix    reference model/resnet_generator_app_v2.py:466-470    (fwd + bwd)
// Index rules follow SURVEY.md Appendix C.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

// torch.linspace(0, 1, n)[i] in fp32: step = fl(1/(n-1)); fl(step*i) below the midpoint,
// fl(1 - step*(n-1-i)) (single rounding) above it.
__device__ __forceinline__ float linspace01(int i, int n) {
  const float step = __fdiv_rn(1.0f, static_cast<float>(n - 1));
  return (i < n / 2) ? __fmul_rn(step, static_cast<float>(i)) : __fmaf_rn(-step, static_cast<float>(n - 1 - i), 1.0f);
}

// ------------------------------------------------------------------------------------------ bbox_mask
__global__ void bbox_mask_kernel(const float* __restrict__ bbox, int BO, int H, int W, float* __restrict__ out) {
  const long long total = 1LL * BO * H * W;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int bo = static_cast<int>(i / (1LL * W * H));
    const float x0 = __ldg(bbox + bo * 4), y0 = __ldg(bbox + bo * 4 + 1);
    const float ww = __ldg(bbox + bo * 4 + 2), hh = __ldg(bbox + bo * 4 + 3);
    const float X = __fdiv_rn(__fsub_rn(linspace01(x, W), x0), ww);
    const float Y = __fdiv_rn(__fsub_rn(linspace01(y, H), y0), hh);
    const bool outside = (X < 0.f) | (X > 1.f) | (Y < 0.f) | (Y > 1.f);
    out[i] = outside ? 0.f : 1.f;
  }
}

int bbox_mask(const float* bbox, int BO, int H, int W, float* out, cudaStream_t stream) {
  if (!bbox || !out || BO <= 0 || H < 2 || W < 2) { set_error("bbox_mask: bad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * BO * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bbox_mask_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(bbox, BO, H, W, out);
  return check_launch("bbox_mask_kernel");
}

// ------------------------------------------------------------------------------------------ masks_to_layout
struct GridTaps {
  int x0, y0;
  float nw, ne, sw, se;
};
__device__ __forceinline__ GridTaps grid_taps(const float* __restrict__ box, int x, int y, int S, int M) {
  const float bx = __ldg(box), by = __ldg(box + 1), bw = __ldg(box + 2), bh = __ldg(box + 3);
  const float X = __fdiv_rn(__fsub_rn(linspace01(x, S), bx), bw);
  const float Y = __fdiv_rn(__fsub_rn(linspace01(y, S), by), bh);
  const float gx = __fsub_rn(__fmul_rn(X, 2.0f), 1.0f), gy = __fsub_rn(__fmul_rn(Y, 2.0f), 1.0f);
  // grid_sample unnormalise, align_corners=False: ((g + 1) * size - 1) / 2
  const float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(M)), 1.0f), 2.0f);
  const float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(M)), 1.0f), 2.0f);
  const float fx = floorf(ix), fy = floorf(iy);
  GridTaps t;
  // keep the integer conversion safe for far-away (padding) boxes
  t.x0 = (fx < -2.f) ? -2 : (fx > M + 1.f ? M + 1 : static_cast<int>(fx));
  t.y0 = (fy < -2.f) ? -2 : (fy > M + 1.f ? M + 1 : static_cast<int>(fy));
  const float ex = fx + 1.0f, ey = fy + 1.0f;
  t.nw = (ex - ix) * (ey - iy);
  t.ne = (ix - fx) * (ey - iy);
  t.sw = (ex - ix) * (iy - fy);
  t.se = (ix - fx) * (iy - fy);
  return t;
}

__global__ void masks_to_layout_fwd_kernel(const float* __restrict__ bbox, const float* __restrict__ masks, int BO, int M,
                                           int S, float* __restrict__ out) {
  const long long total = 1LL * BO * S * S;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % S);
    const int y = static_cast<int>((i / S) % S);
    const int bo = static_cast<int>(i / (1LL * S * S));
    const GridTaps t = grid_taps(bbox + bo * 4, x, y, S, M);
    const float* m = masks + 1LL * bo * M * M;
    float v = 0.f;
    const bool xin0 = t.x0 >= 0 && t.x0 < M, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < M;
    const bool yin0 = t.y0 >= 0 && t.y0 < M, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < M;
    if (yin0 && xin0) v += __ldg(m + t.y0 * M + t.x0) * t.nw;
    if (yin0 && xin1) v += __ldg(m + t.y0 * M + t.x0 + 1) * t.ne;
    if (yin1 && xin0) v += __ldg(m + (t.y0 + 1) * M + t.x0) * t.sw;
    if (yin1 && xin1) v += __ldg(m + (t.y0 + 1) * M + t.x0 + 1) * t.se;
    out[i] = v;
  }
}

// one block per (b,o): scatter into a shared M x M tile, then store
__global__ void masks_to_layout_bwd_kernel(const float* __restrict__ bbox, const float* __restrict__ dout, int M, int S,
                                           float* __restrict__ dmasks) {
  extern __shared__ float tile[];
  const int bo = blockIdx.x;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int x = i % S, y = i / S;
    const GridTaps t = grid_taps(bbox + bo * 4, x, y, S, M);
    const float g = __ldg(dout + 1LL * bo * S * S + i);
    const bool xin0 = t.x0 >= 0 && t.x0 < M, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < M;
    const bool yin0 = t.y0 >= 0 && t.y0 < M, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < M;
    if (yin0 && xin0) atomicAdd(tile + t.y0 * M + t.x0, g * t.nw);
    if (yin0 && xin1) atomicAdd(tile + t.y0 * M + t.x0 + 1, g * t.ne);
    if (yin1 && xin0) atomicAdd(tile + (t.y0 + 1) * M + t.x0, g * t.sw);
    if (yin1 && xin1) atomicAdd(tile + (t.y0 + 1) * M + t.x0 + 1, g * t.se);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) dmasks[1LL * bo * M * M + i] = tile[i];
}

int masks_to_layout_fwd(const float* bbox, const float* masks, int BO, int M, int S, float* out, cudaStream_t stream) {
  if (!bbox || !masks || !out || BO <= 0 || M <= 0 || S < 2) { set_error("masks_to_layout_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * BO * S * S;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  masks_to_layout_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(bbox, masks, BO, M, S, out);
  return check_launch("masks_to_layout_fwd_kernel");
}
int masks_to_layout_bwd(const float* bbox, const float* dout, int BO, int M, int S, float* dmasks, cudaStream_t stream) {
  if (!bbox || !dout || !dmasks || BO <= 0 || M <= 0 || M > 64 || S < 2) { set_error("masks_to_layout_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  masks_to_layout_bwd_kernel<<<BO, 256, sizeof(float) * M * M, stream>>>(bbox, dout, M, S, dmasks);
  return check_launch("masks_to_layout_bwd_kernel");
}

// ------------------------------------------------------------------------------------------ bilinear resize
// torch upsample_bilinear2d, align_corners=False: src = max(scale*(dst+0.5)-0.5, 0), scale = in/out
struct Lin1 { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lin1 lin_src(int dst, int in, int out) {
  const float scale = static_cast<float>(in) / static_cast<float>(out);
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Lin1 r;
  r.i0 = static_cast<int>(src);
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + ((r.i0 < in - 1) ? 1 : 0);
  r.l1 = src - static_cast<float>(r.i0);
  r.l0 = 1.0f - r.l1;
  return r;
}

// in (B,O,hi,wi) -> out (B,O,h,w) [pixel_major = 0] or (B,h,w,O) [pixel_major = 1]
__global__ void mask_resize_fwd_kernel(const float* __restrict__ in, int B, int O, int hi, int wi, int h, int w,
                                       int pixel_major, float* __restrict__ out) {
  const long long total = 1LL * B * O * h * w;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int b, o, y, x;
    if (pixel_major) {
      o = static_cast<int>(i % O); x = static_cast<int>((i / O) % w); y = static_cast<int>((i / (1LL * O * w)) % h);
      b = static_cast<int>(i / (1LL * O * w * h));
    } else {
      x = static_cast<int>(i % w); y = static_cast<int>((i / w) % h); o = static_cast<int>((i / (1LL * w * h)) % O);
      b = static_cast<int>(i / (1LL * w * h * O));
    }
    const float* src = in + (1LL * b * O + o) * hi * wi;
    float v;
    if (hi == h && wi == w) {
      v = __ldg(src + y * wi + x);
    } else {
      const Lin1 ly = lin_src(y, hi, h), lx = lin_src(x, wi, w);
      v = ly.l0 * (lx.l0 * __ldg(src + ly.i0 * wi + lx.i0) + lx.l1 * __ldg(src + ly.i0 * wi + lx.i1)) +
          ly.l1 * (lx.l0 * __ldg(src + ly.i1 * wi + lx.i0) + lx.l1 * __ldg(src + ly.i1 * wi + lx.i1));
    }
    out[i] = v;
  }
}

// weight with which output index j reads input index i along one axis (0 if it does not)
__device__ __forceinline__ float lin_weight(int j, int i, int in, int out) {
  const Lin1 l = lin_src(j, in, out);
  float wgt = 0.f;
  if (l.i0 == i) wgt += l.l0;
  if (l.i1 == i) wgt += l.l1;
  return wgt;
}

// gather-form backward (fixed summation order): thread <-> one input pixel (b,o,yi,xi)
__global__ void mask_resize_bwd_kernel(const float* __restrict__ dout, int B, int O, int hi, int wi, int h, int w,
                                       int pixel_major, float* __restrict__ din) {
  const long long total = 1LL * B * O * hi * wi;
  const float ry = static_cast<float>(h) / static_cast<float>(hi), rx = static_cast<float>(w) / static_cast<float>(wi);
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int xi = static_cast<int>(i % wi), yi = static_cast<int>((i / wi) % hi);
    const int o = static_cast<int>((i / (1LL * wi * hi)) % O), b = static_cast<int>(i / (1LL * wi * hi * O));
    float acc = 0.f;
    if (hi == h && wi == w) {
      acc = pixel_major ? __ldg(dout + ((1LL * b * h + yi) * w + xi) * O + o) : __ldg(dout + ((1LL * b * O + o) * h + yi) * w + xi);
    } else {
      // outputs whose source coordinate lies within (i-1, i+1); edge inputs also collect the clamped range
      int jy0 = static_cast<int>(floorf((yi - 1.0f + 0.5f) * ry - 0.5f)) - 1, jy1 = static_cast<int>(ceilf((yi + 1.0f + 0.5f) * ry - 0.5f)) + 1;
      int jx0 = static_cast<int>(floorf((xi - 1.0f + 0.5f) * rx - 0.5f)) - 1, jx1 = static_cast<int>(ceilf((xi + 1.0f + 0.5f) * rx - 0.5f)) + 1;
      if (yi == 0) jy0 = 0;
      if (yi == hi - 1) jy1 = h - 1;
      if (xi == 0) jx0 = 0;
      if (xi == wi - 1) jx1 = w - 1;
      jy0 = max(jy0, 0); jy1 = min(jy1, h - 1); jx0 = max(jx0, 0); jx1 = min(jx1, w - 1);
      for (int jy = jy0; jy <= jy1; ++jy) {
        const float wy = lin_weight(jy, yi, hi, h);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int jx = jx0; jx <= jx1; ++jx) {
          const float wx = lin_weight(jx, xi, wi, w);
          if (wx == 0.f) continue;
          const float g = pixel_major ? __ldg(dout + ((1LL * b * h + jy) * w + jx) * O + o)
                                      : __ldg(dout + ((1LL * b * O + o) * h + jy) * w + jx);
          row += wx * g;
        }
        acc += wy * row;
      }
    }
    din[i] = acc;
  }
}

int mask_resize_fwd(const float* in, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* out,
                    cudaStream_t stream) {
  if (!in || !out || B <= 0 || O <= 0 || hi <= 0 || wi <= 0 || h <= 0 || w <= 0) { set_error("mask_resize_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * B * O * h * w;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mask_resize_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, B, O, hi, wi, h, w, pixel_major, out);
  return check_launch("mask_resize_fwd_kernel");
}
int mask_resize_bwd(const float* dout, int B, int O, int hi, int wi, int h, int w, int pixel_major, float* din,
                    cudaStream_t stream) {
  if (!dout || !din || B <= 0 || O <= 0 || hi <= 0 || wi <= 0 || h <= 0 || w <= 0) { set_error("mask_resize_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * B * O * hi * wi;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mask_resize_bwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dout, B, O, hi, wi, h, w, pixel_major, din);
  return check_launch("mask_resize_bwd_kernel");
}

// ------------------------------------------------------------------------------------------ stage_mask_mix
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// stage (B,h,w,NC) NHWC conv output; y (B,O); alpha (NC); bmask, hard (B,O,S,S); out (B,O,h,w)
__global__ void stage_mix_fwd_kernel(const float* __restrict__ stage, const long long* __restrict__ y,
                                     const float* __restrict__ alpha, const float* __restrict__ bmask,
                                     const float* __restrict__ hard, int B, int O, int h, int w, int NC, int S,
                                     float* __restrict__ out) {
  const long long total = 1LL * B * O * h * w;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i % w), py = static_cast<int>((i / w) % h);
    const int o = static_cast<int>((i / (1LL * w * h)) % O), b = static_cast<int>(i / (1LL * w * h * O));
    const int cls = static_cast<int>(__ldg(y + b * O + o));
    const float sel = __ldg(stage + ((1LL * b * h + py) * w + x) * NC + cls);
    // nearest: src = min(floor(dst * (S / h)), S - 1)
    const int sy = min(static_cast<int>(floorf(py * (static_cast<float>(S) / h))), S - 1);
    const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(S) / w))), S - 1);
    const float* bm = bmask + (1LL * b * O + o) * S * S;
    const float hd = __ldg(hard + (1LL * b * O + o) * S * S + sy * S + sx);
    float soft;
    if (S == h && S == w) {
      soft = __ldg(bm + py * S + x);
    } else {
      const Lin1 ly = lin_src(py, S, h), lx = lin_src(x, S, w);
      soft = ly.l0 * (lx.l0 * __ldg(bm + ly.i0 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i0 * S + lx.i1)) +
             ly.l1 * (lx.l0 * __ldg(bm + ly.i1 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i1 * S + lx.i1));
    }
    const float a = sigmoidf_(__ldg(alpha + cls));
    const float seman = sigmoidf_(sel) * hd;
    out[i] = soft * (1.0f - a) + seman * a;
  }
}

// dstage (B,h,w,NC) zero-initialised (atomics: two objects of one image may share a class);
// dalpha (NC) zero-initialised; dsoft (B,O,h,w) = dout * (1 - a)  (then mask_resize_bwd -> dbmask)
__global__ void stage_mix_bwd_kernel(const float* __restrict__ stage, const long long* __restrict__ y,
                                     const float* __restrict__ alpha, const float* __restrict__ bmask,
                                     const float* __restrict__ hard, const float* __restrict__ dout, int B, int O, int h,
                                     int w, int NC, int S, float* __restrict__ dstage, float* __restrict__ dalpha,
                                     float* __restrict__ dsoft) {
  // one block per (b,o) so that the alpha gradient reduces locally
  const int bo = blockIdx.x;
  const int b = bo / O;
  const int cls = static_cast<int>(__ldg(y + bo));
  const float a = sigmoidf_(__ldg(alpha + cls));
  const float* bm = bmask + 1LL * bo * S * S;
  float da = 0.f;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int x = i % w, py = i / w;
    const float g = __ldg(dout + 1LL * bo * h * w + i);
    const float sel = __ldg(stage + ((1LL * b * h + py) * w + x) * NC + cls);
    const int sy = min(static_cast<int>(floorf(py * (static_cast<float>(S) / h))), S - 1);
    const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(S) / w))), S - 1);
    const float hd = __ldg(hard + 1LL * bo * S * S + sy * S + sx);
    float soft;
    if (S == h && S == w) {
      soft = __ldg(bm + py * S + x);
    } else {
      const Lin1 ly = lin_src(py, S, h), lx = lin_src(x, S, w);
      soft = ly.l0 * (lx.l0 * __ldg(bm + ly.i0 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i0 * S + lx.i1)) +
             ly.l1 * (lx.l0 * __ldg(bm + ly.i1 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i1 * S + lx.i1));
    }
    const float sg = sigmoidf_(sel);
    const float seman = sg * hd;
    dsoft[1LL * bo * h * w + i] = g * (1.0f - a);
    const float dsel = g * a * hd * sg * (1.0f - sg);
    if (dsel != 0.f) atomicAdd(dstage + ((1LL * b * h + py) * w + x) * NC + cls, dsel);
    da += g * (seman - soft);
  }
  da = warp_sum(da);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += red[k];
    atomicAdd(dalpha + cls, t * a * (1.0f - a));
  }
}

int stage_mix_fwd(const float* stage, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  int B, int O, int h, int w, int NC, int S, float* out, cudaStream_t stream) {
  if (!stage || !y || !alpha || !bmask || !hard || !out || B <= 0 || O <= 0 || h <= 0 || w <= 0 || NC <= 0 || S <= 0) { set_error("stage_mix_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * B * O * h * w;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stage_mix_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(stage, y, alpha, bmask, hard, B, O, h, w, NC, S, out);
  return check_launch("stage_mix_fwd_kernel");
}
int stage_mix_bwd(const float* stage, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  const float* dout, int B, int O, int h, int w, int NC, int S, float* dstage, float* dalpha,
                  float* dsoft, cudaStream_t stream) {
  if (!stage || !y || !alpha || !bmask || !hard || !dout || !dstage || !dalpha || !dsoft || B <= 0 || O <= 0) { set_error("stage_mix_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  stage_mix_bwd_kernel<<<B * O, 256, 0, stream>>>(stage, y, alpha, bmask, hard, dout, B, O, h, w, NC, S, dstage, dalpha, dsoft);
  return check_launch("stage_mix_bwd_kernel");
}

}  // namespace l2i

// ------------------------------------------------------------------------------------------------
// Gathered mask head + stage-mask mixing (reference resnet_generator_app_v2.py:646/651 + :466-470).  The reference
// computes all 184 class channels of the stage mask with a 1x1 convolution and then gathers the o channels of the
// image's object classes; here only those o channels are formed, fused with the mixing:
//     sel[b,o,p] = bias[y[b,o]] + sum_c W[y[b,o], c] * t[b,p,c]          (t: the head's 100-channel features)
//     out[b,o,p] = bilinear(bmask)[b,o,p] * (1 - a) + sigmoid(sel) * nearest(hard)[b,o,p] * a,    a = sigmoid(alpha[y])
// One warp per pixel: lanes stride the channels; the O dot products are reduced with the transposed butterfly; the image's
// O weight rows sit in shared memory.  Backward: dt, dW / dbias / dalpha (class-indexed, atomics), dsoft.
// ------------------------------------------------------------------------------------------------
namespace l2i {

struct ClassMixParams {
  const float* t; const float* Wc; const float* bc; const long long* y; const float* alpha; const float* bmask;
  const float* hard; const float* sel_in; const float* dout;
  float* sel; float* out; float* dt; float* dW; float* db; float* dalpha; float* dsoft;
  int B, O, h, w, C, NC, S, pix_per_block;
};

__device__ __forceinline__ float class_mix_soft(const float* __restrict__ bm, int py, int x, int h, int w, int S) {
  if (S == h && S == w) return __ldg(bm + py * S + x);
  const Lin1 ly = lin_src(py, S, h), lx = lin_src(x, S, w);
  return ly.l0 * (lx.l0 * __ldg(bm + ly.i0 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i0 * S + lx.i1)) +
         ly.l1 * (lx.l0 * __ldg(bm + ly.i1 * S + lx.i0) + lx.l1 * __ldg(bm + ly.i1 * S + lx.i1));
}

template <int OM>
__global__ void __launch_bounds__(256) class_mix_fwd_kernel(const ClassMixParams p) {
  extern __shared__ float s_w[];                           // [OM][C] rows of this image's object classes
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < OM * p.C; i += blockDim.x) {
    const int o = i / p.C, c = i - o * p.C;
    s_w[i] = (o < p.O) ? __ldg(p.Wc + static_cast<size_t>(__ldg(p.y + b * p.O + o)) * p.C + c) : 0.f;
  }
  __syncthreads();
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * p.pix_per_block, p1 = min(hw, p0 + p.pix_per_block);
  constexpr int LPO = 32 / OM;                             // lanes that end up holding object o's total
  const int o_mine = lane / LPO;
  const bool holder = (lane % LPO) == 0 && o_mine < p.O;
  int cls = 0;
  float a = 0.f, bias = 0.f;
  if (holder) {
    cls = static_cast<int>(__ldg(p.y + b * p.O + o_mine));
    a = sigmoidf_(__ldg(p.alpha + cls));
    bias = p.bc ? __ldg(p.bc + cls) : 0.f;
  }
  for (int pix = p0 + warp; pix < p1; pix += (blockDim.x >> 5)) {
    const float* tp = p.t + (static_cast<size_t>(b) * hw + pix) * p.C;
    float part[OM];
#pragma unroll
    for (int o = 0; o < OM; ++o) part[o] = 0.f;
    for (int c = lane; c < p.C; c += 32) {
      const float tv = __ldg(tp + c);
#pragma unroll
      for (int o = 0; o < OM; ++o) part[o] = fmaf(tv, s_w[o * p.C + c], part[o]);
    }
    const float tot = warp_transpose_sum<OM>(part, lane);
    if (holder) {
      const int py = pix / p.w, x = pix - py * p.w;
      const float sel = tot + bias;
      const int sy = min(static_cast<int>(floorf(py * (static_cast<float>(p.S) / p.h))), p.S - 1);
      const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(p.S) / p.w))), p.S - 1);
      const size_t bo = static_cast<size_t>(b) * p.O + o_mine;
      const float hd = __ldg(p.hard + bo * p.S * p.S + sy * p.S + sx);
      const float soft = class_mix_soft(p.bmask + bo * p.S * p.S, py, x, p.h, p.w, p.S);
      p.sel[bo * hw + pix] = sel;
      p.out[bo * hw + pix] = soft * (1.0f - a) + sigmoidf_(sel) * hd * a;
    }
  }
}

template <int OM>
__global__ void __launch_bounds__(256) class_mix_bwd_kernel(const ClassMixParams p) {
  extern __shared__ float sh[];                            // s_w [OM][C], then s_dw [OM][C]
  float* s_w = sh;
  float* s_dw = sh + OM * p.C;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < OM * p.C; i += blockDim.x) {
    const int o = i / p.C, c = i - o * p.C;
    s_w[i] = (o < p.O) ? __ldg(p.Wc + static_cast<size_t>(__ldg(p.y + b * p.O + o)) * p.C + c) : 0.f;
    s_dw[i] = 0.f;
  }
  __syncthreads();
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * p.pix_per_block, p1 = min(hw, p0 + p.pix_per_block);
  // lane o (< O) owns object o's per-pixel scalars
  const bool owner = lane < p.O;
  int cls = 0;
  float a = 0.f;
  if (owner) {
    cls = static_cast<int>(__ldg(p.y + b * p.O + lane));
    a = sigmoidf_(__ldg(p.alpha + cls));
  }
  float da = 0.f, dbs = 0.f;
  for (int pix = p0 + warp; pix < p1; pix += (blockDim.x >> 5)) {
    float dsel = 0.f;
    if (owner) {
      const int py = pix / p.w, x = pix - py * p.w;
      const size_t bo = static_cast<size_t>(b) * p.O + lane;
      const float g = __ldg(p.dout + bo * hw + pix);
      const float sg = sigmoidf_(__ldg(p.sel_in + bo * hw + pix));
      const int sy = min(static_cast<int>(floorf(py * (static_cast<float>(p.S) / p.h))), p.S - 1);
      const int sx = min(static_cast<int>(floorf(x * (static_cast<float>(p.S) / p.w))), p.S - 1);
      const float hd = __ldg(p.hard + bo * p.S * p.S + sy * p.S + sx);
      const float soft = class_mix_soft(p.bmask + bo * p.S * p.S, py, x, p.h, p.w, p.S);
      p.dsoft[bo * hw + pix] = g * (1.0f - a);
      dsel = g * a * hd * sg * (1.0f - sg);
      da += g * (sg * hd - soft);
      dbs += dsel;
    }
    const float* tp = p.t + (static_cast<size_t>(b) * hw + pix) * p.C;
    float* dp = p.dt + (static_cast<size_t>(b) * hw + pix) * p.C;
    float ds[OM];                                          // every lane gets all objects' d sel (all 32 lanes take part)
#pragma unroll
    for (int o = 0; o < OM; ++o) ds[o] = __shfl_sync(0xffffffffu, dsel, o);
    for (int c = lane; c < p.C; c += 32) {
      const float tv = __ldg(tp + c);
      float acc = 0.f;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        acc = fmaf(ds[o], s_w[o * p.C + c], acc);
        if (ds[o] != 0.f) atomicAdd(&s_dw[o * p.C + c], ds[o] * tv);
      }
      dp[c] = acc;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.O * p.C; i += blockDim.x) {
    const int o = i / p.C, c = i - o * p.C;
    const float v = s_dw[o * p.C + c];
    if (v != 0.f) atomicAdd(p.dW + static_cast<size_t>(__ldg(p.y + b * p.O + o)) * p.C + c, v);
  }
  if (owner) {
    // the 8 warps of the block each hold partial sums for object `lane`
    atomicAdd(p.dalpha + cls, da * a * (1.0f - a));
    if (p.db) atomicAdd(p.db + cls, dbs);
  }
}

static int class_mix_check(const ClassMixParams& p) {
  if (!p.t || !p.Wc || !p.y || !p.alpha || !p.bmask || !p.hard || p.B <= 0 || p.O <= 0 || p.O > 32 || p.h <= 0 || p.w <= 0 ||
      p.C <= 0 || p.NC <= 0 || p.S <= 0 || p.C > 1024) {
    set_error("class_mix: bad arguments (O <= 32, C <= 1024)");
    return L2I_ERR_BAD_ARG;
  }
  return L2I_OK;
}

int class_mix_fwd(const float* t, const float* Wc, const float* bc, const long long* y, const float* alpha, const float* bmask,
                  const float* hard, int B, int O, int h, int w, int C, int NC, int S, float* sel, float* out,
                  cudaStream_t stream) {
  ClassMixParams p{};
  p.t = t; p.Wc = Wc; p.bc = bc; p.y = y; p.alpha = alpha; p.bmask = bmask; p.hard = hard; p.sel = sel; p.out = out;
  p.B = B; p.O = O; p.h = h; p.w = w; p.C = C; p.NC = NC; p.S = S;
  int rc = class_mix_check(p);
  if (rc) return rc;
  if (!sel || !out) { set_error("class_mix_fwd: null output"); return L2I_ERR_BAD_ARG; }
  const int hw = h * w;
  int chunks = (148 * 8 + B - 1) / B;
  if (chunks > (hw + 7) / 8) chunks = (hw + 7) / 8;
  p.pix_per_block = (hw + chunks - 1) / chunks;
  dim3 grid((hw + p.pix_per_block - 1) / p.pix_per_block, B);
  const int om = O <= 8 ? 8 : (O <= 16 ? 16 : 32);
  const size_t smem = sizeof(float) * om * C;
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(class_mix_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
    cudaFuncSetAttribute(class_mix_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024 * 4);
    configured.done();
  }
  if (om == 8) class_mix_fwd_kernel<8><<<grid, 256, smem, stream>>>(p);
  else if (om == 16) class_mix_fwd_kernel<16><<<grid, 256, smem, stream>>>(p);
  else class_mix_fwd_kernel<32><<<grid, 256, smem, stream>>>(p);
  return check_launch("class_mix_fwd_kernel");
}

int class_mix_bwd(const float* t, const float* Wc, const long long* y, const float* alpha, const float* bmask, const float* hard,
                  const float* sel, const float* dout, int B, int O, int h, int w, int C, int NC, int S, float* dt, float* dW,
                  float* db, float* dalpha, float* dsoft, cudaStream_t stream) {
  ClassMixParams p{};
  p.t = t; p.Wc = Wc; p.y = y; p.alpha = alpha; p.bmask = bmask; p.hard = hard; p.sel_in = sel; p.dout = dout;
  p.dt = dt; p.dW = dW; p.db = db; p.dalpha = dalpha; p.dsoft = dsoft;
  p.B = B; p.O = O; p.h = h; p.w = w; p.C = C; p.NC = NC; p.S = S;
  int rc = class_mix_check(p);
  if (rc) return rc;
  if (!sel || !dout || !dt || !dW || !dalpha || !dsoft || C > 512) { set_error("class_mix_bwd: bad arguments (C <= 512)"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(dW, 0, sizeof(float) * static_cast<size_t>(NC) * C, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dalpha, 0, sizeof(float) * NC, stream);
  if (e == cudaSuccess && db) e = cudaMemsetAsync(db, 0, sizeof(float) * NC, stream);
  if (e != cudaSuccess) { set_error("class_mix_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  const int hw = h * w;
  int chunks = (148 * 4 + B - 1) / B;
  if (chunks > (hw + 63) / 64) chunks = (hw + 63) / 64;
  if (chunks < 1) chunks = 1;
  p.pix_per_block = (hw + chunks - 1) / chunks;
  dim3 grid((hw + p.pix_per_block - 1) / p.pix_per_block, B);
  const int om = O <= 8 ? 8 : (O <= 16 ? 16 : 32);
  const size_t smem = sizeof(float) * 2 * om * C;
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(class_mix_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32 * 512 * 4);
    cudaFuncSetAttribute(class_mix_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 512 * 4);
    configured.done();
  }
  if (om == 8) class_mix_bwd_kernel<8><<<grid, 256, smem, stream>>>(p);
  else if (om == 16) class_mix_bwd_kernel<16><<<grid, 256, smem, stream>>>(p);
  else class_mix_bwd_kernel<32><<<grid, 256, smem, stream>>>(p);
  return check_launch("class_mix_bwd_kernel");
}

}  // namespace l2i

// ------------------------------------------------------------------------------------------------
// Mask-regression trunk (reference model/mask_regression.py:66-99): after each 3x3 conv comes
// InstanceNorm2d(256) (affine=False, biased variance, eps 1e-5) -> ReLU -> [bilinear x2, align_corners=False].
// inorm_relu_fwd writes the NEXT convolution's bf16 operand pair directly; inorm_relu_bwd takes the gradient
// w.r.t. that (up-sampled) tensor back to the conv output.  One block per sample, one thread per channel
// (coalesced over channels), so the per-(sample, channel) statistics never leave the thread.
// ------------------------------------------------------------------------------------------------
namespace l2i {

// torch upsample_bilinear2d, align_corners=False, scale 2: src = max((dst + 0.5) / 2 - 0.5, 0)
__device__ __forceinline__ void up2_taps(int dst, int in, int& i0, int& i1, float& l1) {
  const float src = fmaxf((static_cast<float>(dst) + 0.5f) * 0.5f - 0.5f, 0.f);
  i0 = static_cast<int>(src);
  i1 = min(i0 + 1, in - 1);
  l1 = src - static_cast<float>(i0);
}

__global__ void __launch_bounds__(256) inorm_stats_kernel(const float* __restrict__ x, int HW, int C, float eps,
                                                          float* __restrict__ stats) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = x + static_cast<size_t>(n) * HW * C + c;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < HW; ++i) s += __ldg(p + static_cast<size_t>(i) * C);
    const float mean = s / HW;
    for (int i = 0; i < HW; ++i) { const float d = __ldg(p + static_cast<size_t>(i) * C) - mean; q = fmaf(d, d, q); }   // two-pass variance
    const float var = q / HW;
    stats[(static_cast<size_t>(n) * C + c) * 2] = mean;
    stats[(static_cast<size_t>(n) * C + c) * 2 + 1] = rsqrtf(var + eps);
  }
}

// thread = (sample, output pixel, 8-channel group)
__global__ void __launch_bounds__(256) inorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                          int N, int H, int W, int C, int up,
                                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int cpad) {
  const int Ho = H << up, Wo = W << up;
  const int groups = cpad >> 3;
  const long long total = 1LL * N * Ho * Wo * groups;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long opix = i / groups;
    const int wo = static_cast<int>(opix % Wo);
    const int ho = static_cast<int>((opix / Wo) % Ho);
    const int n = static_cast<int>(opix / (1LL * Wo * Ho));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    int y0 = ho, y1 = ho, x0 = wo, x1 = wo;
    float ly = 0.f, lx = 0.f;
    if (up) { up2_taps(ho, H, y0, y1, ly); up2_taps(wo, W, x0, x1, lx); }
    const float wts[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
    const int ntap = up ? 4 : 1;
    for (int t = 0; t < ntap; ++t) {
      const float* p = x + ((static_cast<size_t>(n) * H + ys[t]) * W + xs[t]) * C + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        if (c < C) {
          const float mean = __ldg(stats + (static_cast<size_t>(n) * C + c) * 2);
          const float rstd = __ldg(stats + (static_cast<size_t>(n) * C + c) * 2 + 1);
          const float yv = fmaxf((__ldg(p + j) - mean) * rstd, 0.f);
          v[j] = fmaf(up ? wts[t] : 1.f, yv, v[j]);
        }
      }
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(v[j], ah, al);
      split_bf16(v[j + 1], bh, bl);
      ph[j >> 1] = pack_bf16x2(ah, bh);
      pl[j >> 1] = pack_bf16x2(al, bl);
    }
    *reinterpret_cast<uint4*>(hi + opix * cpad + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo + opix * cpad + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// block = sample, thread = channel.  Phase 1: dz = relu'(y) * (transpose of the x2 up-sampling applied to da),
// stored in dx, with its two per-channel sums; phase 2: dx = rstd * (dz - mean(dz) - xh * mean(dz * xh)).
__global__ void __launch_bounds__(256) inorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                        const float* __restrict__ da, int H, int W, int C, int up,
                                                        float* __restrict__ dx) {
  const int n = blockIdx.x;
  const int Ho = H << up, Wo = W << up;
  const int HW = H * W;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = stats[(static_cast<size_t>(n) * C + c) * 2], rstd = stats[(static_cast<size_t>(n) * C + c) * 2 + 1];
    const float* xp = x + static_cast<size_t>(n) * HW * C + c;
    const float* gp = da + static_cast<size_t>(n) * Ho * Wo * C + c;
    float* op = dx + static_cast<size_t>(n) * HW * C + c;
    float s1 = 0.f, s2 = 0.f;
    for (int h = 0; h < H; ++h) {
      for (int w = 0; w < W; ++w) {
        float g = 0.f;
        if (up) {
          for (int qy = max(2 * h - 1, 0); qy <= min(2 * h + 2, Ho - 1); ++qy) {
            int y0, y1; float ly;
            up2_taps(qy, H, y0, y1, ly);
            const float wy = (y0 == h ? 1.f - ly : 0.f) + (y1 == h ? ly : 0.f);
            if (wy == 0.f) continue;
            for (int qx = max(2 * w - 1, 0); qx <= min(2 * w + 2, Wo - 1); ++qx) {
              int x0, x1; float lx;
              up2_taps(qx, W, x0, x1, lx);
              const float wx = (x0 == w ? 1.f - lx : 0.f) + (x1 == w ? lx : 0.f);
              if (wx != 0.f) g = fmaf(wy * wx, __ldg(gp + (static_cast<size_t>(qy) * Wo + qx) * C), g);
            }
          }
        } else {
          g = __ldg(gp + (static_cast<size_t>(h) * W + w) * C);
        }
        const float xh = (__ldg(xp + (static_cast<size_t>(h) * W + w) * C) - mean) * rstd;
        const float dz = xh > 0.f ? g : 0.f;
        op[(static_cast<size_t>(h) * W + w) * C] = dz;
        s1 += dz;
        s2 = fmaf(dz, xh, s2);
      }
    }
    const float m1 = s1 / HW, m2 = s2 / HW;
    for (int i = 0; i < HW; ++i) {
      const float xh = (__ldg(xp + static_cast<size_t>(i) * C) - mean) * rstd;
      op[static_cast<size_t>(i) * C] = rstd * (op[static_cast<size_t>(i) * C] - m1 - xh * m2);
    }
  }
}

int inorm_relu_fwd(const float* x, int N, int H, int W, int C, int up2, float eps, float* stats, void* hi, void* lo,
                   int cpad, cudaStream_t stream) {
  if (!x || !stats || !hi || !lo || N <= 0 || H <= 0 || W <= 0 || C <= 0 || cpad < C || cpad % 8) {
    set_error("inorm_relu_fwd: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  inorm_stats_kernel<<<N, 256, 0, stream>>>(x, H * W, C, eps, stats);
  int rc = check_launch("inorm_stats_kernel");
  if (rc) return rc;
  const int up = up2 ? 1 : 0;
  const long long total = 1LL * N * (H << up) * (W << up) * (cpad >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  inorm_apply_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, stats, N, H, W, C, up, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                  reinterpret_cast<__nv_bfloat16*>(lo), cpad);
  return check_launch("inorm_apply_kernel");
}

int inorm_relu_bwd(const float* x, const float* stats, const float* da, int N, int H, int W, int C, int up2, float* dx,
                   cudaStream_t stream) {
  if (!x || !stats || !da || !dx || N <= 0 || H <= 0 || W <= 0 || C <= 0) {
    set_error("inorm_relu_bwd: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  inorm_bwd_kernel<<<N, 256, 0, stream>>>(x, stats, da, H, W, C, up2 ? 1 : 0, dx);
  return check_launch("inorm_bwd_kernel");
}

}  // namespace l2i
