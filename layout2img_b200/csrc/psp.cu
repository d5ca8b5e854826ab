// Pyramid pooling of the res4 mask head (reference model/resnet_generator_app_v2.py:724-752, PSPModule):
//   priors_s = ReLU(BN(conv1x1(AdaptiveAvgPool_s(feats)))),  s in {1, 2, 3, 6}
//   cat      = [ up(priors_1), up(priors_2), up(priors_3), up(priors_6), feats ]   (bilinear, align_corners=True)
//   bottle   = conv3x3(cat) ...
// Three HBM-bound kernels replace ~15 library launches (adaptive pools, four up-samplings whose backward
// serialises thousands of atomics on a 1x1 target, the concat and the fp32 -> pair split of its 528 channels):
//   psp_pool_fwd/bwd   : all four adaptive average pools in one pass over feats   -> pooled [B, 50, C]
//   psp_concat_fwd     : writes the 3x3 conv's bf16 operand pair [B,H,W,4*CP+C] straight from priors + feats
//   psp_concat_bwd     : gradient of the four up-samplings as per-thread gathers + one atomic per (band, cell)
// Cell order inside the 50: s=1 (1 cell), s=2 (4), s=3 (9), s=6 (36), row-major inside a stage.
#include "common.cuh"
#include "kernels.h"

namespace l2i {

__constant__ int kPspSize[4] = {1, 2, 3, 6};
__constant__ int kPspOff[4] = {0, 1, 5, 14};
static constexpr int kPspCells = 50;

__device__ __forceinline__ int bin_start(int i, int in, int s) { return (i * in) / s; }              // floor
__device__ __forceinline__ int bin_end(int i, int in, int s) { return ((i + 1) * in + s - 1) / s; }  // ceil

// block = (cell, image); thread = channel (strided); fp32 sums of <= H*W pixels
__global__ void __launch_bounds__(128) psp_pool_fwd_kernel(const float* __restrict__ x, int H, int W, int C,
                                                           float* __restrict__ pooled) {
  const int cell = blockIdx.x, b = blockIdx.y;
  int st = 3;
  if (cell < 1) st = 0; else if (cell < 5) st = 1; else if (cell < 14) st = 2;
  const int s = kPspSize[st], local = cell - kPspOff[st];
  const int iy = local / s, ix = local % s;
  const int h0 = bin_start(iy, H, s), h1 = bin_end(iy, H, s), w0 = bin_start(ix, W, s), w1 = bin_end(ix, W, s);
  const float inv = 1.0f / static_cast<float>((h1 - h0) * (w1 - w0));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int h = h0; h < h1; ++h) {
      const float* row = x + ((static_cast<size_t>(b) * H + h) * W) * C + c;
      float racc = 0.f;
      for (int w = w0; w < w1; ++w) racc += __ldg(row + static_cast<size_t>(w) * C);
      acc += racc;
    }
    pooled[(static_cast<size_t>(b) * kPspCells + cell) * C + c] = acc * inv;
  }
}

// dx[b,h,w,c] = base[b,h,w,base_off + c] (optional) + sum over the cells that contain (h,w) of dpooled / area.
// block = (image row h, image b); the (<= 2 per axis and stage) bins covering each coordinate are tabulated in
// shared memory once per block; thread <-> (w, float4 channel group).
__global__ void __launch_bounds__(256) psp_pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ base,
                                                           int base_stride, int base_off, int H, int W, int C,
                                                           float* __restrict__ dx) {
  extern __shared__ unsigned char s_raw[];
  int* s_wbin = reinterpret_cast<int*>(s_raw);                    // [4][W][2]
  float* s_winv = reinterpret_cast<float*>(s_wbin + 8 * W);        // [4][W][2]
  __shared__ int s_hbin[4][2];
  __shared__ float s_hinv[4][2];
  const int h = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < 4 * (W + 1); i += blockDim.x) {
    const int st = i / (W + 1), pos = i % (W + 1);
    const int s = kPspSize[st];
    const bool is_h = pos == W;
    const int coord = is_h ? h : pos, extent = is_h ? H : W;
    int n = 0, bins[2] = {-1, -1};
    float inv[2] = {0.f, 0.f};
    const int c0 = (coord * s) / extent;
    for (int k = max(c0 - 1, 0); k <= min(c0 + 1, s - 1) && n < 2; ++k) {
      const int lo = bin_start(k, extent, s), hi = bin_end(k, extent, s);
      if (coord >= lo && coord < hi) { bins[n] = k; inv[n] = 1.0f / static_cast<float>(hi - lo); ++n; }
    }
    for (int k = 0; k < 2; ++k) {
      if (is_h) { s_hbin[st][k] = bins[k]; s_hinv[st][k] = inv[k]; }
      else { s_wbin[(st * W + pos) * 2 + k] = bins[k]; s_winv[(st * W + pos) * 2 + k] = inv[k]; }
    }
  }
  __syncthreads();
  const int c4n = C >> 2;
  const float* dp = dpooled + static_cast<size_t>(b) * kPspCells * C;
  for (int i = threadIdx.x; i < W * c4n; i += blockDim.x) {
    const int w = i / c4n, c = (i - w * c4n) * 4;
    const size_t pix = (static_cast<size_t>(b) * H + h) * W + w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (base) acc = __ldg(reinterpret_cast<const float4*>(base + pix * base_stride + base_off + c));
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int s = kPspSize[st];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int iy = s_hbin[st][a];
        if (iy < 0) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ix = s_wbin[(st * W + w) * 2 + e];
          if (ix < 0) continue;
          const float wt = s_hinv[st][a] * s_winv[(st * W + w) * 2 + e];
          const float4 g = __ldg(reinterpret_cast<const float4*>(dp + static_cast<size_t>(kPspOff[st] + iy * s + ix) * C + c));
          acc.x = fmaf(wt, g.x, acc.x); acc.y = fmaf(wt, g.y, acc.y); acc.z = fmaf(wt, g.z, acc.z); acc.w = fmaf(wt, g.w, acc.w);
        }
      }
    }
    *reinterpret_cast<float4*>(dx + pix * C + c) = acc;
  }
}

// torch upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1)
__device__ __forceinline__ void ac_taps(int dst, int in, int out, int& i0, int& i1, float& l1) {
  const float scale = (out > 1) ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(dst);
  i0 = static_cast<int>(src);
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - static_cast<float>(i0);
}

// thread = (pixel, channel quad of the concat); CP = channels per prior (multiple of 4), C = feats channels
__global__ void __launch_bounds__(256) psp_concat_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ priors,
                                                             int B, int H, int W, int C, int CP,
                                                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                             int cpad) {
  const int ctot = 4 * CP + C;
  const int quads = cpad >> 2;
  const long long total = 1LL * B * H * W * quads;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int q = static_cast<int>(i % quads);
    const long long pix = i / quads;
    const int w = static_cast<int>(pix % W);
    const int h = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (1LL * W * H));
    const int c = q * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < 4 * CP) {
      const int st = c / CP, cc = c - st * CP;
      const int s = kPspSize[st];
      int y0, y1, x0, x1;
      float ly, lx;
      ac_taps(h, s, H, y0, y1, ly);
      ac_taps(w, s, W, x0, x1, lx);
      const float* base = priors + (static_cast<size_t>(b) * kPspCells + kPspOff[st]) * CP + cc;
      const float4 p00 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * s + x0) * CP));
      const float4 p01 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(y0 * s + x1) * CP));
      const float4 p10 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * s + x0) * CP));
      const float4 p11 = __ldg(reinterpret_cast<const float4*>(base + static_cast<size_t>(y1 * s + x1) * CP));
      const float hy = 1.f - ly, hx = 1.f - lx;
      // same association as torch: h0lambda * (w0lambda * a + w1lambda * b) + h1lambda * (w0lambda * c + w1lambda * d)
      v[0] = hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
      v[1] = hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
      v[2] = hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
      v[3] = hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
    } else if (c < ctot) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(feats + pix * C + (c - 4 * CP)));
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    __nv_bfloat16 ah, al, bh, bl, ch, cl, dh, dl;
    split_bf16(v[0], ah, al); split_bf16(v[1], bh, bl); split_bf16(v[2], ch, cl); split_bf16(v[3], dh, dl);
    *reinterpret_cast<uint2*>(hi + pix * cpad + c) = make_uint2(pack_bf16x2(ah, bh), pack_bf16x2(ch, dh));
    *reinterpret_cast<uint2*>(lo + pix * cpad + c) = make_uint2(pack_bf16x2(al, bl), pack_bf16x2(cl, dl));
  }
}

// block = (row band, image); thread = (stage, channel) pair; per-thread cell accumulators in shared memory
__global__ void __launch_bounds__(416) psp_concat_bwd_kernel(const float* __restrict__ dcat, int H, int W, int CP, int cstride,
                                                             int rows_per_band, float* __restrict__ dpriors) {
  extern __shared__ float s_acc[];                     // [36][blockDim.x]
  const int t = threadIdx.x;
  const int b = blockIdx.y;
  const int st = t / CP, cc = t - st * CP;
  const bool active = st < 4;
  const int s = active ? kPspSize[st] : 1;
  for (int k = 0; k < 36; ++k) s_acc[k * blockDim.x + t] = 0.f;
  const int h_begin = blockIdx.x * rows_per_band, h_end = min(H, h_begin + rows_per_band);
  if (active) {
    for (int h = h_begin; h < h_end; ++h) {
      int y0, y1;
      float ly;
      ac_taps(h, s, H, y0, y1, ly);
      const float* row = dcat + ((static_cast<size_t>(b) * H + h) * W) * cstride + st * CP + cc;
      for (int w = 0; w < W; ++w) {
        int x0, x1;
        float lx;
        ac_taps(w, s, W, x0, x1, lx);
        const float g = __ldg(row + static_cast<size_t>(w) * cstride);
        const float hy = 1.f - ly, hx = 1.f - lx;
        s_acc[(y0 * s + x0) * blockDim.x + t] += hy * hx * g;
        s_acc[(y0 * s + x1) * blockDim.x + t] += hy * lx * g;
        s_acc[(y1 * s + x0) * blockDim.x + t] += ly * hx * g;
        s_acc[(y1 * s + x1) * blockDim.x + t] += ly * lx * g;
      }
    }
    for (int k = 0; k < s * s; ++k)
      atomicAdd(dpriors + (static_cast<size_t>(b) * kPspCells + kPspOff[st] + k) * CP + cc, s_acc[k * blockDim.x + t]);
  }
}

int psp_pool_fwd(const float* x, int B, int H, int W, int C, float* pooled, cudaStream_t stream) {
  if (!x || !pooled || B <= 0 || H < 6 || W < 6 || C <= 0) { set_error("psp_pool_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  psp_pool_fwd_kernel<<<dim3(kPspCells, B), 128, 0, stream>>>(x, H, W, C, pooled);
  return check_launch("psp_pool_fwd_kernel");
}

int psp_pool_bwd(const float* dpooled, const float* base, int base_stride, int base_off, int B, int H, int W, int C,
                 float* dx, cudaStream_t stream) {
  if (!dpooled || !dx || B <= 0 || H < 6 || W < 6 || C <= 0 || (C & 3) || W > 512 || (base && ((base_stride | base_off) & 3))) {
    set_error("psp_pool_bwd: bad arguments (C and the base offsets must be multiples of 4)");
    return L2I_ERR_BAD_ARG;
  }
  psp_pool_bwd_kernel<<<dim3(H, B), 256, 16 * W * sizeof(float), stream>>>(dpooled, base, base_stride, base_off, H, W, C, dx);
  return check_launch("psp_pool_bwd_kernel");
}

int psp_concat_fwd(const float* feats, const float* priors, int B, int H, int W, int C, int CP, void* hi, void* lo, int cpad,
                   cudaStream_t stream) {
  if (!feats || !priors || !hi || !lo || B <= 0 || H < 2 || W < 2 || (C & 3) || (CP & 3) || cpad % 8 || cpad < 4 * CP + C) {
    set_error("psp_concat_fwd: bad arguments (C=%d CP=%d cpad=%d)", C, CP, cpad);
    return L2I_ERR_BAD_ARG;
  }
  const long long total = 1LL * B * H * W * (cpad >> 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  psp_concat_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(feats, priors, B, H, W, C, CP,
                                                                     reinterpret_cast<__nv_bfloat16*>(hi),
                                                                     reinterpret_cast<__nv_bfloat16*>(lo), cpad);
  return check_launch("psp_concat_fwd_kernel");
}

int psp_concat_bwd(const float* dcat, int B, int H, int W, int CP, int cstride, float* dpriors, cudaStream_t stream) {
  if (!dcat || !dpriors || B <= 0 || H < 2 || W < 2 || CP <= 0 || 4 * CP > 416 || cstride < 4 * CP) {
    set_error("psp_concat_bwd: bad arguments (CP=%d cstride=%d)", CP, cstride);
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaMemsetAsync(dpriors, 0, sizeof(float) * B * kPspCells * CP, stream);
  if (e != cudaSuccess) { set_error("psp_concat_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  const int threads = 416;
  const size_t smem = sizeof(float) * 36 * threads;
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(psp_concat_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured.done();
  }
  const int rows = 4;
  psp_concat_bwd_kernel<<<dim3((H + rows - 1) / rows, B), threads, smem, stream>>>(dcat, H, W, CP, cstride, rows, dpriors);
  return check_launch("psp_concat_bwd_kernel");
}

}  // namespace l2i
