// ISLA normalisation (SpatialAdaptiveSynBatchNorm2d, reference model/norm_module.py:152-189) and its
// batch-norm statistics (model/sync_batchnorm/batchnorm.py:48-53 single-device branch == F.batch_norm),
// forward and backward, NHWC fp32.  Arithmetic follows SURVEY.md Appendix B:
//
//   xh   = (x - mean_c) * invstd_c                      (batch stats in train, running stats in eval)
//   S    = sum_o m_o + 1e-6
//   G    = sum_o m_o gamma_oc / S + 1 ;  Bt = sum_o m_o beta_oc / S
//   out  = G * xh + Bt          -> optionally ReLU, nearest x2 up-sampling, bf16 (hi, lo) pair
//
// The reference materialises two (b,o,C,h,w) products per call (2.1 GB each at the last layer);
// here x is read once per pass, the (b,h,w,o) mask once per pixel and gamma/beta from shared memory.
// HBM-bound: algorithmic bytes per element = 4 (x) + 4 or 4*4 (pair at 1x / 2x resolution).
#include "common.cuh"
#include "kernels.h"

namespace l2i {

static constexpr float kMaskEps = 1e-6f;

// ------------------------------------------------------------------------------------------
// per-channel sum / sum of squares over all pixels.  thread <-> 4 consecutive channels,
// fp32 partials over short runs folded into fp64, one fp64 atomic per (block, channel).
// ------------------------------------------------------------------------------------------
__global__ void bn_stats_kernel(const float* __restrict__ x, long long pixels, int C, double* __restrict__ sums) {
  const int cg = C >> 2;                                  // float4 groups per pixel
  const int tpp = blockDim.x / cg > 0 ? blockDim.x / cg : 1;  // pixels handled per block iteration
  const int g = threadIdx.x % cg;
  const int prow = threadIdx.x / cg;
  const bool active = prow < tpp;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  const long long stride = 1LL * gridDim.x * tpp;
  long long p = 1LL * blockIdx.x * tpp + prow;
  while (active && p < pixels) {
    float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
#pragma unroll 4
    for (int it = 0; it < 16 && p < pixels; ++it, p += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + p * C) + g);
      fs[0] += v.x; fs[1] += v.y; fs[2] += v.z; fs[3] += v.w;
      fq[0] += v.x * v.x; fq[1] += v.y * v.y; fq[2] += v.z * v.z; fq[3] += v.w * v.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j] += fs[j]; q[j] += fq[j]; }
  }
  extern __shared__ double sh_d[];                        // [tpp][C][2]
  if (active) {
    double* mine = sh_d + (static_cast<size_t>(prow) * C + g * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { mine[j * 2] = s[j]; mine[j * 2 + 1] = q[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    double a = 0;
    for (int r = 0; r < tpp; ++r) a += sh_d[static_cast<size_t>(r) * C * 2 + i];
    atomicAdd(sums + i, a);
  }
}

// scalar fallback for channel counts that are not a multiple of 4 (none on the named path)
__global__ void bn_stats_scalar_kernel(const float* __restrict__ x, long long pixels, int C, double* __restrict__ sums) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (long long p = blockIdx.x; p < pixels; p += gridDim.x) {
    const float v = __ldg(x + p * C + c);
    s += v; q += static_cast<double>(v) * v;
  }
  atomicAdd(sums + 2 * c, s);
  atomicAdd(sums + 2 * c + 1, q);
}

// mean / invstd from the fp64 sums (+ F.batch_norm's running-stat update, momentum 0.1, unbiased var)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[2 * c] / count;
  double var = sums[2 * c + 1] / count - m * m;
  if (var < 0) var = 0;
  mean_invstd[c] = static_cast<float>(m);
  mean_invstd[C + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(m);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// eval mode: mean / invstd from the running statistics
__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean_invstd[c] = rm[c];
  mean_invstd[C + c] = 1.0f / sqrtf(rv[c] + eps);
}

int bn_stats(const float* x, long long pixels, int C, double* sums, cudaStream_t stream) {
  if (!x || !sums || pixels <= 0 || C <= 0) { set_error("bn_stats: bad arguments"); return L2I_ERR_BAD_ARG; }
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream);
  if (e != cudaSuccess) { set_error("bn_stats: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  if ((C & 3) == 0 && C <= 1024) {
    const int cg = C >> 2;
    const int threads = cg >= 256 ? cg : 256;
    const int tpp = threads / cg;
    long long blocks = (pixels + tpp * 16 - 1) / (tpp * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    const size_t smem = sizeof(double) * 2 * C * tpp;
    bn_stats_kernel<<<static_cast<int>(blocks), threads, smem, stream>>>(x, pixels, C, sums);
  } else {
    long long bx = pixels < 1184 ? pixels : 1184;
    dim3 grid(static_cast<int>(bx), (C + 127) / 128);
    bn_stats_scalar_kernel<<<grid, 128, 0, stream>>>(x, pixels, C, sums);
  }
  return check_launch("bn_stats_kernel");
}

int bn_finalize(const double* sums, double count, int C, float eps, float momentum, float* running_mean,
                float* running_var, float* mean_invstd, cudaStream_t stream) {
  if (!sums || !mean_invstd || C <= 0 || count <= 0) { set_error("bn_finalize: bad arguments"); return L2I_ERR_BAD_ARG; }
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, count, C, eps, momentum, running_mean, running_var, mean_invstd);
  return check_launch("bn_finalize_kernel");
}

int bn_eval_stats(const float* rm, const float* rv, int C, float eps, float* mean_invstd, cudaStream_t stream) {
  if (!rm || !rv || !mean_invstd || C <= 0) { set_error("bn_eval_stats: bad arguments"); return L2I_ERR_BAD_ARG; }
  bn_eval_stats_kernel<<<(C + 127) / 128, 128, 0, stream>>>(rm, rv, C, eps, mean_invstd);
  return check_launch("bn_eval_stats_kernel");
}

// ------------------------------------------------------------------------------------------
// forward apply.  grid = (pixel chunks, channel chunks, images); gamma/beta of this image and
// channel chunk live in shared memory; thread <-> (pixel, 8 consecutive channels).
// O == 0 : plain affine batch norm  out = xh * aff_w + aff_b   (final.0, mask heads)
// ------------------------------------------------------------------------------------------
static constexpr int kIslaCc = 256;   // channels per block (shared memory = O * 256 * 2 floats)

struct IslaFwdParams {
  const float* x;           // [B,H,W,C]
  const float* mean_invstd; // [2C]
  const float* mask;        // [B,H,W,O] pixel-major, or null when O == 0
  const float* gamma;       // [B,O,C]
  const float* beta;        // [B,O,C]
  const float* aff_w;       // [C] or null
  const float* aff_b;       // [C] or null
  float* out;               // [B,H,W,C] fp32 (pre-ReLU) or null
  __nv_bfloat16* hi;        // [B,H<<up,W<<up,cpad] or null
  __nv_bfloat16* lo;
  int B, H, W, C, O, cpad, relu, up;
};

__global__ void __launch_bounds__(256) isla_fwd_kernel(const IslaFwdParams p) {
  extern __shared__ float sh[];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kIslaCc;
  const int cc = min(kIslaCc, p.C - c0);
  float* s_gam = sh;                       // [O][cc]
  float* s_bet = sh + p.O * kIslaCc;
  for (int i = threadIdx.x; i < p.O * cc; i += blockDim.x) {
    const int o = i / cc, c = i - o * cc;
    s_gam[o * kIslaCc + c] = __ldg(p.gamma + (static_cast<size_t>(b) * p.O + o) * p.C + c0 + c);
    s_bet[o * kIslaCc + c] = __ldg(p.beta + (static_cast<size_t>(b) * p.O + o) * p.C + c0 + c);
  }
  __syncthreads();
  // 8-channel groups in this chunk; a pair's padding channels (up to cpad) are written as zeros
  const int ccp = p.hi ? min(kIslaCc, p.cpad - c0) : cc;
  const int groups = (max(cc, ccp) + 7) >> 3;
  const int hw = p.H * p.W;
  const int Ho = p.H << p.up, Wo = p.W << p.up;
  const long long items = 1LL * hw * groups;
  for (long long it = 1LL * blockIdx.x * blockDim.x + threadIdx.x; it < items; it += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(it % groups);
    const int pix = static_cast<int>(it / groups);
    const int c = g * 8;                   // within chunk
    const size_t gp = static_cast<size_t>(b) * hw + pix;
    const float* xp = p.x + gp * p.C + c0 + c;
    float xv[8], gm[8], bt[8];
    const bool full = (c + 8 <= cc) && ((p.C & 3) == 0);
    if (full) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xp));
      const float4 d = __ldg(reinterpret_cast<const float4*>(xp) + 1);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = d.x; xv[5] = d.y; xv[6] = d.z; xv[7] = d.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = (c + j < cc) ? __ldg(xp + j) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { gm[j] = 0.f; bt[j] = 0.f; }
    float S = kMaskEps;
    if (p.O > 0) {
      const float* mp = p.mask + gp * p.O;
      for (int o = 0; o < p.O; ++o) {
        const float m = __ldg(mp + o);
        S += m;
        const float* sg = s_gam + o * kIslaCc + c;
        const float* sb = s_bet + o * kIslaCc + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          gm[j] = fmaf(m, sg[j], gm[j]);
          bt[j] = fmaf(m, sb[j], bt[j]);
        }
      }
    }
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = c0 + c + j;
      float v = 0.f;
      if (c + j < cc) {
        const float xh = (xv[j] - __ldg(p.mean_invstd + ch)) * __ldg(p.mean_invstd + p.C + ch);
        if (p.O > 0) {
          v = (gm[j] / S + 1.0f) * xh + bt[j] / S;
        } else {
          v = xh;
          if (p.aff_w) v = v * __ldg(p.aff_w + ch) + __ldg(p.aff_b + ch);
        }
      }
      y[j] = v;
    }
    if (p.out) {
      float* op = p.out + gp * p.C + c0 + c;
      if (full) {
        *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
        *(reinterpret_cast<float4*>(op) + 1) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
        for (int j = 0; j < 8 && c + j < cc; ++j) op[j] = y[j];
      }
    }
    if (p.hi) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float a = y[j], d = y[j + 1];
        if (p.relu) { a = fmaxf(a, 0.f); d = fmaxf(d, 0.f); }
        __nv_bfloat16 ah, al, dh, dl;
        split_bf16(a, ah, al);
        split_bf16(d, dh, dl);
        ph[j >> 1] = pack_bf16x2(ah, dh);
        pl[j >> 1] = pack_bf16x2(al, dl);
      }
      if (c0 + c < p.cpad) {
        const int h = pix / p.W, w = pix - h * p.W;
        const uint4 vh = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        const uint4 vl = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        const int rep = 1 << p.up;
        for (int dy = 0; dy < rep; ++dy)
          for (int dx = 0; dx < rep; ++dx) {
            const size_t op = ((static_cast<size_t>(b) * Ho + (h << p.up) + dy) * Wo + (w << p.up) + dx) * p.cpad + c0 + c;
            *reinterpret_cast<uint4*>(p.hi + op) = vh;
            *reinterpret_cast<uint4*>(p.lo + op) = vl;
          }
      }
    }
  }
}

int isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, int B, int H, int W, int C, int O, float* out, void* hi, void* lo,
             int cpad, int relu, int up2, cudaStream_t stream) {
  if (!x || !mean_invstd || B <= 0 || H <= 0 || W <= 0 || C <= 0 || O < 0 || (!out && !hi)) { set_error("isla_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  if (O > 0 && (!mask || !gamma || !beta)) { set_error("isla_fwd: mask/gamma/beta required when O > 0"); return L2I_ERR_BAD_ARG; }
  if (hi && (!lo || cpad % 8 || cpad < C)) { set_error("isla_fwd: bad pair arguments"); return L2I_ERR_BAD_ARG; }
  if (O > 48) { set_error("isla_fwd: at most 48 objects per image supported (got %d)", O); return L2I_ERR_UNSUPPORTED; }
  IslaFwdParams p;
  p.x = x; p.mean_invstd = mean_invstd; p.mask = mask; p.gamma = gamma; p.beta = beta; p.aff_w = aff_w; p.aff_b = aff_b;
  p.out = out; p.hi = reinterpret_cast<__nv_bfloat16*>(hi); p.lo = reinterpret_cast<__nv_bfloat16*>(lo);
  p.B = B; p.H = H; p.W = W; p.C = C; p.O = O; p.cpad = cpad; p.relu = relu; p.up = up2 ? 1 : 0;
  const int chunks = (C + kIslaCc - 1) / kIslaCc;
  const long long items = 1LL * H * W * ((min(C, kIslaCc) + 7) / 8);
  long long bx = (items + 255) / 256;
  const long long cap = (148LL * 32 + 1LL * B * chunks - 1) / (1LL * B * chunks);   // ~32 CTAs per SM in flight over the launch
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const size_t smem = sizeof(float) * 2 * O * kIslaCc;
  static DeviceOnce configured;
  if (configured.need()) {
    cudaFuncSetAttribute(isla_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 2 * kIslaCc * 4);
    configured.done();
  }
  dim3 grid(static_cast<int>(bx), chunks, B);
  isla_fwd_kernel<<<grid, 256, smem, stream>>>(p);
  return check_launch("isla_fwd_kernel");
}

// ------------------------------------------------------------------------------------------
// backward.  Three passes:
//   A  pixel-major   : g = relu'(out) * (sum of the 2x2 up-sampled dout)  -> gbuf ; dmask (reduce over c)
//   B  channel-major : dgamma, dbeta (reduce over pixels of one image), sum dxh, sum dxh*xh (per channel)
//   C  elementwise   : dx = (G*g - mean(dxh) - xh*mean(dxh*xh)) * invstd      [train]
//                      dx = G*g*invstd                                         [eval]
// ------------------------------------------------------------------------------------------
struct IslaBwdParams {
  const float* x; const float* mean_invstd; const float* mask; const float* gamma; const float* beta;
  const float* aff_w; const float* aff_b;
  const float* dout;        // [B,H<<up,W<<up,C]
  float* gbuf;              // [B,H,W,C]
  float* dmask;             // [B,H,W,O]
  float* dgamma;            // [B,O,C]  (zero-initialised, atomics)
  float* dbeta;
  double* csum;             // [2C] sum dxh, sum dxh*xh  (zero-initialised);  O == 0: dbias, dweight sums
  float* dx;                // [B,H,W,C]
  double count;
  int B, H, W, C, O, relu, up, train;
};

// pass A: LPP lanes per pixel (32, or 16 when C <= 64), lanes stride the channels (float4 per lane per step)
template <int OM, int LPP>
__global__ void __launch_bounds__(256, OM <= 8 ? 3 : 2) isla_bwd_a_kernel(const IslaBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPP, l = lane % LPP;
  constexpr int PPW = 32 / LPP;
  const int wpb = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const long long pixels = 1LL * p.B * p.H * p.W;
  const int hw = p.H * p.W;
  const int Ho = p.H << p.up, Wo = p.W << p.up;
  const long long pixel_groups = (pixels + PPW - 1) / PPW;
  for (long long pg = 1LL * blockIdx.x * wpb + warp; pg < pixel_groups; pg += 1LL * gridDim.x * wpb) {
    const long long gp = pg * PPW + sub;
    const bool pok = gp < pixels;
    const long long gpc = pok ? gp : pixels - 1;
    const int b = static_cast<int>(gpc / hw);
    const int pix = static_cast<int>(gpc - 1LL * b * hw);
    const int h = pix / p.W, w = pix - h * p.W;
    float S = kMaskEps;
    const float* mp = p.mask ? p.mask + gpc * p.O : nullptr;
    for (int o = 0; o < p.O; ++o) S += __ldg(mp + o);
    const float invS = 1.0f / S;
    float dm[OM];
    float common = 0.f;
#pragma unroll
    for (int o = 0; o < OM; ++o) dm[o] = 0.f;
    for (int c = l * 4; c < p.C; c += LPP * 4) {
      const float4 xv4 = __ldg(reinterpret_cast<const float4*>(p.x + gpc * p.C + c));
      const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
      float dsum[4] = {0, 0, 0, 0};
      const int rep = 1 << p.up;
      for (int dy = 0; dy < rep; ++dy)
        for (int dx = 0; dx < rep; ++dx) {
          const size_t op = ((static_cast<size_t>(b) * Ho + (h << p.up) + dy) * Wo + (w << p.up) + dx) * p.C + c;
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(p.dout + op));
          dsum[0] += d4.x; dsum[1] += d4.y; dsum[2] += d4.z; dsum[3] += d4.w;
        }
      float G[4] = {0, 0, 0, 0}, Bt[4] = {0, 0, 0, 0};
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        if (o < p.O) {
          const float m = __ldg(mp + o);
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + (static_cast<size_t>(b) * p.O + o) * p.C + c));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + (static_cast<size_t>(b) * p.O + o) * p.C + c));
          G[0] = fmaf(m, g4.x, G[0]); G[1] = fmaf(m, g4.y, G[1]); G[2] = fmaf(m, g4.z, G[2]); G[3] = fmaf(m, g4.w, G[3]);
          Bt[0] = fmaf(m, b4.x, Bt[0]); Bt[1] = fmaf(m, b4.y, Bt[1]); Bt[2] = fmaf(m, b4.z, Bt[2]); Bt[3] = fmaf(m, b4.w, Bt[3]);
        }
      }
      // d m_o * S = sum_c g xh gamma_oc + sum_c g beta_oc - sum_c g (xh (Gamma - 1) + B), and
      // xh (Gamma - 1) + B = out - xh: the subtracted term is the same for every object (`common`)
      float gv[4], gx[4], dxh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xv[j] - __ldg(p.mean_invstd + c + j)) * __ldg(p.mean_invstd + p.C + c + j);
        float outv;
        if (p.O > 0) {
          outv = (G[j] * invS + 1.0f) * xh + Bt[j] * invS;
        } else {
          outv = xh;
          if (p.aff_w) outv = outv * __ldg(p.aff_w + c + j) + __ldg(p.aff_b + c + j);
        }
        gv[j] = (p.relu && !(outv > 0.f)) ? 0.f : dsum[j];
        gx[j] = gv[j] * xh;
        common = fmaf(gv[j], outv - xh, common);
        // d xh = Gamma * g (ISLA) or aff_w * g (affine BN): parked in dx for passes B and C
        dxh[j] = gv[j] * ((p.O > 0) ? (G[j] * invS + 1.0f) : (p.aff_w ? __ldg(p.aff_w + c + j) : 1.0f));
      }
      if (pok) {
        *reinterpret_cast<float4*>(p.gbuf + gp * p.C + c) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        *reinterpret_cast<float4*>(p.dx + gp * p.C + c) = make_float4(dxh[0], dxh[1], dxh[2], dxh[3]);
      }
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        if (o < p.O) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + (static_cast<size_t>(b) * p.O + o) * p.C + c));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta + (static_cast<size_t>(b) * p.O + o) * p.C + c));
          const float go[4] = {g4.x, g4.y, g4.z, g4.w}, bo[4] = {b4.x, b4.y, b4.z, b4.w};
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) acc = fmaf(gx[j], go[j], fmaf(gv[j], bo[j], acc));
          dm[o] += acc;
        }
      }
    }
    if (p.O > 0) {
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        if (o < p.O) {
          float v = dm[o] - common;
#pragma unroll
          for (int off = LPP / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (l == 0 && pok) p.dmask[gp * p.O + o] = v * invS;
        }
      }
    }
  }
}

// pass B: block = (pixel segment, channel chunk, image); thread <-> channel, loops over pixels
template <int OM>
__global__ void __launch_bounds__(128) isla_bwd_b_kernel(const IslaBwdParams p) {
  const int b = blockIdx.z;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const int hw = p.H * p.W;
  const int seg = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * seg, p1 = min(hw, p0 + seg);
  float accg[OM], accb[OM];
#pragma unroll
  for (int o = 0; o < OM; ++o) { accg[o] = 0.f; accb[o] = 0.f; }
  const float mean = __ldg(p.mean_invstd + c), invstd = __ldg(p.mean_invstd + p.C + c);
  const float aw = (p.O == 0 && p.aff_w) ? __ldg(p.aff_w + c) : 1.0f;
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0, d2 = 0;
#pragma unroll 4
  for (int pix = p0; pix < p1; ++pix) {
    const size_t gp = static_cast<size_t>(b) * hw + pix;
    const float g = __ldg(p.gbuf + gp * p.C + c);
    const float xh = (__ldg(p.x + gp * p.C + c) - mean) * invstd;
    if (p.O > 0) {
      float m[OM];
      float S = kMaskEps;
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        m[o] = (o < p.O) ? __ldg(p.mask + gp * p.O + o) : 0.f;   // warp-uniform address: one broadcast load
        S += m[o];
      }
      const float invS = 1.0f / S;
      const float gx = g * xh * invS, gs = g * invS;
#pragma unroll
      for (int o = 0; o < OM; ++o) { accg[o] = fmaf(gx, m[o], accg[o]); accb[o] = fmaf(gs, m[o], accb[o]); }
      const float dxh = __ldg(p.dx + gp * p.C + c);            // Gamma * g, written by pass A
      s1 += dxh; s2 += dxh * xh;
    } else {
      s1 += g; s2 += g * xh;           // dbias, dweight of the affine form
    }
    if (((pix - p0) & 63) == 63) { d1 += s1; d2 += s2; s1 = 0.f; s2 = 0.f; }
  }
  d1 += s1; d2 += s2;
#pragma unroll
  for (int o = 0; o < OM; ++o)
    if (o < p.O) {
      atomicAdd(p.dgamma + (static_cast<size_t>(b) * p.O + o) * p.C + c, accg[o]);
      atomicAdd(p.dbeta + (static_cast<size_t>(b) * p.O + o) * p.C + c, accb[o]);
    }
  atomicAdd(p.csum + 2 * c, d1);
  atomicAdd(p.csum + 2 * c + 1, d2);
  (void)aw;
}

// pass C: dx = (dxh - mean(dxh) - xh * mean(dxh * xh)) * invstd in place (dx holds dxh from pass A).
// thread <-> (pixel, 4 channels); a pure streaming pass.
__global__ void __launch_bounds__(256) isla_bwd_c_kernel(const IslaBwdParams p) {
  const int cg = p.C >> 2;
  const long long items = 1LL * p.B * p.H * p.W * cg;
  for (long long it = 1LL * blockIdx.x * blockDim.x + threadIdx.x; it < items; it += 1LL * gridDim.x * blockDim.x) {
    const int g4 = static_cast<int>(it % cg);
    const long long gp = it / cg;
    const int c = g4 * 4;
    const float4 xv4 = __ldg(reinterpret_cast<const float4*>(p.x + gp * p.C + c));
    const float4 dv4 = *reinterpret_cast<const float4*>(p.dx + gp * p.C + c);
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, dv[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mean = __ldg(p.mean_invstd + c + j), invstd = __ldg(p.mean_invstd + p.C + c + j);
      const float xh = (xv[j] - mean) * invstd;
      float dxh = dv[j];
      if (p.train) {
        double m1 = p.csum[2 * (c + j)], m2 = p.csum[2 * (c + j) + 1];
        if (p.O == 0) {                  // csum holds (sum g, sum g*xh) = (d bias, d weight) of the affine form
          const float scale = p.aff_w ? __ldg(p.aff_w + c + j) : 1.0f;
          m1 *= scale; m2 *= scale;
        }
        dxh = dxh - static_cast<float>(m1 / p.count) - xh * static_cast<float>(m2 / p.count);
      }
      r[j] = dxh * invstd;
    }
    *reinterpret_cast<float4*>(p.dx + gp * p.C + c) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

int isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* dout, int B, int H, int W, int C, int O, int relu,
             int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta, double* csum, float* dx,
             int phase, double count, cudaStream_t stream) {
  if (!x || !mean_invstd || !dout || !gbuf || !csum || !dx || B <= 0 || C <= 0 || (C & 3)) { set_error("isla_bwd: bad arguments (C must be a multiple of 4)"); return L2I_ERR_BAD_ARG; }
  if (O > 0 && (!mask || !gamma || !beta || !dmask || !dgamma || !dbeta)) { set_error("isla_bwd: null ISLA operand"); return L2I_ERR_BAD_ARG; }
  if (O > 48) { set_error("isla_bwd: at most 48 objects per image supported"); return L2I_ERR_UNSUPPORTED; }
  IslaBwdParams p;
  p.x = x; p.mean_invstd = mean_invstd; p.mask = mask; p.gamma = gamma; p.beta = beta; p.aff_w = aff_w; p.aff_b = aff_b;
  p.dout = dout; p.gbuf = gbuf; p.dmask = dmask; p.dgamma = dgamma; p.dbeta = dbeta; p.csum = csum; p.dx = dx;
  p.count = count > 0 ? count : static_cast<double>(B) * H * W;   // > 0: global count of a cross-rank batch norm
  if (phase < 0 || phase > 2) { set_error("isla_bwd: phase must be 0 (all), 1 (reductions) or 2 (dx)"); return L2I_ERR_BAD_ARG; }
  p.B = B; p.H = H; p.W = W; p.C = C; p.O = O; p.relu = relu; p.up = up2 ? 1 : 0; p.train = train;
  const long long pixels = 1LL * B * H * W;
  if (phase == 2) {            // csum has been reduced across ranks by the caller; dx holds Gamma * g from phase 1
    const long long items = pixels * (C >> 2);
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    isla_bwd_c_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p);
    return check_launch("isla_bwd_c_kernel");
  }
  cudaError_t e = cudaMemsetAsync(csum, 0, sizeof(double) * 2 * C, stream);
  if (e == cudaSuccess && O > 0) e = cudaMemsetAsync(dgamma, 0, sizeof(float) * B * O * C, stream);
  if (e == cudaSuccess && O > 0) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * B * O * C, stream);
  if (e != cudaSuccess) { set_error("isla_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  {
    long long blocks = (pixels + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    // lanes per pixel: each lane should own >= 4 float4 channel groups, so that the per-pixel shuffle reduction
    // of the O mask gradients is amortised (C = 64 -> 4 lanes, 8 pixels per warp; C >= 512 -> a whole warp)
    int lpp = C / 16;
    lpp = lpp < 4 ? 4 : (lpp > 32 ? 32 : lpp);
    while (lpp & (lpp - 1)) lpp &= lpp - 1;              // power of two
    blocks = (pixels * lpp / 32 + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    const int nb = static_cast<int>(blocks);
#define L2I_ISLA_A(OM)                                                                   \
    switch (lpp) {                                                                       \
      case 4: isla_bwd_a_kernel<OM, 4><<<nb, 256, 0, stream>>>(p); break;                \
      case 8: isla_bwd_a_kernel<OM, 8><<<nb, 256, 0, stream>>>(p); break;                \
      case 16: isla_bwd_a_kernel<OM, 16><<<nb, 256, 0, stream>>>(p); break;              \
      default: isla_bwd_a_kernel<OM, 32><<<nb, 256, 0, stream>>>(p); break;              \
    }
    if (O <= 8) { L2I_ISLA_A(8) } else if (O <= 16) { L2I_ISLA_A(16) } else { L2I_ISLA_A(48) }
#undef L2I_ISLA_A
    int rc = check_launch("isla_bwd_a_kernel");
    if (rc) return rc;
  }
  {
    const int threads = C <= 64 ? 64 : 128;
    const int chunks = (C + threads - 1) / threads;
    const long long base = 1LL * B * chunks * threads;
    int segs = static_cast<int>((148LL * 1024 + base - 1) / base);
    const int hw = H * W;
    if (segs > (hw + 15) / 16) segs = (hw + 15) / 16;
    if (segs < 1) segs = 1;
    dim3 grid(segs, chunks, B);
    if (O <= 8) isla_bwd_b_kernel<8><<<grid, threads, 0, stream>>>(p);
    else if (O <= 16) isla_bwd_b_kernel<16><<<grid, threads, 0, stream>>>(p);
    else isla_bwd_b_kernel<48><<<grid, threads, 0, stream>>>(p);
    int rc = check_launch("isla_bwd_b_kernel");
    if (rc) return rc;
  }
  if (phase == 1) return L2I_OK;
  {
    const long long items = pixels * (C >> 2);
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    isla_bwd_c_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(p);
    return check_launch("isla_bwd_c_kernel");
  }
}

}  // namespace l2i
