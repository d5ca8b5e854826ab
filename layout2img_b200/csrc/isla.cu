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
// ISLA apply and its backward.  One layout for all three kernels: a WARP owns 32 * CPT consecutive channels of one
// pixel (CPT = 1 or 2 channels per lane: 128- / 256-byte coalesced loads, bf16x2 stores), a block is
// warps_c (channel direction) x warps_p (pixel lanes) warps and walks the pixels of ONE image; each lane keeps the
// image's gamma / beta of its channels in registers, the O mask values of a pixel are warp-uniform (broadcast) loads.
//   forward   y = (G/S + 1) xh + Bt/S -> [fp32] and/or ReLU'd bf16 pair (nearest x2 optional): x read once
//   backward  2 passes, nothing but the results is written:
//     reduce  g = relu'(y) dout (y recomputed from x with the forward's exact instruction sequence);
//             dgamma, dbeta (sum over pixels), sum dxh, sum dxh*xh (per channel), dmask (sum over channels: a
//             transposed warp butterfly turns 32 per-lane partials into one total per lane in 31 shuffles)
//     dx      dx = (dxh - mean(dxh) - xh mean(dxh xh)) invstd   [train]   |   dxh invstd   [eval]
//   traffic per element: forward 4 B + pair; backward 2 x (x + dout) + dx = 20 B (the 3-pass form moved 40 B).
// O == 0 : plain / affine batch norm  y = xh * aff_w + aff_b   (final.0, mask heads)
// ------------------------------------------------------------------------------------------
struct IslaFwdParams {
  const float* x;           // [B,H,W,C]
  const float* mean_invstd; // [2C]
  const float* mask;        // [B,H,W,O] pixel-major, or null when O == 0
  const float* gamma;       // [B,O,C]
  const float* beta;        // [B,O,C]
  const float* aff_w;       // [C] or null
  const float* aff_b;       // [C] or null
  const float* chan_scale;  // [B,C] or null: O == 0 only, y = (xh * aff_w + aff_b) * chan_scale (Dropout2d keep-mask)
  float* out;               // [B,H,W,C] fp32 (ReLU'd when relu & 2) or null
  __nv_bfloat16* hi;        // [B,H<<up,W<<up,cpad] or null
  __nv_bfloat16* lo;
  int B, H, W, C, O, cpad, relu, up;
  int warps_c, warps_p;
};

struct IslaBwdParams {
  const float* x; const float* mean_invstd; const float* mask; const float* gamma; const float* beta;
  const float* aff_w; const float* aff_b; const float* chan_scale;
  const float* dout;        // [B,H<<up,W<<up,C]
  float* dmask;             // [B,H,W,O]  (zero-initialised when the channels span several blocks)
  float* dgamma;            // [B,O,C]  (zero-initialised, atomics)
  float* dbeta;
  double* csum;             // [2C] sum dxh, sum dxh*xh  (zero-initialised);  O == 0: dbias, dweight sums
  float* dx;                // [B,H,W,C]
  double count;
  int B, H, W, C, O, relu, up, train;
  int warps_c, warps_p, seg;   // seg: pixels per block (reduce pass)
};

// the one expression of the modulation; every kernel below must produce bit-identical y for the ReLU mask
__device__ __forceinline__ float isla_y(float G, float Bt, float invS, float xh) {
  return fmaf(fmaf(G, invS, 1.0f), xh, Bt * invS);
}

// per-lane constants of the lane's CPT channels for image b
template <int OM, int CPT>
struct IslaLane {
  static constexpr int OMX = OM > 0 ? OM : 1;
  float mean[CPT], invstd[CPT], gam[OMX][CPT], bet[OMX][CPT], aw[CPT], ab[CPT], ks[CPT];
  __device__ __forceinline__ void load(const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ aff_w,
                                       const float* __restrict__ aff_b, const float* __restrict__ chan_scale, int b, int c,
                                       int C, int O) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const bool ok = c + j < C;
      mean[j] = ok ? __ldg(mean_invstd + c + j) : 0.f;
      invstd[j] = ok ? __ldg(mean_invstd + C + c + j) : 0.f;
      aw[j] = (ok && aff_w) ? __ldg(aff_w + c + j) : 1.f;
      ab[j] = (ok && aff_b) ? __ldg(aff_b + c + j) : 0.f;
      ks[j] = (ok && chan_scale) ? __ldg(chan_scale + static_cast<size_t>(b) * C + c + j) : 1.f;
#pragma unroll
      for (int o = 0; o < OMX; ++o) {
        const bool oo = OM > 0 && ok && o < O;
        gam[o][j] = oo ? __ldg(gamma + (static_cast<size_t>(b) * O + o) * C + c + j) : 0.f;
        bet[o][j] = oo ? __ldg(beta + (static_cast<size_t>(b) * O + o) * C + c + j) : 0.f;
      }
    }
  }
};

// the O mask values of one pixel (warp-uniform address) and 1 / (sum + 1e-6)
template <int OM>
__device__ __forceinline__ float isla_masks(const float* __restrict__ mp, int O, float (&m)[OM > 0 ? OM : 1]) {
  float S = kMaskEps;
  if constexpr (OM > 0) {
    if ((O & 3) == 0) {
#pragma unroll
      for (int o = 0; o < OM; o += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < O) v = __ldg(reinterpret_cast<const float4*>(mp + o));
        m[o] = v.x; m[o + 1] = v.y; m[o + 2] = v.z; m[o + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int o = 0; o < OM; ++o) m[o] = (o < O) ? __ldg(mp + o) : 0.f;
    }
#pragma unroll
    for (int o = 0; o < OM; ++o) S += m[o];
  }
  return __frcp_rn(S);
}

template <int CPT>
__device__ __forceinline__ void load_cpt(const float* __restrict__ p, bool ok, float (&v)[CPT]) {
  if constexpr (CPT == 2) {
    float2 t = make_float2(0.f, 0.f);
    if (ok) t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
    v[0] = ok ? __ldg(p) : 0.f;
  }
}

// G = sum_o m_o gamma_o, Bt = sum_o m_o beta_o for the lane's channels; objects whose mask is exactly 0 at this pixel are
// skipped (their terms are exactly 0; the masks are box-shaped, so most are) -- the test is warp-uniform.
template <int OM, int CPT>
__device__ __forceinline__ void isla_modulation(const IslaLane<OM, CPT>& L, const float (&m)[OM > 0 ? OM : 1], float (&G)[CPT],
                                                float (&Bt)[CPT]) {
#pragma unroll
  for (int j = 0; j < CPT; ++j) { G[j] = 0.f; Bt[j] = 0.f; }
  if constexpr (OM > 0) {
#pragma unroll
    for (int o = 0; o < OM; ++o) {
      if (m[o] != 0.f) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) { G[j] = fmaf(m[o], L.gam[o][j], G[j]); Bt[j] = fmaf(m[o], L.bet[o][j], Bt[j]); }
      }
    }
  }
}

static constexpr int kIslaU = 4;        // pixels whose loads are in flight together per warp (forward / dx kernels)

template <int OM, int CPT>
__global__ void __launch_bounds__(256, 2) isla_fwd_kernel(const IslaFwdParams p) {
  using Lane = IslaLane<OM, CPT>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = warp % p.warps_c, wp = warp / p.warps_c;
  const int b = blockIdx.z;
  const int c = ((blockIdx.y * p.warps_c + wc) * 32 + lane) * CPT;
  const bool c_ok = c < p.C;                        // C is a multiple of CPT
  const bool c_pad = p.hi && c < p.cpad;            // a pair's padding channels are written as zeros
  Lane L;
  L.load(p.mean_invstd, p.gamma, p.beta, p.aff_w, p.aff_b, p.chan_scale, b, c, p.C, p.O);
  const int hw = p.H * p.W;
  const int Ho = p.H << p.up, Wo = p.W << p.up;
  const int rep = 1 << p.up;
  const float* __restrict__ xb = p.x + static_cast<size_t>(b) * hw * p.C + c;
  const float* __restrict__ mb = p.mask ? p.mask + static_cast<size_t>(b) * hw * p.O : nullptr;
  float* __restrict__ outb = p.out ? p.out + static_cast<size_t>(b) * hw * p.C + c : nullptr;
  __nv_bfloat16* __restrict__ hib = p.hi;
  __nv_bfloat16* __restrict__ lob = p.lo;
  const int stride = gridDim.x * p.warps_p;
  for (int base = blockIdx.x * p.warps_p + wp; base < hw; base += stride * kIslaU) {
    // ---- all loads of kIslaU pixels first (the stores below may alias them as far as the compiler knows)
    float xv[kIslaU][CPT], m[kIslaU][Lane::OMX], invS[kIslaU];
#pragma unroll
    for (int u = 0; u < kIslaU; ++u) {
      const int pix = min(base + u * stride, hw - 1);
      load_cpt<CPT>(xb + static_cast<size_t>(pix) * p.C, c_ok, xv[u]);
      invS[u] = 1.0f;
      if constexpr (OM > 0) invS[u] = isla_masks<OM>(mb + static_cast<size_t>(pix) * p.O, p.O, m[u]);
    }
#pragma unroll
    for (int u = 0; u < kIslaU; ++u) {
      const int pix = base + u * stride;
      if (pix >= hw) break;
      float G[CPT], Bt[CPT], y[CPT];
      isla_modulation<OM, CPT>(L, m[u], G, Bt);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float xh = (xv[u][j] - L.mean[j]) * L.invstd[j];
        if constexpr (OM > 0) y[j] = isla_y(G[j], Bt[j], invS[u], xh); else y[j] = fmaf(xh, L.aw[j], L.ab[j]) * L.ks[j];
        if (!c_ok) y[j] = 0.f;
      }
      if (outb && c_ok) {
        float* op = outb + static_cast<size_t>(pix) * p.C;
        const bool r2 = (p.relu & 2) != 0;
        if constexpr (CPT == 2) *reinterpret_cast<float2*>(op) = make_float2(r2 ? fmaxf(y[0], 0.f) : y[0], r2 ? fmaxf(y[1], 0.f) : y[1]);
        else op[0] = r2 ? fmaxf(y[0], 0.f) : y[0];
      }
      if (c_pad) {
        __nv_bfloat16 h[CPT], l[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) split_bf16((p.relu & 1) ? fmaxf(y[j], 0.f) : y[j], h[j], l[j]);
        if (rep == 1) {                         // same resolution: the output pixel index is the input's
          const size_t op = (static_cast<size_t>(b) * hw + pix) * p.cpad + c;
          if constexpr (CPT == 2) {
            *reinterpret_cast<uint32_t*>(hib + op) = pack_bf16x2(h[0], h[1]);
            *reinterpret_cast<uint32_t*>(lob + op) = pack_bf16x2(l[0], l[1]);
          } else {
            hib[op] = h[0];
            lob[op] = l[0];
          }
        } else {
          const int hh = pix / p.W, ww = pix - hh * p.W;
          const size_t o00 = ((static_cast<size_t>(b) * Ho + 2 * hh) * Wo + 2 * ww) * p.cpad + c;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const size_t op = o00 + (static_cast<size_t>(q >> 1) * Wo + (q & 1)) * p.cpad;
            if constexpr (CPT == 2) {
              *reinterpret_cast<uint32_t*>(hib + op) = pack_bf16x2(h[0], h[1]);
              *reinterpret_cast<uint32_t*>(lob + op) = pack_bf16x2(l[0], l[1]);
            } else {
              hib[op] = h[0];
              lob[op] = l[0];
            }
          }
        }
      }
    }
  }
}

// block shape for a channel extent: warps along the channels (<= 8), pixel lanes, channel chunks (grid.y)
static void isla_shape(int cext, int cpt, int* warps_c, int* warps_p, int* chunks) {
  const int need = (cext + 32 * cpt - 1) / (32 * cpt);
  int wc = need < 8 ? need : 8;
  if (wc > 4 && wc < 8) wc = 4;                   // 5..7 -> 4 warps x 2 chunks (keeps 8 warps per block busy)
  if (wc == 3) wc = 4;
  *warps_c = wc;
  *warps_p = 8 / wc;
  *chunks = (need + wc - 1) / wc;
}

int isla_fwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* chan_scale, int B, int H, int W, int C, int O, float* out,
             void* hi, void* lo, int cpad, int relu, int up2, cudaStream_t stream) {
  if (!x || !mean_invstd || B <= 0 || H <= 0 || W <= 0 || C <= 0 || O < 0 || (!out && !hi)) { set_error("isla_fwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  if (O > 0 && (!mask || !gamma || !beta)) { set_error("isla_fwd: mask/gamma/beta required when O > 0"); return L2I_ERR_BAD_ARG; }
  if (hi && (!lo || cpad % 8 || cpad < C)) { set_error("isla_fwd: bad pair arguments"); return L2I_ERR_BAD_ARG; }
  if (O > 32) { set_error("isla_fwd: at most 32 objects per image supported (got %d)", O); return L2I_ERR_UNSUPPORTED; }
  IslaFwdParams p;
  p.x = x; p.mean_invstd = mean_invstd; p.mask = mask; p.gamma = gamma; p.beta = beta; p.aff_w = aff_w; p.aff_b = aff_b;
  p.chan_scale = (O == 0) ? chan_scale : nullptr;
  p.out = out; p.hi = reinterpret_cast<__nv_bfloat16*>(hi); p.lo = reinterpret_cast<__nv_bfloat16*>(lo);
  p.B = B; p.H = H; p.W = W; p.C = C; p.O = O; p.cpad = cpad; p.relu = relu; p.up = up2 ? 1 : 0;
  const int cpt = ((C & 1) == 0 && O <= 8) ? 2 : 1;      // O > 8: the per-lane gamma / beta registers allow one channel
  int chunks;
  isla_shape(hi ? (cpad > C ? cpad : C) : C, cpt, &p.warps_c, &p.warps_p, &chunks);
  const int hw = H * W;
  long long bx = (hw + p.warps_p * 4 - 1) / (p.warps_p * 4);          // >= 4 pixels per warp
  const long long cap = (148LL * 12 + 1LL * B * chunks - 1) / (1LL * B * chunks);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<int>(bx), chunks, B);
  const int threads = 32 * p.warps_c * p.warps_p;
  if (O == 0) {
    if (cpt == 2) isla_fwd_kernel<0, 2><<<grid, threads, 0, stream>>>(p); else isla_fwd_kernel<0, 1><<<grid, threads, 0, stream>>>(p);
  } else if (O <= 8) {
    if (cpt == 2) isla_fwd_kernel<8, 2><<<grid, threads, 0, stream>>>(p); else isla_fwd_kernel<8, 1><<<grid, threads, 0, stream>>>(p);
  } else if (O <= 16) {
    isla_fwd_kernel<16, 1><<<grid, threads, 0, stream>>>(p);
  } else {
    isla_fwd_kernel<32, 1><<<grid, threads, 0, stream>>>(p);
  }
  return check_launch("isla_fwd_kernel");
}

// ---- backward, pass 1: every reduction ---------------------------------------------------------------------------
template <int OM, int CPT>
struct IslaPix {            // one pixel's operands for a lane: x, dout (2x2 sum when up-sampled) and the O masks
  float xv[CPT], dv[CPT], m[OM > 0 ? OM : 1];
};

template <int OM, int CPT>
__device__ __forceinline__ void isla_load_pix(IslaPix<OM, CPT>& q, const float* __restrict__ xb, const float* __restrict__ db,
                                              const float* __restrict__ mb, int pix, int C, int O, int W, int rep, bool c_ok) {
  load_cpt<CPT>(xb + static_cast<size_t>(pix) * C, c_ok, q.xv);
  if (rep == 1) {
    load_cpt<CPT>(db + static_cast<size_t>(pix) * C, c_ok, q.dv);
  } else {
    const int hh = pix / W, ww = pix - hh * W;
#pragma unroll
    for (int j = 0; j < CPT; ++j) q.dv[j] = 0.f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        float t[CPT];
        load_cpt<CPT>(db + ((static_cast<size_t>(hh) * 2 + dy) * (2 * W) + ww * 2 + dx) * C, c_ok, t);
#pragma unroll
        for (int j = 0; j < CPT; ++j) q.dv[j] += t[j];
      }
  }
  if constexpr (OM > 0) {
    if ((O & 3) == 0) {
#pragma unroll
      for (int o = 0; o < OM; o += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < O) v = __ldg(reinterpret_cast<const float4*>(mb + static_cast<size_t>(pix) * O + o));
        q.m[o] = v.x; q.m[o + 1] = v.y; q.m[o + 2] = v.z; q.m[o + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int o = 0; o < OM; ++o) q.m[o] = (o < O) ? __ldg(mb + static_cast<size_t>(pix) * O + o) : 0.f;
    }
  }
}

template <int OM>
__device__ __forceinline__ float isla_inv_s(const float (&m)[OM > 0 ? OM : 1]) {
  float S = kMaskEps;
  if constexpr (OM > 0) {
#pragma unroll
    for (int o = 0; o < OM; ++o) S += m[o];
  }
  return __frcp_rn(S);
}

template <int OM, int CPT>
__global__ void __launch_bounds__(256, 2) isla_bwd_reduce_kernel(const IslaBwdParams p) {
  using Lane = IslaLane<OM, CPT>;
  constexpr int OMX = Lane::OMX;
  extern __shared__ float sh[];                     // s_dm [seg * O]  then (reused) the block reductions
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = warp % p.warps_c, wp = warp / p.warps_c;
  const int b = blockIdx.z;
  const int cl = (wc * 32 + lane) * CPT;            // channel inside the block's chunk
  const int cb = 32 * CPT * p.warps_c;              // channels per block
  const int c = blockIdx.y * cb + cl;
  const bool c_ok = c < p.C;
  Lane L;
  L.load(p.mean_invstd, p.gamma, p.beta, p.aff_w, p.aff_b, p.chan_scale, b, c, p.C, p.O);
  const int hw = p.H * p.W;
  const int rep = 1 << p.up;
  const int p0 = blockIdx.x * p.seg, p1 = min(hw, p0 + p.seg);
  if constexpr (OM > 0) {
    for (int i = threadIdx.x; i < (p1 - p0) * p.O; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
  }
  const float* __restrict__ xb = p.x + static_cast<size_t>(b) * hw * p.C + c;
  const float* __restrict__ db = p.dout + static_cast<size_t>(b) * hw * rep * rep * p.C + c;
  const float* __restrict__ mb = p.mask ? p.mask + static_cast<size_t>(b) * hw * p.O : nullptr;
  float accg[OMX][CPT], accb[OMX][CPT], s1[CPT], s2[CPT];
  double d1[CPT], d2[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    s1[j] = s2[j] = 0.f; d1[j] = d2[j] = 0.0;
#pragma unroll
    for (int o = 0; o < OMX; ++o) accg[o][j] = accb[o][j] = 0.f;
  }
  int since_flush = 0;
  // the warp walks pixels p0 + wp, p0 + wp + warps_p, ...; the next pixel's operands are loaded before the current
  // one is processed (rolling prefetch)
  IslaPix<OM, CPT> nxt;
  int pix = p0 + wp;
  if (pix < p1) isla_load_pix<OM, CPT>(nxt, xb, db, mb, pix, p.C, p.O, p.W, rep, c_ok);
  for (; pix < p1; pix += p.warps_p) {
    const IslaPix<OM, CPT> cur = nxt;
    if (pix + p.warps_p < p1) isla_load_pix<OM, CPT>(nxt, xb, db, mb, pix + p.warps_p, p.C, p.O, p.W, rep, c_ok);
    const float invS = isla_inv_s<OM>(cur.m);
    float G[CPT], Bt[CPT];
    isla_modulation<OM, CPT>(L, cur.m, G, Bt);
    float g[CPT], gx[CPT], common = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const float xh = (cur.xv[j] - L.mean[j]) * L.invstd[j];
      float y, fac;
      if constexpr (OM > 0) {
        y = isla_y(G[j], Bt[j], invS, xh);
        fac = fmaf(G[j], invS, 1.0f);
      } else {
        y = fmaf(xh, L.aw[j], L.ab[j]) * L.ks[j];
        fac = L.ks[j];                               // csum of the affine form = (sum g ks, sum g ks xh) = (d bias, d weight)
      }
      g[j] = (!c_ok || (p.relu && !(y > 0.f))) ? 0.f : cur.dv[j];
      gx[j] = g[j] * xh;
      const float dxh = g[j] * fac;
      s1[j] += dxh;
      s2[j] = fmaf(dxh, xh, s2[j]);
      // d m_o * S = sum_c (g xh gamma_oc + g beta_oc) - sum_c g (y - xh): the last term is the same for every object
      common = fmaf(g[j], y - xh, common);
    }
    if constexpr (OM > 0) {
      float part[OM];
#pragma unroll
      for (int o = 0; o < OM; ++o) {
        float acc = -common;
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc = fmaf(gx[j], L.gam[o][j], fmaf(g[j], L.bet[o][j], acc));
        part[o] = acc;
        if (cur.m[o] != 0.f) {                       // warp-uniform: objects that do not touch this pixel contribute 0
          const float mo = cur.m[o] * invS;
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            accg[o][j] = fmaf(gx[j], mo, accg[o][j]);
            accb[o][j] = fmaf(g[j], mo, accb[o][j]);
          }
        }
      }
      const float tot = warp_transpose_sum<OM>(part, lane);
      const int o = lane / (32 / OM);
      if ((lane & (32 / OM - 1)) == 0 && o < p.O) atomicAdd(&sh[(pix - p0) * p.O + o], tot * invS);
    }
    if (++since_flush == 64) {                      // fold the fp32 running sums into fp64 every 64 pixels
#pragma unroll
      for (int j = 0; j < CPT; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = s2[j] = 0.f; }
      since_flush = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < CPT; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; }
  __syncthreads();
  // ---- dmask of this block's pixels
  if constexpr (OM > 0) {
    float* dm = p.dmask + (static_cast<size_t>(b) * hw + p0) * p.O;
    if (gridDim.y == 1) {
      for (int i = threadIdx.x; i < (p1 - p0) * p.O; i += blockDim.x) dm[i] = sh[i];
    } else {
      for (int i = threadIdx.x; i < (p1 - p0) * p.O; i += blockDim.x) atomicAdd(dm + i, sh[i]);
    }
    __syncthreads();
  }
  // ---- per-channel results: combine the block's pixel-lane warps in shared memory, one global atomic per value
  double* s_cs = reinterpret_cast<double*>(sh);                       // [2][cb]
  float* s_gb = reinterpret_cast<float*>(s_cs + 2 * cb);              // [2 * O][cb]
  for (int i = threadIdx.x; i < 2 * cb; i += blockDim.x) s_cs[i] = 0.0;
  if constexpr (OM > 0)
    for (int i = threadIdx.x; i < 2 * p.O * cb; i += blockDim.x) s_gb[i] = 0.f;
  __syncthreads();
  if (c_ok) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      atomicAdd(&s_cs[cl + j], d1[j]);
      atomicAdd(&s_cs[cb + cl + j], d2[j]);
      if constexpr (OM > 0) {
#pragma unroll
        for (int o = 0; o < OM; ++o)
          if (o < p.O) {
            atomicAdd(&s_gb[(2 * o) * cb + cl + j], accg[o][j]);
            atomicAdd(&s_gb[(2 * o + 1) * cb + cl + j], accb[o][j]);
          }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cb; i += blockDim.x) {
    const int ch = blockIdx.y * cb + i;
    if (ch >= p.C) continue;
    atomicAdd(p.csum + 2 * ch, s_cs[i]);
    atomicAdd(p.csum + 2 * ch + 1, s_cs[cb + i]);
    if constexpr (OM > 0) {
      for (int o = 0; o < p.O; ++o) {
        atomicAdd(p.dgamma + (static_cast<size_t>(b) * p.O + o) * p.C + ch, s_gb[(2 * o) * cb + i]);
        atomicAdd(p.dbeta + (static_cast<size_t>(b) * p.O + o) * p.C + ch, s_gb[(2 * o + 1) * cb + i]);
      }
    }
  }
}

// ---- backward, pass 2: dx (x and dout are read a second time; g and dxh are recomputed, never stored)
template <int OM, int CPT>
__global__ void __launch_bounds__(256, 2) isla_bwd_dx_kernel(const IslaBwdParams p) {
  using Lane = IslaLane<OM, CPT>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = warp % p.warps_c, wp = warp / p.warps_c;
  const int b = blockIdx.z;
  const int c = ((blockIdx.y * p.warps_c + wc) * 32 + lane) * CPT;
  if (c >= p.C) return;
  Lane L;
  L.load(p.mean_invstd, p.gamma, p.beta, p.aff_w, p.aff_b, p.chan_scale, b, c, p.C, p.O);
  float m1[CPT], m2[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    m1[j] = m2[j] = 0.f;
    if (p.train) {
      double a = p.csum[2 * (c + j)], q = p.csum[2 * (c + j) + 1];
      if (OM == 0) { a *= L.aw[j]; q *= L.aw[j]; }     // csum holds (d bias, d weight); d xh = aff_w * g
      m1[j] = static_cast<float>(a / p.count);
      m2[j] = static_cast<float>(q / p.count);
    }
  }
  const int hw = p.H * p.W;
  const int rep = 1 << p.up;
  const float* __restrict__ xb = p.x + static_cast<size_t>(b) * hw * p.C + c;
  const float* __restrict__ db = p.dout + static_cast<size_t>(b) * hw * rep * rep * p.C + c;
  const float* __restrict__ mb = p.mask ? p.mask + static_cast<size_t>(b) * hw * p.O : nullptr;
  float* __restrict__ ob = p.dx + static_cast<size_t>(b) * hw * p.C + c;
  const int stride = gridDim.x * p.warps_p;
  for (int base = blockIdx.x * p.warps_p + wp; base < hw; base += stride * kIslaU) {
    IslaPix<OM, CPT> q[kIslaU];
#pragma unroll
    for (int u = 0; u < kIslaU; ++u)
      isla_load_pix<OM, CPT>(q[u], xb, db, mb, min(base + u * stride, hw - 1), p.C, p.O, p.W, rep, true);
#pragma unroll
    for (int u = 0; u < kIslaU; ++u) {
      const int pix = base + u * stride;
      if (pix >= hw) break;
      const float invS = isla_inv_s<OM>(q[u].m);
      float G[CPT], Bt[CPT], r[CPT];
      isla_modulation<OM, CPT>(L, q[u].m, G, Bt);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float xh = (q[u].xv[j] - L.mean[j]) * L.invstd[j];
        float y, fac;
        if constexpr (OM > 0) {
          y = isla_y(G[j], Bt[j], invS, xh);
          fac = fmaf(G[j], invS, 1.0f);
        } else {
          y = fmaf(xh, L.aw[j], L.ab[j]) * L.ks[j];
          fac = L.aw[j] * L.ks[j];
        }
        const float g = (p.relu && !(y > 0.f)) ? 0.f : q[u].dv[j];
        float dxh = g * fac;
        if (p.train) dxh = dxh - m1[j] - xh * m2[j];
        r[j] = dxh * L.invstd[j];
      }
      float* op = ob + static_cast<size_t>(pix) * p.C;
      if constexpr (CPT == 2) *reinterpret_cast<float2*>(op) = make_float2(r[0], r[1]); else op[0] = r[0];
    }
  }
}

int isla_bwd(const float* x, const float* mean_invstd, const float* mask, const float* gamma, const float* beta,
             const float* aff_w, const float* aff_b, const float* chan_scale, const float* dout, int B, int H, int W, int C,
             int O, int relu, int up2, int train, float* gbuf, float* dmask, float* dgamma, float* dbeta, double* csum,
             float* dx, int phase, double count, cudaStream_t stream) {
  (void)gbuf;                  // no intermediate tensor is written any more; the argument is kept for ABI stability
  if (!x || !mean_invstd || !dout || !csum || !dx || B <= 0 || C <= 0) { set_error("isla_bwd: bad arguments"); return L2I_ERR_BAD_ARG; }
  if (O > 0 && (!mask || !gamma || !beta || !dmask || !dgamma || !dbeta)) { set_error("isla_bwd: null ISLA operand"); return L2I_ERR_BAD_ARG; }
  if (O > 32) { set_error("isla_bwd: at most 32 objects per image supported"); return L2I_ERR_UNSUPPORTED; }
  if (phase < 0 || phase > 2) { set_error("isla_bwd: phase must be 0 (all), 1 (reductions) or 2 (dx)"); return L2I_ERR_BAD_ARG; }
  IslaBwdParams p;
  p.x = x; p.mean_invstd = mean_invstd; p.mask = mask; p.gamma = gamma; p.beta = beta; p.aff_w = aff_w; p.aff_b = aff_b;
  p.chan_scale = (O == 0) ? chan_scale : nullptr;
  p.dout = dout; p.dmask = dmask; p.dgamma = dgamma; p.dbeta = dbeta; p.csum = csum; p.dx = dx;
  p.count = count > 0 ? count : static_cast<double>(B) * H * W;   // > 0: global count of a cross-rank batch norm
  p.B = B; p.H = H; p.W = W; p.C = C; p.O = O; p.relu = relu; p.up = up2 ? 1 : 0; p.train = train; p.seg = 0;
  const int hw = H * W;
  if (phase != 2) {
    // ---- pass 1
    const int cpt = ((C & 1) == 0 && O <= 8) ? 2 : 1;
    int chunks;
    isla_shape(C, cpt, &p.warps_c, &p.warps_p, &chunks);
    const int cb = 32 * cpt * p.warps_c;
    cudaError_t e = cudaMemsetAsync(csum, 0, sizeof(double) * 2 * C, stream);
    if (e == cudaSuccess && O > 0) e = cudaMemsetAsync(dgamma, 0, sizeof(float) * B * O * C, stream);
    if (e == cudaSuccess && O > 0) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * B * O * C, stream);
    if (e == cudaSuccess && O > 0 && chunks > 1) e = cudaMemsetAsync(dmask, 0, sizeof(float) * B * hw * O, stream);
    if (e != cudaSuccess) { set_error("isla_bwd: memset: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
    // pixels per block: ~4 blocks per SM over the launch, at least 8 butterflies per pixel-lane warp, bounded by the
    // shared-memory dmask accumulator (seg * O floats <= 32 KB)
    const int pixg = O > 0 ? 32 / (O <= 8 ? 8 : (O <= 16 ? 16 : 32)) : 1;
    long long want_blocks = (148LL * 4 + 1LL * B * chunks - 1) / (1LL * B * chunks);
    int seg = static_cast<int>((hw + want_blocks - 1) / want_blocks);
    const int min_seg = p.warps_p * pixg * 8;
    if (seg < min_seg) seg = min_seg;
    const int max_seg = O > 0 ? 8192 / O : 1 << 20;
    if (seg > max_seg) seg = max_seg;
    if (seg > hw) seg = hw;
    p.seg = seg;
    const size_t smem_dm = sizeof(float) * seg * (O > 0 ? O : 0);
    const size_t smem_red = sizeof(double) * 2 * cb + sizeof(float) * 2 * O * cb;
    const size_t smem = smem_dm > smem_red ? smem_dm : smem_red;
    dim3 grid((hw + seg - 1) / seg, chunks, B);
    const int threads = 32 * p.warps_c * p.warps_p;
    static DeviceOnce configured;
    if (configured.need()) {
      cudaFuncSetAttribute(isla_bwd_reduce_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(isla_bwd_reduce_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      configured.done();
    }
    if (O == 0) {
      if (cpt == 2) isla_bwd_reduce_kernel<0, 2><<<grid, threads, smem, stream>>>(p); else isla_bwd_reduce_kernel<0, 1><<<grid, threads, smem, stream>>>(p);
    } else if (O <= 8) {
      if (cpt == 2) isla_bwd_reduce_kernel<8, 2><<<grid, threads, smem, stream>>>(p); else isla_bwd_reduce_kernel<8, 1><<<grid, threads, smem, stream>>>(p);
    } else if (O <= 16) {
      isla_bwd_reduce_kernel<16, 1><<<grid, threads, smem, stream>>>(p);
    } else {
      isla_bwd_reduce_kernel<32, 1><<<grid, threads, smem, stream>>>(p);
    }
    int rc = check_launch("isla_bwd_reduce_kernel");
    if (rc) return rc;
    if (phase == 1) return L2I_OK;
  }
  {
    // ---- pass 2 (phase 2: csum has been reduced across ranks by the caller)
    const int cpt = ((C & 1) == 0 && O <= 8) ? 2 : 1;      // O > 8: the per-lane gamma / beta registers allow one channel
    int chunks;
    isla_shape(C, cpt, &p.warps_c, &p.warps_p, &chunks);
    long long bx = (hw + p.warps_p * 4 - 1) / (p.warps_p * 4);
    const long long cap = (148LL * 12 + 1LL * B * chunks - 1) / (1LL * B * chunks);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(static_cast<int>(bx), chunks, B);
    const int threads = 32 * p.warps_c * p.warps_p;
    if (O == 0) {
      if (cpt == 2) isla_bwd_dx_kernel<0, 2><<<grid, threads, 0, stream>>>(p); else isla_bwd_dx_kernel<0, 1><<<grid, threads, 0, stream>>>(p);
    } else if (O <= 8) {
      if (cpt == 2) isla_bwd_dx_kernel<8, 2><<<grid, threads, 0, stream>>>(p); else isla_bwd_dx_kernel<8, 1><<<grid, threads, 0, stream>>>(p);
    } else if (O <= 16) {
      isla_bwd_dx_kernel<16, 1><<<grid, threads, 0, stream>>>(p);
    } else {
      isla_bwd_dx_kernel<32, 1><<<grid, threads, 0, stream>>>(p);
    }
    return check_launch("isla_bwd_dx_kernel");
  }
}

}  // namespace l2i
