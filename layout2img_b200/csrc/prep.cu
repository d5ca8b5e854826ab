// Operand preparation for the tensor-core convolutions: fp32 -> (hi, lo) bf16 pairs.
//  * weights: torch layout [Cout][Cin][kh][kw] -> [Cout][tap][CinPad] (forward, K-major) and
//    [Cin][tap'][CoutPad] with tap' = taps-1-tap (data gradient = correlation with the flipped,
//    transposed filter), optionally divided by the spectral norm sigma read from device memory
//    (nn.utils.spectral_norm's W/sigma; reference resnet_generator_app_v2.py:681-686).
//  * activations: NHWC fp32 -> NHWC bf16 pair with fused ReLU and nearest x2 up-sampling
//    (ResBlock.residual, resnet_generator_app_v2.py:653-663; D blocks rcnn_discriminator_app.py:327-334).
#include "common.cuh"
#include "kernels.h"

namespace l2i {

__global__ void weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ sigma, int cout, int cin,
                                   int taps, __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                                   int cin_pad, __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo,
                                   int cout_pad) {
  const float inv = sigma ? 1.0f / __ldg(sigma) : 1.0f;
  const long long n_f = 1LL * cout * taps * cin_pad;
  const long long n_d = d_hi ? 1LL * cin * taps * cout_pad : 0;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n_f + n_d; i += 1LL * gridDim.x * blockDim.x) {
    if (i < n_f) {
      const int ci = static_cast<int>(i % cin_pad);
      const int tap = static_cast<int>((i / cin_pad) % taps);
      const int co = static_cast<int>(i / (1LL * cin_pad * taps));
      float v = 0.f;
      if (ci < cin) v = __ldg(w + (1LL * co * cin + ci) * taps + tap) * inv;
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      f_hi[i] = h;
      f_lo[i] = l;
    } else {
      const long long j = i - n_f;
      const int co = static_cast<int>(j % cout_pad);
      const int tap = static_cast<int>((j / cout_pad) % taps);
      const int ci = static_cast<int>(j / (1LL * cout_pad * taps));
      float v = 0.f;
      if (co < cout) v = __ldg(w + (1LL * co * cin + ci) * taps + (taps - 1 - tap)) * inv;
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      d_hi[j] = h;
      d_lo[j] = l;
    }
  }
}

int weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* f_hi, void* f_lo, int cin_pad,
                void* d_hi, void* d_lo, int cout_pad, cudaStream_t stream) {
  if (!w || !f_hi || !f_lo || cout <= 0 || cin <= 0 || (taps != 1 && taps != 9) || cin_pad < cin || cin_pad % 8) {
    set_error("weight_prep: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  if (d_hi && (!d_lo || cout_pad < cout || cout_pad % 8)) { set_error("weight_prep: bad dgrad arguments"); return L2I_ERR_BAD_ARG; }
  const long long total = 1LL * cout * taps * cin_pad + (d_hi ? 1LL * cin * taps * cout_pad : 0);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  weight_prep_kernel<<<static_cast<int>(blocks), threads, 0, stream>>>(
      w, sigma, cout, cin, taps, reinterpret_cast<__nv_bfloat16*>(f_hi), reinterpret_cast<__nv_bfloat16*>(f_lo), cin_pad,
      reinterpret_cast<__nv_bfloat16*>(d_hi), reinterpret_cast<__nv_bfloat16*>(d_lo), cout_pad);
  return check_launch("weight_prep_kernel");
}

// one thread per 8 output channels of one output pixel: 16-byte stores to both halves
__global__ void act_split_kernel(const float* __restrict__ x, int N, int H, int W, int C, int relu, int up,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int cpad) {
  const int Ho = H << up, Wo = W << up;
  const int groups = cpad >> 3;
  const long long total = 1LL * N * Ho * Wo * groups;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long pix = i / groups;
    const int wo = static_cast<int>(pix % Wo);
    const int ho = static_cast<int>((pix / Wo) % Ho);
    const int n = static_cast<int>(pix / (1LL * Wo * Ho));
    const float* src = x + ((1LL * n * H + (ho >> up)) * W + (wo >> up)) * C + g * 8;
    float v[8];
    if ((C & 3) == 0 && g * 8 + 8 <= C) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (g * 8 + j < C) ? __ldg(src + j) : 0.f;
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a = v[j], b = v[j + 1];
      if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(a, ah, al);
      split_bf16(b, bh, bl);
      ph[j >> 1] = pack_bf16x2(ah, bh);
      pl[j >> 1] = pack_bf16x2(al, bl);
    }
    *reinterpret_cast<uint4*>(hi + pix * cpad + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo + pix * cpad + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

int act_split(const float* x, int N, int H, int W, int C, int relu, int up2, void* hi, void* lo, int cpad,
              cudaStream_t stream) {
  if (!x || !hi || !lo || N <= 0 || H <= 0 || W <= 0 || C <= 0 || cpad < C || cpad % 8) {
    set_error("act_split: bad arguments (N=%d H=%d W=%d C=%d cpad=%d)", N, H, W, C, cpad);
    return L2I_ERR_BAD_ARG;
  }
  const int up = up2 ? 1 : 0;
  const long long total = 1LL * N * (H << up) * (W << up) * (cpad >> 3);
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 32) blocks = 148 * 32;
  act_split_kernel<<<static_cast<int>(blocks), threads, 0, stream>>>(x, N, H, W, C, relu, up, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                    reinterpret_cast<__nv_bfloat16*>(lo), cpad);
  return check_launch("act_split_kernel");
}

}  // namespace l2i
