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

// Tile = 32 output channels x 32 input channels x all taps, staged in shared memory: the torch-layout rows
// are read coalesced (32 * taps contiguous floats per output channel) and both operand layouts are written
// in 64-byte runs (32 consecutive bf16 along Cin for the forward operand, along Cout for the dgrad operand).
static constexpr int kWpTile = 32;
template <int taps>
__device__ __forceinline__ void weight_prep_body(const float* __restrict__ w, const float* __restrict__ sigma, int cout, int cin,
                                                 __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo, int cin_pad,
                                                 __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, int cout_pad,
                                                 int tile_x, int tile_y, float* tile) {
  const int pitch = kWpTile * taps + 1;
  const int ci0 = tile_x * kWpTile, co0 = tile_y * kWpTile;
  const float inv = sigma ? 1.0f / __ldg(sigma) : 1.0f;
  const int nci = max(0, min(kWpTile, cin - ci0));        // valid input channels in this tile
  for (int i = threadIdx.x; i < kWpTile * kWpTile * taps; i += blockDim.x) {
    const int co_l = i / (kWpTile * taps), rem = i - co_l * (kWpTile * taps);
    float v = 0.f;
    if (co0 + co_l < cout && rem < nci * taps) v = __ldg(w + (static_cast<size_t>(co0 + co_l) * cin + ci0) * taps + rem) * inv;
    tile[co_l * pitch + rem] = v;                         // rem = ci_l * taps + tap
  }
  __syncthreads();
  // forward operand [cout][tap][cin_pad]: pairs of consecutive input channels per thread
  for (int i = threadIdx.x; i < kWpTile * taps * (kWpTile / 2); i += blockDim.x) {
    const int cp = i % (kWpTile / 2);
    const int tap = (i / (kWpTile / 2)) % taps;
    const int co_l = i / ((kWpTile / 2) * taps);
    const int co = co0 + co_l, ci = ci0 + 2 * cp;
    if (co >= cout || ci >= cin_pad) continue;
    __nv_bfloat16 ah, al, bh, bl;
    split_bf16(tile[co_l * pitch + (2 * cp) * taps + tap], ah, al);
    split_bf16(tile[co_l * pitch + (2 * cp + 1) * taps + tap], bh, bl);
    const size_t o = (static_cast<size_t>(co) * taps + tap) * cin_pad + ci;
    *reinterpret_cast<uint32_t*>(f_hi + o) = pack_bf16x2(ah, bh);
    *reinterpret_cast<uint32_t*>(f_lo + o) = pack_bf16x2(al, bl);
  }
  if (d_hi) {
    // data-gradient operand [cin][taps-1-tap][cout_pad]: pairs of consecutive output channels per thread
    for (int i = threadIdx.x; i < kWpTile * taps * (kWpTile / 2); i += blockDim.x) {
      const int cp = i % (kWpTile / 2);
      const int tap = (i / (kWpTile / 2)) % taps;
      const int ci_l = i / ((kWpTile / 2) * taps);
      const int ci = ci0 + ci_l, co = co0 + 2 * cp;
      if (ci >= cin || co >= cout_pad) continue;
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(tile[(2 * cp) * pitch + ci_l * taps + tap], ah, al);
      split_bf16(tile[(2 * cp + 1) * pitch + ci_l * taps + tap], bh, bl);
      const size_t o = (static_cast<size_t>(ci) * taps + (taps - 1 - tap)) * cout_pad + co;
      *reinterpret_cast<uint32_t*>(d_hi + o) = pack_bf16x2(ah, bh);
      *reinterpret_cast<uint32_t*>(d_lo + o) = pack_bf16x2(al, bl);
    }
  }
}
template <int taps>
__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ sigma, int cout, int cin,
                   __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo, int cin_pad,
                   __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo, int cout_pad) {
  extern __shared__ float tile[];                         // [32][32 * taps + 1]
  weight_prep_body<taps>(w, sigma, cout, cin, f_hi, f_lo, cin_pad, d_hi, d_lo, cout_pad, blockIdx.x, blockIdx.y, tile);
}

// Channel padding of an operand pair (must match layout2img_b200/ops.py pad8): a multiple of 8; counts <= 8 are padded
// to a full 64-wide K chunk and tails <= 40 of a 64-chunk are rounded up (TMA boxes that are mostly out of bounds in the
// channel dimension are slow).
__host__ __device__ inline int wp_pad8(int c) {
  if (c <= 8) return 64;
  const int c8 = (c + 7) / 8 * 8, tail = c8 % 64;
  return (tail > 0 && tail <= 40) ? c8 + (64 - tail) : c8;
}
// Elements of one half (hi or lo) of the forward / data-gradient operand, rounded to 64 elements (128 B) so that every
// slice of the per-call bf16 buffer satisfies TMA's global-address alignment.
__host__ __device__ inline long long wp_fwd_elems(int cout, int cin, int taps) { return ((1LL * cout * taps * wp_pad8(cin)) + 63) & ~63LL; }
__host__ __device__ inline long long wp_dg_elems(int cout, int cin, int taps) { return ((1LL * cin * taps * wp_pad8(cout)) + 63) & ~63LL; }

// Grouped form: the operand pairs of EVERY convolution weight of a network call in one launch per tap count.  items:
// (module, tile_x, tile_y).  The module's slice of the bf16 buffer holds f_hi, f_lo, then (want_dgrad) d_hi, d_lo.
template <int taps>
__global__ void __launch_bounds__(256)
weight_prep_group_kernel(const SnEntry* __restrict__ tab, const int4* __restrict__ items, const float* __restrict__ f32,
                         __nv_bfloat16* __restrict__ bf, int want_dgrad) {
  extern __shared__ float tile[];
  const int4 it = items[blockIdx.x];
  const SnEntry e = tab[it.x];
  const int cout = e.R, cin = e.cin;
  const long long nf = wp_fwd_elems(cout, cin, taps), nd = wp_dg_elems(cout, cin, taps);
  __nv_bfloat16* f_hi = bf + e.bf_off;
  __nv_bfloat16* f_lo = f_hi + nf;
  __nv_bfloat16* d_hi = want_dgrad ? f_lo + nf : nullptr;
  __nv_bfloat16* d_lo = want_dgrad ? d_hi + nd : nullptr;
  weight_prep_body<taps>(e.W, e.has_sn ? f32 + e.f32_off : nullptr, cout, cin, f_hi, f_lo, wp_pad8(cin), d_hi, d_lo,
                         wp_pad8(cout), it.y, it.z, tile);
}

int weight_prep_group(const void* table, const int* items9, int n9, const int* items1, int n1, const float* f32, void* bf16,
                      int want_dgrad, cudaStream_t stream) {
  if (!table || !bf16 || (n9 > 0 && !items9) || (n1 > 0 && !items1)) { set_error("weight_prep_group: bad arguments"); return L2I_ERR_BAD_ARG; }
  const SnEntry* tab = reinterpret_cast<const SnEntry*>(table);
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(bf16);
  int rc;
  if (n9 > 0) {
    weight_prep_group_kernel<9><<<n9, 256, sizeof(float) * kWpTile * (kWpTile * 9 + 1), stream>>>(
        tab, reinterpret_cast<const int4*>(items9), f32, bf, want_dgrad);
    if ((rc = check_launch("weight_prep_group_kernel<9>"))) return rc;
  }
  if (n1 > 0) {
    weight_prep_group_kernel<1><<<n1, 256, sizeof(float) * kWpTile * (kWpTile * 1 + 1), stream>>>(
        tab, reinterpret_cast<const int4*>(items1), f32, bf, want_dgrad);
    if ((rc = check_launch("weight_prep_group_kernel<1>"))) return rc;
  }
  return L2I_OK;
}

int weight_prep(const float* w, const float* sigma, int cout, int cin, int taps, void* f_hi, void* f_lo, int cin_pad,
                void* d_hi, void* d_lo, int cout_pad, cudaStream_t stream) {
  if (!w || !f_hi || !f_lo || cout <= 0 || cin <= 0 || (taps != 1 && taps != 9) || cin_pad < cin || cin_pad % 8) {
    set_error("weight_prep: bad arguments");
    return L2I_ERR_BAD_ARG;
  }
  if (d_hi && (!d_lo || cout_pad < cout || cout_pad % 8)) { set_error("weight_prep: bad dgrad arguments"); return L2I_ERR_BAD_ARG; }
  // tiles cover the padded extents so that the padding channels of both operands are written (as zeros)
  const int gx = (cin_pad + kWpTile - 1) / kWpTile;
  const int gy = ((d_hi ? cout_pad : cout) + kWpTile - 1) / kWpTile;
  const size_t smem = sizeof(float) * kWpTile * (kWpTile * taps + 1);
  if (taps == 9)
    weight_prep_kernel<9><<<dim3(gx, gy), 256, smem, stream>>>(
        w, sigma, cout, cin, reinterpret_cast<__nv_bfloat16*>(f_hi), reinterpret_cast<__nv_bfloat16*>(f_lo), cin_pad,
        reinterpret_cast<__nv_bfloat16*>(d_hi), reinterpret_cast<__nv_bfloat16*>(d_lo), cout_pad);
  else
    weight_prep_kernel<1><<<dim3(gx, gy), 256, smem, stream>>>(
        w, sigma, cout, cin, reinterpret_cast<__nv_bfloat16*>(f_hi), reinterpret_cast<__nv_bfloat16*>(f_lo), cin_pad,
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

// ------------------------------------------------------------------------------------------------
// Block-level operand preparation for the discriminator / generator residual blocks
// (reference rcnn_discriminator_app.py:294-344): one read of a tensor produces every operand
// pair the block's convolutions need, and the per-channel sums that are the bias gradients.
// ------------------------------------------------------------------------------------------------
namespace l2i {

__device__ __forceinline__ void load8(const float* src, int c0, int C, float (&v)[8]) {
  if ((C & 3) == 0 && c0 + 8 <= C) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? __ldg(src + j) : 0.f;
  }
}

__device__ __forceinline__ void store_pair8(const float (&v)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    __nv_bfloat16 ah, al, bh, bl;
    split_bf16(v[j], ah, al);
    split_bf16(v[j + 1], bh, bl);
    ph[j >> 1] = pack_bf16x2(ah, bh);
    pl[j >> 1] = pack_bf16x2(al, bl);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// a = relu?(x) at full resolution; b = b_scale * x (b_mode 1) or b_scale * (2x2 sum of x) (b_mode 2), never ReLU'd.
// One thread: 8 channels of one pixel (b_mode 0/1) or of one 2x2 quad (b_mode 2).
__global__ void act_split2_kernel(const float* __restrict__ x, int N, int H, int W, int C, int relu_a,
                                  __nv_bfloat16* __restrict__ a_hi, __nv_bfloat16* __restrict__ a_lo, int b_mode,
                                  float b_scale, __nv_bfloat16* __restrict__ b_hi, __nv_bfloat16* __restrict__ b_lo,
                                  int cpad) {
  const int groups = cpad >> 3;
  const int sh = (b_mode == 2) ? 1 : 0;
  const int Hq = H >> sh, Wq = W >> sh;
  const long long total = 1LL * N * Hq * Wq * groups;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long q = i / groups;
    const int wq = static_cast<int>(q % Wq);
    const int hq = static_cast<int>((q / Wq) % Hq);
    const int n = static_cast<int>(q / (1LL * Wq * Hq));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int reps = sh ? 2 : 1;
    for (int dy = 0; dy < reps; ++dy) {
      for (int dx = 0; dx < reps; ++dx) {
        const long long pix = (1LL * n * H + (hq << sh) + dy) * W + (wq << sh) + dx;
        float v[8];
        load8(x + pix * C + g * 8, g * 8, C, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
        if (relu_a) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (a_hi) store_pair8(v, a_hi + pix * cpad + g * 8, a_lo + pix * cpad + g * 8);
      }
    }
    if (b_mode) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= b_scale;
      store_pair8(acc, b_hi + q * cpad + g * 8, b_lo + q * cpad + g * 8);
    }
  }
}

int act_split2(const float* x, int N, int H, int W, int C, int relu_a, void* a_hi, void* a_lo, int b_mode, float b_scale,
               void* b_hi, void* b_lo, int cpad, cudaStream_t stream) {
  if (!x || (!a_hi != !a_lo) || (!a_hi && !b_mode) || N <= 0 || H <= 0 || W <= 0 || C <= 0 || cpad < C || cpad % 8 || b_mode < 0 || b_mode > 2 ||
      (b_mode && (!b_hi || !b_lo)) || (b_mode == 2 && ((H | W) & 1))) {
    set_error("act_split2: bad arguments (N=%d H=%d W=%d C=%d cpad=%d b_mode=%d)", N, H, W, C, cpad, b_mode);
    return L2I_ERR_BAD_ARG;
  }
  const int sh = (b_mode == 2) ? 1 : 0;
  const long long total = 1LL * N * (H >> sh) * (W >> sh) * (cpad >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  act_split2_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      x, N, H, W, C, relu_a, reinterpret_cast<__nv_bfloat16*>(a_hi), reinterpret_cast<__nv_bfloat16*>(a_lo), b_mode,
      b_scale, reinterpret_cast<__nv_bfloat16*>(b_hi), reinterpret_cast<__nv_bfloat16*>(b_lo), cpad);
  return check_launch("act_split2_kernel");
}

// Per-channel sums: thread t of a 256-thread block owns channel group t % groups of pixel slot t / groups.
// MODE 0: g fp32 [P, C] -> pair `lo` at the same resolution, optional pair `up` = up_scale * nearest-x2(g),
//         colsum[c] += sum_p g[p, c].          (gradient arriving at a residual block's output)
// MODE 1: (hi, lo) pair [P, cpad] -> colsum only.   (bias gradient of a conv whose output gradient is a pair)
template <int MODE>
__global__ void __launch_bounds__(256)
colsum_split_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ in_hi,
                    const __nv_bfloat16* __restrict__ in_lo, int N, int H, int W, int C, int cpad,
                    __nv_bfloat16* __restrict__ lo_hi, __nv_bfloat16* __restrict__ lo_lo, float up_scale,
                    __nv_bfloat16* __restrict__ up_hi, __nv_bfloat16* __restrict__ up_lo, float* __restrict__ colsum) {
  extern __shared__ float s_sum[];          // [cpad]
  const int groups = cpad >> 3;
  const int ppb = 256 / groups;
  const int grp = threadIdx.x % groups;
  const int slot = threadIdx.x / groups;
  for (int i = threadIdx.x; i < cpad; i += 256) s_sum[i] = 0.f;
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const long long P = 1LL * N * H * W;
  if (slot < ppb) {
    for (long long pix = 1LL * blockIdx.x * ppb + slot; pix < P; pix += 1LL * gridDim.x * ppb) {
      float v[8];
      if (MODE == 0) {
        load8(g + pix * C + grp * 8, grp * 8, C, v);
        if (lo_hi) store_pair8(v, lo_hi + pix * cpad + grp * 8, lo_lo + pix * cpad + grp * 8);
        if (up_hi) {
          const int w = static_cast<int>(pix % W);
          const int h = static_cast<int>((pix / W) % H);
          const long long n = pix / (1LL * W * H);
          float u[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) u[j] = v[j] * up_scale;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const long long op = (n * (2 * H) + 2 * h + dy) * (2 * W) + 2 * w + dx;
              store_pair8(u, up_hi + op * cpad + grp * 8, up_lo + op * cpad + grp * 8);
            }
        }
      } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(in_hi + pix * cpad + grp * 8));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(in_lo + pix * cpad + grp * 8));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] = __uint_as_float(aw[j] << 16) + __uint_as_float(bw[j] << 16);
          v[2 * j + 1] = __uint_as_float(aw[j] & 0xFFFF0000u) + __uint_as_float(bw[j] & 0xFFFF0000u);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_sum[grp * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) atomicAdd(colsum + c, s_sum[c]);
}

static int colsum_grid(long long P, int ppb) {
  long long blocks = (P + ppb - 1) / ppb;
  // enough pixels per thread that the shared/global atomics are noise, enough blocks to fill the chip
  long long cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

int grad_split(const float* g, int N, int H, int W, int C, void* lo_hi, void* lo_lo, float up_scale, void* up_hi,
               void* up_lo, float* colsum, int cpad, cudaStream_t stream) {
  if (!g || N <= 0 || H <= 0 || W <= 0 || C <= 0 || cpad < C || cpad % 8 || cpad > 2048 || !colsum ||
      (lo_hi && !lo_lo) || (up_hi && !up_lo)) {
    set_error("grad_split: bad arguments (N=%d H=%d W=%d C=%d cpad=%d)", N, H, W, C, cpad);
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * C, stream);
  if (e != cudaSuccess) { set_error("grad_split: memset failed: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  const int ppb = 256 / (cpad >> 3);
  colsum_split_kernel<0><<<colsum_grid(1LL * N * H * W, ppb), 256, cpad * sizeof(float), stream>>>(
      g, nullptr, nullptr, N, H, W, C, cpad, reinterpret_cast<__nv_bfloat16*>(lo_hi),
      reinterpret_cast<__nv_bfloat16*>(lo_lo), up_scale, reinterpret_cast<__nv_bfloat16*>(up_hi),
      reinterpret_cast<__nv_bfloat16*>(up_lo), colsum);
  return check_launch("colsum_split_kernel<0>");
}

int pair_colsum(const void* hi, const void* lo, long long pixels, int C, int cpad, float* colsum, cudaStream_t stream) {
  if (!hi || !lo || pixels <= 0 || C <= 0 || cpad < C || cpad % 8 || cpad > 2048 || !colsum) {
    set_error("pair_colsum: bad arguments (pixels=%lld C=%d cpad=%d)", pixels, C, cpad);
    return L2I_ERR_BAD_ARG;
  }
  cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * C, stream);
  if (e != cudaSuccess) { set_error("pair_colsum: memset failed: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  const int ppb = 256 / (cpad >> 3);
  colsum_split_kernel<1><<<colsum_grid(pixels, ppb), 256, cpad * sizeof(float), stream>>>(
      nullptr, reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo), 1, 1,
      static_cast<int>(pixels), C, cpad, nullptr, nullptr, 1.f, nullptr, nullptr, colsum);
  return check_launch("colsum_split_kernel<1>");
}

}  // namespace l2i
