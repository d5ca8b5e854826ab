// Internal (C++) interface of the tensor-core convolution kernels; the C ABI lives in api.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace l2i {

struct ConvFwdParams {          // device-side view
  int N, H, W, cin_pad, cout, taps;
  int TW, TH, TN, tiles_w, tiles_h, kchunks;
  int n_pass, pass_len;         // K loop = n_pass passes of pass_len (tap, chunk) iterations, one TMEM buffer each
  int m_tiles, n_tiles;         // persistent tile list: tile t -> (t / n_tiles, t % n_tiles)
  const __nv_bfloat16* mask_hi; // [N,H,W,mask_cpad]: output element is zeroed where mask_hi <= 0 (ReLU backward), or null
  int mask_cpad;
  int pool;                     // 0 none, 1 = 2x2 average, 2 = 2x2 sum of the conv output (stored at H/2 x W/2)
  float res_scale;
  int pair_maps;                // 1: tm_a_hi / tm_b_hi are (hi, lo) pair maps (one TMA instruction per operand tile)
  int ncat;                     // 1: a_hi x [b_hi ; b_lo] as one N = 2 BN instruction (2 MMAs per k16 step instead of 3)
  int sched_slot;               // >= 0: tiles are handed out by a device-wide counter (slot of g_tile_sched); -1: static stride
  const float* bias;            // [cout] or null
  const float* residual;        // [N,H,W,cout] (res_shift=0) or [N,H/2,W/2,cout] nearest-x2 (res_shift=1), or null
  int res_shift;
  float* out;                   // [N,H,W,cout] fp32 or null
  __nv_bfloat16* out_hi;        // [N,H,W,cout_pad] split copy (optionally ReLU'd) or null
  __nv_bfloat16* out_lo;
  int cout_pad, relu_split;
  float out_scale;
};

struct ConvFwdArgs {            // host-side call
  int N, H, W, cin_pad, cout, taps;
  const void *x_hi, *x_lo;      // [N,H,W,cin_pad] bf16
  const void *w_hi, *w_lo;      // [cout][taps][cin_pad] bf16
  const float* bias;
  const float* residual;
  int res_shift;
  float* out;
  void *out_hi, *out_lo;
  int cout_pad, relu_split;
  float out_scale;
  const void* mask_hi;
  int mask_cpad, pool;
  float res_scale;
};

struct ConvWgradParams {
  int N, H, W, cin, cout, taps;
  int TW, TH, TN, tiles_w, tiles_h, pix_blocks, blocks_per_split, cin_tiles, atomic, tap_pairs, ncat;
  float* dw;                    // [cout][taps][cin] fp32
};

struct ConvWgradArgs {
  int N, H, W, cin, cin_pad, cout, cout_pad, taps;
  const void *dy_hi, *dy_lo;    // [N,H,W,cout_pad] bf16
  const void *x_hi, *x_lo;      // [N,H,W,cin_pad] bf16
  float* dw;
};

int conv_fwd_tc(const ConvFwdArgs& a, cudaStream_t stream);
int conv_wgrad_tc(const ConvWgradArgs& a, cudaStream_t stream);

}  // namespace l2i
