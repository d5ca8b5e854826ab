// 3x3 (pad 1) and 1x1 convolution as an implicit GEMM on the sm_100a tensor cores.
//
// Replaces the cuDNN calls behind nn.Conv2d on the hot path (reference
// model/resnet_generator_app_v2.py:633-639,659-669 and model/rcnn_discriminator_app.py:297-326).
//
// Numerics: every fp32 operand x is carried as a bf16 pair (hi, lo) with x ~= hi + lo
// (16 significant bits, fp32 exponent range).  A product a*b is accumulated in fp32 TMEM as
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo -- three bf16 tcgen05.mma per K step.  Measured against
// the fp32 reference this is ~60x more accurate than single-pass TF32.
//
// Accumulation: the tensor core adds each MMA result into the fp32 TMEM accumulator with
// truncation (measured: a systematic -1.6e-8 relative bias per non-zero add, i.e. -2.7e-5 after the
// 1728 adds of a K=9216 reduction -- 4x the fp32 reference's own noise).  So (i) the two small cross
// terms go to their own accumulator (their truncation is 2^-8 smaller), and (ii) the K loop is cut
// into passes of at most 144 k16 steps; each pass starts a fresh (main, cross) accumulator pair in
// one of two TMEM buffers, and the epilogue warps add the finished pass into fp32 registers with
// round-to-nearest while the tensor core already runs the next pass (or the next tile).
//
// Data layout: activations NHWC (channel count padded to a multiple of 8), one tensor per
// half of the pair.  Forward / data-gradient:
//   GEMM M = 128 output pixels (a TW x TH x TN patch), N = BN output channels,
//   K = taps x Cin, walked as (tap, 64-channel chunk).  The A tile of a tap is ONE 4-D TMA box
//   at spatially shifted coordinates; the zero padding of the convolution is TMA's
//   out-of-bounds fill.  Weights are pre-arranged [Cout][tap][Cin] (K-major).
// Weight gradient:
//   GEMM M = 128 output channels, N = BN input channels, K = pixels (64 per stage), both
//   operands MN-major straight out of the NHWC tensors; one tap per CTA, split-K over pixels.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, then the epilogue warps
// (TMEM -> registers -> global): 8 in the forward/dgrad kernel (two per TMEM lane quarter, each taking half
// of the tile's columns -- with one warp per SM sub-partition the epilogue, not the tensor core, paced every
// layer with K <= 2304), 4 in the weight-gradient kernel.
#include <stdlib.h>
#include <atomic>
#include "common.cuh"
#include "conv_tc.h"

namespace l2i {

static constexpr int kThreads = 192;       // weight-gradient kernel
static constexpr int kFwdThreads = 320;    // forward / data-gradient kernel: 2 + 8 warps
static constexpr int kBK = 64;          // bf16 elements per smem row (128 B, SWIZZLE_128B)
static constexpr int kTileBytes = 128 * kBK * 2;   // one 128-row operand tile: 16 KB
static constexpr int kPassLen = 36;     // k-iterations (of 4 k16 steps) accumulated inside TMEM before the epilogue takes over

// ------------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------------
template <int BN, bool HALO>
struct FwdCfg {
  // plain mode: one ring; a stage = the (A hi, A lo, B hi, B lo) tiles of one (tap, 64-channel chunk)
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = 2 * kTileBytes + 2 * kBBytes;
  // halo mode (3x3 on an 8 x 16 pixel patch): ring A holds the (16+2) x 16 pixel halo patch of one 64-channel
  // chunk ONCE for all nine taps (a tap's A operand is a shifted window into it); ring B streams the nine
  // weight tiles.  L2 -> SM traffic per chunk: 72 KB + 9 B-tiles instead of 9 x (32 KB + B-tile).
  static constexpr int kHaloRows = 18 * 16;                         // patch rows of 128 B
  static constexpr int kHaloBytes = kHaloRows * 128;                // per half (hi or lo): 36 KB
  static constexpr int kAStages = 2;
  static constexpr int kBStages = (BN == 128) ? 2 : 4;
  static constexpr int kBStageBytes = 2 * kBBytes;
  static constexpr int kRingBytes = HALO ? (kAStages * 2 * kHaloBytes + kBStages * kBStageBytes) : (kStages * kStageBytes);
  static constexpr int kSmem = kRingBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 512;      // whole TMEM: one CTA per SM (shared memory already enforces it)
};

// acc (+)= main + cross for the 32-column chunk at `taddr` (lane = GEMM row); cross tile BN columns later
template <int BN>
__device__ __forceinline__ void add_pass_chunk(uint32_t taddr, bool first, float* acc) {
  uint32_t r[32], x[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_32x32(taddr + BN, x);
  tmem_ld_wait();
  if (first) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]) + __uint_as_float(x[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]) + __uint_as_float(x[j]);
  }
}

// Persistent kernel: gridDim.x CTAs (one per SM) walk the (m_tile, n_tile) list.
// The TMA producer and the MMA issuer run ahead across tile boundaries (the smem ring never drains);
// with two TMEM accumulator buffers the epilogue of tile i overlaps the main loop of tile i+1.
//
// Tile order: the producer lane is the CTA's scheduler.  It takes tile numbers from a device-wide counter (work
// stealing; p.sched_slot >= 0) or with stride gridDim.x (p.sched_slot < 0) and hands them to the MMA issuer and the
// epilogue warps through a small shared-memory queue.  With the counter a CTA that starts late -- its SM was held by
// another kernel, e.g. an NCCL all-reduce overlapping the backward pass -- finds the list already drained by its
// peers and exits, where a static stride would leave its whole share of tiles for a second wave (measured: the
// overlapped all-reduces cost as much as exposed ones).  The last producer to retire zeroes the slot.
constexpr int kSchedSlots = 1024;              // launches take slots round-robin; a slot is busy for one kernel's lifetime
__device__ int g_tile_sched[2 * kSchedSlots];  // [slot] = {next tile, CTAs finished}
constexpr int kSched = 4;                      // depth of the per-CTA tile queue
template <int BN, bool HALO>
__global__ void __launch_bounds__(kFwdThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                const ConvFwdParams p) {
  using Cfg = FwdCfg<BN, HALO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // barriers: plain: full/empty [kStages]; halo: A full/empty [kAStages] then B full/empty [kBStages]
  constexpr int kNBarA = HALO ? Cfg::kAStages : Cfg::kStages;
  constexpr int kNBarB = HALO ? Cfg::kBStages : 0;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes);
  uint64_t* empty_bar = full_bar + kNBarA;
  uint64_t* bfull_bar = empty_bar + kNBarA;
  uint64_t* bempty_bar = bfull_bar + kNBarB;
  uint64_t* tfull_bar = bempty_bar + kNBarB;          // [2] accumulator buffer complete
  uint64_t* tempty_bar = tfull_bar + 2;               // [2] accumulator buffer drained by the epilogue
  uint64_t* sfull_bar = tempty_bar + 2;               // [kSched] tile queue entry written by the producer lane
  uint64_t* sempty_bar = sfull_bar + kSched;          // [kSched] ... read by the MMA issuer and every epilogue thread
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + kSched);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [kSched]
  static_assert((kNBarA + kNBarB) * 2 * 8 + 4 * 8 + 2 * kSched * 8 + 4 + kSched * 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_iters = HALO ? p.kchunks : p.taps * p.kchunks;   // halo: one iteration = one chunk, all nine taps
  const int total_tiles = p.m_tiles * p.n_tiles;

  // the first tile number is requested before the set-up below, which hides the atomic's round trip
  int* const sched = p.sched_slot >= 0 ? g_tile_sched + 2 * p.sched_slot : nullptr;
  int t_first = blockIdx.x;
  if (sched && threadIdx.x == 0) t_first = atomicAdd(sched, 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    for (int s = 0; s < kNBarA; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kNBarB; ++s) {
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kFwdThreads - 64);
    }
    for (int s = 0; s < kSched; ++s) {
      mbar_init(&sfull_bar[s], 1);
      mbar_init(&sempty_bar[s], kFwdThreads - 64 + 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile queue.  Producer side: fetch() = next tile number of this CTA, publish() = hand it to the consumers
  // (a number >= total_tiles tells them to stop).  Consumer side: next_tile().
  int sq = 0;                                       // queue position of this thread (producer or consumer)
  auto fetch = [&](int cur) -> int { return sched ? atomicAdd(sched, 1) : cur + static_cast<int>(gridDim.x); };
  // producer lane, after its last fetch: this CTA will not touch the counter again.  The CTA that counts gridDim.x - 1
  // others before it leaves the slot zeroed for the launch that reuses it (the consumers are still busy with the last
  // tile, so this costs nothing at the end of the kernel).
  auto retire = [&]() {
    if (!sched) return;
    __threadfence();
    if (atomicAdd(sched + 1, 1) == static_cast<int>(gridDim.x) - 1) {
      sched[0] = 0;
      sched[1] = 0;
    }
  };
  auto publish = [&](int t) {
    const int s = sq % kSched;
    mbar_wait(&sempty_bar[s], ((sq / kSched) & 1) ^ 1);
    sched_tile[s] = t;
    mbar_arrive(&sfull_bar[s]);                     // release: the store above is visible to whoever sees this phase
    ++sq;
  };
  auto next_tile = [&]() -> int {
    const int s = sq % kSched;
    mbar_wait(&sfull_bar[s], (sq / kSched) & 1);
    const int t = sched_tile[s];
    mbar_arrive(&sempty_bar[s]);
    ++sq;
    return t;
  };

  if (warp == 0) {
    if (lane == 0 && HALO) {
      // order: A(j+1) is requested before the nine B tiles of chunk j, so the patch of the next chunk (or tile)
      // lands while the tensor core works on the current one
      uint8_t* b_ring = smem + Cfg::kAStages * 2 * Cfg::kHaloBytes;
      int ja = 0, jb = 0;                         // A-ring / B-ring counters, run on across tiles
      auto load_a = [&](int t, int kc) {
        const int mt = t / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.TW;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = mt / (p.tiles_w * p.tiles_h);
        const int s = ja % Cfg::kAStages;
        mbar_wait(&empty_bar[s], ((ja / Cfg::kAStages) & 1) ^ 1);
        uint8_t* st = smem + s * 2 * Cfg::kHaloBytes;
        mbar_expect_tx(&full_bar[s], 2 * Cfg::kHaloBytes);
        tma_load_4d(st, &tm_a_hi, &full_bar[s], kc * kBK, w0 - 1, h0 - 1, n0);
        tma_load_4d(st + Cfg::kHaloBytes, &tm_a_lo, &full_bar[s], kc * kBK, w0 - 1, h0 - 1, n0);
        ++ja;
      };
      int t = t_first;
      if (t < total_tiles) load_a(t, 0);
      for (;;) {
        publish(t);
        if (t >= total_tiles) break;
        const int t_next = fetch(t);
        const int co0 = (t % p.n_tiles) * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (kc + 1 < p.kchunks) load_a(t, kc + 1);
          else if (t_next < total_tiles) load_a(t_next, 0);
          for (int tap = 0; tap < 9; ++tap, ++jb) {
            const int s = jb % Cfg::kBStages;
            mbar_wait(&bempty_bar[s], ((jb / Cfg::kBStages) & 1) ^ 1);
            uint8_t* st = b_ring + s * Cfg::kBStageBytes;
            mbar_expect_tx(&bfull_bar[s], Cfg::kBStageBytes);
            tma_load_2d(st, &tm_b_hi, &bfull_bar[s], tap * p.cin_pad + kc * kBK, co0);
            tma_load_2d(st + Cfg::kBBytes, &tm_b_lo, &bfull_bar[s], tap * p.cin_pad + kc * kBK, co0);
          }
        }
        t = t_next;
      }
      retire();
    } else if (lane == 0) {
      int ring = 0;                               // stage counter, runs on across tiles
      for (int t = t_first;;) {
        publish(t);
        if (t >= total_tiles) break;
        const int t_next = fetch(t);              // requested before this tile's loads: the atomic's latency hides behind them
        const int nt = t % p.n_tiles;
        const int mt = t / p.n_tiles;
        const int w0 = (mt % p.tiles_w) * p.TW;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
        const int co0 = nt * BN;
        for (int it = 0; it < k_iters; ++it, ++ring) {
          const int s = ring % Cfg::kStages;
          const uint32_t ph = (ring / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int tap = it / p.kchunks;
          const int kc = it - tap * p.kchunks;
          const int dr = (p.taps == 9) ? (tap / 3 - 1) : 0;
          const int ds = (p.taps == 9) ? (tap % 3 - 1) : 0;
          uint8_t* st = smem + s * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if (p.pair_maps) {
            // (hi, lo) are the two slices of one buffer: one 5-D / 3-D box fetches both halves of an operand tile
            tma_load_5d(st, &tm_a_hi, &full_bar[s], kc * kBK, w0 + ds, h0 + dr, n0, 0);
            tma_load_3d(st + 2 * kTileBytes, &tm_b_hi, &full_bar[s], tap * p.cin_pad + kc * kBK, co0, 0);
          } else {
            tma_load_4d(st, &tm_a_hi, &full_bar[s], kc * kBK, w0 + ds, h0 + dr, n0);
            tma_load_4d(st + kTileBytes, &tm_a_lo, &full_bar[s], kc * kBK, w0 + ds, h0 + dr, n0);
            tma_load_2d(st + 2 * kTileBytes, &tm_b_hi, &full_bar[s], tap * p.cin_pad + kc * kBK, co0);
            tma_load_2d(st + 2 * kTileBytes + Cfg::kBBytes, &tm_b_lo, &full_bar[s], tap * p.cin_pad + kc * kBK, co0);
          }
        }
        t = t_next;
      }
      retire();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 2 * BN, 0, 0);
      int ring = 0, ring_b = 0, pass_i = 0;       // smem stage counters / accumulator pass counter (run across tiles)
      (void)ring_b;
      for (int t = next_tile(); t < total_tiles; t = next_tile()) {
        int it = 0;
        for (int ps = 0; ps < p.n_pass; ++ps, ++pass_i) {
          const int buf = pass_i & 1;
          mbar_wait(&tempty_bar[buf], ((pass_i >> 1) & 1) ^ 1);   // epilogue has drained this buffer
          tc_fence_after();
          const uint32_t d_main = tmem_base + buf * (2 * BN);
          const uint32_t d_cross = d_main + BN;
          const int it_end = min(it + p.pass_len, k_iters);
          uint32_t fresh = 0;                       // 0 for the first MMA into each accumulator of the pass
          if constexpr (HALO) {
            const uint32_t b_ring = smem_u32(smem + Cfg::kAStages * 2 * Cfg::kHaloBytes);
            for (; it < it_end; ++it, ++ring) {     // ring counts chunks here, ring_b the weight tiles
              const int sa = ring % Cfg::kAStages;
              mbar_wait(&full_bar[sa], (ring / Cfg::kAStages) & 1);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(smem + sa * 2 * Cfg::kHaloBytes);
              const uint32_t a_lo = a_hi + Cfg::kHaloBytes;
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap, ++ring_b) {
                const int sb = ring_b % Cfg::kBStages;
                mbar_wait(&bfull_bar[sb], (ring_b / Cfg::kBStages) & 1);
                tc_fence_after();
                // A operand of this tap = rows (h + tap/3) * 16 + (w + tap%3) of the halo patch: an 8-row group per
                // image row (8 consecutive patch rows), groups 16 rows = 2048 B apart
                const uint32_t a_off = ((tap / 3) * 16 + (tap % 3)) * 128;
                const uint32_t b_hi = b_ring + sb * Cfg::kBStageBytes;
                const uint32_t b_lo = b_hi + Cfg::kBBytes;
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  const uint64_t dah = umma_desc_sw128(a_hi + a_off + k * 32, 16, 2048);
                  const uint64_t dal = umma_desc_sw128(a_lo + a_off + k * 32, 16, 2048);
                  const uint64_t dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
                  const uint64_t dbl = umma_desc_sw128(b_lo + k * 32, 16, 1024);
                  umma_bf16(d_cross, dal, dbh, idesc, fresh);
                  umma_bf16(d_cross, dah, dbl, idesc, 1u);
                  umma_bf16(d_main, dah, dbh, idesc, fresh);
                  fresh = 1u;
                }
                umma_commit(&bempty_bar[sb]);
              }
              umma_commit(&empty_bar[sa]);          // the patch is free once all nine taps have read it
            }
          } else
          for (; it < it_end; ++it, ++ring) {
            const int s = ring % Cfg::kStages;
            const uint32_t ph = (ring / Cfg::kStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * Cfg::kStageBytes);
            const uint32_t a_lo = a_hi + kTileBytes;
            const uint32_t b_hi = a_hi + 2 * kTileBytes;
            const uint32_t b_lo = b_hi + Cfg::kBBytes;
            if (p.ncat) {
              // [main | cross] (+)= a_hi x [b_hi ; b_lo]: the lo tile follows the hi tile in shared memory and the cross
              // accumulator follows the main one in TMEM, so ONE N = 2 BN instruction forms both products and a_hi is
              // read from shared memory once instead of twice; then cross += a_lo x b_hi.
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t dah = umma_desc_sw128(a_hi + k * 32, 16, 1024);
                const uint64_t dal = umma_desc_sw128(a_lo + k * 32, 16, 1024);
                const uint64_t dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
                umma_bf16(d_main, dah, dbh, idesc2, fresh);
                umma_bf16(d_cross, dal, dbh, idesc, 1u);
                fresh = 1u;
              }
            } else {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t dah = umma_desc_sw128(a_hi + k * 32, 16, 1024);
                const uint64_t dal = umma_desc_sw128(a_lo + k * 32, 16, 1024);
                const uint64_t dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
                const uint64_t dbl = umma_desc_sw128(b_lo + k * 32, 16, 1024);
                umma_bf16(d_cross, dal, dbh, idesc, fresh);
                umma_bf16(d_cross, dah, dbl, idesc, 1u);
                umma_bf16(d_main, dah, dbh, idesc, fresh);
                fresh = 1u;
              }
            }
            umma_commit(&empty_bar[s]);             // frees the smem stage once these MMAs have read it
          }
          umma_commit(&tfull_bar[buf]);             // this pass's accumulators are complete
        }
      }
    }
  } else {
    // ---- epilogue: warps 2..9; warp w reads TMEM lanes [32q, 32q+32), q = w % 4 (GEMM rows = pixels of the
    // patch) and the column half (w - 2) / 4 of the tile
    constexpr int HN = BN / 2;
    const int q = warp & 3;
    const int c_half = ((warp - 2) >> 2) * HN;
    const int m = q * 32 + lane;
    const int tw = m % p.TW;
    const int th = (m / p.TW) % p.TH;
    const int tn = m / (p.TW * p.TH);
    const bool vec_ok = (p.cout & 3) == 0;
    const bool writer = (p.pool == 0) || (((tw | th) & 1) == 0);
    const float pool_scale = (p.pool == 1) ? 0.25f : 1.0f;
    int pass_i = 0;
    float acc[HN];
    for (int t = next_tile(); t < total_tiles; t = next_tile()) {
      const int nt = t % p.n_tiles;
      const int mt = t / p.n_tiles;
      const int w = (mt % p.tiles_w) * p.TW + tw;
      const int h = ((mt / p.tiles_w) % p.tiles_h) * p.TH + th;
      const int n_img = (mt / (p.tiles_w * p.tiles_h)) * p.TN + tn;
      const int co0 = nt * BN + c_half;
      const bool row_ok = n_img < p.N;
      const size_t pix = (static_cast<size_t>(n_img) * p.H + h) * p.W + w;          // conv output pixel
      size_t opix = pix;                                                            // stored pixel
      if (p.pool) opix = (static_cast<size_t>(n_img) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      size_t rpix = opix;                                                           // residual pixel
      if (p.res_shift) rpix = (static_cast<size_t>(n_img) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
      // ---- gather the passes of this tile into registers (fp32 round-to-nearest adds)
      for (int ps = 0; ps < p.n_pass; ++ps, ++pass_i) {
        const int buf = pass_i & 1;
        mbar_wait(&tfull_bar[buf], (pass_i >> 1) & 1);
        tc_fence_after();
        const uint32_t t_base = tmem_base + buf * (2 * BN) + c_half + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < HN; c += 32) {
          if (co0 + c < p.cout) add_pass_chunk<BN>(t_base + c, ps == 0, acc + c);
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[buf]);              // 256 arrivals hand the buffer back to the MMA issuer
      }
      // ---- finalize and store (the tensor core is already working on the next pass / tile)
#pragma unroll
      for (int c = 0; c < HN; c += 32) {
        const int cbase = co0 + c;
        if (cbase >= (p.out_hi ? p.cout_pad : p.cout)) continue;                    // warp-uniform; pad channels of a pair are written as zeros
        float* v = acc + c;
        // (acc + bias) * out_scale, ReLU-derivative mask
        uint32_t mk[16];
        const bool use_mask = p.mask_hi != nullptr;
        if (use_mask) {
          const __nv_bfloat16* mp = p.mask_hi + pix * p.mask_cpad + cbase;
          if (row_ok && cbase + 32 <= p.mask_cpad) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const uint4 u = __ldg(reinterpret_cast<const uint4*>(mp) + g4);
              mk[g4 * 4] = u.x; mk[g4 * 4 + 1] = u.y; mk[g4 * 4 + 2] = u.z; mk[g4 * 4 + 3] = u.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              uint32_t lo16 = 0, hi16 = 0;
              if (row_ok && cbase + j < p.mask_cpad) lo16 = __bfloat16_as_ushort(mp[j]);
              if (row_ok && cbase + j + 1 < p.mask_cpad) hi16 = __bfloat16_as_ushort(mp[j + 1]);
              mk[j >> 1] = lo16 | (hi16 << 16);
            }
          }
        }
        float bz[32];
        if (p.bias && vec_ok && cbase + 32 <= p.cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));   // warp-uniform address
            bz[j] = b4.x; bz[j + 1] = b4.y; bz[j + 2] = b4.z; bz[j + 3] = b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) bz[j] = (p.bias && cbase + j < p.cout) ? __ldg(p.bias + cbase + j) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = cbase + j;
          float x = v[j];
          if (co < p.cout && row_ok) {
            x = (x + bz[j]) * p.out_scale;
            if (use_mask) {
              const uint32_t bits = (mk[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
              if ((bits & 0x8000u) || (bits & 0x7FFFu) == 0) x = 0.f;              // saved activation <= 0
            }
          } else {
            x = 0.f;
          }
          v[j] = x;
        }
        if (p.pool) {                                                               // 2x2 pooling across lanes
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j];
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            x += __shfl_xor_sync(0xffffffffu, x, p.TW);
            v[j] = x * pool_scale;
          }
        }
        if (!row_ok || !writer) continue;
        if (p.residual) {
          const float* rp = p.residual + rpix * p.cout + cbase;
          if (vec_ok && cbase + 32 <= p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp + j));
              v[j] = fmaf(p.res_scale, r4.x, v[j]); v[j + 1] = fmaf(p.res_scale, r4.y, v[j + 1]);
              v[j + 2] = fmaf(p.res_scale, r4.z, v[j + 2]); v[j + 3] = fmaf(p.res_scale, r4.w, v[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cbase + j < p.cout) v[j] = fmaf(p.res_scale, __ldg(rp + j), v[j]);
          }
        }
        if (p.out) {
          float* o = p.out + opix * p.cout + cbase;
          if (vec_ok && cbase + 32 <= p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (cbase + j < p.cout) o[j] = v[j];
          }
        }
        if (p.out_hi) {
          // split (optionally ReLU'd) copy for a following convolution; channel stride cout_pad
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a = v[j], b = v[j + 1];
            if (p.relu_split) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            __nv_bfloat16 ah, al, bh, bl;
            split_bf16(a, ah, al);
            split_bf16(b, bh, bl);
            hi[j >> 1] = pack_bf16x2(ah, bh);
            lo[j >> 1] = pack_bf16x2(al, bl);
          }
          __nv_bfloat16* oh = p.out_hi + opix * p.cout_pad + cbase;
          __nv_bfloat16* ol = p.out_lo + opix * p.cout_pad + cbase;
          // cout_pad is a multiple of 8 and cbase a multiple of 32: 16-byte groups of 8 channels
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (cbase + g * 8 < p.cout_pad) {
              *reinterpret_cast<uint4*>(oh + g * 8) = make_uint4(hi[g * 4], hi[g * 4 + 1], hi[g * 4 + 2], hi[g * 4 + 3]);
              *reinterpret_cast<uint4*>(ol + g * 8) = make_uint4(lo[g * 4], lo[g * 4 + 1], lo[g * 4 + 2], lo[g * 4 + 3]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------
template <int BN>
struct WgCfg {
  static constexpr int kStages = 3;                 // (a 4th stage for BN = 64 was measured: no gain)
  static constexpr int kABytes = 128 * 64 * 2;      // 128 channels x 64 pixels (two 64-channel boxes)
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
  static constexpr int kTmemCols = 4 * BN;   // two (main, cross) buffers
};

// SWAP = false: M = 128 output channels (dy), N = BN input channels (x shifted by this CTA's tap).
// SWAP = true (chosen when Cout <= 64, where the M = 128 tile would be half padding): the roles are exchanged,
//   M = 128 rows of x -- two 64-channel atoms that are either the two halves of a 128-channel slice of one tap
//   (Cin >= 128) or the SAME 64 channels shifted by two different taps (Cin <= 64) -- and N = BN output channels.
template <int BN, bool SWAP>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_dy_hi, const __grid_constant__ CUtensorMap tm_dy_lo,
                  const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                  const ConvWgradParams p) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int co_t = blockIdx.x / p.cin_tiles;
  const int ci_t = blockIdx.x - co_t * p.cin_tiles;
  const int co0 = co_t * (SWAP ? BN : 128), ci0 = ci_t * (SWAP ? 128 : BN);
  // taps of the two x atoms (SWAP) / of the CTA (plain); a missing second tap (9 is odd) re-reads the first
  const int tap = (SWAP && p.tap_pairs) ? 2 * blockIdx.y : blockIdx.y;
  const int tap1 = (SWAP && p.tap_pairs) ? min(tap + 1, p.taps - 1) : tap;
  const int dr = (p.taps == 9) ? (tap / 3 - 1) : 0;
  const int ds = (p.taps == 9) ? (tap % 3 - 1) : 0;
  const int dr1 = (p.taps == 9) ? (tap1 / 3 - 1) : 0;
  const int ds1 = (p.taps == 9) ? (tap1 % 3 - 1) : 0;
  const int pb_begin = blockIdx.z * p.blocks_per_split;
  const int pb_end = min(pb_begin + p.blocks_per_split, p.pix_blocks);
  const int k_iters = pb_end - pb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dy_hi);
    tma_prefetch_desc(&tm_dy_lo);
    tma_prefetch_desc(&tm_x_hi);
    tma_prefetch_desc(&tm_x_lo);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // passes of at most kPassLen k-iterations (144 k16 steps), nearly equal
  const int n_pass = (k_iters + kPassLen - 1) / kPassLen;
  const int pass_len = n_pass ? (k_iters + n_pass - 1) / n_pass : 0;

  if (k_iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < k_iters; ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int pb = pb_begin + it;
          const int w0 = (pb % p.tiles_w) * p.TW;
          const int h0 = ((pb / p.tiles_w) % p.tiles_h) * p.TH;
          const int n0 = (pb / (p.tiles_w * p.tiles_h)) * p.TN;
          uint8_t* st = smem + s * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if constexpr (SWAP) {
            // A = x atoms
            const int cj = p.tap_pairs ? 0 : 64;      // channel step between the two atoms
            tma_load_4d(st, &tm_x_hi, &full_bar[s], ci0, w0 + ds, h0 + dr, n0);
            tma_load_4d(st + 8192, &tm_x_hi, &full_bar[s], ci0 + cj, w0 + ds1, h0 + dr1, n0);
            tma_load_4d(st + Cfg::kABytes, &tm_x_lo, &full_bar[s], ci0, w0 + ds, h0 + dr, n0);
            tma_load_4d(st + Cfg::kABytes + 8192, &tm_x_lo, &full_bar[s], ci0 + cj, w0 + ds1, h0 + dr1, n0);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {         // B = dy atoms
              tma_load_4d(st + 2 * Cfg::kABytes + j * 8192, &tm_dy_hi, &full_bar[s], co0 + j * 64, w0, h0, n0);
              tma_load_4d(st + 2 * Cfg::kABytes + Cfg::kBBytes + j * 8192, &tm_dy_lo, &full_bar[s], co0 + j * 64, w0, h0, n0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_4d(st + j * 8192, &tm_dy_hi, &full_bar[s], co0 + j * 64, w0, h0, n0);
              tma_load_4d(st + Cfg::kABytes + j * 8192, &tm_dy_lo, &full_bar[s], co0 + j * 64, w0, h0, n0);
            }
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              tma_load_4d(st + 2 * Cfg::kABytes + j * 8192, &tm_x_hi, &full_bar[s], ci0 + j * 64, w0 + ds, h0 + dr, n0);
              tma_load_4d(st + 2 * Cfg::kABytes + Cfg::kBBytes + j * 8192, &tm_x_lo, &full_bar[s], ci0 + j * 64,
                          w0 + ds, h0 + dr, n0);
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
        constexpr uint32_t idesc2 = umma_idesc_bf16(128, 2 * BN, 1, 1);
        int it = 0;
        for (int ps = 0; ps < n_pass; ++ps) {
          const int buf = ps & 1;
          mbar_wait(&tempty_bar[buf], ((ps >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_main = tmem_base + buf * (2 * BN);
          const uint32_t d_cross = d_main + BN;
          const int it_end = min(it + pass_len, k_iters);
          uint32_t fresh = 0;
          for (; it < it_end; ++it) {
            const int s = it % Cfg::kStages;
            const uint32_t ph = (it / Cfg::kStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * Cfg::kStageBytes);
            const uint32_t a_lo = a_hi + Cfg::kABytes;
            const uint32_t b_hi = a_hi + 2 * Cfg::kABytes;
            const uint32_t b_lo = b_hi + Cfg::kBBytes;
            if (p.ncat) {       // as in the forward kernel: [main | cross] (+)= a_hi x [b_hi ; b_lo], cross += a_lo x b_hi
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 16 pixels per MMA
                const uint64_t dah = umma_desc_sw128(a_hi + k * 2048, 8192, 1024);
                const uint64_t dal = umma_desc_sw128(a_lo + k * 2048, 8192, 1024);
                const uint64_t dbh = umma_desc_sw128(b_hi + k * 2048, 8192, 1024);
                umma_bf16(d_main, dah, dbh, idesc2, fresh);
                umma_bf16(d_cross, dal, dbh, idesc, 1u);
                fresh = 1u;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 16 pixels per MMA
                const uint64_t dah = umma_desc_sw128(a_hi + k * 2048, 8192, 1024);
                const uint64_t dal = umma_desc_sw128(a_lo + k * 2048, 8192, 1024);
                const uint64_t dbh = umma_desc_sw128(b_hi + k * 2048, 8192, 1024);
                const uint64_t dbl = umma_desc_sw128(b_lo + k * 2048, 8192, 1024);
                umma_bf16(d_cross, dal, dbh, idesc, fresh);
                umma_bf16(d_cross, dah, dbl, idesc, 1u);
                umma_bf16(d_main, dah, dbh, idesc, fresh);
                fresh = 1u;
              }
            }
            umma_commit(&empty_bar[s]);
          }
          umma_commit(&tfull_bar[buf]);
        }
      }
    } else {
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
      float acc[BN];
      for (int ps = 0; ps < n_pass; ++ps) {
        const int buf = ps & 1;
        mbar_wait(&tfull_bar[buf], (ps >> 1) & 1);
        tc_fence_after();
        const uint32_t t_base = tmem_base + buf * (2 * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          if ((SWAP ? co0 : ci0) + c < (SWAP ? p.cout : p.cin)) add_pass_chunk<BN>(t_base + c, ps == 0, acc + c);
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[buf]);
      }
      if constexpr (SWAP) {
        // row m = q * 32 + lane of the accumulator: x atom m / 64 (second tap or second channel half), channel m % 64
        const int m = q * 32 + lane;
        const int my_tap = p.tap_pairs ? tap + (m >> 6) : tap;
        const int ci = p.tap_pairs ? (m & 63) : ci0 + m;
        if (my_tap < p.taps && ci < p.cin) {
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            if (co0 + c >= p.cout) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cc = co0 + c + j;
              if (cc < p.cout) {
                float* o = p.dw + (static_cast<size_t>(cc) * p.taps + my_tap) * p.cin + ci;   // lanes: consecutive ci
                if (p.atomic) atomicAdd(o, acc[c + j]); else *o = acc[c + j];
              }
            }
          }
        }
      } else if (co < p.cout) {
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          if (ci0 + c >= p.cin) continue;
          float* o = p.dw + (static_cast<size_t>(co) * p.taps + tap) * p.cin + ci0 + c;
          if (p.atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (ci0 + c + j < p.cin) atomicAdd(o + j, acc[c + j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (ci0 + c + j < p.cin) o[j] = acc[c + j];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC bf16 activation map: dims (C, W, H, N), box (64, bw, bh, bn), 128B swizzle, zero OOB fill
static int make_act_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int Cpad, int bw, int bh, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return L2I_ERR_DRIVER; }
  cuuint64_t dims[4] = {(cuuint64_t)Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)Cpad * 2, (cuuint64_t)W * Cpad * 2, (cuuint64_t)H * W * Cpad * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act N=%d H=%d W=%d C=%d box=%d,%d,%d) failed: %d", N, H, W, Cpad, bw, bh, bn, (int)r);
    return L2I_ERR_DRIVER;
  }
  return L2I_OK;
}

// (hi, lo) pair of NHWC maps as ONE 5-D map: dims (C, W, H, N, 2), the last stride = distance between the halves
static int make_act_pair_map(CUtensorMap* m, const void* hi, long long pair_stride, int N, int H, int W, int Cpad, int bw,
                             int bh, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return L2I_ERR_DRIVER; }
  cuuint64_t dims[5] = {(cuuint64_t)Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 2};
  cuuint64_t strides[4] = {(cuuint64_t)Cpad * 2, (cuuint64_t)W * Cpad * 2, (cuuint64_t)H * W * Cpad * 2, (cuuint64_t)pair_stride};
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, 2};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(hi), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(act pair) failed: %d", (int)r); return L2I_ERR_DRIVER; }
  return L2I_OK;
}

// (hi, lo) pair of weight maps as ONE 3-D map: dims (K, rows, 2)
static int make_w_pair_map(CUtensorMap* m, const void* hi, long long pair_stride, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return L2I_ERR_DRIVER; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)pair_stride};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(hi), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weight pair) failed: %d", (int)r); return L2I_ERR_DRIVER; }
  return L2I_OK;
}

// weight map: [rows][K] bf16 K-major, box (64, box_rows)
static int make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return L2I_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weight rows=%d K=%d) failed: %d", rows, K, (int)r);
    return L2I_ERR_DRIVER;
  }
  return L2I_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// pick a (tw, th, tn) patch of exactly `pixels` pixels
static int pick_tile(int H, int W, int pixels, int* tw, int* th, int* tn) {
  if (!is_pow2(H) || !is_pow2(W)) return L2I_ERR_UNSUPPORTED;
  int w = W < 16 ? W : 16;
  int h = pixels / w;
  if (h > H) h = H;
  *tw = w; *th = h; *tn = pixels / (w * h);
  return L2I_OK;
}

static int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[dev & 63];
  if (!n) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n = v;
  }
  return n;
}

template <int BN, bool HALO>
static int launch_fwd(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                      const CUtensorMap& b_lo, const ConvFwdParams& p, int grid, cudaStream_t stream) {
  static DeviceOnce configured;
  if (configured.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel<BN, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FwdCfg<BN, HALO>::kSmem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_fwd): %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
    configured.done();
  }
  conv_fwd_kernel<BN, HALO><<<grid, kFwdThreads, FwdCfg<BN, HALO>::kSmem, stream>>>(a_hi, a_lo, b_hi, b_lo, p);
  return check_launch(HALO ? "conv_fwd_kernel<halo>" : "conv_fwd_kernel");
}

// L2I_CONV_HALO=1 enables the halo-patch variant (off by default: measured on B200 it cuts L2 -> SM traffic per
// tile from 432 KB to 221 KB at Cin = Cout = 64 but is 5-15 % SLOWER than the plain ring -- those layers are paced
// by shared-memory bandwidth (TMA writes + UMMA operand reads) and the epilogue, not by L2).
// (The descriptor's matrix-base-offset field stays 0 for the shifted windows: on sm_100a the 128B swizzle is a
// function of the shared-memory address.)
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

int conv_fwd_tc(const ConvFwdArgs& a, cudaStream_t stream) {
  if (a.taps != 9 && a.taps != 1) { set_error("conv: taps must be 9 (3x3 pad 1) or 1 (1x1), got %d", a.taps); return L2I_ERR_UNSUPPORTED; }
  if (a.cin_pad % 8 || a.cin_pad <= 0 || a.cout <= 0 || a.N <= 0) { set_error("conv: bad channel/batch sizes (cin_pad=%d cout=%d N=%d)", a.cin_pad, a.cout, a.N); return L2I_ERR_BAD_ARG; }
  if (!a.x_hi || !a.x_lo || !a.w_hi || !a.w_lo || (!a.out && !a.out_hi)) { set_error("conv: null operand pointer"); return L2I_ERR_BAD_ARG; }
  if (a.out_hi && (a.cout_pad % 8 || a.cout_pad < a.cout)) { set_error("conv: cout_pad must be a multiple of 8 >= cout"); return L2I_ERR_BAD_ARG; }
  if (a.pool < 0 || a.pool > 2) { set_error("conv: pool must be 0 (none), 1 (2x2 average) or 2 (2x2 sum)"); return L2I_ERR_BAD_ARG; }
  if ((a.res_shift || a.pool) && (((a.H | a.W) & 1) || a.H < 2 || a.W < 2)) { set_error("conv: pooling / upsampled residual need even H, W"); return L2I_ERR_BAD_ARG; }
  if (a.res_shift && a.pool) { set_error("conv: pooled output with an upsampled residual is not supported"); return L2I_ERR_UNSUPPORTED; }
  if (a.mask_hi && (a.mask_cpad % 8 || a.mask_cpad < a.cout)) { set_error("conv: mask channel stride must be a multiple of 8 >= cout"); return L2I_ERR_BAD_ARG; }
  ConvFwdParams p;
  p.N = a.N; p.H = a.H; p.W = a.W; p.cin_pad = a.cin_pad; p.cout = a.cout; p.taps = a.taps;
  if (pick_tile(a.H, a.W, 128, &p.TW, &p.TH, &p.TN) != L2I_OK) { set_error("conv: H=%d W=%d must be powers of two", a.H, a.W); return L2I_ERR_UNSUPPORTED; }
  static const int halo_enabled = env_int("L2I_CONV_HALO", 0);
  const bool halo = halo_enabled && a.taps == 9 && a.H >= 16 && a.W >= 8 && a.cin_pad >= 64;
  if (halo) { p.TW = 8; p.TH = 16; p.TN = 1; }
  p.tiles_w = a.W / p.TW; p.tiles_h = a.H / p.TH;
  const int tiles_n = (a.N + p.TN - 1) / p.TN;
  p.kchunks = (a.cin_pad + kBK - 1) / kBK;
  {
    const int k_iters = halo ? p.kchunks : a.taps * p.kchunks;
    const int max_len = halo ? kPassLen / 9 : kPassLen;
    p.n_pass = (k_iters + max_len - 1) / max_len;
    p.pass_len = (k_iters + p.n_pass - 1) / p.n_pass;
  }
  p.bias = a.bias; p.residual = a.residual; p.res_shift = a.res_shift; p.res_scale = a.res_scale; p.out = a.out;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(a.out_hi); p.out_lo = reinterpret_cast<__nv_bfloat16*>(a.out_lo);
  p.cout_pad = a.cout_pad; p.relu_split = a.relu_split; p.out_scale = a.out_scale;
  p.mask_hi = reinterpret_cast<const __nv_bfloat16*>(a.mask_hi); p.mask_cpad = a.mask_cpad; p.pool = a.pool;
  static const int ncat_enabled = env_int("L2I_CONV_NCAT", 1);
  p.ncat = ncat_enabled;
  // L2I_CONV_DYNAMIC=1 hands the tiles out through the device-wide counter.  Off by default: measured on B200 it is
  // 0.3 % slower than the static stride on one GPU and gains nothing at 4 GPUs with the all-reduces overlapping the
  // backward pass (88.6 vs 89.0 ms / step) -- the NCCL kernels do not displace enough CTAs for long enough to matter.
  static const int dynamic_enabled = env_int("L2I_CONV_DYNAMIC", 0);
  static std::atomic<unsigned> next_slot{0};
  p.sched_slot = dynamic_enabled ? static_cast<int>(next_slot.fetch_add(1, std::memory_order_relaxed) % kSchedSlots) : -1;
  p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  int BN = (a.cout > 64) ? 128 : 64;
  // few pixels, many channels (1024 -> 1024 at 4 x 4: 8 x 8 tiles of 128 x 128 for 148 SMs; the 308-wide linears):
  // with 128-wide tiles less than half of the SMs would get a tile, each with the whole K loop; 64-wide tiles double
  // the CTAs that work (measured: 96 -> 78 us on 1024 -> 1024 at 4 x 4, 89.2 -> 87.8 ms per step;
  // L2I_CONV_NARROW_TILES=0 restores the 128-wide tiles)
  static const int narrow_enabled = env_int("L2I_CONV_NARROW_TILES", 1);
  if (narrow_enabled && BN == 128 && 2LL * p.m_tiles * ((a.cout + 127) / 128) <= sm_count()) BN = 64;
  p.n_tiles = (a.cout + BN - 1) / BN;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  // halo mode: the A box is the (16 + 2) x 16 pixel patch around the 8 x 16 output tile (x from w0 - 1)
  const int bw = halo ? 16 : p.TW, bh = halo ? 18 : p.TH;
  // when both operands' halves are slices of one buffer each (they are, for every pair this library's callers
  // build), one tensor map per operand covers (hi, lo) and a k-iteration is 2 TMA instructions instead of 4
  static const int pair_enabled = env_int("L2I_CONV_PAIR_MAPS", 1);
  const long long sx = reinterpret_cast<const char*>(a.x_lo) - reinterpret_cast<const char*>(a.x_hi);
  const long long sw = reinterpret_cast<const char*>(a.w_lo) - reinterpret_cast<const char*>(a.w_hi);
  const long long x_bytes = 2LL * a.N * a.H * a.W * a.cin_pad, w_bytes = 2LL * a.cout * a.taps * a.cin_pad;
  p.pair_maps = (pair_enabled && !halo && sx >= x_bytes && sw >= w_bytes && (sx % 16) == 0 && (sw % 16) == 0 &&
                 sx < (1LL << 40) && sw < (1LL << 40)) ? 1 : 0;
  if (p.pair_maps) {
    if ((rc = make_act_pair_map(&ta_hi, a.x_hi, sx, a.N, a.H, a.W, a.cin_pad, bw, bh, p.TN))) return rc;
    if ((rc = make_w_pair_map(&tb_hi, a.w_hi, sw, a.cout, a.taps * a.cin_pad, BN))) return rc;
    ta_lo = ta_hi; tb_lo = tb_hi;
  } else {
    if ((rc = make_act_map(&ta_hi, a.x_hi, a.N, a.H, a.W, a.cin_pad, bw, bh, p.TN))) return rc;
    if ((rc = make_act_map(&ta_lo, a.x_lo, a.N, a.H, a.W, a.cin_pad, bw, bh, p.TN))) return rc;
    if ((rc = make_w_map(&tb_hi, a.w_hi, a.cout, a.taps * a.cin_pad, BN))) return rc;
    if ((rc = make_w_map(&tb_lo, a.w_lo, a.cout, a.taps * a.cin_pad, BN))) return rc;
  }
  const long long total = 1LL * p.m_tiles * p.n_tiles;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  if (halo) {
    if (BN == 128) return launch_fwd<128, true>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
    return launch_fwd<64, true>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
  }
  if (BN == 128) return launch_fwd<128, false>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
  return launch_fwd<64, false>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
}

template <int BN, bool SWAP>
static int launch_wgrad(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                        const CUtensorMap& b_lo, const ConvWgradParams& p, dim3 grid, cudaStream_t stream) {
  static DeviceOnce configured;
  if (configured.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<BN, SWAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<BN>::kSmem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv_wgrad): %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
    configured.done();
  }
  conv_wgrad_kernel<BN, SWAP><<<grid, kThreads, WgCfg<BN>::kSmem, stream>>>(a_hi, a_lo, b_hi, b_lo, p);
  return check_launch(SWAP ? "conv_wgrad_kernel<swap>" : "conv_wgrad_kernel");
}

int conv_wgrad_tc(const ConvWgradArgs& a, cudaStream_t stream) {
  if (a.taps != 9 && a.taps != 1) { set_error("wgrad: taps must be 9 or 1"); return L2I_ERR_UNSUPPORTED; }
  if (a.cin_pad % 8 || a.cout_pad % 8 || a.cin > a.cin_pad || a.cout > a.cout_pad || a.N <= 0) { set_error("wgrad: bad sizes"); return L2I_ERR_BAD_ARG; }
  if (!a.dy_hi || !a.dy_lo || !a.x_hi || !a.x_lo || !a.dw) { set_error("wgrad: null pointer"); return L2I_ERR_BAD_ARG; }
  ConvWgradParams p;
  p.N = a.N; p.H = a.H; p.W = a.W; p.cin = a.cin; p.cout = a.cout; p.taps = a.taps;
  if (pick_tile(a.H, a.W, 64, &p.TW, &p.TH, &p.TN) != L2I_OK) { set_error("wgrad: H=%d W=%d must be powers of two", a.H, a.W); return L2I_ERR_UNSUPPORTED; }
  p.tiles_w = a.W / p.TW; p.tiles_h = a.H / p.TH;
  const int tiles_n = (a.N + p.TN - 1) / p.TN;
  p.pix_blocks = p.tiles_w * p.tiles_h * tiles_n;
  // Cout <= 64: exchange the operand roles (M = 128 rows of x, N = 64 output channels) so that the M = 128 tile is
  // not half padding; with Cin <= 64 as well, the two x atoms are two taps of the same channels
  const bool swap = a.cout <= 64;
  p.tap_pairs = (swap && a.cin <= 64) ? 1 : 0;
  const int BN = swap ? 64 : ((a.cin > 64) ? 128 : 64);
  p.cin_tiles = swap ? (p.tap_pairs ? 1 : (a.cin + 127) / 128) : (a.cin + BN - 1) / BN;
  const int co_tiles = swap ? 1 : (a.cout + 127) / 128;
  const int tap_groups = p.tap_pairs ? (a.taps + 1) / 2 : a.taps;
  const int base_ctas = co_tiles * p.cin_tiles * tap_groups;
  // split-K over pixel blocks: choose the split count that minimises (waves of 148 CTAs) x (K iterations per CTA +
  // a fixed per-CTA cost of ~12 iterations for prologue, pipeline fill and the atomic epilogue), >= 8 blocks per split
  int max_splits = (p.pix_blocks + 7) / 8;
  if (max_splits < 1) max_splits = 1;
  int cap = (4 * 148 + base_ctas - 1) / base_ctas;
  if (cap > max_splits) cap = max_splits;
  int splits = 1;
  long long best = -1;
  for (int sp = 1; sp <= cap; ++sp) {
    const long long waves = (1LL * base_ctas * sp + 147) / 148;
    const long long cost = waves * ((p.pix_blocks + sp - 1) / sp + 12);
    if (best < 0 || cost < best) { best = cost; splits = sp; }
  }
  p.blocks_per_split = (p.pix_blocks + splits - 1) / splits;
  splits = (p.pix_blocks + p.blocks_per_split - 1) / p.blocks_per_split;
  p.atomic = splits > 1;
  p.dw = a.dw;
  static const int ncat_enabled = env_int("L2I_CONV_NCAT", 1);
  p.ncat = ncat_enabled;
  if (p.atomic) {
    cudaError_t e = cudaMemsetAsync(a.dw, 0, sizeof(float) * (size_t)a.cout * a.taps * a.cin, stream);
    if (e != cudaSuccess) { set_error("wgrad: memset failed: %s", cudaGetErrorString(e)); return L2I_ERR_LAUNCH; }
  }
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  if ((rc = make_act_map(&ta_hi, a.dy_hi, a.N, a.H, a.W, a.cout_pad, p.TW, p.TH, p.TN))) return rc;
  if ((rc = make_act_map(&ta_lo, a.dy_lo, a.N, a.H, a.W, a.cout_pad, p.TW, p.TH, p.TN))) return rc;
  if ((rc = make_act_map(&tb_hi, a.x_hi, a.N, a.H, a.W, a.cin_pad, p.TW, p.TH, p.TN))) return rc;
  if ((rc = make_act_map(&tb_lo, a.x_lo, a.N, a.H, a.W, a.cin_pad, p.TW, p.TH, p.TN))) return rc;
  dim3 grid(co_tiles * p.cin_tiles, tap_groups, splits);
  if (swap) return launch_wgrad<64, true>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
  if (BN == 128) return launch_wgrad<128, false>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
  return launch_wgrad<64, false>(ta_hi, ta_lo, tb_hi, tb_lo, p, grid, stream);
}

}  // namespace l2i
