// TF32 precision mode of the backbone: fp32 activations and weights in HBM, tcgen05.mma kind::tf32 (fp32 operands read
// straight from shared memory, 10-bit mantissa products, fp32 accumulate), fp32 epilogue.  Same layers as the bf16 path
// (Conv3d -> BatchNorm3d -> ReLU triples, residual adds, max-pools, global average pool: src/i3d.py:101-116, :212-217,
// :262-272, :303-318) for the accuracy mode BASELINE.json asks for beside bf16 (features within 1e-3 of the fp32
// reference instead of 1e-2).  One general kernel per op kind -- this mode trades the layer-specialised bf16 kernels for
// precision, so it is built for coverage (any kernel / stride / padding, cin % 4 == 0) rather than for peak:
//   conv_tf32_kernel<BN, TMA_A>  persistent 128 x BN tiles; warp 0: TMA producer of the weight tile (fp32 [cout, K_pad] map,
//                          32 floats = one 128-byte swizzled row per k-block) and, for Cin % 32 == 0, of the activation
//                          tile (rank-5 fp32 im2col map); warp 1: single-thread MMA issuer; warps 2-5: epilogue (TMEM ->
//                          scale/shift (+residual) (+ReLU) -> TF32-rounded fp32 channels-last, written into a channel
//                          slice of the destination); warps 6-9: activation gather for the other layers (cp.async
//                          16 B = 4 channels with zero fill for padding, into the 128B-swizzled A tile).
//   maxpool3d_f32_kernel   one thread per (output pixel, 4-channel vector)
//   avgpool_f32_kernel     [B, P, C] fp32 -> [B, C] fp32, a warp per 128 channels
//   ingest_ncthw_f32_to_ndhwc4_kernel   the reference's fp32 NCTHW clip -> channels-last with RGB padded to 4 channels
#pragma once

#include "aux_kernels.cuh"
#include "conv_umma.cuh"
#include "head_kernels.cuh"
#include "stem_umma.cuh"
#include "conv_pair.cuh"

namespace vad {

struct Tf32ConvParams {
  int M, N, num_kb;  // num_kb: 32-float k-blocks
  int n_tiles, num_tiles;
  int To, Ho, Wo;
  int Ti, Hi, Wi;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int cin, ntaps;
  int fold;  // 1 (RGB stem, Cin = 4, kw <= 8): a k-block is one (dt, dh) tap x a window of 8 consecutive pixels x 4 channels,
             // 128 contiguous bytes of the input row; weights carry zeros for window pixels >= kw
  long long sN, sT, sH, sW;  // input strides in elements
  int relu;
  int ldo, ldr;  // output / residual row pitch in elements
  const float* in;
  const float* scale;
  const float* shift;
  const float* res;
  float* out;  // already offset by dst_c_off
};

// Round to the nearest TF32 value (10-bit mantissa, ties away from zero).  The tensor core simply drops the low 13
// mantissa bits of an fp32 operand, a truncation whose bias compounds over ~50 layers of non-negative activations
// (measured: 6e-3 on the features); every activation this mode stores, and every packed weight, is therefore already
// a TF32 value, so the truncation in the MMA is exact.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <int BN>
struct Tf32Cfg {
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes);
  static constexpr int kThreads = 64 + 128 + 128;
  static constexpr int kGatherLag = kStages - 2 > 6 ? 6 : kStages - 2;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * BN * 4 + (2 * kStages + 4) * 8 + 16 + 1024;
};

// TMA_A: the activation tile comes from a rank-5 fp32 im2col tensor map (128 output pixels x 32 channels per k-block,
// halo zero-filled by the TMA unit) issued by the producer thread next to the weight tile; needs Cin % 32 == 0.  Otherwise
// the four gather warps build the tile with zero-filling cp.async (any Cin % 4 == 0: the RGB stem, Inception's 16/24/48).
template <int BN, bool TMA_A>
__global__ void __launch_bounds__(Tf32Cfg<BN>::kThreads, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Tf32ConvParams p) {
  using Cfg = Tf32Cfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* s_scale = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
    if (TMA_A) tma_prefetch_desc(&tmA);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], TMA_A ? 1 : 1 + 128);  // TMA (expect_tx) (+ one arrival per gather thread)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * BN;
        int wq = 0, hq = 0, dq = 0, nq = 0;
        if (TMA_A) {  // base pixel of the tile's first row, in input coordinates
          int t = (tile / p.n_tiles) * kBlockM;
          const int wo = t % p.Wo; t /= p.Wo;
          const int ho = t % p.Ho; t /= p.Ho;
          const int to = t % p.To; t /= p.To;
          wq = wo * p.sw - p.pw; hq = ho * p.sh - p.ph; dq = to * p.st - p.pt; nq = t;
        }
        int c0 = 0, dw = 0, dh = 0, dt = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          mbar_arrive_expect_tx_a(full0 + s * 8, (uint32_t)(TMA_A ? Cfg::kStageBytes : Cfg::kBBytes));
          if (TMA_A) {
            tma_load_im2col_5d_a(stage0 + s * Cfg::kStageBytes, &tmA, full0 + s * 8, c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh,
                                 (uint16_t)dt);
            c0 += 32;
            if (c0 >= p.cin) {
              c0 = 0;
              if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
            }
          }
          tma_load_2d_a(stage0 + s * Cfg::kStageBytes + Cfg::kABytes, &tmB, full0 + s * 8, kb * 32, n0);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_tf32_m128(BN);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      uint32_t s = 0, ph = 0;
      int tc = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const uint32_t acc = (uint32_t)tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, (((uint32_t)tc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_a(full0 + s * 8, ph);
          tc_fence_after();
          const uint32_t a_lo = (stage0 + s * Cfg::kStageBytes) >> 4;
          const uint64_t adesc = desc_hi | a_lo;
          const uint64_t bdesc = desc_hi | (a_lo + (Cfg::kABytes >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 8 tf32 = 32 B per MMA: +2 in the (addr >> 4) field
            umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit_a(empty0 + s * 8);
          if (kb == p.num_kb - 1) umma_commit_a(tfull0 + acc * 8);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        ++tc;
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue warps
    const int t = threadIdx.x - 64;
    const int q = warp & 3;
    int tc = 0, cached_n0 = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (tile / p.n_tiles) * kBlockM;
      const int acc = tc & 1;
      const uint32_t aph = (tc >> 1) & 1;
      ++tc;
      if (n0 != cached_n0) {
        named_bar_sync(1, 128);
        for (int i = t; i < BN; i += 128) {
          const int n = n0 + i;
          s_scale[i] = (n < p.N) ? p.scale[n] : 0.f;
          s_shift[i] = (n < p.N) ? p.shift[n] : 0.f;
        }
        named_bar_sync(1, 128);
        cached_n0 = n0;
      }
      const int row = m0 + q * 32 + lane;
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool row_ok = row < p.M;
      float* out_row = p.out + (long long)row * p.ldo + n0;
      const float* res_row = p.res ? p.res + (long long)row * p.ldr + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int col = c * 32 + g * 4;
            if (n0 + col < p.N) {
              float4 f;
              f.x = fmaf(__uint_as_float(v[g * 4 + 0]), s_scale[col + 0], s_shift[col + 0]);
              f.y = fmaf(__uint_as_float(v[g * 4 + 1]), s_scale[col + 1], s_shift[col + 1]);
              f.z = fmaf(__uint_as_float(v[g * 4 + 2]), s_scale[col + 2], s_shift[col + 2]);
              f.w = fmaf(__uint_as_float(v[g * 4 + 3]), s_scale[col + 3], s_shift[col + 3]);
              if (res_row) {
                const float4 r = *reinterpret_cast<const float4*>(res_row + col);
                f.x += r.x; f.y += r.y; f.z += r.z; f.w += r.w;
              }
              if (p.relu) {
                f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
              }
              f.x = tf32_rna(f.x); f.y = tf32_rna(f.y); f.z = tf32_rna(f.z); f.w = tf32_rna(f.w);
              *reinterpret_cast<float4*>(out_row + col) = f;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  } else if (!TMA_A) {
    // ------------------------------------------------------------------ activation gather (one tile row per thread)
    constexpr int LAG = Cfg::kGatherLag;
    const int t = threadIdx.x - 192;
    const uint32_t sw_xor = (uint32_t)(t & 7);
    int g = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m = (tile / p.n_tiles) * kBlockM + t;
      const bool row_ok = m < p.M;
      int wo = 0, ho = 0, to = 0, nb = 0;
      if (row_ok) {
        int qd = m;
        wo = qd % p.Wo; qd /= p.Wo;
        ho = qd % p.Ho; qd /= p.Ho;
        to = qd % p.To; qd /= p.To;
        nb = qd;
      }
      const int w_base = wo * p.sw - p.pw, h_base = ho * p.sh - p.ph, t_base = to * p.st - p.pt;
      const float* img = p.in + (long long)nb * p.sN;
      int c = 0, dw = 0, dh = 0, dt = 0, tap = 0;
      bool ok = false;
      const float* src = p.in;
      auto set_tap = [&]() {
        const int wi = w_base + dw, hi = h_base + dh, ti = t_base + dt;
        ok = row_ok && tap < p.ntaps && (unsigned)wi < (unsigned)p.Wi && (unsigned)hi < (unsigned)p.Hi && (unsigned)ti < (unsigned)p.Ti;
        src = ok ? img + ti * p.sT + hi * p.sH + wi * p.sW : p.in;
      };
      set_tap();
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = g % STAGES;
        const uint32_t ph = (g / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t dst_row = smem_u32(stage_base + s * Cfg::kStageBytes) + (uint32_t)t * 128u;
        if (p.fold) {
          // one (dt, dh) tap: 8 consecutive input pixels x 4 channels, one 16-byte chunk per pixel
          const int ti = t_base + dt, hi = h_base + dh;
          const bool rv = row_ok && (unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi;
          const float* row = img + ti * p.sT + hi * p.sH;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int wi = w_base + j;
            const bool okj = rv && (unsigned)wi < (unsigned)p.Wi;
            cp_async_16_zfill(dst_row + (((uint32_t)j ^ sw_xor) << 4), okj ? row + wi * 4 : p.in, okj ? 16u : 0u);
          }
          if (++dh == p.kh) { dh = 0; ++dt; }
          cp_async_commit();
          ++g;
          if (g > LAG) {
            cp_async_wait<LAG>();
            fence_proxy_async_smem();
            mbar_arrive(&full_bar[(g - 1 - LAG) % STAGES]);
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 chunks of 4 channels
          cp_async_16_zfill(dst_row + (((uint32_t)j ^ sw_xor) << 4), src + (ok ? c : 0), ok ? 16u : 0u);
          c += 4;
          if (c == p.cin) {
            c = 0;
            ++tap;
            if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
            set_tap();
          }
        }
        cp_async_commit();
        ++g;
        if (g > LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[(g - 1 - LAG) % STAGES]);
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (int i = (g > LAG ? g - LAG : 0); i < g; ++i) mbar_arrive(&full_bar[i % STAGES]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// TF32 stem (VAD_FLAG_STEM_PLANES): conv kt x kh x (kw <= 8), stride 2 in h and w, RGB(+0) -> 64 channels, + BN + ReLU
// (src/i3d.py:202-209,303-305) on the dedicated-stem design of K2s (stem_umma.cuh), for fp32 operands.
//
// Through the general kernel above the stem is half of the TF32 forward: every (dt, dh) tap re-reads a 16 KB activation
// column and 8 KB of weights through L2.  K2s removes that for bf16 by letting the tensor core read the sliding 8-pixel
// windows of 8 consecutive output columns straight out of raw input-row segments -- a no-swizzle K-major descriptor whose
// rows are 16 bytes apart.  bf16 pixels (r, g, b, 0) are 8 bytes, so stride-2 output columns are 16 bytes apart; fp32
// pixels are 16 bytes and stride 2 would be 32.  So the input is de-interleaved by COLUMN PARITY (vad_tf32_ingest_ncthw_planes):
// window pixel j of output column w' is padded pixel 2 w' + j, i.e. position w' + j / 2 of plane j & 1 -- in each plane
// consecutive output columns are one pixel = 16 bytes apart again, a window is 4 pixels = 64 bytes, and a (dt, dh) tap becomes
// two 16-float "virtual taps" (even plane, odd plane).  Byte for byte that is the bf16 stem's operand geometry (176-byte
// segments, 64-byte windows, two 32-byte MMAs per virtual tap), with kind::tf32 MMAs.
//   * Weights: kt * kh * 2 virtual taps x (64 couts x 64 B) = 280 KB in fp32 -- twice shared memory.  A CTA therefore owns
//     HALF of the output channels (blockIdx.x & 1) and keeps its 140 KB resident; every spatial tile is visited by two CTAs
//     (the activation segments are 13 KB per input frame, their second read comes from L2).
//   * A (dt) stage = four TMA boxes: even / odd input rows (vertical stride 2) x even / odd column plane.
//   * N = 32 MMAs are bound by their operand fetch (A 4 KB + B 1 KB per 128 B/clk: 40 cycles against 16 of math); that is
//     still 3.6 x fewer cycles than the L2-bound general kernel spends.
//   * Epilogue: TMEM -> scale / shift + ReLU -> TF32 rounding -> 128B-swizzled fp32 staging tile (128 pixels x 32 channels) ->
//     one TMA store per warp into the channel half of the fp32 output.
struct StemTf32Params {
  int B, To, Ho, Wo;
  int kt, kh, st, pt, ph;
  int tiles_w, tiles_h;
  int num_tiles;            // B * To * tiles_h * tiles_w (each visited by both channel halves)
  int rows_even, rows_odd;  // box rows: 16 + (#dh taps of that parity) - 1
  int box_bytes[2];         // bytes of one even-row / odd-row box (rows x 176, padded to 128)
  int stage_bytes;          // 2 * (box_bytes[0] + box_bytes[1])
  int n_stages;
  int relu;
  int Ti;                   // input frames: frame taps outside the clip are skipped (pair kernel: only when both tiles of an item
                            // share their output frame, else 0 = no skipping)
  const float* scale;
  const float* shift;
};
constexpr int kStemTf32Threads = 192;
constexpr int kStemTf32TapBytes = 32 * 64;        // one virtual tap of one channel half: 32 couts x 16 floats
constexpr int kStemTf32StagingBytes = 128 * 128;  // 128 pixels x 32 fp32 channels
constexpr int kStemTf32SegBytes = 176;            // 8 windows of 64 B, 16 B apart

__device__ __forceinline__ void tma_load_5d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// tmE / tmOdd: input planes viewed as (x = Wh * 4 floats, plane, H, T, N), boxes of 44 floats x 1 x rows (row stride 2);
// tmW: fp32 weights [64][kt * kh * 32], boxes of 16 floats x 32 rows (SWIZZLE_64B); tmO: output (C, Wo, Ho, To, N), boxes 32 x 8 x 4.
__global__ void __launch_bounds__(kStemTf32Threads, 1)
stem_tf32_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const StemTf32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int ntaps = p.kt * p.kh * 2;                               // virtual taps
  uint8_t* w_smem = smem;                                          // ntaps x 2 KB, resident
  uint8_t* staging = smem + ((ntaps * kStemTf32TapBytes + 1023) & ~1023);   // 2 x 16 KB output staging (1024-aligned)
  uint8_t* stage_base = staging + 2 * kStemTf32StagingBytes;
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 32;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 32);
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kStemMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;
  const int half = (int)(blockIdx.x & 1u);                         // which 32 of the 64 output channels
  const int t_first = (int)(blockIdx.x >> 1), t_step = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);   // two 32-column accumulators
    tmem_relinquish();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    if (t < 32) {
      s_scale[t] = p.scale[half * 32 + t];
      s_shift[t] = p.shift[half * 32 + t];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t box_e = (uint32_t)p.box_bytes[0], box_o = (uint32_t)p.box_bytes[1];
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTf32TapBytes));
      for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemTf32TapBytes, &tmW, w_bar, tap * 16, half * 32);
      const uint32_t tx = (uint32_t)(2 * (p.rows_even + p.rows_odd) * kStemTf32SegBytes);
      uint32_t s = 0, ph = 0;
      for (int tile = t_first; tile < p.num_tiles; tile += t_step) {
        int r = tile;
        const int wb = r % p.tiles_w; r /= p.tiles_w;
        const int hb = r % p.tiles_h; r /= p.tiles_h;
        const int to = r % p.To;
        const int n = r / p.To;
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 4;    // 8 windows, one plane pixel (4 floats) apart
        const int t0 = to * p.st - p.pt;
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;   // taps inside the clip
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t dst = stage0 + s * (uint32_t)p.stage_bytes;
          const uint32_t fb = full0 + s * 8;
          mbar_arrive_expect_tx_a(fb, tx);
          tma_load_5d_a(dst, &tmE, fb, x_start, 0, h_start, t0 + dt, n);
          tma_load_5d_a(dst + box_e, &tmE, fb, x_start, 1, h_start, t0 + dt, n);
          tma_load_5d_a(dst + 2 * box_e, &tmOdd, fb, x_start, 0, h_start + 1, t0 + dt, n);
          tma_load_5d_a(dst + 2 * box_e + box_o, &tmOdd, fb, x_start, 1, h_start + 1, t0 + dt, n);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_tf32_m128(32);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      mbar_wait(w_bar, 0);
      const uint32_t seg16 = kStemTf32SegBytes >> 4;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, kStemTf32SegBytes);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      uint32_t s = 0, ph = 0, tc = 0;
      for (int tile = t_first; tile < p.num_tiles; tile += t_step, ++tc) {
        const uint32_t acc = tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, ((tc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 32u;
        const int t0 = ((tile / (p.tiles_w * p.tiles_h)) % p.To) * p.st - p.pt;
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(full0 + s * 8, ph);
          tc_fence_after();
          const uint32_t st16 = (stage0 + s * (uint32_t)p.stage_bytes) >> 4;
          uint32_t b_lo = w16 + (uint32_t)(dt * p.kh * 2) * (kStemTf32TapBytes >> 4);
          for (int dh = 0; dh < p.kh; ++dh) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              const uint32_t box16 = ((dh & 1) ? 2 * box_e + (uint32_t)pl * box_o : (uint32_t)pl * box_e) >> 4;
              const uint64_t adesc = a_hi | (st16 + box16 + (uint32_t)(dh >> 1) * seg16);
              const uint64_t bdesc = b_hi | b_lo;
              umma_tf32(d_tmem, adesc, bdesc, idesc, (dt > dt_lo || dh || pl) ? 1u : 0u);
              umma_tf32(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              b_lo += kStemTf32TapBytes >> 4;
            }
          }
          umma_commit_a(empty0 + s * 8);
          if (dt == dt_hi - 1) umma_commit_a(tfull0 + acc * 8);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int q = warp & 3;
    const int lrow = q * 32 + lane;  // tile row = TMEM lane: (h_i, w_i) = (lrow / 8, lrow % 8)
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    uint32_t tc = 0;
    for (int tile = t_first; tile < p.num_tiles; tile += t_step, ++tc) {
      int r = tile;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h; r /= p.tiles_h;
      const int to = r % p.To;
      const int n = r / p.To;
      const uint32_t sb = tc & 1u, acc = tc & 1u;
      const uint32_t row_addr = staging0 + sb * kStemTf32StagingBytes + (uint32_t)lrow * 128u;
      // the TMA store that read this staging buffer two tiles ago must have drained it
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32u, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float4 f;
        f.x = fmaf(__uint_as_float(v[g * 4 + 0]), s_scale[g * 4 + 0], s_shift[g * 4 + 0]);
        f.y = fmaf(__uint_as_float(v[g * 4 + 1]), s_scale[g * 4 + 1], s_shift[g * 4 + 1]);
        f.z = fmaf(__uint_as_float(v[g * 4 + 2]), s_scale[g * 4 + 2], s_shift[g * 4 + 2]);
        f.w = fmaf(__uint_as_float(v[g * 4 + 3]), s_scale[g * 4 + 3], s_shift[g * 4 + 3]);
        if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
        f.x = tf32_rna(f.x); f.y = tf32_rna(f.y); f.z = tf32_rna(f.z); f.w = tf32_rna(f.w);
        const uint32_t addr = row_addr + ((((uint32_t)g) ^ xr) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        // this warp's 32 rows = output rows hb*16 + 4q .. +3, columns wb*8 .. +7 (clipped at Ho / Wo by the map)
        tma_store_5d(&tmO, staging0 + sb * kStemTf32StagingBytes + (uint32_t)q * 4096u, half * 32, wb * 8, hb * 16 + q * 4, to, n);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ------------------------------------------------------------------------------------------------
// The same stem on CTA PAIRS (tcgen05 cta_group::2): stem_tf32_kernel's N = 32 MMAs spend 40 cycles fetching operands for 16
// cycles of math.  Here the two CTAs of a cluster take two spatial tiles (M = 256), each still keeps only ITS half of the
// output channels' weights resident (140 KB) -- and the pair MMA reads the B operand's two halves from both CTAs' shared
// memory, so every MMA is N = 64: A 4 KB + B 2 KB per CTA for 32 cycles of math, and every tile is loaded once instead of twice.
// Protocol as in conv_pair.cuh: both CTAs' TMA loads complete on the LEADER's full barrier; the leader's elected thread issues
// the MMAs; tcgen05.commit multicasts the stage release and the accumulator-ready signal into both CTAs; the epilogue warps of
// both CTAs arrive on the leader's accumulator-empty barrier.  Item i = tiles 2 i (rank 0) and 2 i + 1 (rank 1); an odd CTA
// without a tile loads out-of-range boxes (zeros) and stores nothing.
constexpr int kStemTf32PairThreads = 192;

__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__global__ void __launch_bounds__(kStemTf32PairThreads, 1)
stem_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                      const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const StemTf32Params p) {
  extern __shared__ __align__(1024) uint8_t stem_pair_smem[];
  uint8_t* smem = stem_pair_smem;
  if (smem_u32(smem) & 1023u) __trap();
  const int crank = (int)cluster_ctarank();
  const int ntaps = p.kt * p.kh * 2;                               // virtual taps
  uint8_t* w_smem = smem;                                          // this CTA's 32 output channels of every tap, resident
  uint8_t* staging = smem + ((ntaps * kStemTf32TapBytes + 1023) & ~1023);   // ONE 32 KB staging tile: two [128 px x 32 ch] halves
  uint8_t* stage_base = staging + 2 * kStemTf32StagingBytes;
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);  // used in the leader only
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kStemMaxStages;            // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                    // [2] leader only: arrivals from both CTAs' epilogue warps
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;
  const int i_first = (int)(blockIdx.x >> 1), i_step = (int)(gridDim.x >> 1);
  const int n_items = (p.num_tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * 4);   // four epilogue warps in each CTA
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    // this CTA's half of the weights (rows 32 crank .. + 31 of every virtual tap)
    mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTf32TapBytes));
    for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemTf32TapBytes, &tmW, w_bar, tap * 16, crank * 32);
    mbar_wait(w_bar, 0);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 128);   // two 64-column accumulators per CTA
    tmem_relinquish_pair();
  }
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    if (t < 64) {
      s_scale[t] = p.scale[t];
      s_shift[t] = p.shift[t];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers, TMEM and resident weights are in place before anything is signalled or multiplied
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t box_e = (uint32_t)p.box_bytes[0], box_o = (uint32_t)p.box_bytes[1];
  auto tile_coords = [&](int tile, int& wb, int& hb, int& to, int& n) {
    int r = tile;
    wb = r % p.tiles_w; r /= p.tiles_w;
    hb = r % p.tiles_h; r /= p.tiles_h;
    to = r % p.To;
    n = r / p.To;   // == B for the tile-less odd CTA of the last item: every box out of range
  };
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread, both CTAs)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t lfull0 = mapa_u32(full0, 0);
      const uint32_t tx = (uint32_t)(2 * (p.rows_even + p.rows_odd) * kStemTf32SegBytes);
      uint32_t s = 0, ph = 0;
      for (int item = i_first; item < n_items; item += i_step) {
        int wb, hb, to, n;
        tile_coords(2 * item + crank, wb, hb, to, n);
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 4;
        const int t0 = to * p.st - p.pt;
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;   // taps inside the clip
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t dst = stage0 + s * (uint32_t)p.stage_bytes;
          const uint32_t fb = lfull0 + s * 8;
          if (crank == 0) mbar_arrive_expect_tx_a(full0 + s * 8, 2u * tx);   // both CTAs' bytes
          tma_load_5d_pair(dst, &tmE, fb, x_start, 0, h_start, t0 + dt, n);
          tma_load_5d_pair(dst + box_e, &tmE, fb, x_start, 1, h_start, t0 + dt, n);
          tma_load_5d_pair(dst + 2 * box_e, &tmOdd, fb, x_start, 0, h_start + 1, t0 + dt, n);
          tma_load_5d_pair(dst + 2 * box_e + box_o, &tmOdd, fb, x_start, 1, h_start + 1, t0 + dt, n);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (crank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_tf32_m256(64);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      const uint32_t seg16 = kStemTf32SegBytes >> 4;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, kStemTf32SegBytes);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      uint32_t s = 0, ph = 0, tc = 0;
      for (int item = i_first; item < n_items; item += i_step, ++tc) {
        const uint32_t acc = tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, ((tc >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained it
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64u;
        const int t0 = (((2 * item) / (p.tiles_w * p.tiles_h)) % p.To) * p.st - p.pt;
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(full0 + s * 8, ph);
          tc_fence_after();
          const uint32_t st16 = (stage0 + s * (uint32_t)p.stage_bytes) >> 4;
          uint32_t b_lo = w16 + (uint32_t)(dt * p.kh * 2) * (kStemTf32TapBytes >> 4);
          for (int dh = 0; dh < p.kh; ++dh) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
              const uint32_t box16 = ((dh & 1) ? 2 * box_e + (uint32_t)pl * box_o : (uint32_t)pl * box_e) >> 4;
              const uint64_t adesc = a_hi | (st16 + box16 + (uint32_t)(dh >> 1) * seg16);
              const uint64_t bdesc = b_hi | b_lo;
              umma_tf32_pair(d_tmem, adesc, bdesc, idesc, (dt > dt_lo || dh || pl) ? 1u : 0u);
              umma_tf32_pair(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              b_lo += kStemTf32TapBytes >> 4;
            }
          }
          umma_commit_pair(empty0 + s * 8);                             // frees the slot in both CTAs
          if (dt == dt_hi - 1) umma_commit_pair(tfull0 + acc * 8);      // both CTAs' accumulators complete
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5 (both CTAs)
    const int q = warp & 3;
    const int lrow = q * 32 + lane;
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    const uint32_t ltempty0 = mapa_u32(smem_u32(tmem_empty_bar), 0);
    uint32_t tc = 0;
    for (int item = i_first; item < n_items; item += i_step, ++tc) {
      const int tile = 2 * item + crank;
      int wb, hb, to, n;
      tile_coords(tile, wb, hb, to, n);
      const uint32_t acc = tc & 1u;
      const uint32_t row_addr = staging0 + (uint32_t)lrow * 128u;
      // single staging tile: the store issued one tile ago (a whole tile time back) must have been read out
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1u);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u, v0);
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);   // hands this CTA's accumulator back to the leader
      auto half_tile = [&](const uint32_t (&v)[32], int hc) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int sc = hc * 32 + g * 4;
          float4 f;
          f.x = fmaf(__uint_as_float(v[g * 4 + 0]), s_scale[sc + 0], s_shift[sc + 0]);
          f.y = fmaf(__uint_as_float(v[g * 4 + 1]), s_scale[sc + 1], s_shift[sc + 1]);
          f.z = fmaf(__uint_as_float(v[g * 4 + 2]), s_scale[sc + 2], s_shift[sc + 2]);
          f.w = fmaf(__uint_as_float(v[g * 4 + 3]), s_scale[sc + 3], s_shift[sc + 3]);
          if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
          f.x = tf32_rna(f.x); f.y = tf32_rna(f.y); f.z = tf32_rna(f.z); f.w = tf32_rna(f.w);
          const uint32_t addr = row_addr + (uint32_t)hc * kStemTf32StagingBytes + ((((uint32_t)g) ^ xr) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
        }
      };
      half_tile(v0, 0);
      half_tile(v1, 1);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (tile < p.num_tiles) {
          tma_store_5d(&tmO, staging0 + (uint32_t)q * 4096u, 0, wb * 8, hb * 16 + q * 4, to, n);
          tma_store_5d(&tmO, staging0 + kStemTf32StagingBytes + (uint32_t)q * 4096u, 32, wb * 8, hb * 16 + q * 4, to, n);
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while the peer may still signal into this CTA or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 128);
  }
}

// fp32 NCTHW -> column-parity planes [B, T, H, 2, Wh, 4] with Wh = (W + 8) / 2: one thread per output pixel slot
__global__ void ingest_ncthw_f32_to_planes_kernel(const float* __restrict__ x, int B, long long th, int W, float4* __restrict__ out) {
  const int Wh = (W + 8) >> 1;
  const long long total = (long long)B * th * 2 * Wh;
  const long long plane_sz = th * W;   // one channel of one clip
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)(i % Wh);
    long long r = i / Wh;
    const int pl = (int)(r & 1);
    r >>= 1;                           // (b * T + t) * H + h
    const long long b = r / th, row = r - b * th;
    const int xs = 2 * pos + pl - 3;   // source column of padded pixel 2 pos + pl
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xs >= 0 && xs < W) {
      const float* px = x + b * 3 * plane_sz + row * W + xs;
      o.x = tf32_rna(px[0]); o.y = tf32_rna(px[plane_sz]); o.z = tf32_rna(px[2 * plane_sz]);
    }
    out[i] = o;
  }
}

struct PoolF32Params {
  const float* in;
  float* out;  // already offset by dst_c_off
  int B, Ti, Hi, Wi, C;
  int To, Ho, Wo;
  int kt, kh, kw, st, sh, sw;
  int pt, ph, pw;
  int pad_zero;  // 1: out-of-range taps contribute 0 (SAME-padding port), 0: they are ignored (-inf, torch MaxPool3d)
  int ldo;
};

// un-padded windows that fit the input (I3Res50's maxpool1 (2,3,3)/2 and maxpool2 (2,1,1)/(2,1,1)): compile-time extents, every
// 128-bit load of a thread issued before the first max (the general kernel below chains load -> max and spends its time in
// 64-bit index arithmetic: 0.52 / 0.37 ms at 64 clip-crops against 0.28 / 0.19 ms of HBM time)
template <int KT, int KH, int KW>
__global__ void __launch_bounds__(256) maxpool3d_f32_fixed_kernel(const PoolF32Params p) {
  const int cv = p.C >> 2;
  const long long total = (long long)p.B * p.To * p.Ho * p.Wo * cv;
  const long long sW = p.C, sH = (long long)p.Wi * p.C, sT = sH * p.Hi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const long long m_out = m;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const int to = (int)(m % p.To); m /= p.To;
    const float* base = p.in + (m * p.Ti + (long long)to * p.st) * sT + (long long)ho * p.sh * sH + (long long)wo * p.sw * sW + v * 4;
    float4 x[KT * KH * KW];
#pragma unroll
    for (int dt = 0; dt < KT; ++dt)
#pragma unroll
      for (int dh = 0; dh < KH; ++dh)
#pragma unroll
        for (int dw = 0; dw < KW; ++dw) {
          const uint4 u = ld_stream_16(base + dt * sT + dh * sH + dw * sW);
          x[(dt * KH + dh) * KW + dw] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
        }
    float4 acc = x[0];
#pragma unroll
    for (int k = 1; k < KT * KH * KW; ++k) {
      acc.x = fmaxf(acc.x, x[k].x); acc.y = fmaxf(acc.y, x[k].y); acc.z = fmaxf(acc.z, x[k].z); acc.w = fmaxf(acc.w, x[k].w);
    }
    *reinterpret_cast<float4*>(p.out + m_out * p.ldo + v * 4) = acc;
  }
}

__global__ void __launch_bounds__(256) maxpool3d_f32_kernel(const PoolF32Params p) {
  const int cv = p.C >> 2;
  const long long total = (long long)p.B * p.To * p.Ho * p.Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const long long m_out = m;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const int to = (int)(m % p.To); m /= p.To;
    const long long b = m;
    const float ninf = __int_as_float(0xff800000);
    float4 acc = make_float4(ninf, ninf, ninf, ninf);
    bool any_oob = false;
    for (int dt = 0; dt < p.kt; ++dt) {
      const int ti = to * p.st - p.pt + dt;
      for (int dh = 0; dh < p.kh; ++dh) {
        const int hi = ho * p.sh - p.ph + dh;
        for (int dw = 0; dw < p.kw; ++dw) {
          const int wi = wo * p.sw - p.pw + dw;
          if ((unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi) {
            const float4 x = *reinterpret_cast<const float4*>(p.in + ((((b * p.Ti + ti) * p.Hi + hi) * p.Wi + wi) * (long long)p.C) + v * 4);
            acc.x = fmaxf(acc.x, x.x); acc.y = fmaxf(acc.y, x.y); acc.z = fmaxf(acc.z, x.z); acc.w = fmaxf(acc.w, x.w);
          } else {
            any_oob = true;
          }
        }
      }
    }
    if (any_oob && p.pad_zero) {
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    }
    *reinterpret_cast<float4*>(p.out + m_out * p.ldo + v * 4) = acc;
  }
}

// in [B, P, C] fp32 -> out [B, C] fp32: lane = pg * 8 + cv reads channels [c0 + 4 cv, +4) at positions pg, pg + 4, ...
__global__ void __launch_bounds__(256) avgpool_f32_kernel(const float* __restrict__ in, int B, int P, int C, float* __restrict__ out) {
  const int warps_per_clip = C >> 5;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)B * warps_per_clip) return;
  const int b = (int)(gw / warps_per_clip);
  const int c0 = (int)(gw % warps_per_clip) * 32 + (lane & 7) * 4;
  const int pg = lane >> 3;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* base = in + (long long)b * P * C + c0;
  for (int pos = pg; pos < P; pos += 4) {
    const float4 x = *reinterpret_cast<const float4*>(base + (long long)pos * C);
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
  }
  float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
  }
  if (pg == 0) {
    const float inv = 1.f / (float)P;
    *reinterpret_cast<float4*>(out + (long long)b * C + c0) = make_float4(a[0] * inv, a[1] * inv, a[2] * inv, a[3] * inv);
  }
}

// x [B, 3, T, H, W] fp32 (what the reference hands its model, extract_features.py:86) -> [B, T, H, W, 4] fp32, channel 3 = 0
__global__ void __launch_bounds__(256) ingest_ncthw_f32_to_ndhwc4_kernel(const float* __restrict__ x, int B, long long thw, float4* __restrict__ out) {
  const long long total = (long long)B * thw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / thw, r = i - b * thw;
    const float* src = x + b * 3 * thw + r;
    out[i] = make_float4(tf32_rna(src[0]), tf32_rna(src[thw]), tf32_rna(src[2 * thw]), 0.f);
  }
}

}  // namespace vad
