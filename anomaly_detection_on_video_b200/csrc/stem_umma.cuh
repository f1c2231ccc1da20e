// K2s: the I3D stem (conv 5x7x7, stride 2, 3 -> 64 channels + BN + ReLU; src/i3d.py:202-209,303-305) as
// a dedicated tcgen05 kernel.
//
// The generic implicit-GEMM kernel loads one 128-pixel x 64-byte im2col column per (dt, dh) tap, i.e. it
// pulls every input pixel ~35 times through L2 -- the stem is then L2-bandwidth bound, not tensor bound.
// Here an output tile is a 16 (w) x 8 (h) block of one output frame, so for a fixed dt the seven dh taps
// read overlapping input rows: rows 2*ho0 - 3 + dh + 2*i, i = 0..7.  All even dh share ONE TMA box of 11
// rows (stride 2), all odd dh one box of 10 rows; tap dh is the same shared-memory box at a row offset of
// (dh / 2) * 16 pixels = 1024 B (a multiple of the 512-byte SWIZZLE_64B atom, so UMMA descriptors can
// point straight into the box).  L2 -> SM traffic drops from 35 x 8 KB to 5 x 21 KB per tile, and the
// 140 KB of weights stay resident in shared memory for the whole persistent CTA.
//
// A operand: the padded input [N, T, H, Wp, 4] bf16 viewed through a rank-5 *tiled* tensor map as
// (C' = 32, W' = Wo, H, T, N): pixel w' is the 8-pixel x 4-channel window at padded column 2 * w'
// (W' stride 16 B < 64 B extent: windows overlap), kw folded into the contraction like the generic path.
// TMA zero-fills every out-of-range row / frame, which implements the conv padding.
#pragma once

#include "conv_umma.cuh"

namespace vad {

struct StemParams {
  int B, To, Ho, Wo;
  int kt, kh, st, pt, ph;  // sh == sw == 2, kw folded
  int tiles_w, tiles_h, num_tiles;
  int rows_even, rows_odd;  // box rows: 8 + (#taps of that parity) - 1
  int off_odd;              // byte offset of the odd-row box inside a stage
  int stage_bytes;
  int n_stages;
  int relu;
  int ldo;
  int dbg;  // VAD_STEM_DEBUG bit mask (bottleneck hunting only): 1 = no global stores, 2 = no MMA issue, 4 = no A loads
  const float* scale;
  const float* shift;
  __nv_bfloat16* out;
};

constexpr int kStemThreads = 192;
constexpr int kStemTapBytes = 64 * 64;  // one tap of weights: 64 output channels x 32 bf16

__global__ void __launch_bounds__(kStemThreads, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                 const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int ntaps = p.kt * p.kh;
  uint8_t* w_smem = smem;                                  // ntaps x 4 KB, resident
  uint8_t* stage_base = smem + ntaps * kStemTapBytes;      // n_stages x stage_bytes
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);  // two 64-column accumulators
    tmem_relinquish();
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform loop)
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTapBytes));
      for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemTapBytes, &tmW, w_bar, tap * 32, 0);
    }
    __syncwarp();
    const uint32_t tx = (uint32_t)((p.rows_even + p.rows_odd) * 1024);
    int kc = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int r = tile;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h; r /= p.tiles_h;
      const int to = r % p.To;
      const int n = r / p.To;
      const int h_start = 2 * (hb * 8) - p.ph;
      for (int dt = 0; dt < p.kt; ++dt, ++kc) {
        const int s = kc % S;
        mbar_wait(&empty_bar[s], ((kc / S) & 1) ^ 1);
        uint8_t* dst = stage_base + s * p.stage_bytes;
        const int ti = to * p.st - p.pt + dt;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[s], tx);
          tma_load_5d(dst, &tmE, &full_bar[s], 0, wb * 16, h_start, ti, n);
          tma_load_5d(dst + p.off_odd, &tmOdd, &full_bar[s], 0, wb * 16, h_start + 1, ti, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop)
    constexpr uint32_t idesc = umma_idesc_bf16_m128(64);
    mbar_wait(w_bar, 0);
    int kc = 0, tc = 0;
    const uint32_t w_addr = smem_u32(w_smem);
    const uint64_t d_hi = umma_desc_kmajor<64>(0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      const int acc = tc & 1;
      mbar_wait(&tmem_empty_bar[acc], ((tc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
      for (int dt = 0; dt < p.kt; ++dt, ++kc) {
        const int s = kc % S;
        mbar_wait(&full_bar[s], (kc / S) & 1);
        tc_fence_after();
        const uint32_t st_addr = smem_u32(stage_base + s * p.stage_bytes);
        if (elect_one_sync()) {
          uint32_t b_lo = (w_addr + (uint32_t)(dt * p.kh) * kStemTapBytes) >> 4;
          for (int dh = 0; dh < p.kh; ++dh, b_lo += kStemTapBytes >> 4) {
            const uint32_t a_lo = (st_addr + ((dh & 1) ? (uint32_t)p.off_odd : 0u) + (uint32_t)(dh >> 1) * 1024u) >> 4;
            const uint64_t adesc = d_hi | a_lo;
            const uint64_t bdesc = d_hi | b_lo;
            if (dt | dh) umma_f16_c<true>(d_tmem, adesc, bdesc, idesc);
            else         umma_f16_c<false>(d_tmem, adesc, bdesc, idesc);
            umma_f16_c<true>(d_tmem, adesc + 2, bdesc + 2, idesc);
          }
          umma_commit(&empty_bar[s]);
          if (dt == p.kt - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int lrow = q * 32 + lane;  // tile row = TMEM lane: (h_i, w_i) = (lrow / 16, lrow % 16)
    int tc = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      int r = tile;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h; r /= p.tiles_h;
      const int to = r % p.To;
      const int n = r / p.To;
      const int ho = hb * 8 + (lrow >> 4);
      const int wo = wb * 16 + (lrow & 15);
      const bool ok = ho < p.Ho && wo < p.Wo;
      const int acc = tc & 1;
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* out_px = p.out + ((((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo) * (long long)p.ldo;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(out_px + col) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ------------------------------------------------------------------------------------------------
// v2: no im2col copy at all.  An output tile is 8 (w) x 16 (h); for one input row the 8 windows of
// 8 pixels x 4 channels start 16 B apart and overlap, and together cover one contiguous 176-byte segment
// of the padded row.  A K-major SWIZZLE_NONE UMMA descriptor addresses the operand as
//     addr(row r, 16-byte K chunk j) = start + (r % 8) * 16 + (r / 8) * SBO + j * LBO
// so with LBO = 16 B (the window stride) and SBO = the 176-byte segment pitch, the tensor core reads the
// sliding windows straight out of the raw segment: row (h_i, w_i) = h_i * 8 + w_i sees bytes
// [16 * w_i, 16 * w_i + 64) of segment h_i.  A (dt) stage is two TMA boxes of raw segments (even / odd
// input rows, 19 + 18 segments = 6.5 KB) instead of 35 KB of im2col columns.
struct StemV2Params {
  StemParams s;
  int seg_bytes;  // bytes of one raw row segment (176 for stride 2)
};

__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

constexpr int kStemV2MaxStages = 8;

__global__ void __launch_bounds__(kStemThreads, 1)
stem_umma_v2_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                    const __grid_constant__ CUtensorMap tmW, const StemV2Params pp) {
  const StemParams& p = pp.s;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int ntaps = p.kt * p.kh;
  uint8_t* w_smem = smem;
  uint8_t* stage_base = smem + ntaps * kStemTapBytes;
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + kStemV2MaxStages;
  uint64_t* tmem_full_bar = empty_bar + kStemV2MaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
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
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTapBytes));
      for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemTapBytes, &tmW, w_bar, tap * 32, 0);
      const uint32_t tx = (uint32_t)((p.rows_even + p.rows_odd) * pp.seg_bytes);
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int r = tile;
        const int wb = r % p.tiles_w; r /= p.tiles_w;
        const int hb = r % p.tiles_h; r /= p.tiles_h;
        const int to = r % p.To;
        const int n = r / p.To;
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 8;  // 8 windows x (2 px x 4 ch) elements
        const int t0 = to * p.st - p.pt;
        for (int dt = 0; dt < p.kt; ++dt) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t dst = stage0 + s * (uint32_t)p.stage_bytes;
          const uint32_t fb = full0 + s * 8;
          if (p.dbg & 4) {
            mbar_arrive_a(fb);
          } else {
            mbar_arrive_expect_tx_a(fb, tx);
            tma_load_4d_a(dst, &tmE, fb, x_start, h_start, t0 + dt, n);
            tma_load_4d_a(dst + (uint32_t)p.off_odd, &tmOdd, fb, x_start, h_start + 1, t0 + dt, n);
          }
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    // Software pipelined like the generic kernel: the wait for the next dt stage sits inside the 2 * kh MMAs
    // of the current one (tools/umma_bench.cu: N = 64 reaches its 48-cycle floor this way).
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(64);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      mbar_wait(w_bar, 0);
      const uint32_t seg = (uint32_t)pp.seg_bytes;
      const uint32_t w_addr = smem_u32(w_smem);
      // descriptor high words are loop invariant; only the 14-bit start-address field moves
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, seg);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      const int kh = (p.dbg & 2) ? 0 : p.kh;
      const int wait_dh = (kh * 3) / 4;  // wait for the next stage after this many taps
      uint32_t s = 0, ph = 0;
      int tc = 0;
      mbar_wait_a(full0, 0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        const uint32_t acc = (uint32_t)tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, (((uint32_t)tc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64u;
        const bool last_tile = tile + (int)gridDim.x >= p.num_tiles;
        for (int dt = 0; dt < p.kt; ++dt) {
          uint32_t ns = s + 1, nph = ph;
          if (ns == (uint32_t)S) { ns = 0; nph ^= 1u; }
          const bool do_wait = !(last_tile && dt == p.kt - 1);
          const uint32_t st_addr = stage0 + s * (uint32_t)p.stage_bytes;
          uint32_t b_lo = (w_addr + (uint32_t)(dt * p.kh) * kStemTapBytes) >> 4;
          const uint32_t a_even = st_addr >> 4, a_odd = (st_addr + (uint32_t)p.off_odd) >> 4, seg16 = seg >> 4;
          auto tap = [&](int dh) {
            const uint32_t a_lo = ((dh & 1) ? a_odd : a_even) + (uint32_t)(dh >> 1) * seg16;
            const uint64_t adesc = a_hi | a_lo;
            const uint64_t bdesc = b_hi | b_lo;
            umma_f16(d_tmem, adesc, bdesc, idesc, (dt | dh) ? 1u : 0u);
            umma_f16_c<true>(d_tmem, adesc + 2, bdesc + 2, idesc);
            b_lo += kStemTapBytes >> 4;
          };
          if (kh == 7) {
#pragma unroll
            for (int dh = 0; dh < 7; ++dh) {
              tap(dh);
              if (dh == 4 && do_wait) { mbar_wait_a(full0 + ns * 8, nph); tc_fence_after(); }
            }
          } else {
            for (int dh = 0; dh < kh; ++dh) {
              tap(dh);
              if (dh == wait_dh && do_wait) { mbar_wait_a(full0 + ns * 8, nph); tc_fence_after(); }
            }
            if (kh <= wait_dh && do_wait) { mbar_wait_a(full0 + ns * 8, nph); tc_fence_after(); }
          }
          umma_commit_a(empty0 + s * 8);
          if (dt == p.kt - 1) umma_commit_a(tfull0 + acc * 8);
          s = ns; ph = nph;
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int lrow = q * 32 + lane;  // tile row = TMEM lane: (h_i, w_i) = (lrow / 8, lrow % 8)
    int tc = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      int r = tile;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h; r /= p.tiles_h;
      const int to = r % p.To;
      const int n = r / p.To;
      const int ho = hb * 16 + (lrow >> 3);
      const int wo = wb * 8 + (lrow & 7);
      const bool ok = ho < p.Ho && wo < p.Wo && !(p.dbg & 1);
      const int acc = tc & 1;
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* out_px = p.out + ((((long long)n * p.To + to) * p.Ho + ho) * p.Wo + wo) * (long long)p.ldo;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(out_px + col) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace vad
