// K2s: the I3D stem (conv 5x7x7, stride 2, 3 -> 64 channels + BN + ReLU; src/i3d.py:202-209,303-305), with the
// temporal half of maxpool1 (src/i3d.py:212-214,306) optionally fused into its epilogue, as a dedicated
// tcgen05 kernel.
//
// The generic implicit-GEMM kernel loads one 128-pixel x 64-byte im2col column per (dt, dh) tap, i.e. it pulls
// every input pixel ~35 times through L2 -- the stem is then L2-bandwidth bound, not tensor bound.  Here an
// output tile is an 8 (w) x 16 (h) block of one output frame and NO im2col copy is made at all:
//
//  * Input layout [N, T, H, Wp, 4] bf16 (RGB + zero channel, 3 zero pixels left of every row).  For one input
//    row the 8 windows of 8 pixels x 4 channels (= 64 B, kw folded into the contraction) of 8 consecutive
//    output columns start 16 B apart, overlap, and together cover one contiguous 176-byte segment.
//  * A K-major SWIZZLE_NONE UMMA descriptor addresses its operand as
//        addr(row r, 16-byte K chunk j) = start + (r % 8) * 16 + (r / 8) * SBO + j * LBO
//    so with LBO = 16 B (the window stride) and SBO = the 176-byte segment pitch the tensor core reads the
//    sliding windows straight out of the raw segments: row (h_i, w_i) = h_i * 8 + w_i sees bytes
//    [16 * w_i, 16 * w_i + 64) of segment h_i.
//  * For a fixed dt all even dh taps share ONE TMA box of raw segments (input rows 2*ho0 - 3 + 2i, stride 2)
//    and all odd taps another; tap dh is the same box at a segment offset dh / 2.  A (dt) stage is
//    19 + 18 segments = 6.5 KB instead of 35 KB of im2col columns; TMA zero-fills out-of-range rows / frames,
//    which implements the conv padding.
//  * The 35 x 4 KB of weights stay resident in shared memory for the whole persistent CTA.
//
// Epilogue: TMEM -> registers -> BN scale/shift + ReLU -> bf16 -> 128B-swizzled staging tile in shared memory ->
// one TMA store per warp (32 rows = 4 output rows x 8 columns x 128 B).  With pool_t == 2 a CTA computes output
// frames 2k and 2k+1 of the same spatial tile back to back; the first is parked in the staging tile, the second
// is max-reduced into it (bf16 max is exact, and max over (2,3,3) separates into max over t then over (3,3)),
// and only the pooled frame is written: the stem's 2 GB of output per 160 clips and the pool's re-read halve.
#pragma once

#include <cuda_bf16.h>

#include "conv_umma.cuh"

namespace vad {

struct StemParams {
  int B, To, Ho, Wo;        // conv output extent (To: frames the convolution produces)
  int To_out;               // frames written: To, or To / 2 with pool_t == 2
  int pool_t;               // 1, or 2: max over output frames (2k, 2k+1) in the epilogue
  int kt, kh, st, pt, ph;   // sh == sw == 2, kw folded
  int tiles_w, tiles_h;
  int num_units;            // B * To_out * tiles_h * tiles_w; a unit is pool_t tiles
  int rows_even, rows_odd;  // box rows: 16 + (#taps of that parity) - 1
  int off_odd;              // byte offset of the odd-row box inside a stage
  int stage_bytes;
  int seg_bytes;            // bytes of one raw row segment (176 for stride 2)
  int n_stages;
  int w_stream;             // 0: all kt x kh taps of weights resident in shared memory (<= 150 KB: the 5x7x7 stem);
                            // 1: the kh taps of frame tap dt travel with dt's stage (InceptionI3d's 7x7x7: 196 KB)
  int off_w;                // w_stream: byte offset of the kh x 4 KB of weights inside a stage (1024-aligned)
  int relu;
  int Ti;   // input frames when frame taps outside the clip may be skipped (stem_pair.cuh), else 0
  int dbg;  // VAD_STEM_DEBUG bit mask (bottleneck hunting only): 1 = no global stores, 2 = no MMA issue, 4 = no A loads,
            // multi-frame kernel also: 8 = no tcgen05.ld, 16 = no tcgen05.st zeroing, 32 = no epilogue math / staging
  const float* scale;
  const float* shift;
  long long* clk_out;  // VAD_STEM_CLOCKS=1: {SM cycles, nanoseconds} spent by CTA 0's MMA thread (debug)
};

constexpr int kStemThreads = 192;
constexpr int kStemMfThreads = 320;  // multi-frame kernel: 8 epilogue warps
constexpr int kStemTapBytes = 64 * 64;        // one tap of weights: 64 output channels x 32 bf16
constexpr int kStemMaxStages = 8;
constexpr int kStemStagingBytes = 128 * 128;  // one output tile: 128 pixels x 64 channels bf16

__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

__global__ void __launch_bounds__(kStemThreads, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int ntaps = p.kt * p.kh;
  uint8_t* w_smem = smem;                                      // ntaps x 4 KB, resident (nothing when w_stream)
  uint8_t* staging = smem + (p.w_stream ? 0 : ntaps * kStemTapBytes);  // 2 x 16 KB output staging (1024-aligned)
  uint8_t* stage_base = staging + 2 * kStemStagingBytes;       // n_stages x stage_bytes
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kStemMaxStages;
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
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);  // one arrival per epilogue warp
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
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      if (!p.w_stream) {
        mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTapBytes));
        for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemTapBytes, &tmW, w_bar, tap * 32, 0);
      }
      griddep_wait();  // the weights are constants; the clips come from the preceding preprocessing kernel
      const uint32_t tx = (uint32_t)((p.rows_even + p.rows_odd) * p.seg_bytes) + (p.w_stream ? (uint32_t)(p.kh * kStemTapBytes) : 0u);
      uint32_t s = 0, ph = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int r = unit;
        const int wb = r % p.tiles_w; r /= p.tiles_w;
        const int hb = r % p.tiles_h; r /= p.tiles_h;
        const int tu = r % p.To_out;
        const int n = r / p.To_out;
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 8;  // 8 windows x (2 px x 4 ch) elements
        for (int sub = 0; sub < p.pool_t; ++sub) {
          const int t0 = (tu * p.pool_t + sub) * p.st - p.pt;
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
              if (p.w_stream)
                for (int dh = 0; dh < p.kh; ++dh)
                  tma_load_2d_a(dst + (uint32_t)p.off_w + (uint32_t)dh * kStemTapBytes, &tmW, fb, (dt * p.kh + dh) * 32, 0);
            }
            if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
          }
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
      if (!p.w_stream) mbar_wait(w_bar, 0);
      const uint32_t seg = (uint32_t)p.seg_bytes;
      const uint32_t w_addr = smem_u32(w_smem);
      // descriptor high words are loop invariant; only the 14-bit start-address field moves
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, seg);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      const int kh = (p.dbg & 2) ? 0 : p.kh;
      const int wait_dh = (kh * 3) / 4;  // wait for the next stage after this many taps
      uint32_t s = 0, ph = 0;
      uint32_t tc = 0;
      mbar_wait_a(full0, 0);
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const bool last_unit = unit + (int)gridDim.x >= p.num_units;
        for (int sub = 0; sub < p.pool_t; ++sub, ++tc) {
          const uint32_t acc = tc & 1u;
          mbar_wait_a(tempty0 + acc * 8, ((tc >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 64u;
          const bool last_tile = last_unit && sub == p.pool_t - 1;
          for (int dt = 0; dt < p.kt; ++dt) {
            uint32_t ns = s + 1, nph = ph;
            if (ns == (uint32_t)S) { ns = 0; nph ^= 1u; }
            const bool do_wait = !(last_tile && dt == p.kt - 1);
            const uint32_t st_addr = stage0 + s * (uint32_t)p.stage_bytes;
            uint32_t b_lo = (p.w_stream ? st_addr + (uint32_t)p.off_w : w_addr + (uint32_t)(dt * p.kh) * kStemTapBytes) >> 4;
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
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5
    griddep_wait();
    const int q = warp & 3;
    const int lrow = q * 32 + lane;  // tile row = TMEM lane: (h_i, w_i) = (lrow / 8, lrow % 8)
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    uint32_t tc = 0, sb = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, sb ^= 1u) {
      int r = unit;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h; r /= p.tiles_h;
      const int tu = r % p.To_out;
      const int n = r / p.To_out;
      const uint32_t row_addr = staging0 + sb * kStemStagingBytes + (uint32_t)lrow * 128u;
      // the TMA store that read this staging buffer two units ago must have drained it
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      for (int sub = 0; sub < p.pool_t; ++sub, ++tc) {
        const uint32_t acc = tc & 1u;
        mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(taddr, v0);
        tmem_ld_32x32(taddr + 32u, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);  // accumulator is in registers: hand it back
        auto chunk = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            const uint32_t addr = row_addr + ((((uint32_t)col >> 3) ^ xr) << 4);
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
            uint32_t o[4] = {pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])};
            if (sub > 0) {
              uint32_t e[4];
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "r"(addr));
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&o[j]),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&e[j]));
                o[j] = *reinterpret_cast<const uint32_t*>(&m);
              }
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        };
        chunk(v0, 0);
        chunk(v1, 1);
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        // this warp's 32 rows = output rows hb*16 + 4q .. +3, columns wb*8 .. +7 (clipped at Ho / Wo by the map)
        if (!(p.dbg & 1))
          tma_store_5d(&tmO, staging0 + sb * kStemStagingBytes + (uint32_t)q * 4096u, 0, wb * 8, hb * 16 + q * 4, tu, n);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();  // bulk stores fully complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-frame stem (kt = 5, st = 2, pt = 2, To <= 8): input-frame stationary.
//
// An N = 64 MMA is bound by its operand fetch (A 4 KB + B 2 KB per MMA through a 128 B/clk shared-memory port:
// 48 cycles against the pipe's 32).  But one input frame ti feeds 2-3 output frames (to = (ti + 2 - dt) / 2 for
// every dt of ti's parity), with the same A windows and different weights.  So the CTA keeps ALL output frames
// of its spatial tile in TMEM (8 x 64 columns = the whole 512), walks the input frames once, and issues per
// (dh, K half) a single MMA with N = 64 x (#output frames fed): the B operand is the weights of dt = 4, 2, 0
// (even ti) or 3, 1 (odd ti) stacked along N -- they are laid out that way in shared memory --, D is the
// matching run of consecutive accumulator blocks.  A is fetched once per 128/192 output columns instead of
// once per 64, out-of-range input frames are skipped instead of multiplied as zeros, and every input frame is
// loaded once per spatial tile instead of 2.5 times.  Since an MMA's accumulate flag covers all its columns, a
// block cannot be (re)initialised by its first MMA: the epilogue zeroes each block (tcgen05.st) right after
// draining it, and every MMA accumulates.
struct StemMfParams {
  StemParams s;
  int Ti;       // input frames
  int ti_max;   // last input frame any output frame reads: min(Ti - 1, 2 * (To - 1) + 2)
};

__device__ __forceinline__ void tmem_st_zero_32x32(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kStemMfThreads, 1)
stem_umma_mf_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const StemMfParams pp) {
  const StemParams& p = pp.s;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int ntaps = 5 * p.kh;
  uint8_t* w_smem = smem;                                      // even dt: [dh][dt = 4, 2, 0] x 4 KB, then odd dt: [dh][dt = 3, 1]
  uint8_t* staging = smem + ntaps * kStemTapBytes;             // 2 x 16 KB output staging (1024-aligned)
  uint8_t* stage_base = staging + 2 * kStemStagingBytes;       // n_stages x stage_bytes
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* acc_full_bar = empty_bar + kStemMaxStages;   // [8] output frame complete
  uint64_t* acc_empty_bar = acc_full_bar + 8;            // [8] accumulator block drained and zeroed
  uint64_t* w_bar = acc_empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  uint8_t* stage_tab = reinterpret_cast<uint8_t*>(tmem_slot + 4);  // 32 rows x 32 B, see the MMA issuer

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;
  const uint32_t w_odd_off = (uint32_t)(p.kh * 3) * kStemTapBytes;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 8; ++a) {
      mbar_init(&acc_full_bar[a], 1);
      mbar_init(&acc_empty_bar[a], 8);  // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);  // eight 64-column accumulator blocks: one per output frame
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
  if (warp >= 2) {
    // every MMA accumulates: start from zeroed accumulators
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 2) >> 2) * 32);
    for (int c = 0; c < 8; ++c) tmem_st_zero_32x32(t0 + (uint32_t)(c * 64));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemTapBytes));
      for (int dt = 0; dt < 5; ++dt)
        for (int dh = 0; dh < p.kh; ++dh) {
          uint8_t* dst = (dt & 1) ? w_smem + w_odd_off + (uint32_t)(dh * 2 + (3 - dt) / 2) * kStemTapBytes
                                  : w_smem + (uint32_t)(dh * 3 + (4 - dt) / 2) * kStemTapBytes;
          tma_load_2d(dst, &tmW, w_bar, (dt * p.kh + dh) * 32, 0);
        }
      griddep_wait();  // weights are constants; the clips come from the preceding preprocessing kernel
      const uint32_t tx = (uint32_t)((p.rows_even + p.rows_odd) * p.seg_bytes);
      uint32_t s = 0, ph = 0;
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        int r = unit;
        const int wb = r % p.tiles_w; r /= p.tiles_w;
        const int hb = r % p.tiles_h;
        const int n = r / p.tiles_h;
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 8;
        for (int ti = 0; ti <= pp.ti_max; ++ti) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t dst = stage0 + s * (uint32_t)p.stage_bytes;
          const uint32_t fb = full0 + s * 8;
          if (p.dbg & 4) {
            mbar_arrive_a(fb);
          } else {
            mbar_arrive_expect_tx_a(fb, tx);
            tma_load_4d_a(dst, &tmE, fb, x_start, h_start, ti, n);
            tma_load_4d_a(dst + (uint32_t)p.off_odd, &tmOdd, fb, x_start, h_start + 1, ti, n);
          }
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    // The tensor pipe queues only about one MMA beyond the one it executes, so whatever this thread does between two
    // UTCHMMA issues must fit in one MMA time (48-96 cycles) or the pipe starves.  Everything that depends only on
    // the input-frame index (accumulator run, weight block, which accumulators to wait for / hand over) comes from a
    // small table built once in shared memory; the next stage's table row, operand addresses and barrier waits are
    // fetched between the MMAs of the current stage (the empty asm statements pin that placement).
    struct Row { int d_off, nb, b_base, wait0, wait1, commit0, commit1, pad; };
    Row* tab = reinterpret_cast<Row*>(stage_tab);
    if (lane <= pp.ti_max) {
      const int ti = lane;
      const int odd = ti & 1;
      const int c = (ti - odd) >> 1;
      int lo = odd ? c : c - 1;
      const int top = c + 1 < p.To - 1 ? c + 1 : p.To - 1;
      int skip = 0;
      if (lo < 0) { lo = 0; skip = 1; }
      Row rw;
      rw.d_off = lo * 64;
      rw.nb = top - lo + 1;
      rw.b_base = (int)(((odd ? w_odd_off : 0u) + (uint32_t)skip * kStemTapBytes) >> 4);
      // accumulator blocks this frame touches first: dt = 0 of output frame c + 1 (even ti); at ti = 0 also frame 0
      rw.wait0 = (!odd && ti == 0) ? 0 : -1;
      rw.wait1 = (!odd && c + 1 <= p.To - 1) ? c + 1 : -1;
      // output frames whose last contribution this frame is: dt = 4 of frame c - 1 (even ti >= 2); at ti_max the rest
      const int ev_max = pp.ti_max & ~1;
      rw.commit0 = (!odd && ti >= 2) ? c - 1 : -1;
      rw.commit1 = -1;
      if (ti == pp.ti_max) {
        const int first_left = (ev_max - 2) / 2 + 1;  // frames (ev_max - 2) / 2 and below were finished by even frames
        if (!odd && ti >= 2) { rw.commit1 = (first_left + 0 <= p.To - 1 && first_left > c - 1) ? first_left : -1; }
        else rw.commit0 = first_left <= p.To - 1 ? first_left : -1;
      }
      rw.pad = 0;
      tab[ti] = rw;
    }
    __syncwarp();
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t afull0 = smem_u32(acc_full_bar), aempty0 = smem_u32(acc_empty_bar);
      mbar_wait(w_bar, 0);
      const uint32_t seg = (uint32_t)p.seg_bytes, seg16 = seg >> 4;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, seg);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      const int kh = (p.dbg & 2) ? 0 : p.kh;
      const uint32_t off_odd16 = (uint32_t)p.off_odd >> 4, stage16 = (uint32_t)p.stage_bytes >> 4;
      long long clk0 = 0, ns0 = 0;
      if (p.clk_out && blockIdx.x == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
      }
      uint32_t s = 0, ph = 0, uc = 0;
      Row cur = tab[0];
      uint32_t a_even = stage0 >> 4;
      // prologue of the software pipeline: the first stage's operands and accumulators
      mbar_wait_a(full0, 0);
      if (cur.wait0 >= 0) mbar_wait_a(aempty0 + (uint32_t)cur.wait0 * 8, 1u);
      if (cur.wait1 >= 0) mbar_wait_a(aempty0 + (uint32_t)cur.wait1 * 8, 1u);
      tc_fence_after();
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
        const bool last_unit = unit + (int)gridDim.x >= p.num_units;
        for (int ti = 0; ti <= pp.ti_max; ++ti) {
          const bool last_stage = last_unit && ti == pp.ti_max;
          const int nti = ti == pp.ti_max ? 0 : ti + 1;
          const uint32_t nuc = ti == pp.ti_max ? uc + 1 : uc;
          uint32_t ns = s + 1, nph = ph;
          if (ns == (uint32_t)S) { ns = 0; nph ^= 1u; }
          const uint32_t a_odd = a_even + off_odd16;
          const uint32_t d_tmem = tmem_base + (uint32_t)cur.d_off;
          const uint32_t b_lo = w16 + (uint32_t)cur.b_base;
          Row nxt;
          auto taps = [&](auto n_c, auto odd_c, int dh0, int dh1) {
            constexpr uint32_t idesc = umma_idesc_bf16_m128(decltype(n_c)::value);
            constexpr uint32_t per_dh = (decltype(odd_c)::value ? 2u : 3u) * (kStemTapBytes >> 4);
#pragma unroll
            for (int dh = 0; dh < 7; ++dh) {
              if (dh >= dh0 && dh < dh1 && dh < kh) {
                const uint32_t a_lo = ((dh & 1) ? a_odd : a_even) + (uint32_t)(dh >> 1) * seg16;
                const uint64_t adesc = a_hi | a_lo;
                const uint64_t bdesc = b_hi | (b_lo + (uint32_t)dh * per_dh);
                umma_f16_c<true>(d_tmem, adesc, bdesc, idesc);
                umma_f16_c<true>(d_tmem, adesc + 2, bdesc + 2, idesc);
              }
            }
          };
          auto stage = [&](auto n_c, auto odd_c) {
            taps(n_c, odd_c, 0, 2);
            asm volatile("" ::: "memory");
            nxt = tab[nti];                       // next stage's table row
            taps(n_c, odd_c, 2, 4);
            asm volatile("" ::: "memory");
            if (!last_stage) {                    // accumulators the next stage touches first
              if (nxt.wait0 >= 0) mbar_wait_a(aempty0 + (uint32_t)nxt.wait0 * 8, (nuc & 1u) ^ 1u);
              if (nxt.wait1 >= 0) mbar_wait_a(aempty0 + (uint32_t)nxt.wait1 * 8, (nuc & 1u) ^ 1u);
            }
            taps(n_c, odd_c, 4, 5);
            asm volatile("" ::: "memory");
            if (!last_stage) {                    // next stage's operands
              mbar_wait_a(full0 + ns * 8, nph);
              tc_fence_after();
            }
            taps(n_c, odd_c, 5, 7);
          };
          if (ti & 1) {
            if (cur.nb == 2) stage(std::integral_constant<int, 128>{}, std::true_type{});
            else             stage(std::integral_constant<int, 64>{}, std::true_type{});
          } else {
            if (cur.nb == 3)      stage(std::integral_constant<int, 192>{}, std::false_type{});
            else if (cur.nb == 2) stage(std::integral_constant<int, 128>{}, std::false_type{});
            else                  stage(std::integral_constant<int, 64>{}, std::false_type{});
          }
          umma_commit_a(empty0 + s * 8);
          if (cur.commit0 >= 0) umma_commit_a(afull0 + (uint32_t)cur.commit0 * 8);
          if (cur.commit1 >= 0) umma_commit_a(afull0 + (uint32_t)cur.commit1 * 8);
          cur = nxt;
          s = ns; ph = nph;
          a_even = ns == 0 ? (stage0 >> 4) : a_even + stage16;
        }
      }
      if (p.clk_out && blockIdx.x == 0) {
        long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        p.clk_out[0] = clock64() - clk0;
        p.clk_out[1] = ns1 - ns0;
        p.clk_out[2] = uc;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    // two warps per TMEM lane quarter, 32 of the 64 channels each; the pair shares one 32-row x 128-byte slice of the
    // staging tile (named barrier 1 + q) and warp 2 + q issues the TMA store
    griddep_wait();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool issuer = half == 0;
    const int lrow = q * 32 + lane;
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    uint32_t uc = 0, sb = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++uc) {
      int r = unit;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h;
      const int n = r / p.tiles_h;
      for (int to = 0; to < p.To; ++to) {
        const int sub = p.pool_t == 2 ? (to & 1) : 0;
        const bool dropped = p.pool_t == 2 && to == p.To - 1 && !(to & 1);  // odd frame count: MaxPool3d floors
        const uint32_t row_addr = staging0 + sb * kStemStagingBytes + (uint32_t)lrow * 128u;
        if (sub == 0 && !dropped) {
          // the TMA store that read this staging buffer two outputs ago must have drained it
          if (issuer && lane == 0) tma_store_wait_read<1>();
          named_bar_sync(1 + q, 64);
        }
        mbar_wait(&acc_full_bar[to], uc & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(to * 64 + half * 32);
        uint32_t v[32];
        if (!(p.dbg & 8)) {
          tmem_ld_32x32(taddr, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (!(p.dbg & 16)) {
          tmem_st_zero_32x32(taddr);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[to]);  // drained and zeroed: the next unit may accumulate into it
        if (dropped || (p.dbg & 32)) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = half * 32 + g * 8;
          const uint32_t addr = row_addr + ((((uint32_t)col >> 3) ^ xr) << 4);
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
            if (p.relu) f[j] = fmaxf(f[j], 0.f);
          }
          uint32_t o[4] = {pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7])};
          if (sub > 0) {
            uint32_t e[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "r"(addr));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&o[j]),
                                               *reinterpret_cast<const __nv_bfloat162*>(&e[j]));
              o[j] = *reinterpret_cast<const uint32_t*>(&m);
            }
          }
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        }
        if (sub == p.pool_t - 1) {
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
          named_bar_sync(1 + q, 64);
          if (issuer && lane == 0) {
            if (!(p.dbg & 1))
              tma_store_5d(&tmO, staging0 + sb * kStemStagingBytes + (uint32_t)q * 4096u, 0, wb * 8, hb * 16 + q * 4, to / p.pool_t, n);
            tma_store_commit();
          }
          sb ^= 1u;
        }
      }
    }
    if (issuer && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vad
