// K2f: the tail of a layer1 bottleneck in ONE launch -- conv2 (1,3,3) 64 -> 64 + BN + ReLU, then conv3 1x1x1 64 -> 256
// + BN + residual + ReLU (src/i3d.py:104-119) -- so that conv2's output never reaches HBM and conv2's tensor work hides
// under conv3, which is bound by the 256-channel residual read and output write.
//
// Per tile (8 w x 16 h pixels of one frame = 128 GEMM rows, as in conv_s3x3.cuh):
//   conv2   36 MMAs (N = 64) over ONE 10 x 18 halo box (two stages)           -> TMEM acc2 (64 columns, double buffered)
//   E1      acc2 -> BN2 + ReLU -> bf16 -> 128B-swizzled shared-memory tile A3 (exactly the tile conv_s3x3 TMA-stores)
//   conv3   4 MMAs (N = 256, K = 64) with A = A3, B = resident W3           -> TMEM acc3 (256 columns)
//   E2      acc3 -> BN3 + residual + ReLU -> bf16, in four 64-channel chunks through 16 KB swizzled staging tiles
// Warp roles: 0 = halo-box producer, 1 = MMA issuer, 2 = chunk DMA, 3..10 = epilogue (TMEM lane quarter = warp % 4).
// The chunk DMA thread owns a ring of three staging tiles: it TMA-loads the residual chunk (64 ch x 8 w x 16 h) three
// chunks ahead, the epilogue warps add their accumulator columns in place and arrive on the tile's "done" barrier, the
// DMA thread TMA-stores the tile, waits until the store engine has read it and re-arms it with the residual of chunk
// c + 3.  No epilogue thread ever waits for a store, and global traffic is whole 128-byte lines in both directions
// (a first version loaded / stored each thread's pixel row straight from registers: 16-byte accesses at a 512-byte
// stride, 31 sectors per request -- 0.7 / 1.1 ms per launch, three times the HBM time).
//
// DS = true is the first block of the layer (src/i3d.py:262-272): the residual is itself a 1x1x1 conv + BN of the block
// input x.  conv3 then contracts K = 128 = [A3 | X] against [s3 * W3 | sd * Wd] (BN scales folded into bf16 weights by
// fold_tail_weights_kernel at bind time), the epilogue adds (b3 + bd): the 256-channel downsample tensor is neither
// written nor read and its kernel disappears.  With three weight tiles resident there is no room for a staging ring:
// the A3 and X tiles themselves stage the output chunks once conv3 has consumed them (chunk n -> tile n & 1), and the
// DMA thread reloads X for the next tile after the last store has been read.
//
// RES mode is bit-identical to conv_s3x3 followed by the generic conv3 (same MMA order per accumulator, same epilogue
// arithmetic); DS mode differs from the unfused path at bf16 resolution (the residual is no longer rounded to bf16).
#pragma once

#include "conv_s3x3.cuh"

namespace vad {

struct TailParams {
  int F, H, W;            // frames (clips x T), height, width
  int tiles_w, tiles_h, num_tiles;
  int relu2, relu3;
  // BN scale / shift BY VALUE: the kernel parameter block lives in the constant bank, so the epilogue's per-column
  // operands come through the constant cache instead of shared memory (as broadcast LDS.128 they were 60 % of the
  // kernel's shared-memory wavefronts, and the shared-memory data pipe -- MMA operand fetch + staging -- is what bounds it)
  float s2[64], b2[64];     // conv2 BN
  float s3[256], b3[256];   // conv3 BN (DS: s3 unused -- folded into the weights --, b3 = conv3 shift + downsample shift)
  long long* dbg;           // VAD_TAIL_DEBUG: clock64() stamps of CTA 0's role threads for its tiles 8..11 ([4 tiles][32 events]), else null
};
// event slots: 0 IN issued | 1 conv3 operands ready, 2 conv3 issued, 3 conv2 operands ready, 4 conv2 issued |
// 5 + 3 n: DMA chunk n done-wait over, stored, store read | 17 acc2_full seen, 18 E1 done, 19 acc3_full seen, 20 + 2 n: chunk n ready, chunk n done
#define TAIL_STAMP(tc_, ev_) do { if (p.dbg && blockIdx.x == 0 && (tc_) >= 8u && (tc_) < 12u) p.dbg[((tc_) - 8u) * 32u + (ev_)] = clock64(); } while (0)

constexpr int kTailW2Bytes = 9 * 64 * 128;      // 72 KB: nine taps x (64 cout x 64 cin)
constexpr int kTailW3Bytes = 256 * 128;         // 32 KB: 256 cout x 64 cin
constexpr int kTailInBytes = kS3HaloBytes;      // 23 KB: ONE 10 x 18 halo box per tile (conv_s3x3.cuh)
constexpr int kTailChunkBytes = 128 * 128;      // 16 KB: 128 pixels x 64 channels (A3, X, one staging tile)
constexpr int kTailThreads = 96 + 8 * 32;
constexpr int kTailL2Ahead = 2;                 // tiles of halo box / X / residual pulled into L2 ahead of their smem loads
// IN_STAGES halo-box stages, RING staging tiles.  RES: the ring carries the residual in and the result out, its depth is
// the residual prefetch distance; DS: the ring only stages the output.
template <bool DS, int IN_STAGES, int RING>
struct TailCfg {
  static_assert(IN_STAGES == 1 || IN_STAGES == 2, "one or two halo-box stages");
  static_assert(RING >= 2 && RING <= 5, "staging ring depth");
  static constexpr int kOffW3 = kTailW2Bytes;
  static constexpr int kOffWd = kOffW3 + kTailW3Bytes;                    // DS only
  static constexpr int kOffIn = kOffWd + (DS ? kTailW3Bytes : 0);
  static constexpr int kOffA3 = kOffIn + IN_STAGES * kTailInBytes;
  static constexpr int kOffX = kOffA3 + kTailChunkBytes;                  // DS only: block-input tile X
  static constexpr int kOffRing = kOffX + (DS ? kTailChunkBytes : 0);
  static constexpr int kOffBar = kOffRing + RING * kTailChunkBytes;
  static constexpr int kNumBars = 13 + 2 * RING;
  static constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// [256][128] bf16 <- [ s3[n] * W3[n][0..63] | sd[n] * Wd[n][0..63] ]
__global__ void fold_tail_weights_kernel(const __nv_bfloat16* __restrict__ w3, const __nv_bfloat16* __restrict__ wd,
                                         const float* __restrict__ s3, const float* __restrict__ sd, int ld3, int ldd,
                                         __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 128) return;
  const int n = i >> 7, k = i & 127;
  const float v = k < 64 ? __bfloat162float(w3[(size_t)n * ld3 + k]) * s3[n] : __bfloat162float(wd[(size_t)n * ldd + (k - 64)]) * sd[n];
  out[i] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}

// tmA: conv2 input (64, W, H, F), box 64 x 10 x 18;  tmW2: conv2 weights, box 64 x 64;  tmW3: conv3 weights (DS: the folded
// [W3 | Wd] matrix), box 64 x 256;  tmXR: box 64 x 8 x 16 over the block input X (DS) or the 256-channel residual (RES);
// tmO: box 64 x 8 x 16 over the 256-channel output.
template <bool DS, int IN_STAGES, int RING>
__global__ void __launch_bounds__(kTailThreads, 1)
conv_tail_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmXR,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ TailParams p) {
  using Cfg = TailCfg<DS, IN_STAGES, RING>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* w_bar = bars + 0;
  uint64_t* in_full = bars + 1;      // [2]
  uint64_t* in_empty = bars + 3;     // [2]
  uint64_t* acc2_full = bars + 5;    // [2]
  uint64_t* acc2_empty = bars + 7;   // [2]
  uint64_t* a3_full = bars + 9;
  uint64_t* a3_empty = bars + 10;
  uint64_t* acc3_full = bars + 11;
  uint64_t* acc3_empty = bars + 12;
  uint64_t* x_full = a3_full;        // DS: the X tile lands on the same barrier the epilogue warps arrive on (8 + 1 arrivals)
  uint64_t* x_empty = a3_empty;      //     and is released with A3 by the commit that follows conv3
  uint64_t* st_ready = bars + 13;            // [RING] RES: residual chunk landed in staging tile b; DS: tile b may be rewritten
  uint64_t* st_done = bars + 13 + RING;      // [RING] the epilogue warps have finished staging tile b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW3);
    tma_prefetch_desc(&tmXR);
    tma_prefetch_desc(&tmO);
    mbar_init(w_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&in_full[a], 1);
      mbar_init(&in_empty[a], 1);
      mbar_init(&acc2_full[a], 1);
      mbar_init(&acc2_empty[a], 8);
    }
    mbar_init(a3_full, DS ? 9 : 8);
    mbar_init(a3_empty, 1);
    mbar_init(acc3_full, 1);
    mbar_init(acc3_empty, 8);
    for (int b = 0; b < RING; ++b) {
      mbar_init(&st_ready[b], 1);
      mbar_init(&st_done[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);   // acc2: columns 0..127 (two stages), acc3: columns 256..511
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();

  auto tile_coords = [&](int tile, int& wb, int& hb, int& f) {
    int r = tile;
    wb = r % p.tiles_w; r /= p.tiles_w;
    hb = r % p.tiles_h;
    f = r / p.tiles_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ halo-box (and, DS, X tile) producer: one elected thread
    if (elect_one_sync()) {
      const uint32_t in_full_a = smem_u32(in_full), in_empty_a = smem_u32(in_empty);
      const uint32_t x_full_a = smem_u32(x_full), x_empty_a = smem_u32(x_empty);
      const uint32_t in0 = smem_u32(smem + Cfg::kOffIn), x0 = smem_u32(smem + Cfg::kOffX);
      {
        const uint32_t wb = smem_u32(w_bar), w2 = smem_u32(smem), w3 = smem_u32(smem + Cfg::kOffW3);
        mbar_arrive_expect_tx_a(wb, (uint32_t)(kTailW2Bytes + kTailW3Bytes + (DS ? kTailW3Bytes : 0)));
        for (int tap = 0; tap < 9; ++tap) tma_load_2d_a(w2 + (uint32_t)tap * 8192u, &tmW2, wb, tap * 64, 0);
        tma_load_2d_a(w3, &tmW3, wb, 0, 0);
        if (DS) tma_load_2d_a(smem_u32(smem + Cfg::kOffWd), &tmW3, wb, 64, 0);
      }
      griddep_wait();  // weights are constants; the activations come from preceding kernels
      for (int a = 1; a < kTailL2Ahead; ++a) {
        const int t = blockIdx.x + a * gridDim.x;
        if (t < p.num_tiles) {
          int wb, hb, f;
          tile_coords(t, wb, hb, f);
          tma_prefetch_l2_4d(&tmA, 0, wb * 8 - 1, hb * 16 - 1, f);
          if (DS) tma_prefetch_l2_4d(&tmXR, 0, wb * 8, hb * 16, f);
        }
      }
      uint32_t tc = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        int wb, hb, f;
        {
          const int t = tile + kTailL2Ahead * (int)gridDim.x;
          if (t < p.num_tiles) {
            tile_coords(t, wb, hb, f);
            tma_prefetch_l2_4d(&tmA, 0, wb * 8 - 1, hb * 16 - 1, f);
            if (DS) tma_prefetch_l2_4d(&tmXR, 0, wb * 8, hb * 16, f);
          }
        }
        tile_coords(tile, wb, hb, f);
        const uint32_t st = IN_STAGES == 2 ? (tc & 1u) : 0u;
        const uint32_t use = IN_STAGES == 2 ? (tc >> 1) : tc;
        mbar_wait_a(in_empty_a + st * 8, (use & 1u) ^ 1u);
        mbar_arrive_expect_tx_a(in_full_a + st * 8, (uint32_t)(kS3HaloRows * 128));
        tma_load_4d_b(in0 + st * (uint32_t)kTailInBytes, &tmA, in_full_a + st * 8, 0, wb * 8 - 1, hb * 16 - 1, f);
        TAIL_STAMP(tc, 0);
        if (DS) {
          mbar_wait_a(x_empty_a, (tc & 1u) ^ 1u);   // conv3 of the previous tile has read X (and A3)
          mbar_arrive_expect_tx_a(x_full_a, (uint32_t)kTailChunkBytes);
          tma_load_4d_b(x0, &tmXR, x_full_a, 0, wb * 8, hb * 16, f);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one_sync()) {
      constexpr uint32_t idesc2 = umma_idesc_bf16_m128(64);
      constexpr uint32_t idesc3 = umma_idesc_bf16_m128(256);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint64_t desc_halo = umma_desc_sw128_sbo(1280u);   // 8-row groups of the 10-pixel-wide halo box
      const uint32_t in_full_a = smem_u32(in_full), in_empty_a = smem_u32(in_empty);
      const uint32_t acc2_full_a = smem_u32(acc2_full), acc2_empty_a = smem_u32(acc2_empty);
      const uint32_t a3_full_a = smem_u32(a3_full), a3_empty_a = smem_u32(a3_empty);
      const uint32_t acc3_full_a = smem_u32(acc3_full), acc3_empty_a = smem_u32(acc3_empty);
      const uint32_t w2_16 = smem_u32(smem) >> 4, w3_16 = smem_u32(smem + Cfg::kOffW3) >> 4, wd_16 = smem_u32(smem + Cfg::kOffWd) >> 4;
      const uint32_t in16 = smem_u32(smem + Cfg::kOffIn) >> 4, a3_16 = smem_u32(smem + Cfg::kOffA3) >> 4, x16 = smem_u32(smem + Cfg::kOffX) >> 4;
      mbar_wait(w_bar, 0);
      auto conv2 = [&](uint32_t t2) {
        const uint32_t acc = t2 & 1u;
        const uint32_t st = IN_STAGES == 2 ? acc : 0u;
        const uint32_t use = IN_STAGES == 2 ? (t2 >> 1) : t2;
        mbar_wait_a(in_full_a + st * 8, use & 1u);
        mbar_wait_a(acc2_empty_a + acc * 8, ((t2 >> 1) & 1u) ^ 1u);
        tc_fence_after();
        TAIL_STAMP(t2, 3);
        const uint32_t d_tmem = tmem_base + acc * 64u;
        const uint32_t a16 = in16 + st * (uint32_t)(kTailInBytes >> 4);
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            // tap (dh, dw): the halo box read from row dh * 10 + dw on (the swizzle follows absolute address bits)
            const uint64_t adesc = desc_halo | (a16 + (uint32_t)((dh * 10 + dw) * 8));
            const uint64_t bdesc = desc_hi | (w2_16 + (uint32_t)((dh * 3 + dw) * 512));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dh == 0 && dw == 0 && k == 0) umma_f16_c<false>(d_tmem, adesc, bdesc, idesc2);
              else                              umma_f16_c<true>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc2);
            }
          }
        }
        umma_commit_a(in_empty_a + st * 8);
        umma_commit_a(acc2_full_a + acc * 8);
        TAIL_STAMP(t2, 4);
      };
      if ((int)blockIdx.x < p.num_tiles) conv2(0);
      uint32_t tc = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        // conv3 of this tile: the epilogue warps have written A3 (and, DS, the block-input tile X has landed)
        mbar_wait_a(acc3_empty_a, (tc & 1u) ^ 1u);
        mbar_wait_a(a3_full_a, tc & 1u);
        tc_fence_after();
        TAIL_STAMP(tc, 1);
        const uint32_t d3 = tmem_base + 256u;
        {
          const uint64_t adesc = desc_hi | a3_16, bdesc = desc_hi | w3_16;
          umma_f16_c<false>(d3, adesc, bdesc, idesc3);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_f16_c<true>(d3, adesc + 2 * k, bdesc + 2 * k, idesc3);
        }
        if (DS) {
          const uint64_t adesc = desc_hi | x16, bdesc = desc_hi | wd_16;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_c<true>(d3, adesc + 2 * k, bdesc + 2 * k, idesc3);
        }
        umma_commit_a(a3_empty_a);
        umma_commit_a(acc3_full_a);
        TAIL_STAMP(tc, 2);
        if (tile + (int)gridDim.x < p.num_tiles) conv2(tc + 1);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------------------------ chunk DMA (one elected thread)
    if (elect_one_sync()) {
      griddep_wait();
      const uint32_t st_ready_a = smem_u32(st_ready), st_done_a = smem_u32(st_done);
      const uint32_t r0 = smem_u32(smem + Cfg::kOffRing);
      const int my_tiles = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int chunks = 4 * my_tiles;
      auto load_res = [&](int c) {   // RES: residual chunk c -> staging tile c % RING; a later tile's chunk -> L2
        int wb, hb, f;
        tile_coords(blockIdx.x + (c >> 2) * gridDim.x, wb, hb, f);
        const uint32_t b = (uint32_t)(c % RING);
        mbar_arrive_expect_tx_a(st_ready_a + b * 8, (uint32_t)kTailChunkBytes);
        tma_load_4d_b(r0 + b * kTailChunkBytes, &tmXR, st_ready_a + b * 8, (c & 3) * 64, wb * 8, hb * 16, f);
        const int ca = c + 4 * kTailL2Ahead;
        if (ca < chunks) {
          tile_coords(blockIdx.x + (ca >> 2) * gridDim.x, wb, hb, f);
          tma_prefetch_l2_4d(&tmXR, (ca & 3) * 64, wb * 8, hb * 16, f);
        }
      };
      if (!DS) {
        for (int ca = RING; ca < 4 * kTailL2Ahead && ca < chunks; ++ca) {
          int wb, hb, f;
          tile_coords(blockIdx.x + (ca >> 2) * gridDim.x, wb, hb, f);
          tma_prefetch_l2_4d(&tmXR, (ca & 3) * 64, wb * 8, hb * 16, f);
        }
        for (int c = 0; c < RING && c < chunks; ++c) load_res(c);
      }
      for (int c = 0; c < chunks; ++c) {
        int wb, hb, f;
        tile_coords(blockIdx.x + (c >> 2) * gridDim.x, wb, hb, f);
        const uint32_t b = (uint32_t)(c % RING);
        mbar_wait_a(st_done_a + b * 8, (uint32_t)(c / RING) & 1u);
        TAIL_STAMP((uint32_t)(c >> 2), 5 + 3 * (c & 3));
        tma_store_4d(&tmO, r0 + b * kTailChunkBytes, (c & 3) * 64, wb * 8, hb * 16, f);
        tma_store_commit();
        TAIL_STAMP((uint32_t)(c >> 2), 6 + 3 * (c & 3));
        tma_store_wait_read<0>();
        TAIL_STAMP((uint32_t)(c >> 2), 7 + 3 * (c & 3));
        if (DS) mbar_arrive_a(st_ready_a + b * 8);
        else if (c + RING < chunks) load_res(c + RING);
      }
      tma_store_wait<0>();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 3..10
    // The body is instantiated per column half so that every BN scale / shift is a constant-bank operand with a static
    // address (FFMA R, R, c[0][imm], c[0][imm]): no load instruction at all for them.
    const int q = warp & 3;            // TMEM lane quarter
    const int lrow = q * 32 + lane;    // tile row = (h, w) = (lrow / 8, lrow % 8)
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t a3_row = smem_u32(smem + Cfg::kOffA3) + (uint32_t)lrow * 128u;
    const uint32_t ring_row = smem_u32(smem + Cfg::kOffRing) + (uint32_t)lrow * 128u;
    auto body = [&](auto half_c) {
      constexpr int half = decltype(half_c)::value;  // which 32 of every 64 columns
      uint32_t tc = 0;
      uint32_t c = 0;  // chunk counter of this CTA
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        // ---- E1: conv2 accumulator -> BN2 + ReLU -> bf16 -> A3
        const uint32_t acc = tc & 1u;
        mbar_wait(&acc2_full[acc], (tc >> 1) & 1u);
        tc_fence_after();
        if (warp == 3 && lane == 0) TAIL_STAMP(tc, 17);
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u + (uint32_t)(half * 32), v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[acc]);
        mbar_wait(a3_empty, (tc & 1u) ^ 1u);          // conv3 of the previous tile has read A3
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          constexpr int col0 = half * 32;
          const int col = col0 + g * 8;
          const uint32_t addr = a3_row + ((((uint32_t)col >> 3) ^ xr) << 4);
          float fv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) fv[j] = fmaf(__uint_as_float(v[g * 8 + j]), p.s2[col + j], p.b2[col + j]);
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = p.relu2 ? pack_bf16x2_relu(fv[2 * j], fv[2 * j + 1]) : pack_bf16x2(fv[2 * j], fv[2 * j + 1]);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a3_full);
        if (warp == 3 && lane == 0) TAIL_STAMP(tc, 18);
        // ---- E2: conv3 accumulator -> BN3 (+ residual) + ReLU -> bf16 -> staging tile -> (DMA thread) TMA store
        mbar_wait(acc3_full, tc & 1u);
        tc_fence_after();
        if (warp == 3 && lane == 0) TAIL_STAMP(tc, 19);
#pragma unroll
        for (int n = 0; n < 4; ++n, ++c) {
          const uint32_t b = c % RING, u = c / RING;
          // RES: the residual chunk of use u has landed; DS: the store of use u - 1 has been read (passes at once for u = 0)
          mbar_wait(&st_ready[b], DS ? ((u & 1u) ^ 1u) : (u & 1u));
          if (warp == 3 && lane == 0) TAIL_STAMP(tc, 20 + 2 * n);
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(n * 64 + half * 32), v);
          const uint32_t row = ring_row + b * kTailChunkBytes;
          uint4 rr[4];
          if (!DS) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t addr = row + ((((uint32_t)(half * 4 + g)) ^ xr) << 4);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[g].x), "=r"(rr[g].y), "=r"(rr[g].z), "=r"(rr[g].w) : "r"(addr));
            }
          }
          tmem_ld_wait();
          if (n == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc3_empty);   // accumulator drained: conv3 of the next tile may start
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = n * 64 + half * 32 + g * 8;
            const uint32_t addr = row + ((((uint32_t)(half * 4 + g)) ^ xr) << 4);
            float fv[8];
            if (DS) {
#pragma unroll
              for (int j = 0; j < 8; ++j) fv[j] = __uint_as_float(v[g * 8 + j]) + p.b3[col + j];
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) fv[j] = fmaf(__uint_as_float(v[g * 8 + j]), p.s3[col + j], p.b3[col + j]);
              fv[0] += bf16_lo(rr[g].x); fv[1] += bf16_hi(rr[g].x);
              fv[2] += bf16_lo(rr[g].y); fv[3] += bf16_hi(rr[g].y);
              fv[4] += bf16_lo(rr[g].z); fv[5] += bf16_hi(rr[g].z);
              fv[6] += bf16_lo(rr[g].w); fv[7] += bf16_hi(rr[g].w);
            }
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = p.relu3 ? pack_bf16x2_relu(fv[2 * j], fv[2 * j + 1]) : pack_bf16x2(fv[2 * j], fv[2 * j + 1]);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
          fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) mbar_arrive(&st_done[b]);
          if (warp == 3 && lane == 0) TAIL_STAMP(tc, 21 + 2 * n);
        }
      }
    };
    if (((warp - 3) >> 2) == 0) body(std::integral_constant<int, 0>{});
    else                        body(std::integral_constant<int, 1>{});
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vad
