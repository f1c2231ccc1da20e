// K2s on CTA PAIRS, for stems whose taps do not fit in one CTA's shared memory (InceptionI3d's 7x7x7 / 2: 49 taps x 4 KB = 196 KB).
//
// stem_umma_kernel handles that case by streaming the kh taps of frame tap dt with dt's stage (w_stream): 28 KB of weights
// per 6.5 KB of activations per stage, i.e. the stem becomes L2-bound (ncu / bench round 2: 3.0 ms per 160 clip-crops, 0.47 of
// the tensor peak against 0.78 for I3Res50's resident-weight stem).  Here a cluster of two CTAs takes two spatial tiles
// (M = 256 rows, tcgen05 cta_group::2); each CTA keeps the weights of HALF of the 64 output channels resident (98 KB), and the
// pair MMA reads the two halves of its B operand from both CTAs' shared memory: same N = 64 MMAs as before, no weight traffic
// after the prologue.  Same operand geometry as stem_umma_kernel (raw row segments, no-swizzle sliding-window descriptor),
// same protocol as conv_pair.cuh (loads of both CTAs complete on the leader's full barrier, the leader issues, commits are
// multicast, both CTAs' epilogue warps hand the accumulator back to the leader).  No temporal-pool fusion (pool_t == 1).
// Contracts in the same (dt, dh, K half) order per accumulator element as stem_umma_kernel: bit-identical.
#pragma once

#include "conv_pair.cuh"
#include "stem_umma.cuh"

namespace vad {

constexpr int kStemPairThreads = 192;
constexpr int kStemPairTapBytes = 32 * 64;   // one tap of this CTA's 32 output channels: 32 rows x 32 bf16

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// tmE / tmOdd / tmO as for stem_umma_kernel; tmWh: the weight matrix with 32-row boxes (one channel half of one tap).
// Item i = tiles 2 i (rank 0) and 2 i + 1 (rank 1) of p.num_units tiles; a tile-less odd CTA loads zeros and stores nothing.
__global__ void __launch_bounds__(kStemPairThreads, 1)
stem_umma_pair_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmOdd,
                      const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmO, const StemParams p) {
  extern __shared__ __align__(1024) uint8_t stem_bf16_pair_smem[];
  uint8_t* smem = stem_bf16_pair_smem;
  if (smem_u32(smem) & 1023u) __trap();
  const int crank = (int)cluster_ctarank();
  const int ntaps = p.kt * p.kh;
  uint8_t* w_smem = smem;                                                    // ntaps x 2 KB: this CTA's channel half, resident
  uint8_t* staging = smem + ((ntaps * kStemPairTapBytes + 1023) & ~1023);    // 2 x 16 KB output staging
  uint8_t* stage_base = staging + 2 * kStemStagingBytes;
  float* s_scale = reinterpret_cast<float*>(stage_base + p.n_stages * p.stage_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);            // used in the leader only
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kStemMaxStages;                      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                              // [2] leader only
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.n_stages;
  const int i_first = (int)(blockIdx.x >> 1), i_step = (int)(gridDim.x >> 1);
  const int n_items = (p.num_units + 1) >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmOdd);
    tma_prefetch_desc(&tmWh);
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
    mbar_arrive_expect_tx(w_bar, (uint32_t)(ntaps * kStemPairTapBytes));
    for (int tap = 0; tap < ntaps; ++tap) tma_load_2d(w_smem + tap * kStemPairTapBytes, &tmWh, w_bar, tap * 32, crank * 32);
    mbar_wait(w_bar, 0);   // weights are constants: no dependency on the preceding kernel
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();

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
      const uint32_t tx = (uint32_t)((p.rows_even + p.rows_odd) * p.seg_bytes);
      griddep_wait();   // the clips come from the preceding preprocessing kernel
      uint32_t s = 0, ph = 0;
      for (int item = i_first; item < n_items; item += i_step) {
        int wb, hb, to, n;
        tile_coords(2 * item + crank, wb, hb, to, n);
        const int h_start = 2 * (hb * 16) - p.ph;
        const int x_start = wb * 8 * 8;   // 8 windows x (2 px x 4 ch) elements
        const int t0 = to * p.st - p.pt;
        // frame taps outside the clip contribute zeros: skipped (the SAME-padded 7x7x7 stem spends 6 of its 56 (to, dt) pairs there)
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t dst = stage0 + s * (uint32_t)p.stage_bytes;
          const uint32_t fb = lfull0 + s * 8;
          if (crank == 0) mbar_arrive_expect_tx_a(full0 + s * 8, 2u * tx);   // both CTAs' bytes
          tma_load_4d_pair(dst, &tmE, fb, x_start, h_start, t0 + dt, n);
          tma_load_4d_pair(dst + (uint32_t)p.off_odd, &tmOdd, fb, x_start, h_start + 1, t0 + dt, n);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (crank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m256(64);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      const uint32_t seg = (uint32_t)p.seg_bytes, seg16 = seg >> 4;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      const uint64_t a_hi = umma_desc_kmajor_noswizzle(0, 16u, seg);
      const uint64_t b_hi = umma_desc_kmajor<64>(0);
      uint32_t s = 0, ph = 0, tc = 0;
      for (int item = i_first; item < n_items; item += i_step, ++tc) {
        const uint32_t acc = tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, ((tc >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained it
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64u;
        int wb_, hb_, to_, n_;
        tile_coords(2 * item, wb_, hb_, to_, n_);
        const int t0 = to_ * p.st - p.pt;
        const int dt_lo = (p.Ti && t0 < 0) ? -t0 : 0, dt_hi = (p.Ti && t0 + p.kt > p.Ti) ? p.Ti - t0 : p.kt;
        for (int dt = dt_lo; dt < dt_hi; ++dt) {
          mbar_wait_a(full0 + s * 8, ph);
          tc_fence_after();
          const uint32_t st16 = (stage0 + s * (uint32_t)p.stage_bytes) >> 4;
          const uint32_t a_even = st16, a_odd = st16 + ((uint32_t)p.off_odd >> 4);
          uint32_t b_lo = w16 + (uint32_t)(dt * p.kh) * (kStemPairTapBytes >> 4);
          for (int dh = 0; dh < p.kh; ++dh) {
            const uint64_t adesc = a_hi | (((dh & 1) ? a_odd : a_even) + (uint32_t)(dh >> 1) * seg16);
            const uint64_t bdesc = b_hi | b_lo;
            umma_f16_pair(d_tmem, adesc, bdesc, idesc, (dt > dt_lo || dh) ? 1u : 0u);
            umma_f16_pair_acc(d_tmem, adesc + 2, bdesc + 2, idesc);
            b_lo += kStemPairTapBytes >> 4;
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
    griddep_wait();
    const int q = warp & 3;
    const int lrow = q * 32 + lane;  // tile row = TMEM lane: (h_i, w_i) = (lrow / 8, lrow % 8)
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    const uint32_t ltempty0 = mapa_u32(smem_u32(tmem_empty_bar), 0);
    uint32_t tc = 0;
    for (int item = i_first; item < n_items; item += i_step, ++tc) {
      const int tile = 2 * item + crank;
      int wb, hb, to, n;
      tile_coords(tile, wb, hb, to, n);
      const uint32_t acc = tc & 1u, sb = tc & 1u;
      const uint32_t row_addr = staging0 + sb * kStemStagingBytes + (uint32_t)lrow * 128u;
      // the TMA store that read this staging buffer two tiles ago must have drained it
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(taddr, v0);
      tmem_ld_32x32(taddr + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);   // hands this CTA's accumulator back to the leader
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
          const uint32_t o0 = pack_bf16x2(f[0], f[1]), o1 = pack_bf16x2(f[2], f[3]), o2 = pack_bf16x2(f[4], f[5]), o3 = pack_bf16x2(f[6], f[7]);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      };
      chunk(v0, 0);
      chunk(v1, 1);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (tile < p.num_units) tma_store_5d(&tmO, staging0 + sb * kStemStagingBytes + (uint32_t)q * 4096u, 0, wb * 8, hb * 16 + q * 4, to, n);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 128);
  }
}

}  // namespace vad
