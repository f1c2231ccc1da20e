// K2s3: spatial (1,3,3) convolutions, stride 1, pad (0,1,1), 64 -> 64 channels -- conv2 of the layer1 bottlenecks
// (src/i3d.py:85-92) -- with the nine taps read out of ONE shared-memory copy of the activation tile plus its halo.
//
// Through the generic kernel this layer loads nine im2col columns (9 x 16 KB per 128 output pixels); with N = 64
// that is 85 B/clk/SM of L2 -> SM traffic, above what L2 can deliver to 148 SMs, so the layer is L2-bound at half
// the tensor rate.  Here an M tile is 8 (w) x 16 (h) output pixels of one frame and ONE rank-4 TMA box (64 ch, 10 w,
// 18 h, 1 frame) lands as 180 rows of 128 B (SWIZZLE_128B), row = h * 10 + w, TMA's out-of-range zero fill being the
// spatial padding.  Tap (dh, dw) is that same tile read from row dh * 10 + dw on, eight-row groups 1280 B apart
// (SBO).  The descriptor's start address is then NOT a multiple of the 1024-byte swizzle atom, and neither is the
// group stride: this works because the tensor core applies the 128-byte swizzle to absolute shared-memory address
// bits, exactly as TMA does when it writes the box, with the descriptor's matrix-base-offset field left at 0
// (measured on B200 in round 2: bit-identical to three dw-shifted 8 x 18 boxes whose taps start on atom boundaries;
// with base offset = start row & 7 the results are wrong).  A traffic drops from 147 KB (im2col) to 23 KB per tile;
// the 9 x 8 KB of weights stay resident in shared memory.  One stage = one tile = 36 MMAs.
// Epilogue as in the stem: BN + ReLU -> bf16 -> 128B-swizzled staging tile -> TMA store (4 rows x 8 columns per warp
// pair), clipped at the frame border by the tensor map.
#pragma once

#include "conv_umma.cuh"
#include "stem_umma.cuh"

namespace vad {

struct S3x3Params {
  int F, H, W;            // frames (clips x T), height, width
  int tiles_w, tiles_h, num_tiles;
  int relu;
  const float* scale;
  const float* shift;
};

constexpr int kS3HaloRows = 18 * 10;                                   // one halo box: 18 rows x 10 pixels x 64 channels bf16
constexpr int kS3HaloBytes = (kS3HaloRows * 128 + 1023) / 1024 * 1024;  // 23 KB stage (1024-byte multiple)
constexpr int kS3StageBytes = kS3HaloBytes;
constexpr int kS3Stages = 3;
constexpr int kS3WBytes = 9 * 64 * 128;              // resident weights: 9 taps x (64 cout x 64 cin)
constexpr int kS3Threads = 64 + 8 * 32;
constexpr int kS3SmemBytes = kS3WBytes + 2 * kStemStagingBytes + kS3Stages * kS3StageBytes + 2 * 64 * 4 + (2 * kS3Stages + 5) * 8 + 16 + 1024;

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// SWIZZLE_128B K-major descriptor (everything but the start address) with an arbitrary stride between 8-row groups;
// matrix base offset 0 (see the header comment)
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>(1) << 16;                       // LBO (ignored)
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;          // SBO
  d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(kS3Threads, 1)
conv_s3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO, const S3x3Params p) {
  constexpr int kStageBytes = kS3StageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* w_smem = smem;                                     // tap (dh, dw) at (dh * 3 + dw) * 8 KB
  uint8_t* staging = smem + kS3WBytes;                        // 2 x 16 KB
  uint8_t* stage_base = staging + 2 * kStemStagingBytes;      // kS3Stages x 54 KB
  float* s_scale = reinterpret_cast<float*>(stage_base + kS3Stages * kS3StageBytes);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + kS3Stages;
  uint64_t* tmem_full_bar = empty_bar + kS3Stages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kS3Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);
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
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      {
        const uint32_t wb = smem_u32(w_bar), w0 = smem_u32(w_smem);
        mbar_arrive_expect_tx_a(wb, (uint32_t)kS3WBytes);
        for (int tap = 0; tap < 9; ++tap) tma_load_2d_a(w0 + (uint32_t)tap * 8192u, &tmW, wb, tap * 64, 0);
      }
      griddep_wait();  // weights are constants; the activations come from the preceding kernel
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int r = tile;
        const int wb = r % p.tiles_w; r /= p.tiles_w;
        const int hb = r % p.tiles_h;
        const int f = r / p.tiles_h;
        mbar_wait_a(empty0 + s * 8, ph ^ 1u);
        const uint32_t dst = stage0 + s * (uint32_t)kStageBytes;
        const uint32_t fb = full0 + s * 8;
        mbar_arrive_expect_tx_a(fb, (uint32_t)(kS3HaloRows * 128));
        tma_load_4d_b(dst, &tmA, fb, 0, wb * 8 - 1, hb * 16 - 1, f);
        if (++s == kS3Stages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(64);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint64_t desc_halo = umma_desc_sw128_sbo(1280u);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      mbar_wait(w_bar, 0);
      uint32_t s = 0, ph = 0, tc = 0;
      mbar_wait_a(full0, 0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        const uint32_t acc = tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, ((tc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 64u;
        const bool last_tile = tile + (int)gridDim.x >= p.num_tiles;
        uint32_t ns = s + 1, nph = ph;
        if (ns == kS3Stages) { ns = 0; nph ^= 1u; }
        const uint32_t a16 = (stage0 + s * (uint32_t)kStageBytes) >> 4;
        bool ready = last_tile;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            // tap (dh, dw): box dw, shifted by dh rows of 8 pixels (1024 B = one swizzle atom)
            // tap (dh, dw): the halo box read from row dh * 10 + dw on (8 x 16-byte units per row), groups 1280 B apart
            const uint64_t adesc = desc_halo | (a16 + (uint32_t)((dh * 10 + dw) * 8));
            const uint64_t bdesc = desc_hi | (w16 + (uint32_t)((dh * 3 + dw) * 512));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dh == 0 && dw == 0 && k == 0) umma_f16_c<false>(d_tmem, adesc, bdesc, idesc);
              else                              umma_f16_c<true>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc);
            }
            if (dh == 2 && dw == 0 && !last_tile) {  // after 28 of the 36 MMAs: probe the next tile's operands
              ready = mbar_try_wait_a(full0 + ns * 8, nph);
              tc_fence_after();
            }
          }
        }
        umma_commit_a(empty0 + s * 8);
        umma_commit_a(tfull0 + acc * 8);
        if (!ready) {
          mbar_wait_a(full0 + ns * 8, nph);
          tc_fence_after();
        }
        s = ns; ph = nph;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..9
    griddep_wait();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool issuer = half == 0;
    const int lrow = q * 32 + lane;
    const uint32_t xr = (uint32_t)(lrow & 7);
    const uint32_t staging0 = smem_u32(staging);
    uint32_t tc = 0, sb = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc, sb ^= 1u) {
      int r = tile;
      const int wb = r % p.tiles_w; r /= p.tiles_w;
      const int hb = r % p.tiles_h;
      const int f = r / p.tiles_h;
      const uint32_t row_addr = staging0 + sb * kStemStagingBytes + (uint32_t)lrow * 128u;
      if (issuer && lane == 0) tma_store_wait_read<1>();  // the store that read this staging buffer two tiles ago
      named_bar_sync(1 + q, 64);
      const uint32_t acc = tc & 1u;
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64u + (uint32_t)(half * 32);
      uint32_t v[32];
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = half * 32 + g * 8;
        const uint32_t addr = row_addr + ((((uint32_t)col >> 3) ^ xr) << 4);
        float fv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          fv[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
          if (p.relu) fv[j] = fmaxf(fv[j], 0.f);
        }
        const uint32_t o0 = pack_bf16x2(fv[0], fv[1]), o1 = pack_bf16x2(fv[2], fv[3]);
        const uint32_t o2 = pack_bf16x2(fv[4], fv[5]), o3 = pack_bf16x2(fv[6], fv[7]);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + q, 64);
      if (issuer && lane == 0) {
        tma_store_4d(&tmO, staging0 + sb * kStemStagingBytes + (uint32_t)q * 4096u, 0, wb * 8, hb * 16 + q * 4, f);
        tma_store_commit();
      }
    }
    if (issuer && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace vad
