// K2t: temporal (3,1,1) convolutions, stride 1, pad (1,0,0) -- conv1 of every "temp_conv" bottleneck
// (src/i3d.py:71-83) -- with the three taps read out of ONE shared-memory copy of the activation tile.
//
// Through the generic kernel such a layer loads its A operand three times (one im2col column per tap); with
// N = 64 / 128 those layers are bound by L2 -> SM traffic, not by the tensor pipe.  Here an M tile is ALL T
// frames x P pixels (T * P = 128; T = 4, P = 32 in layer1, T = 2, P = 64 later): one rank-4 TMA box
// (64 ch, P, T, 1) lands as T * P rows of 128 B (SWIZZLE_128B) behind P zero rows and in front of P zero
// rows, and tap dt is the same tile at a row offset of dt * P -- P * 128 B is a multiple of the 1024-byte swizzle
// atom, so the UMMA descriptor just moves its start address.  The zero rows are the temporal padding.  A stage
// is one 64-channel block: 1 A box + the 3 taps' weight boxes, 12 MMAs behind one barrier pair.
// The contraction order is (channel block, tap) instead of the generic kernel's (tap, channel block), so
// results differ from it in fp32 summation order only (tests compare them at bf16 resolution).
#pragma once

#include "conv_umma.cuh"

namespace vad {

struct ThaloParams {
  int B, T, HW;          // clips, frames (2 or 4), pixels per frame
  int P, logP;           // pixels per tile = 128 / T
  int tiles_per_clip;    // ceil(HW / P)
  int N, Cin;            // output / input channels (Cin % 64 == 0)
  int n_tiles, num_tiles;
  int relu, ldo;
  int resident;          // 1: all 3 * Cin / 64 weight boxes stay in shared memory for the whole persistent CTA
  int n_stages;          // ring depth (<= kMaxStages)
  int a_region;          // P front pad rows + 128 data rows + P back pad rows = 16384 + 256 * P bytes
  int stage_bytes;       // A region (+ the three taps' weight boxes when not resident)
  const float* scale;
  const float* shift;
  __nv_bfloat16* out;    // [B, T, HW, ldo], already offset to the first output channel
};

template <int BN>
struct ThaloCfg {
  static constexpr int kBBytes = BN * 128;         // one tap, one 64-channel block
  static constexpr int kMaxStages = 6;
  static constexpr int kBudget = 222 * 1024;       // stages + resident weights
  static constexpr int kEpiWarps = 8;              // two warps per TMEM lane quarter, half the columns each
  static constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
  static constexpr int kThreads = 64 + kEpiWarps * 32;
  static constexpr int kFixedBytes = 2 * BN * 4 + (2 * kMaxStages + 5) * 8 + 16 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(ThaloCfg<BN>::kThreads, 1)
conv_thalo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ThaloParams p) {
  using Cfg = ThaloCfg<BN>;
  const uint32_t STAGES = (uint32_t)p.n_stages;
  const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_cb = p.Cin / 64;
  const uint32_t pad_bytes = (uint32_t)p.P * 128u;
  const uint32_t w_bytes = p.resident ? (uint32_t)(3 * num_cb) * Cfg::kBBytes : 0u;
  uint8_t* w_smem = smem;                      // resident weights: box (cb, dt) at (cb * 3 + dt) * kBBytes
  uint8_t* stage_base = smem + w_bytes;
  float* s_scale = reinterpret_cast<float*>(stage_base + STAGES * stage_bytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* empty_bar = full_bar + Cfg::kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], Cfg::kEpiWarps);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  // temporal zero padding: P rows in front of and behind the data rows of every stage; TMA never writes them
  for (uint32_t s = 0; s < STAGES; ++s) {
    uint4* front = reinterpret_cast<uint4*>(stage_base + s * stage_bytes);
    uint4* back = reinterpret_cast<uint4*>(stage_base + s * stage_bytes + pad_bytes + 16384);
    for (int i = threadIdx.x; i < (int)(pad_bytes / 16); i += blockDim.x) {
      front[i] = make_uint4(0u, 0u, 0u, 0u);
      back[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      uint32_t s = 0, ph = 0;
      if (p.resident) {  // n_tiles == 1: the whole weight matrix, once
        const uint32_t wb = smem_u32(w_bar), w0 = smem_u32(w_smem);
        mbar_arrive_expect_tx_a(wb, w_bytes);
        for (int cb = 0; cb < num_cb; ++cb)
          for (int dt = 0; dt < 3; ++dt)
            tma_load_2d_a(w0 + (uint32_t)(cb * 3 + dt) * Cfg::kBBytes, &tmB, wb, dt * p.Cin + cb * 64, 0);
      }
      griddep_wait();  // weights are constants; the activations come from the preceding kernel
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * BN;
        const int mt = tile / p.n_tiles;
        const int clip = mt / p.tiles_per_clip;
        const int p0 = (mt - clip * p.tiles_per_clip) * p.P;
        for (int cb = 0; cb < num_cb; ++cb) {
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t a_dst = stage0 + s * stage_bytes;
          const uint32_t fb = full0 + s * 8;
          mbar_arrive_expect_tx_a(fb, 16384u + (p.resident ? 0u : 3u * Cfg::kBBytes));
          tma_load_4d_b(a_dst + pad_bytes, &tmA, fb, cb * 64, p0, 0, clip);
          if (!p.resident) {
#pragma unroll
            for (int dt = 0; dt < 3; ++dt)
              tma_load_2d_a(a_dst + (uint32_t)p.a_region + dt * Cfg::kBBytes, &tmB, fb, dt * p.Cin + cb * 64, n0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(BN);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      const uint32_t pad16 = pad_bytes >> 4;
      uint32_t s = 0, ph = 0;
      int tc = 0;
      const uint32_t w16 = smem_u32(w_smem) >> 4;
      if (p.resident) mbar_wait(w_bar, 0);
      mbar_wait_a(full0, 0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
        const uint32_t acc = (uint32_t)tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, (((uint32_t)tc >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const bool last_tile = tile + (int)gridDim.x >= p.num_tiles;
        for (int cb = 0; cb < num_cb; ++cb) {
          const bool last_stage = cb == num_cb - 1;
          uint32_t ns = s + 1, nph = ph;
          if (ns == STAGES) { ns = 0; nph ^= 1u; }
          const bool do_wait = !(last_stage && last_tile);
          bool ready = !do_wait;
          const uint32_t a_lo = (stage0 + s * stage_bytes) >> 4;
          const uint32_t b_lo = p.resident ? w16 + (uint32_t)(cb * 3) * (Cfg::kBBytes >> 4) : a_lo + ((uint32_t)p.a_region >> 4);
#pragma unroll
          for (int dt = 0; dt < 3; ++dt) {
            const uint64_t adesc = desc_hi | (a_lo + (uint32_t)dt * pad16);   // tap dt: the tile shifted by dt * P rows
            const uint64_t bdesc = desc_hi | (b_lo + (uint32_t)dt * (Cfg::kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dt == 0 && k == 0) umma_f16(d_tmem, adesc, bdesc, idesc, cb ? 1u : 0u);
              else                   umma_f16_c<true>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc);
              if (dt == 2 && k == 0 && do_wait) {  // after 9 of the 12 MMAs: the next stage's operands
                if (last_stage) {
                  ready = mbar_try_wait_a(full0 + ns * 8, nph);
                } else {
                  mbar_wait_a(full0 + ns * 8, nph);
                  ready = true;
                }
                tc_fence_after();
              }
            }
          }
          umma_commit_a(empty0 + s * 8);
          if (last_stage) umma_commit_a(tfull0 + acc * 8);
          if (!ready) {
            mbar_wait_a(full0 + ns * 8, nph);
            tc_fence_after();
          }
          s = ns; ph = nph;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    constexpr int CPW = Cfg::kColsPerWarp;
    griddep_wait();
    const int t = threadIdx.x - 64;
    const int q = warp & 3;
    const int col0 = ((warp - 2) >> 2) * CPW;
    int tc = 0, cached_n0 = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int mt = tile / p.n_tiles;
      const int clip = mt / p.tiles_per_clip;
      const int p0 = (mt - clip * p.tiles_per_clip) * p.P;
      const int acc = tc & 1;
      if (n0 != cached_n0) {
        named_bar_sync(1, Cfg::kEpiWarps * 32);
        for (int i = t; i < BN; i += Cfg::kEpiWarps * 32) {
          const int n = n0 + i;
          s_scale[i] = (n < p.N) ? p.scale[n] : 0.f;
          s_shift[i] = (n < p.N) ? p.shift[n] : 0.f;
        }
        named_bar_sync(1, Cfg::kEpiWarps * 32);
        cached_n0 = n0;
      }
      const int r = q * 32 + lane;          // tile row = (frame, pixel)
      const int tf = r >> p.logP;
      const int px = p0 + (r & (p.P - 1));
      const bool row_ok = px < p.HW;
      __nv_bfloat16* out_row = p.out + (((long long)clip * p.T + tf) * p.HW + px) * p.ldo + n0 + col0;
      mbar_wait(&tmem_full_bar[acc], (tc >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0);
#pragma unroll 1
      for (int c = 0; c < CPW / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            if (n0 + col0 + col < p.N) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col0 + col + j], s_shift[col0 + col + j]);
                if (p.relu) f[j] = fmaxf(f[j], 0.f);
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]);
              o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]);
              o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(out_row + col) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace vad
