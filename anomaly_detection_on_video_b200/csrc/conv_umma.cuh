// K2: conv3d + folded BatchNorm (+ residual) (+ ReLU) as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces the Conv3d -> BatchNorm3d -> ReLU triples of the reference backbone
// (src/i3d.py:101-116 bottleneck, :262-272 downsample, :303-305 stem).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]          m = (clip, to, ho, wo)   output pixel
//                                              n = output channel
//                                              k = (dt, dh, dw, cin)    filter tap x input channel
//   y[m, n] = relu( D * scale[n] + shift[n] + residual[m, n] )   -> bf16, channels-last
//
// Persistent kernel: one CTA per SM walks the 128 x BN output tiles (tile = blockIdx.x + i * gridDim.x,
// n fastest so co-running CTAs share activation tiles in L2).  Warp roles:
//   warp 0      TMA producer: weights (2D tiled map) and, in the TMA modes, the activation tile
//               (2D tiled map for 1x1x1/stride-1 layers, rank-5 im2col map otherwise; the stem uses an
//               im2col map over an overlapping 8-pixel x 4-channel window view, 64-byte rows)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  epilogue (TMEM -> registers -> scale/shift/residual/ReLU -> bf16 -> global)
//   warps 6..9  (GATHER builds only) A producer: cp.async 16 B with zero fill into the 128B-swizzled
//               tile; any stride / padding / channel count % 8, incl. the folded stem window
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) that runs ahead across tile boundaries, and a
// two-stage TMEM accumulator (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of
// tile i+1.
#pragma once

#include "ptx_sm100.cuh"

namespace vad {

enum AMode : int {
  A_TMA_2D = 0,      // 1x1x1, stride 1: A is the [M, Cin] matrix itself
  A_TMA_IM2COL = 1,  // rank-5 im2col tensor map
  A_GATHER = 2       // cp.async gather (any Cin % 8 == 0, incl. the folded stem window)
};

struct ConvParams {
  int M, N, num_kb;
  int n_tiles, num_tiles;
  int To, Ho, Wo;
  int Ti, Hi, Wi;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int cin_eff;  // contraction width of one tap (channels, or 32 for the folded stem window)
  int ntaps;
  long long sN, sT, sH, sW;  // input strides in elements
  int a_mode;
  int relu;
  int ldo;  // output row pitch in elements
  int ldr;  // residual row pitch in elements
  const __nv_bfloat16* in;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
};

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;

template <int BN, int BK, bool GATHER, bool EPI>
struct ConvCfg {
  static_assert(BK == 64 || BK == 32, "BK is one swizzle row: 64 (SW128) or 32 (SW64) bf16");
  static_assert(!GATHER || BK == 64, "the gather producer writes 128-byte swizzled rows");
  static_assert(!EPI || BN <= 128, "the staged epilogue keeps two 128 x BN bf16 tiles in shared memory");
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kABytes = kBlockM * kRowBytes;
  static constexpr int kBBytes = BN * kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // EPI: residual tile prefetched by TMA / output tile written back by TMA, double buffered;
  // laid out as BN/64 sub-tiles of [128 rows x 64 cols] (128-byte rows, SWIZZLE_128B)
  static constexpr int kEpiSubBytes = kBlockM * 128;
  static constexpr int kEpiBufBytes = EPI ? (BN / 64) * kEpiSubBytes : 0;
  static constexpr int kPipeBudget = 196608 - 2 * kEpiBufBytes;
  static constexpr int kStages = (kPipeBudget / kStageBytes) > 16 ? 16 : (kPipeBudget / kStageBytes);
  static constexpr int kThreads = GATHER ? 320 : 192;
  static constexpr int kGatherLag = kStages - 2 > 6 ? 6 : kStages - 2;  // cp.async groups in flight per thread
  static constexpr int kTmemCols = 2 * BN;                              // two accumulator stages
  // stages + scale/shift staging + barriers + tmem slot, + 1024 for manual alignment
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 2 * kEpiBufBytes + 2 * BN * 4 + (2 * kStages + 8) * 8 + 16 + 1024;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major operand tile with ROW_BYTES-wide rows (128: SWIZZLE_128B, 64: SWIZZLE_64B), rows packed
// densely, 8-row groups SBO apart.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);            // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                                // LBO (ignored)   [16,30)
  d |= static_cast<uint64_t>((8 * ROW_BYTES) >> 4) << 32;             // SBO             [32,46)
  d |= static_cast<uint64_t>(1) << 46;                                // descriptor version (sm_100)
  d |= static_cast<uint64_t>(ROW_BYTES == 128 ? 2 : 4) << 61;         // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

template <int BN, int BK, bool GATHER, bool EPI>
__global__ void __launch_bounds__(ConvCfg<BN, BK, GATHER, EPI>::kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const ConvParams p) {
  using Cfg = ConvCfg<BN, BK, GATHER, EPI>;
  constexpr int STAGES = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + STAGES * Cfg::kStageBytes;  // 2 x kEpiBufBytes, 1024-aligned
  float* s_scale = reinterpret_cast<float*>(epi_base + 2 * Cfg::kEpiBufBytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;    // [2] residual tile landed (EPI)
  uint64_t* res_empty_bar = res_full_bar + 2;     // [2] epilogue buffer free again (EPI)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
    if (!GATHER) tma_prefetch_desc(&tmA);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], GATHER ? (1 + 128) : 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);
      mbar_init(&res_full_bar[a], 1);
      mbar_init(&res_empty_bar[a], 4);
    }
    if (EPI) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmO);
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
    // ------------------------------------------------------------------ TMA producer
    // warp-uniform loops; only the issue itself is predicated on one elected lane
    const uint32_t tx = GATHER ? Cfg::kBBytes : Cfg::kStageBytes;
    int kbc = 0, tcp = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcp) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (tile / p.n_tiles) * kBlockM;
      if (EPI) {
        // residual tile of this output tile -> epilogue buffer (free once the stores of the tile that
        // used it two tiles ago have been read out)
        const int rb = tcp & 1;
        mbar_wait(&res_empty_bar[rb], ((tcp >> 1) & 1) ^ 1);
        int nsub = (p.N - n0 + 63) / 64;
        nsub = nsub > BN / 64 ? BN / 64 : nsub;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&res_full_bar[rb], (uint32_t)(nsub * Cfg::kEpiSubBytes));
          for (int j = 0; j < nsub; ++j)
            tma_load_2d(epi_base + rb * Cfg::kEpiBufBytes + j * Cfg::kEpiSubBytes, &tmR, &res_full_bar[rb], n0 + 64 * j, m0);
        }
        __syncwarp();
      }
      int wq = 0, hq = 0, dq = 0, nq = 0;
      if (!GATHER && p.a_mode == A_TMA_IM2COL) {
        int t = m0;
        const int wo = t % p.Wo; t /= p.Wo;
        const int ho = t % p.Ho; t /= p.Ho;
        const int to = t % p.To; t /= p.To;
        wq = wo * p.sw - p.pw;
        hq = ho * p.sh - p.ph;
        dq = to * p.st - p.pt;
        nq = t;
      }
      int c0 = 0, dw = 0, dh = 0, dt = 0;  // walks (tap, channel block) without divisions
      for (int kb = 0; kb < p.num_kb; ++kb, ++kbc) {
        const int s = kbc % STAGES;
        const uint32_t ph = (kbc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = stage_base + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + Cfg::kABytes;
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[s], tx);
          if (!GATHER) {
            if (p.a_mode == A_TMA_2D)
              tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
            else
              tma_load_im2col_5d(a_dst, &tmA, &full_bar[s], c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh, (uint16_t)dt);
          }
          tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
        }
        __syncwarp();
        if (!GATHER && p.a_mode != A_TMA_2D) {
          c0 += BK;
          if (c0 >= p.cin_eff) {
            c0 = 0;
            if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_m128(BN);
    const uint64_t desc_hi = umma_desc_kmajor<Cfg::kRowBytes>(0);  // everything but the start address
    int kbc = 0, tc = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      const int acc = tc & 1;
      const uint32_t aph = (tc >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < p.num_kb; ++kb, ++kbc) {
        const int s = kbc % STAGES;
        const uint32_t ph = (kbc / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = smem_u32(stage_base + s * Cfg::kStageBytes) >> 4;
        const uint64_t adesc = desc_hi | a_lo;
        const uint64_t bdesc = desc_hi | (a_lo + (Cfg::kABytes >> 4));
        if (elect_one_sync()) {
          // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (addr >> 4) field
          if (kb) umma_f16_c<true>(d_tmem, adesc, bdesc, idesc);
          else    umma_f16_c<false>(d_tmem, adesc, bdesc, idesc);
#pragma unroll
          for (int k = 1; k < BK / kUmmaK; ++k) umma_f16_c<true>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc);
          umma_commit(&empty_bar[s]);                                 // frees the smem stage once these MMAs have read it
          if (kb == p.num_kb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue warps 2..5
    const int t = threadIdx.x - 64;  // 0..127
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    int tc = 0, cached_n0 = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tc) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (tile / p.n_tiles) * kBlockM;
      const int acc = tc & 1;
      const uint32_t aph = (tc >> 1) & 1;
      if (n0 != cached_n0) {  // uniform across the 128 epilogue threads
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
      const bool row_ok = row < p.M;
      __nv_bfloat16* out_row = p.out + (long long)row * p.ldo + n0;
      const __nv_bfloat16* res_row = p.res ? p.res + (long long)row * p.ldr + n0 : nullptr;
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (EPI) {
        // residual (TMA-prefetched) is read from, and the bf16 result written back to, the same swizzled
        // shared-memory tile; each warp then TMA-stores its 32 rows (full 128-byte lines, clipped at M / N)
        mbar_wait(&res_full_bar[acc], aph);
        const int lrow = q * 32 + lane;
        const uint32_t buf = smem_u32(epi_base + acc * Cfg::kEpiBufBytes) + (uint32_t)lrow * 128u;
        const uint32_t xr = (uint32_t)(lrow & 7);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            const uint32_t addr = buf + (uint32_t)(col >> 6) * Cfg::kEpiSubBytes + ((((uint32_t)(col & 63) >> 3) ^ xr) << 4);
            uint4 r;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
            f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
            f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
            f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
            f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            const uint32_t o0 = pack_bf16x2(f[0], f[1]), o1 = pack_bf16x2(f[2], f[3]);
            const uint32_t o2 = pack_bf16x2(f[4], f[5]), o3 = pack_bf16x2(f[6], f[7]);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);  // accumulator drained: the MMA warp may start tile i+2
        fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
          int nsub = (p.N - n0 + 63) / 64;
          nsub = nsub > BN / 64 ? BN / 64 : nsub;
          for (int j = 0; j < nsub; ++j)
            tma_store_2d(&tmO, epi_base + acc * Cfg::kEpiBufBytes + j * Cfg::kEpiSubBytes + q * 32 * 128, n0 + 64 * j,
                         m0 + q * 32);
          tma_store_commit();
          tma_store_wait_read<0>();          // smem may be overwritten once the store engine has read it
          mbar_arrive(&res_empty_bar[acc]);  // 4 arrivals (one per epilogue warp) free the buffer
        }
        __syncwarp();
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            if (n0 + col < p.N) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
              if (res_row) {
                const uint4 r = *reinterpret_cast<const uint4*>(res_row + col);
                f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
                f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
                f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
                f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
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
      mbar_arrive(&tmem_empty_bar[acc]);  // 128 arrivals hand the accumulator back to the MMA warp
    }
  } else {
    // ------------------------------------------------------------------ gather warps 6..9 (GATHER only)
    if (GATHER) {
      constexpr int LAG = Cfg::kGatherLag;
      const int t = threadIdx.x - 192;  // 0..127: tile row owned by this thread
      const uint32_t sw_xor = (uint32_t)(t & 7);
      int g = 0;  // k-blocks issued so far (across tiles)
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
        const int w_base = wo * p.sw - p.pw;
        const int h_base = ho * p.sh - p.ph;
        const int t_base = to * p.st - p.pt;
        const __nv_bfloat16* img = p.in + (long long)nb * p.sN;
        // (tap, channel) cursor advanced 8 channels (one 16 B chunk) at a time; source pointer and
        // validity are recomputed only when the tap changes
        int c = 0, dw = 0, dh = 0, dt = 0, tap = 0;
        bool ok = false;
        const __nv_bfloat16* src = p.in;
        auto set_tap = [&]() {
          const int wi = w_base + dw, hi = h_base + dh, ti = t_base + dt;
          ok = row_ok && tap < p.ntaps && (unsigned)wi < (unsigned)p.Wi && (unsigned)hi < (unsigned)p.Hi &&
               (unsigned)ti < (unsigned)p.Ti;
          src = ok ? img + ti * p.sT + hi * p.sH + wi * p.sW : p.in;
        };
        set_tap();
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t dst_row = smem_u32(stage_base + s * Cfg::kStageBytes) + (uint32_t)t * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            cp_async_16_zfill(dst_row + (((uint32_t)j ^ sw_xor) << 4), src + (ok ? c : 0), ok ? 16u : 0u);
            c += 8;
            if (c == p.cin_eff) {
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
  }

  if (EPI && warp >= 2 && warp < 6 && lane == 0) tma_store_wait<0>();  // bulk stores fully complete
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vad
