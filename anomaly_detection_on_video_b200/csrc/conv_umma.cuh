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
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: weights (2D tiled map) and, in the TMA modes, the activation tile
//               (2D tiled map for 1x1x1/stride-1 layers, rank-5 im2col map otherwise)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  epilogue (TMEM -> registers -> scale/shift/residual/ReLU -> bf16 -> global);
//               in gather mode they first act as the A producer (cp.async 16 B with zero fill
//               into the 128B-swizzled tile; any stride / padding / folded stem window)
// Pipeline: STAGES-deep smem ring, full/empty mbarriers, accumulator in TMEM (BN fp32 columns).
#pragma once

#include "ptx_sm100.cuh"

namespace vad {

enum AMode : int {
  A_TMA_2D = 0,      // 1x1x1, stride 1: A is the [M, Cin] matrix itself
  A_TMA_IM2COL = 1,  // rank-5 im2col tensor map (Cin % 64 == 0)
  A_GATHER = 2       // cp.async gather (any Cin % 8 == 0, incl. the folded stem window)
};

struct ConvParams {
  int M, N, num_kb;
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
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kConvThreads = 192;
constexpr int kGatherLag = 2;  // cp.async groups kept in flight by each gather thread

template <int BN>
struct ConvCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // stages + scale/shift staging + barriers + tmem slot, + 1024 for manual alignment
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * BN * 4 + (2 * kStages + 1) * 8 + 16 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvParams p) {
  using Cfg = ConvCfg<BN>;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_tile = blockIdx.x % n_tiles;
  const int m_tile = blockIdx.x / n_tiles;
  const int m0 = m_tile * kBlockM;
  const int n0 = n_tile * BN;
  const bool gather = (p.a_mode == A_GATHER);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
    if (!gather) tma_prefetch_desc(&tmA);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], gather ? (1 + 128) : 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      const int n = n0 + i;
      s_scale[i] = (n < p.N) ? p.scale[n] : 0.f;
      s_shift[i] = (n < p.N) ? p.shift[n] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int wq = 0, hq = 0, dq = 0, nq = 0;
      if (p.a_mode == A_TMA_IM2COL) {
        int t = m0;
        const int wo = t % p.Wo; t /= p.Wo;
        const int ho = t % p.Ho; t /= p.Ho;
        const int to = t % p.To; t /= p.To;
        wq = wo * p.sw - p.pw;
        hq = ho * p.sh - p.ph;
        dq = to * p.st - p.pt;
        nq = t;
      }
      const int kb_per_tap = p.cin_eff / kBlockK;  // only meaningful in the TMA modes
      const uint32_t tx = gather ? Cfg::kBBytes : Cfg::kStageBytes;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = stage_base + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], tx);
        if (p.a_mode == A_TMA_2D) {
          tma_load_2d(a_dst, &tmA, &full_bar[s], kb * kBlockK, m0);
        } else if (p.a_mode == A_TMA_IM2COL) {
          const int tap = kb / kb_per_tap;
          const int c0 = (kb - tap * kb_per_tap) * kBlockK;
          const int dw = tap % p.kw;
          const int dh = (tap / p.kw) % p.kh;
          const int dt = tap / (p.kw * p.kh);
          tma_load_im2col_5d(a_dst, &tmA, &full_bar[s], c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh,
                             (uint16_t)dt);
        }
        tma_load_2d(b_dst, &tmB, &full_bar[s], kb * kBlockK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(stage_base + s * Cfg::kStageBytes);
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ warps 2..5
    const int t = threadIdx.x - 64;  // 0..127
    if (gather) {
      // A producer: thread t owns tile row t (one output pixel)
      const int m = m0 + t;
      const bool row_ok = m < p.M;
      int wo = 0, ho = 0, to = 0, nb = 0;
      if (row_ok) {
        int q = m;
        wo = q % p.Wo; q /= p.Wo;
        ho = q % p.Ho; q /= p.Ho;
        to = q % p.To; q /= p.To;
        nb = q;
      }
      const int w_base = wo * p.sw - p.pw;
      const int h_base = ho * p.sh - p.ph;
      const int t_base = to * p.st - p.pt;
      const __nv_bfloat16* img = p.in + (long long)nb * p.sN;
      const uint32_t sw_xor = (uint32_t)(t & 7);
      const bool tap_uniform = (p.cin_eff % kBlockK) == 0;
      const int total = p.num_kb + kGatherLag;
      for (int it = 0; it < total; ++it) {
        if (it < p.num_kb) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t dst_row = smem_u32(stage_base + s * Cfg::kStageBytes) + (uint32_t)t * 128u;
          if (tap_uniform) {
            const int k0 = it * kBlockK;
            const int tap = k0 / p.cin_eff;
            const int c0 = k0 - tap * p.cin_eff;
            const int dw = tap % p.kw;
            const int dh = (tap / p.kw) % p.kh;
            const int dt = tap / (p.kw * p.kh);
            const int wi = w_base + dw, hi = h_base + dh, ti = t_base + dt;
            const bool ok = row_ok && tap < p.ntaps && (unsigned)wi < (unsigned)p.Wi &&
                            (unsigned)hi < (unsigned)p.Hi && (unsigned)ti < (unsigned)p.Ti;
            const __nv_bfloat16* src = ok ? img + ti * p.sT + hi * p.sH + wi * p.sW + c0 : p.in;
            const uint32_t nbytes = ok ? 16u : 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              cp_async_16_zfill(dst_row + (((uint32_t)j ^ sw_xor) << 4), src + (ok ? j * 8 : 0), nbytes);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = it * kBlockK + j * 8;
              const int tap = k / p.cin_eff;
              const int c = k - tap * p.cin_eff;
              const int dw = tap % p.kw;
              const int dh = (tap / p.kw) % p.kh;
              const int dt = tap / (p.kw * p.kh);
              const int wi = w_base + dw, hi = h_base + dh, ti = t_base + dt;
              const bool ok = row_ok && tap < p.ntaps && (unsigned)wi < (unsigned)p.Wi &&
                              (unsigned)hi < (unsigned)p.Hi && (unsigned)ti < (unsigned)p.Ti;
              const __nv_bfloat16* src = ok ? img + ti * p.sT + hi * p.sH + wi * p.sW + c : p.in;
              cp_async_16_zfill(dst_row + (((uint32_t)j ^ sw_xor) << 4), src, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (it >= kGatherLag) {
          cp_async_wait<kGatherLag>();
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[(it - kGatherLag) % STAGES]);
        }
      }
    }

    // epilogue: warp q may only touch TMEM lanes [32q, 32q + 32)
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __nv_bfloat16* out_row = p.out + (long long)row * p.ldo + n0;
    const __nv_bfloat16* res_row = p.res ? p.res + (long long)row * p.ldr + n0 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace vad
