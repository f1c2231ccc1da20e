// K2p: the conv + folded-BN (+ReLU) implicit GEMM of conv_umma.cuh on CTA PAIRS (tcgen05 cta_group::2).
//
// Same math and the same reference layers as K2 (src/i3d.py:101-116 bottleneck conv1/conv2, :262-272 downsample),
// for the layers whose tile is 128 x 256 there: what bounds those is not the tensor pipe but the L2 -> shared-memory
// path (a 128 x 256 tile pulls 48 KB per 4.2 MFLOP k-block; at the ~12 TB/s the L2 delivers chip-wide that is
// ~1.05 PFLOP/s, which is what K2 measures).  Here two CTAs on the SMs of one TPC run ONE 256 x BN tile:
//   * each CTA loads its own 128 activation rows and only HALF of the weight rows (BN/2); tcgen05.mma.cta_group::2
//     (M = 256), issued by the even CTA's elected thread, reads A from each CTA's own shared memory and the two B
//     halves from both, and writes each CTA's 128 x BN accumulator into that CTA's own TMEM;
//   * so a CTA pulls 32 KB per 4.2 MFLOP (BN = 256): 1.5x the arithmetic intensity on the L2 path, with the
//     accumulator still double buffered (2 x BN TMEM columns per CTA) and the ring 6 stages deep.
// Synchronisation: every CTA's TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor .cta_group::2,
// barrier address mapped into CTA 0 with mapa); tcgen05.commit.cta_group::2 multicasts the slot release and the
// accumulator-ready signal into both CTAs; the epilogue warps of both CTAs arrive remotely on the leader's
// accumulator-empty barrier.  Operand order inside a k-block is identical to K2, so results are bit-identical.
#pragma once

#include "conv_umma.cuh"

namespace vad {

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                                        int d, int n, uint16_t off_w, uint16_t off_h, uint16_t off_d) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(off_w),
      "h"(off_h), "h"(off_d)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, 0, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
               "l"(adesc), "l"(bdesc), "r"(idesc)
               : "memory");
}
// arrive on the mbarrier at this offset in both CTAs once the previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 256 (two CTAs x 128 rows)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

template <int BN, int KPS>
struct PairCfg {
  static_assert(BN == 256 || BN == 128, "pair tile is 256 x BN");
  static constexpr int kABytes = kBlockM * 128;      // this CTA's 128 rows of one k-block
  static constexpr int kBBytes = (BN / 2) * 128;     // this CTA's half of the weight rows of one k-block
  static constexpr int kKbBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = KPS * kKbBytes;  // [A_0 .. A_{KPS-1}][B_0 .. B_{KPS-1}]
  static constexpr int kStages = (196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes);
  static constexpr int kEpiWarps = 8;
  static constexpr int kColsPerWarp = BN / 2;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreads = 64 + kEpiThreads;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * BN * 4 + (2 * kStages + 4) * 8 + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// tmB: weight map with (64 x BN/2) boxes.  Work item w -> (n tile = w % n_tiles, pair of m tiles = w / n_tiles); the
// cluster (blockIdx.x >> 1) walks items with stride gridDim.x / 2; CTA rank r of the pair owns rows (2 * pair + r) * 128.
// A pair whose odd CTA has no rows (odd number of m tiles) still loads: the boxes are out of bounds and arrive as zeros.
template <int BN, int KPS>
__global__ void __launch_bounds__(PairCfg<BN, KPS>::kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using Cfg = PairCfg<BN, KPS>;
  constexpr int STAGES = Cfg::kStages;
  const int crank = (int)cluster_ctarank();
  const int w_first = (int)(blockIdx.x >> 1), w_step = (int)(gridDim.x >> 1), w_total = p.mc_items;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* s_scale = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2] used in the leader only: arrivals from both CTAs' epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * Cfg::kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers and TMEM are set up before anything is signalled into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t lfull0 = mapa_u32(full0, 0);  // the leader's full barriers, shared::cluster address
      uint32_t s = 0, ph = 0;
      griddep_wait();
      for (int tile = w_first; tile < w_total; tile += w_step) {
        const int n0 = (tile % p.n_tiles) * BN + crank * (BN / 2);
        const int m0 = (2 * (tile / p.n_tiles) + crank) * kBlockM;
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
        int c0 = 0, dw = 0, dh = 0, dt = 0;
        for (int kb = 0; kb < p.num_kb; kb += KPS) {
          const int nk = (KPS == 1 || kb + KPS <= p.num_kb) ? KPS : p.num_kb - kb;
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t a_dst = stage0 + s * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + KPS * Cfg::kABytes;
          const uint32_t fb = lfull0 + s * 8;
          if (crank == 0) mbar_arrive_expect_tx_a(full0 + s * 8, (uint32_t)nk * 2u * Cfg::kKbBytes);  // both CTAs' bytes
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              if (p.a_mode == A_TMA_2D) {
                tma_load_2d_pair(a_dst + j * Cfg::kABytes, &tmA, fb, (kb + j) * 64, m0);
              } else {
                tma_load_im2col_5d_pair(a_dst + j * Cfg::kABytes, &tmA, fb, c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh, (uint16_t)dt);
                c0 += 64;
                if (c0 >= p.cin_eff) {
                  c0 = 0;
                  if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
                }
              }
              tma_load_2d_pair(b_dst + j * Cfg::kBBytes, &tmB, fb, (kb + j) * 64, n0);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (crank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m256(BN);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      uint32_t s = 0, ph = 0;
      int tc = 0;
      mbar_wait_a(full0, 0);
      for (int tile = w_first; tile < w_total; tile += w_step) {
        const uint32_t acc = (uint32_t)tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, (((uint32_t)tc >> 1) & 1u) ^ 1u);  // both CTAs' epilogues have drained it
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const bool last_tile = tile + w_step >= w_total;
        for (int kb = 0; kb < p.num_kb; kb += KPS) {
          const int nk = (KPS == 1 || kb + KPS <= p.num_kb) ? KPS : p.num_kb - kb;
          const bool last_stage = kb + KPS >= p.num_kb;
          uint32_t ns = s + 1, nph = ph;
          if (ns == STAGES) { ns = 0; nph ^= 1u; }
          const uint32_t a_lo = (stage0 + s * Cfg::kStageBytes) >> 4;
          const uint32_t b_lo = a_lo + ((KPS * Cfg::kABytes) >> 4);
          const bool do_wait = !(last_stage && last_tile);
          bool ready = !do_wait;
          auto issue = [&](auto nk_c) {
            constexpr int NK = decltype(nk_c)::value;
            constexpr int WAIT_IDX = (NK * 4 * 3) / 4 - 1;
#pragma unroll
            for (int j = 0; j < NK; ++j) {
              const uint64_t adesc = desc_hi | (a_lo + j * (Cfg::kABytes >> 4));
              const uint64_t bdesc = desc_hi | (b_lo + j * (Cfg::kBBytes >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (j == 0 && k == 0) umma_f16_pair(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
                else                  umma_f16_pair_acc(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc);
                if (j * 4 + k == WAIT_IDX && do_wait) {
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
          };
          if (KPS == 1 || nk == KPS) issue(std::integral_constant<int, KPS>{});
          else                       issue(std::integral_constant<int, 1>{});
          umma_commit_pair(empty0 + s * 8);                               // frees the slot in both CTAs
          if (last_stage) { umma_commit_pair(tfull0 + acc * 8); ++tc; }  // both CTAs' accumulators complete
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
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    constexpr int CPW = Cfg::kColsPerWarp;
    griddep_wait();
    const int t = threadIdx.x - 64;
    const int q = warp & 3;
    const int col0 = ((warp - 2) >> 2) * CPW;
    const uint32_t ltempty0 = mapa_u32(smem_u32(tmem_empty_bar), 0);
    int tc = 0, cached_n0 = -1;
    for (int tile = w_first; tile < w_total; tile += w_step) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (2 * (tile / p.n_tiles) + crank) * kBlockM;
      const int acc = tc & 1;
      const uint32_t aph = (tc >> 1) & 1;
      ++tc;
      if (n0 != cached_n0) {
        named_bar_sync(1, Cfg::kEpiThreads);
        for (int i = t; i < BN; i += Cfg::kEpiThreads) {
          const int n = n0 + i;
          s_scale[i] = (n < p.N) ? p.scale[n] : 0.f;
          s_shift[i] = (n < p.N) ? p.shift[n] : 0.f;
        }
        named_bar_sync(1, Cfg::kEpiThreads);
        cached_n0 = n0;
      }
      const int row = m0 + q * 32 + lane;
      mbar_wait(&tmem_full_bar[acc], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0);
      const bool row_ok = row < p.M;
      __nv_bfloat16* out_row = p.out + (long long)row * p.ldo + n0 + col0;
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
              for (int j = 0; j < 8; ++j)
                f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col0 + col + j], s_shift[col0 + col + j]);
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
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);  // hands this CTA's half of the accumulator back to the leader
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still signal into this CTA or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vad
