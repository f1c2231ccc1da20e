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
//
// EPI = true is the residual form (bottleneck conv3 + BN + identity + ReLU, src/i3d.py:112-116) for layers 3 and 4: each
// CTA's 128 x 256 result passes through three 32 KB staging tiles as 128-column halves -- an extra warp TMA-prefetches
// the residual half-tile, the epilogue warps add it in place and TMA-store the sum -- the staged epilogue of K2 with
// twice the weight reuse (K2 runs these layers as 128 x 128 tiles: 64 FLOP per byte pulled from L2, here 128).
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
// kind::tf32: fp32 operands read from shared memory (128-byte rows = 32 floats per k-block, 8 per MMA), fp32 accumulate
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

// instruction descriptor: tf32 x tf32 -> fp32, both operands K-major, M = 256
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m256(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}
// TF32 rounding of a stored activation (ties away from zero): see tf32_kernels.cuh
__device__ __forceinline__ float pair_tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

constexpr int kPairXposeBytes = 8 * 4096 + 128;   // F32 instantiations: per-warp transposition tiles of the direct epilogue

template <int BN, int KPS, bool EPI = false>
struct PairCfg {
  static_assert(BN == 256 || BN == 128, "pair tile is 256 x BN");
  static_assert(!EPI || (BN == 256 && KPS == 1), "staged residual epilogue: 256 x 256 pair tiles");
  static constexpr int kABytes = kBlockM * 128;      // this CTA's 128 rows of one k-block
  static constexpr int kBBytes = (BN / 2) * 128;     // this CTA's half of the weight rows of one k-block
  static constexpr int kKbBytes = kABytes + kBBytes;
  static constexpr int kStageBytes = KPS * kKbBytes;  // [A_0 .. A_{KPS-1}][B_0 .. B_{KPS-1}]
  // EPI: this CTA's 128 x 256 output goes through the staging tiles as two 128 x 128 halves (residual prefetched by
  // TMA into the tile, result written over it, TMA store), three tiles in rotation: each is two [128 rows x 64 cols]
  // SWIZZLE_128B sub-tiles
  static constexpr int kEpiSubBytes = kBlockM * 128;
  static constexpr int kEpiBufBytes = EPI ? 2 * kEpiSubBytes : 0;
  static constexpr int kEpiBufs = 3;
  static constexpr int kPipeBudget = EPI ? 131072 : 196608;
  static constexpr int kStages = (kPipeBudget / kStageBytes) > 8 ? 8 : (kPipeBudget / kStageBytes);
  static constexpr int kEpiWarps = 8;
  static constexpr int kColsPerWarp = EPI ? 64 : BN / 2;
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreads = 64 + kEpiThreads + (EPI ? 32 : 0);  // EPI: one more warp that only prefetches residual tiles
  static constexpr int kTmemCols = 2 * BN;
  // the dynamic shared memory of these kernels is declared 1024-byte aligned (checked at run time): no slack for
  // manual alignment, which the EPI layout (4 x 32 KB stages + 3 x 32 KB staging tiles) could not afford
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBufs * kEpiBufBytes + 2 * BN * 4 + (2 * kStages + 4 + 2 * kEpiBufs) * 8 + 16;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

// tmB: weight map with (64 x BN/2) boxes.  Work item w -> (n tile = w % n_tiles, pair of m tiles = w / n_tiles); the
// cluster (blockIdx.x >> 1) walks items with stride gridDim.x / 2; CTA rank r of the pair owns rows (2 * pair + r) * 128.
// A pair whose odd CTA has no rows (odd number of m tiles) still loads: the boxes are out of bounds and arrive as zeros.
//
// F32 = true is the TF32 precision mode (tf32_api.cuh): the tensor maps carry fp32 elements, a k-block is 32 floats -- the
// same 128-byte rows, so stages, descriptors and barriers are unchanged --, the MMAs are kind::tf32, and the direct epilogue
// reads an fp32 residual and writes TF32-rounded fp32 activations (p.res / p.out then point at floats).
template <int BN, int KPS, bool EPI, bool F32 = false>
__global__ void __launch_bounds__(PairCfg<BN, KPS, EPI>::kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO, const ConvParams p) {
  using Cfg = PairCfg<BN, KPS, EPI>;
  static_assert(!(F32 && EPI), "the TF32 mode uses the direct epilogue");
  constexpr int STAGES = Cfg::kStages;
  constexpr int NB = Cfg::kEpiBufs;
  constexpr int KBE = F32 ? 32 : 64;   // elements of one k-block (one 128-byte row)
  const int crank = (int)cluster_ctarank();
  // SPLIT (BN = 256, direct epilogue, one n tile): the items of the last, partly filled round run as two 128-column halves
  // each, so 3.3 rounds of work take 3.5 rounds instead of 4 (wave quantisation of the 74 CTA pairs)
  constexpr bool SPLIT = BN == 256 && !EPI;
  const int w_first = (int)(blockIdx.x >> 1), w_step = (int)(gridDim.x >> 1);
  const int w_split = SPLIT ? p.pair_split : p.mc_items, w_total = SPLIT ? p.pair_total : p.mc_items;
  // item -> (m pair, first column, half-width?)
  auto item_mpair = [&](int w) { return w < w_split ? w / p.n_tiles : w_split / p.n_tiles + ((w - w_split) >> 1); };
  auto item_n0 = [&](int w) { return w < w_split ? (w % p.n_tiles) * BN : ((w - w_split) & 1) * (BN / 2); };

  extern __shared__ __align__(1024) uint8_t pair_smem[];
  uint8_t* smem = pair_smem;
  if (smem_u32(smem) & 1023u) __trap();  // swizzled TMA boxes and UMMA descriptors need 1024-byte aligned tiles
  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + STAGES * Cfg::kStageBytes;
  float* s_scale = reinterpret_cast<float*>(epi_base + NB * Cfg::kEpiBufBytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2] used in the leader only: arrivals from both CTAs' epilogue warps
  uint64_t* res_full_bar = tmem_empty_bar + 2;   // [NB] residual half-tile landed (EPI)
  uint64_t* res_empty_bar = res_full_bar + NB;   // [NB] staging tile free again (EPI)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty_bar + NB);
  // F32: 8 x 4 KB transposition tiles of the epilogue warps, behind everything else (the launch adds kPairXposeBytes)
  uint8_t* xpose_base = smem + ((Cfg::kSmemBytes + 127) & ~127);

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
    for (int a = 0; a < NB; ++a) {
      mbar_init(&res_full_bar[a], 1);
      mbar_init(&res_empty_bar[a], Cfg::kEpiWarps);
    }
    if (EPI) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmO);
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
        const bool half = SPLIT && tile >= w_split;
        const int b_rows = half ? BN / 4 : BN / 2;  // weight rows this CTA supplies
        const int n0 = item_n0(tile) + crank * b_rows;
        const int m0 = (2 * item_mpair(tile) + crank) * kBlockM;
        const uint32_t stage_tx = (uint32_t)(Cfg::kABytes + b_rows * 128);
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
          if (crank == 0) mbar_arrive_expect_tx_a(full0 + s * 8, (uint32_t)nk * 2u * stage_tx);  // both CTAs' bytes
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              if (p.a_mode == A_TMA_2D) {
                tma_load_2d_pair(a_dst + j * Cfg::kABytes, &tmA, fb, (kb + j) * KBE, m0);
              } else {
                tma_load_im2col_5d_pair(a_dst + j * Cfg::kABytes, &tmA, fb, c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh, (uint16_t)dt);
                c0 += KBE;
                if (c0 >= p.cin_eff) {
                  c0 = 0;
                  if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
                }
              }
              if (SPLIT && p.pair_box_rows == 64) {  // 64-row weight boxes: two for a full-width item, one for a half
                tma_load_2d_pair(b_dst + j * Cfg::kBBytes, &tmB, fb, (kb + j) * KBE, n0);
                if (!half) tma_load_2d_pair(b_dst + j * Cfg::kBBytes + 64 * 128, &tmB, fb, (kb + j) * KBE, n0 + 64);
              } else {
                tma_load_2d_pair(b_dst + j * Cfg::kBBytes, &tmB, fb, (kb + j) * KBE, n0);
              }
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
      constexpr uint32_t idesc_full = F32 ? umma_idesc_tf32_m256(BN) : umma_idesc_bf16_m256(BN);
      constexpr uint32_t idesc_half = F32 ? umma_idesc_tf32_m256(BN / 2) : umma_idesc_bf16_m256(BN / 2);
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
        const uint32_t idesc = (SPLIT && tile >= w_split) ? idesc_half : idesc_full;
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
                if constexpr (F32) umma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | j | k) ? 1u : 0u);
                else if (j == 0 && k == 0) umma_f16_pair(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
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
  } else if (warp < 2 + Cfg::kEpiWarps) {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    constexpr int CPW = Cfg::kColsPerWarp;
    griddep_wait();
    const int t = threadIdx.x - 64;
    const int q = warp & 3;
    const int col0 = ((warp - 2) >> 2) * CPW;
    const uint32_t ltempty0 = mapa_u32(smem_u32(tmem_empty_bar), 0);
    const bool has_res = EPI && p.res != nullptr;
    uint32_t eb = 0, eph = 0;  // staging tile of the current half and its phase (EPI)
    int prev_eb = -1;          // staging tile whose TMA store has been issued but not yet waited for
    int tc = 0, cached_n0 = -1;
    for (int tile = w_first; tile < w_total; tile += w_step) {
      const int n0 = item_n0(tile);
      const int m0 = (2 * item_mpair(tile) + crank) * kBlockM;
      const bool half = SPLIT && tile >= w_split;        // a 128-column item of the tail round: 64 columns per warp
      const int cpw_i = half ? CPW / 2 : CPW;
      const int col0_i = ((warp - 2) >> 2) * cpw_i;
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
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + col0_i);
      if (EPI) {
        // two 128-column halves; in each, this warp owns rows [32q, 32q+32) x columns [col0, col0+64) of the staging tile
        const int lrow = q * 32 + lane;
        const uint32_t xr = (uint32_t)(lrow & 7);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (has_res) mbar_wait(&res_full_bar[eb], eph);
          const uint32_t sub = smem_u32(epi_base + eb * Cfg::kEpiBufBytes) + (uint32_t)(col0 >> 6) * Cfg::kEpiSubBytes;
          const uint32_t buf = sub + (uint32_t)lrow * 128u;
          auto chunk = [&](const uint32_t (&v)[32], int c, auto res_c) {
            constexpr bool RES = decltype(res_c)::value;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = c * 32 + g * 8;  // inside this warp's 64 columns
              const uint32_t addr = buf + ((((uint32_t)col >> 3) ^ xr) << 4);
              const int sc = h * 128 + col0 + col;
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[sc + j], s_shift[sc + j]);
              if (RES) {
                uint4 r;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
                f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
                f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
                f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
                f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
              }
              // ReLU rides in the conversion (cvt.rn.relu.bf16x2.f32): same bits as max(x, 0) followed by the rounding
              uint32_t o0, o1, o2, o3;
              if (p.relu) { o0 = pack_bf16x2_relu(f[0], f[1]); o1 = pack_bf16x2_relu(f[2], f[3]); o2 = pack_bf16x2_relu(f[4], f[5]); o3 = pack_bf16x2_relu(f[6], f[7]); }
              else        { o0 = pack_bf16x2(f[0], f[1]); o1 = pack_bf16x2(f[2], f[3]); o2 = pack_bf16x2(f[4], f[5]); o3 = pack_bf16x2(f[6], f[7]); }
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
            }
          };
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(taddr + (uint32_t)(h * 128), v0);
          tmem_ld_32x32(taddr + (uint32_t)(h * 128 + 32), v1);
          tmem_ld_wait();
          if (h == 1) {
            // both halves are in registers / shared memory: hand this CTA's accumulator back to the leader
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);
          }
          if (has_res) {
            chunk(v0, 0, std::true_type{});
            chunk(v1, 1, std::true_type{});
          } else {
            chunk(v0, 0, std::false_type{});
            chunk(v1, 1, std::false_type{});
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO, epi_base + eb * Cfg::kEpiBufBytes + (col0 >> 6) * Cfg::kEpiSubBytes + q * 32 * 128, n0 + h * 128 + col0,
                         m0 + q * 32);
            tma_store_commit();
            tma_store_wait_read<1>();  // the PREVIOUS half's store has left shared memory (this one's read-out stays off the critical path)
            if (has_res && prev_eb >= 0) mbar_arrive(&res_empty_bar[prev_eb]);  // one arrival per epilogue warp frees that staging tile
          }
          __syncwarp();
          prev_eb = (int)eb;
          if (++eb == NB) { eb = 0; eph ^= 1u; }
        }
        continue;
      }
      const bool row_ok = row < p.M;
      if constexpr (F32) {
        // fp32 rows: a thread's 32 accumulator columns are one full 128-byte line of ITS row, so register-direct accesses
        // would touch 32 different lines per instruction (the layer1 conv3 launches ran at 0.43 of HBM that way).  Each warp
        // transposes through its own 4 KB shared-memory tile (32 rows x 8 sixteen-byte chunks, chunk index XOR row & 7:
        // conflict-free both ways): residual and output move as four whole 128-byte row segments per instruction.
        float* outp = reinterpret_cast<float*>(p.out);
        const float* resp = reinterpret_cast<const float*>(p.res);
        const uint32_t tw = smem_u32(xpose_base) + (uint32_t)(warp - 2) * 4096u;
        const int r_sub = lane >> 3, c_sub = lane & 7;   // coalesced phases: this lane covers row 4 k + r_sub, chunk c_sub
        const uint32_t own = tw + (uint32_t)lane * 128u;
        const uint32_t lx = (uint32_t)(lane & 7);
#pragma unroll 1
        for (int c = 0; c < cpw_i / 32; ++c) {
          const int colbase = n0 + col0_i + c * 32;
          uint32_t v[32];
          tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
          float4 rr[8];
          if (resp) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int row_l = 4 * k + r_sub;
              const long long grow = (long long)m0 + q * 32 + row_l;
              float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (grow < p.M && colbase + c_sub * 4 < p.N) rv = *reinterpret_cast<const float4*>(resp + grow * p.ldr + colbase + c_sub * 4);
              const uint32_t a = tw + (uint32_t)row_l * 128u + ((((uint32_t)c_sub) ^ (uint32_t)(row_l & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(rv.x), "f"(rv.y), "f"(rv.z), "f"(rv.w) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(rr[j].x), "=f"(rr[j].y), "=f"(rr[j].z), "=f"(rr[j].w)
                           : "r"(own + ((((uint32_t)j) ^ lx) << 4)));
            __syncwarp();
          }
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int sc = col0_i + c * 32 + g * 4;
            float4 f;
            f.x = fmaf(__uint_as_float(v[g * 4 + 0]), s_scale[sc + 0], s_shift[sc + 0]);
            f.y = fmaf(__uint_as_float(v[g * 4 + 1]), s_scale[sc + 1], s_shift[sc + 1]);
            f.z = fmaf(__uint_as_float(v[g * 4 + 2]), s_scale[sc + 2], s_shift[sc + 2]);
            f.w = fmaf(__uint_as_float(v[g * 4 + 3]), s_scale[sc + 3], s_shift[sc + 3]);
            if (resp) { f.x += rr[g].x; f.y += rr[g].y; f.z += rr[g].z; f.w += rr[g].w; }
            if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
            f.x = pair_tf32_rna(f.x); f.y = pair_tf32_rna(f.y); f.z = pair_tf32_rna(f.z); f.w = pair_tf32_rna(f.w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(own + ((((uint32_t)g) ^ lx) << 4)), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w)
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row_l = 4 * k + r_sub;
            const long long grow = (long long)m0 + q * 32 + row_l;
            float4 o;
            const uint32_t a = tw + (uint32_t)row_l * 128u + ((((uint32_t)c_sub) ^ (uint32_t)(row_l & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(a));
            if (grow < p.M && colbase + c_sub * 4 < p.N) *reinterpret_cast<float4*>(outp + grow * p.ldo + colbase + c_sub * 4) = o;
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);
        continue;
      }
      __nv_bfloat16* out_row = p.out + (long long)row * p.ldo + n0 + col0_i;
#pragma unroll 1
      for (int c = 0; c < cpw_i / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            if (n0 + col0_i + col < p.N) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col0_i + col + j], s_shift[col0_i + col + j]);
              uint4 o;
              if (p.relu) { o.x = pack_bf16x2_relu(f[0], f[1]); o.y = pack_bf16x2_relu(f[2], f[3]); o.z = pack_bf16x2_relu(f[4], f[5]); o.w = pack_bf16x2_relu(f[6], f[7]); }
              else        { o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]); }
              *reinterpret_cast<uint4*>(out_row + col) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ltempty0 + acc * 8);  // hands this CTA's half of the accumulator back to the leader
    }
  } else {
    // ------------------------------------------------------------------ residual prefetch warp (EPI, both CTAs)
    if (EPI && p.res != nullptr && elect_one_sync()) {
      const uint32_t res_full0 = smem_u32(res_full_bar), res_empty0 = smem_u32(res_empty_bar), epi0 = smem_u32(epi_base);
      uint32_t rb = 0, rph = 0;
      griddep_wait();
      for (int tile = w_first; tile < w_total; tile += w_step) {
        const int n0 = (tile % p.n_tiles) * BN;
        const int m0 = (2 * (tile / p.n_tiles) + crank) * kBlockM;
        for (int h = 0; h < 2; ++h) {
          mbar_wait_a(res_empty0 + rb * 8, rph ^ 1u);
          mbar_arrive_expect_tx_a(res_full0 + rb * 8, (uint32_t)Cfg::kEpiBufBytes);
          tma_load_2d_a(epi0 + rb * Cfg::kEpiBufBytes, &tmR, res_full0 + rb * 8, n0 + h * 128, m0);
          tma_load_2d_a(epi0 + rb * Cfg::kEpiBufBytes + Cfg::kEpiSubBytes, &tmR, res_full0 + rb * 8, n0 + h * 128 + 64, m0);
          if (++rb == NB) { rb = 0; rph ^= 1u; }
        }
      }
    }
    __syncwarp();
  }

  if (EPI && warp >= 2 && warp < 2 + Cfg::kEpiWarps && lane == 0) tma_store_wait<0>();  // bulk stores fully complete
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer may still signal into this CTA or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vad
