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

#include <type_traits>

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
  int mc_items;          // CTA-pair kernels: work items = ceil(m_tiles / 2) * n_tiles (a cluster of two CTAs per item)
  int pair_split;        // conv_pair_kernel<256>: items [0, pair_split) are 256 x 256 tiles; item pair_split + h is the
  int pair_total;        //   (h & 1)-th 128-column half of m-pair pair_split + h / 2 (tail round at half width); total items
  int pair_box_rows;     //   rows of one weight box (128, or 64 when there are half-width items)
  int To, Ho, Wo;
  int Ti, Hi, Wi;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int cin_eff;  // contraction width of one tap (channels, or 32 for the folded stem window)
  int ntaps;
  long long sN, sT, sH, sW;  // input strides in elements
  int a_mode;
  int pool_tp;           // 1: M tiles are (all 4 frames x 32 pixels) and the staged epilogue max-reduces frame pairs (maxpool2 fused)
  int tp_tiles_per_clip; // ceil(H * W / 32) when pool_tp
  int relu;
  int ldo;  // output row pitch in elements
  int ldr;  // residual row pitch in elements
  const __nv_bfloat16* in;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
  // fused sibling 1x1x1 convs (vad_op_desc.dst1 ...; direct epilogue of conv_umma_kernel): output columns
  // [split1, split2) go to out1 (row pitch ldo1), [split2, N) to out2; of each part only the first seg_w columns are stored
  __nv_bfloat16* out1;
  __nv_bfloat16* out2;
  int ldo1, ldo2, split1, split2, seg_w0, seg_w1, seg_w2;
};

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;

template <int BN, int BK, int KPS, bool GATHER, bool EPI>
struct ConvCfg {
  static_assert(BK == 64 || BK == 32 || BK == 16, "BK is one swizzle row: 64 (SW128), 32 (SW64) or 16 (SW32) bf16");
  static_assert(!GATHER || BK == 64, "the gather producer writes 128-byte swizzled rows");
  static_assert(!EPI || BN <= 128, "the staged epilogue keeps two 128 x BN bf16 tiles in shared memory");
  static_assert(KPS == 1 || (KPS == 2 && !GATHER && BK == 64) || (KPS == 4 && !GATHER && BK == 32) || (KPS == 8 && !GATHER && BK == 16),
                "several k-blocks per stage (TMA producers): 2 x 64-wide (SW128), 4 x 32-wide (SW64) or 8 x 16-wide (SW32), i.e. 8 MMAs");
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kABytes = kBlockM * kRowBytes;   // one k-block of A
  static constexpr int kBBytes = BN * kRowBytes;        // one k-block of B
  static constexpr int kKbBytes = kABytes + kBBytes;
  // A stage holds KPS k-blocks behind ONE full/empty barrier pair: [A_0 .. A_{KPS-1}][B_0 .. B_{KPS-1}].
  // The issuing thread can run only about one MMA ahead of the tensor pipe, so whatever it does between two
  // groups of MMAs (barrier wait, commit) is exposed unless a group is >= ~300 cycles of tensor work
  // (tools/umma_bench.cu): 8 MMAs per barrier for BN <= 128, 4 for BN = 256.
  static constexpr int kStageBytes = KPS * kKbBytes;
  // EPI: residual tile prefetched by TMA / output tile written back by TMA, double buffered;
  // laid out as BN/64 sub-tiles of [128 rows x 64 cols] (128-byte rows, SWIZZLE_128B)
  static constexpr int kEpiSubBytes = kBlockM * 128;
  static constexpr int kEpiBufBytes = EPI ? (BN / 64) * kEpiSubBytes : 0;
  // The residual of tile i+kEpiBufs is requested when the epilogue of tile i has released its buffer, so the
  // HBM round trip of that request has kEpiBufs-1 tile times to complete: three buffers, not two.
  static constexpr int kEpiBufs = 3;
  static constexpr int kPipeBudget = 196608 + (EPI ? 32768 : 0) - kEpiBufs * kEpiBufBytes;
  static constexpr int kStages = (kPipeBudget / kStageBytes) > 8 ? 8 : (kPipeBudget / kStageBytes);
  static_assert(kStages >= 2, "need a ring of at least two stages");
  // Epilogue warps: the TMEM lane quarter a warp may read is (warp % 4); for BN >= 128 two warps share a
  // quarter and split the columns, which halves the per-tile epilogue latency of the small-K layers.
  static constexpr int kEpiWarps = BN >= 128 ? 8 : 4;
  static constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
  static constexpr int kEpiThreads = kEpiWarps * 32;
  static constexpr int kThreads = 64 + kEpiThreads + (GATHER ? 128 : 0);
  static constexpr int kGatherLag = kStages - 2 > 6 ? 6 : kStages - 2;  // cp.async groups in flight per thread
  static constexpr int kTmemCols = 2 * BN;                              // two accumulator stages
  // stages + epilogue buffers + scale/shift staging + barriers + tmem slot, + 1024 for manual alignment
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiBufs * kEpiBufBytes + 2 * BN * 4 + (2 * kStages + 4 + 2 * kEpiBufs) * 8 + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// nk (1 .. N, runtime) -> f(integral_constant<nk>): the issue loop is instantiated per k-block count of a stage
template <int N, class F>
__device__ __forceinline__ void call_with_nk(int nk, F& f) {
  if (nk == N) f(std::integral_constant<int, N>{});
  else if constexpr (N > 1) call_with_nk<N - 1>(nk, f);
}

// K-major operand tile with ROW_BYTES-wide rows (128: SWIZZLE_128B, 64: SWIZZLE_64B, 32: SWIZZLE_32B), rows packed
// densely, 8-row groups SBO apart.
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);            // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                                // LBO (ignored)   [16,30)
  d |= static_cast<uint64_t>((8 * ROW_BYTES) >> 4) << 32;             // SBO             [32,46)
  d |= static_cast<uint64_t>(1) << 46;                                // descriptor version (sm_100)
  d |= static_cast<uint64_t>(ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6)) << 61;  // SWIZZLE_128B / _64B / _32B
  return d;
}

__device__ __forceinline__ void tma_load_4d_b(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d_b(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <int BN, int BK, int KPS, bool GATHER, bool EPI>
__global__ void __launch_bounds__(ConvCfg<BN, BK, KPS, GATHER, EPI>::kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmO2, const ConvParams p) {
  // (tmO2: third output of a fused sibling conv in the staged epilogue; its second output travels in the tmR slot, which such
  // an op -- no residual -- does not otherwise use)
  using Cfg = ConvCfg<BN, BK, KPS, GATHER, EPI>;
  constexpr int STAGES = Cfg::kStages;
  // work distribution: item w -> (n tile, m tile), n fastest
  const int w_first = (int)blockIdx.x;
  const int w_step = (int)gridDim.x;
  const int w_total = p.num_tiles;
  auto m_tile_of = [&](int w) { return w / p.n_tiles; };

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* stage_base = smem;
  uint8_t* epi_base = smem + STAGES * Cfg::kStageBytes;  // kEpiBufs x kEpiBufBytes, 1024-aligned
  constexpr int NB = Cfg::kEpiBufs;
  float* s_scale = reinterpret_cast<float*>(epi_base + NB * Cfg::kEpiBufBytes);
  float* s_shift = s_scale + BN;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* res_full_bar = tmem_empty_bar + 2;    // [NB] residual tile landed (EPI)
  uint64_t* res_empty_bar = res_full_bar + NB;    // [NB] epilogue buffer free again (EPI)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty_bar + NB);

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
      mbar_init(&tmem_empty_bar[a], Cfg::kEpiWarps);  // one arrival per epilogue warp
    }
    for (int a = 0; a < NB; ++a) {
      mbar_init(&res_full_bar[a], 1);
      mbar_init(&res_empty_bar[a], Cfg::kEpiWarps);
    }
    if (EPI) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmO);
      if (p.split1) tma_prefetch_desc(&tmO2);
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
  griddep_launch_dependents();  // the next kernel of the stream may begin its own prologue as SMs free up

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // One elected thread owns the whole loop: no per-stage election / reconvergence, and every barrier and
    // stage address is a plain shared-window register.
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t res_full0 = smem_u32(res_full_bar), res_empty0 = smem_u32(res_empty_bar), epi0 = smem_u32(epi_base);
      uint32_t s = 0, ph = 0;    // ring slot and its phase
      uint32_t rb = 0, rph = 0;  // epilogue buffer of this tile and its phase
      griddep_wait();            // activations / residuals of the previous kernel are complete from here on
      for (int tile = w_first; tile < w_total; tile += w_step) {
        const int n0 = (tile % p.n_tiles) * BN;
        const int m0 = m_tile_of(tile) * kBlockM;
        int tp_clip = 0, tp_p0 = 0;  // pool_tp: tile = (clip, 32 pixels) x all 4 frames
        if (EPI && p.pool_tp) {
          const int mt = tile / p.n_tiles;
          tp_clip = mt / p.tp_tiles_per_clip;
          tp_p0 = (mt - tp_clip * p.tp_tiles_per_clip) * 32;
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
        for (int kb = 0; kb < p.num_kb; kb += KPS) {
          const int nk = (KPS == 1 || kb + KPS <= p.num_kb) ? KPS : p.num_kb - kb;
          mbar_wait_a(empty0 + s * 8, ph ^ 1u);
          const uint32_t a_dst = stage0 + s * Cfg::kStageBytes;
          const uint32_t b_dst = a_dst + KPS * Cfg::kABytes;
          const uint32_t fb = full0 + s * 8;
          mbar_arrive_expect_tx_a(fb, (uint32_t)nk * (GATHER ? Cfg::kBBytes : Cfg::kKbBytes));
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            if (j < nk) {
              if (!GATHER) {
                if (EPI && p.pool_tp) {
                  tma_load_4d_b(a_dst + j * Cfg::kABytes, &tmA, fb, (kb + j) * BK, tp_p0, 0, tp_clip);
                } else if (p.a_mode == A_TMA_2D) {
                  tma_load_2d_a(a_dst + j * Cfg::kABytes, &tmA, fb, (kb + j) * BK, m0);
                } else {
                  tma_load_im2col_5d_a(a_dst + j * Cfg::kABytes, &tmA, fb, c0, wq, hq, dq, nq, (uint16_t)dw, (uint16_t)dh,
                                       (uint16_t)dt);
                  c0 += BK;
                  if (c0 >= p.cin_eff) {
                    c0 = 0;
                    if (++dw == p.kw) { dw = 0; if (++dh == p.kh) { dh = 0; ++dt; } }
                  }
                }
              }
              tma_load_2d_a(b_dst + j * Cfg::kBBytes, &tmB, fb, (kb + j) * BK, n0);
            }
          }
          if (EPI && kb == 0 && p.res != nullptr) {
            // residual tile of this output tile -> epilogue buffer (free once the stores of the tile that used it
            // two tiles ago have been read out).  Issued after the tile's first operand stage so that the MMAs
            // never queue behind the epilogue of an older tile.
            mbar_wait_a(res_empty0 + rb * 8, rph ^ 1u);
            int nsub = (p.N - n0 + 63) / 64;
            nsub = nsub > BN / 64 ? BN / 64 : nsub;
            mbar_arrive_expect_tx_a(res_full0 + rb * 8, (uint32_t)(nsub * Cfg::kEpiSubBytes));
            for (int j = 0; j < nsub; ++j) {
              if (p.pool_tp)
                tma_load_4d_b(epi0 + rb * Cfg::kEpiBufBytes + j * Cfg::kEpiSubBytes, &tmR, res_full0 + rb * 8, n0 + 64 * j, tp_p0, 0, tp_clip);
              else
                tma_load_2d_a(epi0 + rb * Cfg::kEpiBufBytes + j * Cfg::kEpiSubBytes, &tmR, res_full0 + rb * 8, n0 + 64 * j, m0);
            }
            if (++rb == NB) { rb = 0; rph ^= 1u; }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // One elected thread, software pipelined: the wait for stage g+1 sits between the MMAs of stage g, so the
    // barrier round trip is covered by tensor work that is already queued.
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m128(BN);
      constexpr int MMAS_PER_KB = BK / kUmmaK;
      const uint64_t desc_hi = umma_desc_kmajor<Cfg::kRowBytes>(0);  // everything but the start address
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      const uint32_t tfull0 = smem_u32(tmem_full_bar), tempty0 = smem_u32(tmem_empty_bar);
      uint32_t s = 0, ph = 0;
      int tc = 0;
      mbar_wait_a(full0, 0);
      for (int tile = w_first; tile < w_total; tile += w_step) {
        const uint32_t acc = (uint32_t)tc & 1u;
        mbar_wait_a(tempty0 + acc * 8, (((uint32_t)tc >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
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
          const bool do_wait = !(last_stage && last_tile);  // false: this CTA's very last stage
          bool ready = !do_wait;
          // NK k-blocks, fully unrolled; the wait for the next stage follows MMA number 3/4 * (NK * MMAS_PER_KB).
          // On a tile's last stage it is only a probe: the accumulator must be handed to the epilogue even if
          // the next tile's operands are late (their producer may itself be waiting for that epilogue).
          auto issue = [&](auto nk_c) {
            constexpr int NK = decltype(nk_c)::value;
            constexpr int WAIT_IDX = (NK * MMAS_PER_KB * 3) / 4 - 1;
#pragma unroll
            for (int j = 0; j < NK; ++j) {
              const uint64_t adesc = desc_hi | (a_lo + j * (Cfg::kABytes >> 4));
              const uint64_t bdesc = desc_hi | (b_lo + j * (Cfg::kBBytes >> 4));
#pragma unroll
              for (int k = 0; k < MMAS_PER_KB; ++k) {
                // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (addr >> 4) field
                if (j == 0 && k == 0) umma_f16(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
                else                  umma_f16_c<true>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc);
                if (j * MMAS_PER_KB + k == WAIT_IDX && do_wait) {
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
          call_with_nk<KPS>(nk, issue);
          umma_commit_a(empty0 + s * 8);               // frees the smem stage once these MMAs have read it
          if (last_stage) { umma_commit_a(tfull0 + acc * 8); ++tc; }  // accumulator complete
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
    // ------------------------------------------------------------------ epilogue warps
    constexpr int CPW = Cfg::kColsPerWarp;
    griddep_wait();  // this role reads (residual) and overwrites activation slots of earlier kernels
    const int t = threadIdx.x - 64;   // 0 .. kEpiThreads-1
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int col0 = ((warp - 2) >> 2) * CPW;  // first tile column of this warp
    const bool has_res = p.res != nullptr;
    int tc = 0, cached_n0 = -1;
    uint32_t eb = 0, eph = 0;  // epilogue buffer of this tile and its phase (EPI)
    int prev_eb = -1;          // buffer whose TMA store has been issued but not yet waited for
    for (int tile = w_first; tile < w_total; tile += w_step) {
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = m_tile_of(tile) * kBlockM;
      const int tc_this = tc++;
      const int acc = tc_this & 1;
      const uint32_t aph = (tc_this >> 1) & 1;
      if (n0 != cached_n0) {  // uniform across the epilogue threads
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
      if (EPI) {
        // The residual (TMA-prefetched) is read from, and the bf16 result written back to, the same swizzled
        // shared-memory tile; each warp then TMA-stores its 32 rows x CPW columns (full 128-byte lines,
        // clipped at M / N by the tensor map).  Without a residual the buffer only stages the store.
        if (has_res) mbar_wait(&res_full_bar[eb], eph);
        const int lrow = q * 32 + lane;
        const uint32_t buf = smem_u32(epi_base + eb * Cfg::kEpiBufBytes) + (uint32_t)lrow * 128u;
        const uint32_t xr = (uint32_t)(lrow & 7);
        auto chunk = [&](const uint32_t (&v)[32], int c, auto res_c) {
          constexpr bool RES = decltype(res_c)::value;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + c * 32 + g * 8;
            const uint32_t addr = buf + (uint32_t)(col >> 6) * Cfg::kEpiSubBytes + ((((uint32_t)(col & 63) >> 3) ^ xr) << 4);
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              f[j] = fmaf(__uint_as_float(v[g * 8 + j]), s_scale[col + j], s_shift[col + j]);
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
#pragma unroll 1
        for (int c = 0; c < CPW / 32; c += 2) {
          // two TMEM loads in flight per wait
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(taddr + (uint32_t)(c * 32), v0);
          tmem_ld_32x32(taddr + (uint32_t)(c * 32 + 32), v1);
          tmem_ld_wait();
          if (has_res) {
            chunk(v0, c, std::true_type{});
            chunk(v1, c + 1, std::true_type{});
          } else {
            chunk(v0, c, std::false_type{});
            chunk(v1, c + 1, std::false_type{});
          }
        }
        tc_fence_before();
        if (p.pool_tp) {
          // maxpool2 fused: tile rows are (frame = q, pixel = lane); frames (0,1) and (2,3) are max-reduced through the
          // staging tile and only the pooled frame is stored.  Warp pairs (q, q ^ 1) of the same column half meet on a
          // named barrier; the even-frame warp reduces and stores.
          if (CPW == 64) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            const int pair_bar = 2 + (q >> 1) * 2 + (col0 >> 6);
            named_bar_sync(pair_bar, 64);
            {
              // both warps of the pair reduce: lane = pixel, the even-frame warp takes the first 32 columns of the
              // 64-column sub-tile, the odd-frame warp the other 32; results land in the even frame's rows
              const uint32_t even_row = (uint32_t)((q & ~1) * 32 + lane);
              const uint32_t sub = smem_u32(epi_base + eb * Cfg::kEpiBufBytes) + (uint32_t)(col0 >> 6) * Cfg::kEpiSubBytes + even_row * 128u;
              const uint32_t xe = even_row & 7u;
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                const uint32_t g = (uint32_t)((q & 1) * 4 + g4);
                const uint32_t a0 = sub + ((g ^ xe) << 4);
                const uint32_t a1 = a0 + 32u * 128u;  // same pixel, next frame (32 rows further: same swizzle phase)
                uint32_t x[4], y[4];
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(a0));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]) : "r"(a1));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&x[j]), *reinterpret_cast<const __nv_bfloat162*>(&y[j]));
                  x[j] = *reinterpret_cast<const uint32_t*>(&m);
                }
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
              }
              fence_proxy_async_smem();
            }
            named_bar_sync(pair_bar, 64);
            if (!(q & 1) && lane == 0) {
              const int mt = tile / p.n_tiles;
              const int clip = mt / p.tp_tiles_per_clip;
              const int p0 = (mt - clip * p.tp_tiles_per_clip) * 32;
              if (n0 + col0 < p.N)
                tma_store_4d_b(&tmO, smem_u32(epi_base + eb * Cfg::kEpiBufBytes) + (uint32_t)(col0 >> 6) * Cfg::kEpiSubBytes + (uint32_t)q * 32u * 128u,
                               n0 + col0, p0, q >> 1, clip);
              tma_store_commit();
              tma_store_wait_read<0>();
            }
            __syncwarp();
            if (lane == 0 && has_res) mbar_arrive(&res_empty_bar[eb]);
            __syncwarp();
          }
          if (++eb == NB) { eb = 0; eph ^= 1u; }
          continue;
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tmem_empty_bar[acc]);  // accumulator drained: the MMA warp may start tile i+2
#pragma unroll
          for (int j = col0 / 64; j < (col0 + CPW) / 64; ++j)
            if (n0 + 64 * j < p.N) {
              // fused sibling 1x1x1 convs: every 64-column sub-tile belongs to one part (they start on multiples of 64) and goes
              // through that part's own tensor map, which clips the columns beyond the part's width
              const int gc = n0 + 64 * j;
              const CUtensorMap* om = &tmO;
              int oc = gc;
              if (p.split1) {
                if (gc >= p.split2)      { om = &tmO2; oc = gc - p.split2; }
                else if (gc >= p.split1) { om = &tmR;  oc = gc - p.split1; }
              }
              tma_store_2d(om, epi_base + eb * Cfg::kEpiBufBytes + j * Cfg::kEpiSubBytes + q * 32 * 128, oc, m0 + q * 32);
            }
          tma_store_commit();
          // The store engine needs a few hundred cycles to read the tile out of shared memory; waiting for THIS store
          // here would put that on every tile's critical path.  Wait for the PREVIOUS tile's store instead and release
          // its buffer (three buffers: the one this warp writes next was released one tile ago).
          tma_store_wait_read<1>();
          if (has_res && prev_eb >= 0) mbar_arrive(&res_empty_bar[prev_eb]);  // one arrival per epilogue warp frees the buffer
        }
        __syncwarp();
        prev_eb = (int)eb;
        if (++eb == NB) { eb = 0; eph ^= 1u; }
        continue;
      }
      const bool row_ok = row < p.M;
      const __nv_bfloat16* res_row = p.res ? p.res + (long long)row * p.ldr + n0 + col0 : nullptr;
      // one 32-column chunk: scale / shift as 128-bit shared-memory loads (the scalar form was 64 dependent LDS per chunk and the
      // FFMAs behind them were this epilogue's top stall), routed to its destination, 16 bytes per store
      auto chunk = [&](const uint32_t (&v)[32], int c) {
        // destination of this 32-column chunk: the op's own, or one of the sibling outputs of a fused 1x1x1 conv (parts start on
        // multiples of 64 columns, so a chunk never straddles two of them)
        const int gc = n0 + col0 + c * 32;   // first output column of the chunk
        __nv_bfloat16* o_base = p.out;
        int o_ld = p.ldo, o_start = 0, o_lim = p.N;
        if (p.split1) {
          if (gc >= p.split2)      { o_base = p.out2; o_ld = p.ldo2; o_start = p.split2; o_lim = p.split2 + p.seg_w2; }
          else if (gc >= p.split1) { o_base = p.out1; o_ld = p.ldo1; o_start = p.split1; o_lim = p.split1 + p.seg_w1; }
          else                     { o_lim = p.seg_w0; }
        }
        __nv_bfloat16* out_chunk = o_base + (long long)row * o_ld + (gc - o_start);
        if (!row_ok) return;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = c * 32 + g * 8;  // relative to col0
          if (gc + g * 8 < o_lim) {
            const float4 sa = *reinterpret_cast<const float4*>(s_scale + col0 + col), sb = *reinterpret_cast<const float4*>(s_scale + col0 + col + 4);
            const float4 ha = *reinterpret_cast<const float4*>(s_shift + col0 + col), hb = *reinterpret_cast<const float4*>(s_shift + col0 + col + 4);
            float f[8];
            f[0] = fmaf(__uint_as_float(v[g * 8 + 0]), sa.x, ha.x); f[1] = fmaf(__uint_as_float(v[g * 8 + 1]), sa.y, ha.y);
            f[2] = fmaf(__uint_as_float(v[g * 8 + 2]), sa.z, ha.z); f[3] = fmaf(__uint_as_float(v[g * 8 + 3]), sa.w, ha.w);
            f[4] = fmaf(__uint_as_float(v[g * 8 + 4]), sb.x, hb.x); f[5] = fmaf(__uint_as_float(v[g * 8 + 5]), sb.y, hb.y);
            f[6] = fmaf(__uint_as_float(v[g * 8 + 6]), sb.z, hb.z); f[7] = fmaf(__uint_as_float(v[g * 8 + 7]), sb.w, hb.w);
            if (res_row) {
              const uint4 r = *reinterpret_cast<const uint4*>(res_row + col);
              f[0] += bf16_lo(r.x); f[1] += bf16_hi(r.x);
              f[2] += bf16_lo(r.y); f[3] += bf16_hi(r.y);
              f[4] += bf16_lo(r.z); f[5] += bf16_hi(r.z);
              f[6] += bf16_lo(r.w); f[7] += bf16_hi(r.w);
            }
            uint4 o;
            if (p.relu) { o.x = pack_bf16x2_relu(f[0], f[1]); o.y = pack_bf16x2_relu(f[2], f[3]); o.z = pack_bf16x2_relu(f[4], f[5]); o.w = pack_bf16x2_relu(f[6], f[7]); }
            else        { o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]); }
            *reinterpret_cast<uint4*>(out_chunk + g * 8) = o;
          }
        }
      };
      if constexpr (CPW >= 64) {
#pragma unroll 1
        for (int c = 0; c < CPW / 32; c += 2) {   // two TMEM loads in flight per wait
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(taddr + (uint32_t)(c * 32), v0);
          tmem_ld_32x32(taddr + (uint32_t)(c * 32 + 32), v1);
          tmem_ld_wait();
          chunk(v0, c);
          chunk(v1, c + 1);
        }
      } else {
        uint32_t v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        chunk(v, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);  // one arrival per epilogue warp hands the accumulator back
    }
  } else {
    // ------------------------------------------------------------------ gather warps (GATHER only)
    if (GATHER) {
      constexpr int LAG = Cfg::kGatherLag;
      griddep_wait();
      const int t = threadIdx.x - (64 + Cfg::kEpiThreads);  // 0..127: tile row owned by this thread
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

  if (EPI && warp >= 2 && warp < 2 + Cfg::kEpiWarps && lane == 0) tma_store_wait<0>();  // bulk stores fully complete
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vad
