// K5: the MGFN scoring head (reference src/models/mgfn/modeling_mgfn.py, src/loss/*.py) on sm_100a.
//
// Activations are fp32, tokens-major channels-last: X[s, t, c] with s = video * ncrops + crop.  Every Conv1d
// (kernel 1 or 3 over t) is ONE tcgen05 GEMM in kind::tf32 (fp32 operands straight from shared memory, fp32
// accumulate in TMEM): the A operand is a rank-3 TMA view (C, T, S) of the activation, so the k = 3 taps are
// the same box at t - 1, t, t + 1 and TMA's out-of-range zero fill is the conv's zero padding; bias, exact
// GELU and the fp32 residual add live in the epilogue.  Everything that is not a contraction (the two
// LayerNorm flavours, the Glance softmax attention, the Focus depth-wise relation conv, the final
// LayerNorm + fc + sigmoid, top-k magnitude selection, the losses) is a small fp32 SIMT kernel.
#pragma once

#include <math.h>

#include "ptx_sm100.cuh"
#include "conv_umma.cuh"

namespace vad {

// ------------------------------------------------------------------------------------------- GEMM (tf32)
struct HeadGemmParams {
  int S, T;          // sequences, tokens per sequence
  int Tb, Sb;        // tile = Sb sequences x Tb tokens, Tb * Sb == 128
  int t_tiles;       // ceil(T / Tb)
  int N;             // output channels
  int Cin, taps;     // contraction = taps * Cin, tap offsets -(taps/2) .. +(taps/2)
  int gelu;
  int kb_per_split;  // split-K (wgrad: K = all tokens, few output tiles): gridDim.z CTAs take kb_per_split k-blocks each and
                     // red.add their partial tile into `out` (zeroed by the caller; no bias / GELU / residual); 0 = no split
  int ldo, ldr;      // row pitches (elements) of out / residual
  const float* bias;       // [N] or null
  const float* res;        // [S*T, ldr] or null
  float* out;              // [S*T, ldo]
};

constexpr int kHeadBK = 32;      // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kHeadStages = 4;
template <int BN>
struct HeadGemmCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kHeadStages * kStageBytes + BN * 4 + (2 * kHeadStages + 1) * 8 + 16 + 1024;
};

// instruction descriptor: tf32 x tf32 -> fp32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m128(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// grid = (S_tiles * t_tiles, N / BN); warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
template <int BN>
__global__ void __launch_bounds__(192, 1)
head_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HeadGemmParams p) {
  using Cfg = HeadGemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  float* s_bias = reinterpret_cast<float*>(smem + kHeadStages * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* empty_bar = full_bar + kHeadStages;
  uint64_t* done_bar = empty_bar + kHeadStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int st = blockIdx.x / p.t_tiles;
  const int tt = blockIdx.x - st * p.t_tiles;
  const int s0 = st * p.Sb, t0 = tt * p.Tb;
  const int kb_per_tap = p.Cin / kHeadBK;
  const int total_kb = p.taps * kb_per_tap;
  const int kb_first = p.kb_per_split ? (int)blockIdx.z * p.kb_per_split : 0;
  const int num_kb = p.kb_per_split ? (total_kb - kb_first < p.kb_per_split ? total_kb - kb_first : p.kb_per_split) : total_kb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kHeadStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) s_bias[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      uint32_t s = 0, ph = 0;
      int tap = kb_first / kb_per_tap, c0 = (kb_first - tap * kb_per_tap) * kHeadBK;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_a(empty0 + s * 8, ph ^ 1u);
        const uint32_t a_dst = stage0 + s * Cfg::kStageBytes;
        const uint32_t fb = full0 + s * 8;
        mbar_arrive_expect_tx_a(fb, (uint32_t)Cfg::kStageBytes);
        tma_load_3d_a(a_dst, &tmA, fb, c0, t0 + tap - p.taps / 2, s0);
        tma_load_2d_a(a_dst + Cfg::kABytes, &tmB, fb, (kb_first + kb) * kHeadBK, n0);
        c0 += kHeadBK;
        if (c0 >= p.Cin) { c0 = 0; ++tap; }
        if (++s == kHeadStages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_tf32_m128(BN);
      const uint64_t desc_hi = umma_desc_kmajor<128>(0);
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), stage0 = smem_u32(stage_base);
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_a(full0 + s * 8, ph);
        tc_fence_after();
        const uint32_t a_lo = (stage0 + s * Cfg::kStageBytes) >> 4;
        const uint64_t adesc = desc_hi | a_lo;
        const uint64_t bdesc = desc_hi | (a_lo + (Cfg::kABytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 8 tf32 = 32 B per MMA: +2 in the (addr >> 4) field
          umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit_a(empty0 + s * 8);
        if (++s == kHeadStages) { s = 0; ph ^= 1u; }
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int s = s0 + r / p.Tb, t = t0 + r % p.Tb;
    const bool ok = s < p.S && t < p.T;
    const long long m = (long long)s * p.T + t;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      if (ok) {
        float* o = p.out + m * p.ldo + n0 + c * 32;
        const float* rr = p.res ? p.res + m * p.ldr + n0 + c * 32 : nullptr;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (n0 + c * 32 + g * 4 < p.N) {
            float4 f;
            f.x = __uint_as_float(v[g * 4 + 0]) + s_bias[c * 32 + g * 4 + 0];
            f.y = __uint_as_float(v[g * 4 + 1]) + s_bias[c * 32 + g * 4 + 1];
            f.z = __uint_as_float(v[g * 4 + 2]) + s_bias[c * 32 + g * 4 + 2];
            f.w = __uint_as_float(v[g * 4 + 3]) + s_bias[c * 32 + g * 4 + 3];
            if (p.kb_per_split) {  // partial tile of a split-K launch: accumulate (fp32 reductions in L2)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + g * 4), "f"(__uint_as_float(v[g * 4 + 0])),
                           "f"(__uint_as_float(v[g * 4 + 1])), "f"(__uint_as_float(v[g * 4 + 2])), "f"(__uint_as_float(v[g * 4 + 3]))
                           : "memory");
              continue;
            }
            if (p.gelu) { f.x = gelu_erf(f.x); f.y = gelu_erf(f.y); f.z = gelu_erf(f.z); f.w = gelu_erf(f.w); }
            if (rr) {
              const float4 a = *reinterpret_cast<const float4*>(rr + g * 4);
              f.x += a.x; f.y += a.y; f.z += a.z; f.w += a.w;
            }
            *reinterpret_cast<float4*>(o + g * 4) = f;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

// ------------------------------------------------------------------------------------------- SIMT pieces
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// video [n_seq, T, C + 1] (features + appended magnitude, src/dataset.py:121-124) -> feat [n_seq*T, C], mag [n_seq*T]
// (the 2049-float row pitch is not 16-byte aligned, so TMA cannot read the features in place)
__global__ void head_split_kernel(const float* __restrict__ video, long long ntok, int C, float* __restrict__ feat,
                                  float* __restrict__ mag) {
  const long long total = ntok * (C / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / (C / 4);
    const int c = (int)(i - tok * (C / 4)) * 4;
    const float* src = video + tok * (C + 1) + c;
    *reinterpret_cast<float4*>(feat + tok * C + c) = make_float4(src[0], src[1], src[2], src[3]);
    if (c == 0) mag[tok] = video[tok * (C + 1) + C];
  }
}

// MGFNFeatureAmplifier (modeling_mgfn.py:66-94): x += mag_ratio * Conv1d(1 -> C, k = 3, pad 1)(magnitude)
__global__ void head_amplify_kernel(float* __restrict__ x, const float* __restrict__ mag, const float* __restrict__ w,
                                    const float* __restrict__ b, float ratio, int S, int T, int C) {
  const long long total = (long long)S * T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long tok = i / C;
    const int t = (int)(tok % T);
    float m = b[c];
    if (t > 0) m = fmaf(w[c * 3 + 0], mag[tok - 1], m);
    m = fmaf(w[c * 3 + 1], mag[tok], m);
    if (t + 1 < T) m = fmaf(w[c * 3 + 2], mag[tok + 1], m);
    x[i] = x[i] + ratio * m;
  }
}

// MGFNLayerNorm (modeling_mgfn.py:36-47): over channels, population variance, divides by (std + eps) -- NOT
// sqrt(var + eps).  One warp per token.
__global__ void head_mgfn_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                           float eps, long long ntok, int C, float* __restrict__ y) {
  const long long tok = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= ntok) return;
  const float* row = x + tok * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += row[c];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = row[c] - mean; v = fmaf(d, d, v); }
  const float inv = 1.f / (sqrtf(warp_sum(v) / (float)C) + eps);
  float* out = y + tok * C;
  for (int c = lane; c < C; c += 32) out[c] = (row[c] - mean) * inv * g[c] + b[c];
}

// GlanceAttention core (modeling_mgfn.py:107-127): per (sequence, head) softmax(q * scale . k^T) . v with
// dim_head = 64.  qkv [ntok, 3 * heads * 64] (q | k | v, channel = head * 64 + d).  One thread per query row,
// keys / values staged 32 at a time in shared memory, online softmax.
__global__ void __launch_bounds__(128) head_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int S, int T,
                                                             int heads, float scale) {
  __shared__ float sk[32][64];
  __shared__ float sv[32][64];
  const int s = blockIdx.z, h = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int inner = heads * 64;
  const long long base = (long long)s * T;
  float q[64], acc[64];
  const bool live = i < T;
  if (live) {
    const float* qp = qkv + (base + i) * 3 * inner + h * 64;
#pragma unroll
    for (int d = 0; d < 64; ++d) { q[d] = qp[d] * scale; acc[d] = 0.f; }
  } else {
#pragma unroll
    for (int d = 0; d < 64; ++d) { q[d] = 0.f; acc[d] = 0.f; }
  }
  float mx = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < T; j0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 64; e += 128) {
      const int j = e >> 6, d = e & 63;
      const bool in = j0 + j < T;
      const float* kp = qkv + (base + j0 + j) * 3 * inner + inner + h * 64 + d;
      sk[j][d] = in ? kp[0] : 0.f;
      sv[j][d] = in ? kp[inner] : 0.f;
    }
    __syncthreads();
    const int nj = T - j0 < 32 ? T - j0 : 32;
    for (int j = 0; j < nj; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int d = 0; d < 64; ++d) sdot = fmaf(q[d], sk[j][d], sdot);
      const float nm = fmaxf(mx, sdot);
      const float corr = __expf(mx - nm);
      const float pj = __expf(sdot - nm);
      l = l * corr + pj;
#pragma unroll
      for (int d = 0; d < 64; ++d) acc[d] = fmaf(acc[d], corr, pj * sv[j][d]);
      mx = nm;
    }
  }
  if (live) {
    float* op = out + (base + i) * inner + h * 64;
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 64; ++d) op[d] = acc[d] * inv;
  }
}

// FocusAttention.rel_pos (modeling_mgfn.py:166-186): depth-wise Conv1d(k, padding k / 2) over t whose weights are
// shared by the channels of one head; the rearrange "b (c h) t -> (b c) h t" makes head = channel % heads.
__global__ void head_relpos_kernel(const float* __restrict__ v, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ out, int S, int T, int C, int heads, int k) {
  const long long total = (long long)S * T * C;
  const int half = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long tok = i / C;
    const int t = (int)(tok % T);
    const int h = c % heads;
    float a = b[h];
    for (int j = 0; j < k; ++j) {
      const int tj = t + j - half;
      if (tj >= 0 && tj < T) a = fmaf(w[h * k + j], v[(tok + j - half) * C + c], a);
    }
    out[i] = a;
  }
}

// nn.LayerNorm(C) + Linear(C, 1) + sigmoid (modeling_mgfn.py:404-409) and the L2 magnitude of the normalised
// feature (modeling_mgfn.py:313-314).  One warp per token.
__global__ void head_final_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                  const float* __restrict__ fcw, const float* __restrict__ fcb, float eps, long long ntok, int C,
                                  float* __restrict__ xln, float* __restrict__ score, float* __restrict__ fmag) {
  const long long tok = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tok >= ntok) return;
  const float* row = x + tok * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += row[c];
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = row[c] - mean; v = fmaf(d, d, v); }
  const float inv = rsqrtf(warp_sum(v) / (float)C + eps);
  float dot = 0.f, sq = 0.f;
  float* out = xln + tok * C;
  for (int c = lane; c < C; c += 32) {
    const float y = (row[c] - mean) * inv * g[c] + b[c];
    out[c] = y;
    dot = fmaf(y, fcw[c], dot);
    sq = fmaf(y, y, sq);
  }
  dot = warp_sum(dot);
  sq = warp_sum(sq);
  if (lane == 0) {
    score[tok] = 1.f / (1.f + __expf(-(dot + fcb[0])));
    fmag[tok] = sqrtf(sq);
  }
}

// magnitude_selection_and_score_prediction, eval mode (modeling_mgfn.py:302-374): crop-mean magnitudes and
// scores per video, top-k magnitudes over t (ties: lowest index first), mean score of the selected snippets,
// and the selected features gathered crop-major ([ncrops, n_videos, k, C]) for the n_videos videos starting at
// video_off of the batch.  One block per video.
__global__ void __launch_bounds__(256) head_select_kernel(const float* __restrict__ score_tok, const float* __restrict__ fmag_tok,
                                                          const float* __restrict__ xln, int n_videos, int ncrops, int T, int C,
                                                          int k, float* __restrict__ scores, float* __restrict__ vid_score,
                                                          int* __restrict__ idx_out, float* __restrict__ sel, int video_off) {
  extern __shared__ float sh[];  // [T] magnitudes
  __shared__ int s_idx[8];
  const int bl = blockIdx.x;          // index among the selected videos (n_videos of them)
  const int b = video_off + bl;       // index in the whole batch
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float m = 0.f, sc = 0.f;
    for (int c = 0; c < ncrops; ++c) {
      const long long tok = ((long long)b * ncrops + c) * T + t;
      m += fmag_tok[tok];
      sc += score_tok[tok];
    }
    sh[t] = m / (float)ncrops;
    scores[(long long)b * T + t] = sc / (float)ncrops;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      int best = -1;
      float bv = -INFINITY;
      for (int t = 0; t < T; ++t) {
        bool used = false;
        for (int u = 0; u < j; ++u) used |= (s_idx[u] == t);
        // torch.topk semantics: NaN ranks above everything, ties keep the lower index; `best < 0` accepts the first
        // unused candidate whatever its value (all -inf / NaN rows), so an index in [0, T) is always selected
        if (!used && (best < 0 || sh[t] > bv || (sh[t] != sh[t] && bv == bv))) { bv = sh[t]; best = t; }
      }
      s_idx[j] = best;
      idx_out[b * k + j] = best;
      acc += scores[(long long)b * T + best];
    }
    vid_score[b] = acc / (float)k;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ncrops * k * C; e += blockDim.x) {
    const int c = e % C;
    const int j = (e / C) % k;
    const int crop = e / (C * k);
    sel[(((long long)crop * n_videos + bl) * k + j) * C + c] = xln[(((long long)b * ncrops + crop) * T + s_idx[j]) * C + c];
  }
}

// Losses (src/loss/base.py:7-48, src/loss/mgfn.py:7-47, modeling_mgfn.py:411-424).  Videos [0, nn) are the normal
// half, [nn, 2 nn) the abnormal half; sel_n / sel_a are the gathered features [ncrops * nn, k, C] (crop-major).
// out[0] = total, [1] = smoothness, [2] = sparsity, [3] = BCE, [4] = contrastive(a, n), [5] = con_n, [6] = con_a.
__global__ void __launch_bounds__(256) head_loss_kernel(const float* __restrict__ scores, const float* __restrict__ vid_score,
                                                        const float* __restrict__ labels, const float* __restrict__ sel_n,
                                                        const float* __restrict__ sel_a, int nn, int ncrops, int T, int C, int k,
                                                        float* __restrict__ l1, float* __restrict__ out, float w_smooth = 8e-4f,
                                                        float w_sparse = 8e-3f, float alpha = 0.001f, float margin = 200.f) {
  __shared__ float red[256];
  auto block_sum = [&](float v) -> float {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
  };
  const int bs = 2 * nn;
  // L1 norms of the selected features: l1[0 .. R*k) normal, l1[R*k .. 2*R*k) abnormal, R = ncrops * nn
  const int R = ncrops * nn;
  for (int e = threadIdx.x; e < 2 * R * k; e += blockDim.x) {
    const float* src = (e < R * k ? sel_n + (long long)e * C : sel_a + (long long)(e - R * k) * C);
    float a = 0.f;
    for (int c = 0; c < C; ++c) a += fabsf(src[c]);
    l1[e] = a;
  }
  __syncthreads();
  float v = 0.f;
  for (int e = threadIdx.x; e < bs * (T - 1); e += blockDim.x) {
    const int b = e / (T - 1), t = e - b * (T - 1);
    const float d = scores[b * T + t + 1] - scores[b * T + t];
    v = fmaf(d, d, v);
  }
  const float smooth = w_smooth * block_sum(v);
  v = 0.f;
  for (int e = threadIdx.x; e < nn * T; e += blockDim.x) v = fmaf(scores[e], scores[e], v);
  const float sparsity = w_sparse * sqrtf(block_sum(v));
  v = 0.f;
  for (int e = threadIdx.x; e < bs; e += blockDim.x) {
    const float pr = vid_score[e], y = labels[e];
    // nn.BCELoss clamps the logs at -100
    v -= y * fmaxf(logf(pr), -100.f) + (1.f - y) * fmaxf(logf(1.f - pr), -100.f);
  }
  const float bce = block_sum(v) / (float)bs;
  auto dist = [&](const float* a, const float* b) {  // torch.pairwise_distance: || a - b + 1e-6 ||_2 over k
    float s = 0.f;
    for (int j = 0; j < k; ++j) { const float d = a[j] - b[j] + 1e-6f; s = fmaf(d, d, s); }
    return sqrtf(s);
  };
  const float* ln = l1;
  const float* la = l1 + R * k;
  v = 0.f;
  for (int e = threadIdx.x; e < R; e += blockDim.x) {
    const float d = fmaxf(margin - dist(la + e * k, ln + e * k), 0.f);
    v = fmaf(d, d, v);
  }
  const float con = block_sum(v) / (float)R;
  const int sep = R / 2;  // int(len(n_feat_magnitude) / 2)
  const int rest = R - sep;
  v = 0.f;
  for (int e = threadIdx.x; e < (sep < rest ? sep : rest); e += blockDim.x) { const float d = dist(ln + (sep + e) * k, ln + e * k); v = fmaf(d, d, v); }
  const float con_n = block_sum(v) / (float)(sep < rest ? sep : rest);
  v = 0.f;
  for (int e = threadIdx.x; e < (sep < rest ? sep : rest); e += blockDim.x) { const float d = dist(la + (sep + e) * k, la + e * k); v = fmaf(d, d, v); }
  const float con_a = block_sum(v) / (float)(sep < rest ? sep : rest);
  if (threadIdx.x == 0) {
    const float mg = bce + alpha * (alpha * con + con_a + con_n);
    out[0] = mg + smooth + sparsity;
    out[1] = smooth; out[2] = sparsity; out[3] = bce; out[4] = con; out[5] = con_n; out[6] = con_a;
  }
}

}  // namespace vad
