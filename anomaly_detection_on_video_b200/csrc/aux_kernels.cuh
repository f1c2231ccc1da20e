// HBM-bound helper kernels of the extraction path (everything that is not a contraction):
//   K0  ingest        fp32 NCTHW clip -> bf16 stem layout            (extract_features.py:83-86)
//   K1  preprocess    u8 frames -> PIL-exact resize -> crops -> standardise -> clip tensors
//                                                                    (src/gtransforms.py:9-73,115-132)
//   K3  max-pool 3D   channels-last, optional SAME padding / channel-slice destination
//                                                                    (src/i3d.py:212-217,306,309)
//   K4  avg-pool      global mean over (T,H,W) -> fp32 features      (src/i3d.py:244,314)
//       segment mean  (n_clips, crops, C) -> (crops, 32, C)          (extract_features.py:159-185)
//       add magnitude append the L2 norm                             (src/dataset.py:121-124)
#pragma once

#include "ptx_sm100.cuh"

namespace vad {

// ------------------------------------------------------------------------------------------- K0
// one thread per stem-layout pixel (8 bytes out); reads are coalesced along W in each channel plane
__global__ void ingest_ncthw_f32_kernel(const float* __restrict__ x, int B, int T, int H, int W, int pad_left,
                                        uint2* __restrict__ out) {
  const int Wp = W + 8;
  const long long total = (long long)B * T * H * Wp;
  const long long plane = (long long)T * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    const long long r = i / Wp;  // (b*T + t)*H + h
    const int w = wp - pad_left;
    uint2 o = make_uint2(0u, 0u);
    if (w >= 0 && w < W) {
      const long long b = r / ((long long)T * H);
      const long long th = r - b * (long long)T * H;  // t*H + h
      const float* px = x + b * 3 * plane + th * W + w;
      o.x = pack_bf16x2(px[0], px[plane]);
      o.y = pack_bf16x2(px[2 * plane], 0.f);
    }
    out[i] = o;
  }
}

// ------------------------------------------------------------------------------------------- K1
struct PreprocParams {
  const uint8_t* frames;  // [n_frames, H, W, 3]
  int n_frames, H, W;
  int rh, rw;             // resized size
  int ksize_h, ksize_v;
  const int* bounds_h;    // [rw][2] (xmin, count)
  const int* coef_h;      // [rw][ksize_h]  22-bit fixed point
  const int* bounds_v;    // [rh][2]
  const int* coef_v;      // [rh][ksize_v]
  int crop, ncrops;
  int tops[10], lefts[10], flips[10];
  int clip_start, fpc;
  int out_mode, pad_left;
  int rows_per_block;     // resized rows one block produces
  int max_src_rows;       // source rows the vertical filter of rows_per_block consecutive rows can touch
  int off_px, off_h, off_src;  // shared-memory byte offsets (16-byte aligned) of s_px / s_h / s_src
  void* out;
};

__device__ __forceinline__ int clip8_q22(int acc) {
  const int v = acc >> 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// grid = (ceil(resized rows / R), frame slots); one block resamples R consecutive output rows of one frame
// (Pillow's order: horizontal pass rounded to u8, then the vertical pass) and writes them into every crop that
// contains them.
//   phase 0  the source rows the R vertical filters need are staged in shared memory with coalesced 128-bit loads
//            (the 3-byte pixels make per-tap global loads byte-sized otherwise)
//   phase 1a horizontal pass of every staged source row, once (consecutive output rows share source rows)
//   phase 1b vertical pass -> u8 resized rows; for the stem layout also the standardised bf16 pixel (r, g, b, 0)
//            through the 256-entry LUT, so the ten crops below are plain 8-byte copies
//   phase 2  crops: two crops at a time, 128 threads each, one 16-byte store (two pixels) per work item
__global__ void __launch_bounds__(256) preprocess_kernel(const PreprocParams p) {
  extern __shared__ __align__(16) uint8_t s_mem[];
  const int R = p.rows_per_block;
  const int row_bytes = (p.rw * 3 + 15) & ~15;
  const int src_row_bytes = p.W * 3;
  const int src_pitch = (src_row_bytes + 15) & ~15;
  float* lut_f = reinterpret_cast<float*>(s_mem);                          // 256 fp32
  __nv_bfloat16* lut_h = reinterpret_cast<__nv_bfloat16*>(lut_f + 256);    // 256 bf16
  uint8_t* s_row = reinterpret_cast<uint8_t*>(lut_h + 256);                // R resized rows, u8, pitch row_bytes
  uint2* s_px = reinterpret_cast<uint2*>(s_mem + p.off_px);                // R x rw standardised bf16 pixels (stem mode)
  uint8_t* s_h = s_mem + p.off_h;                                          // horizontally resampled source rows
  uint8_t* s_src = s_mem + p.off_src;                                      // staged source rows, pitch src_pitch

  const int y0 = blockIdx.x * R;
  const int ny = (p.rh - y0) < R ? (p.rh - y0) : R;
  const int slot = blockIdx.y;
  const int clip_local = slot / p.fpc;
  const int t = slot - clip_local * p.fpc;
  const int clip = p.clip_start + clip_local;
  int L = p.n_frames - clip * p.fpc;      // real frames in this clip (LoopPad, gtransforms.py:115-132)
  L = L > p.fpc ? p.fpc : L;
  const int src_frame = clip * p.fpc + (t % L);
  const uint8_t* frame = p.frames + (long long)src_frame * p.H * p.W * 3;

  {
    // GroupStandardizationTenCrop: t.sub_(114.75).div_(57.375), two fp32 roundings
    const float v = __fdiv_rn(__fsub_rn((float)threadIdx.x, 114.75f), 57.375f);
    lut_f[threadIdx.x] = v;
    lut_h[threadIdx.x] = __float2bfloat16_rn(v);
  }

  const int ymin0 = p.bounds_v[2 * y0];
  const int nsrc = p.bounds_v[2 * (y0 + ny - 1)] + p.bounds_v[2 * (y0 + ny - 1) + 1] - ymin0;
  {
    const uint8_t* g0 = frame + (long long)ymin0 * src_row_bytes;
    const bool vec = ((src_row_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(g0) & 15) == 0);
    if (vec) {
      const int per_row = src_row_bytes >> 4;
      for (int i = threadIdx.x; i < nsrc * per_row; i += blockDim.x) {
        const int r = i / per_row, c = i - r * per_row;
        reinterpret_cast<uint4*>(s_src + r * src_pitch)[c] = reinterpret_cast<const uint4*>(g0 + (long long)r * src_row_bytes)[c];
      }
    } else {
      for (int i = threadIdx.x; i < nsrc * src_row_bytes; i += blockDim.x) {
        const int r = i / src_row_bytes, c = i - r * src_row_bytes;
        s_src[r * src_pitch + c] = g0[(long long)r * src_row_bytes + c];
      }
    }
  }
  __syncthreads();

  // phase 1a: one work item = one pixel (3 channels) of one staged source row
  for (int i = threadIdx.x; i < nsrc * p.rw; i += blockDim.x) {
    const int r = i / p.rw;
    const int x = i - r * p.rw;
    const int xmin = p.bounds_h[2 * x];
    const int xcnt = p.bounds_h[2 * x + 1];
    const int* kh = p.coef_h + x * p.ksize_h;
    const uint8_t* src = s_src + r * src_pitch + xmin * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int j = 0; j < xcnt; ++j) {
      const int k = kh[j];
      a0 += (int)src[3 * j] * k;
      a1 += (int)src[3 * j + 1] * k;
      a2 += (int)src[3 * j + 2] * k;
    }
    uint8_t* dst = s_h + r * row_bytes + x * 3;
    dst[0] = (uint8_t)clip8_q22(a0);
    dst[1] = (uint8_t)clip8_q22(a1);
    dst[2] = (uint8_t)clip8_q22(a2);
  }
  __syncthreads();

  // phase 1b: one work item = one pixel of one output row
  const bool stem = p.out_mode == 1;
  for (int i = threadIdx.x; i < ny * p.rw; i += blockDim.x) {
    const int yy = i / p.rw;
    const int x = i - yy * p.rw;
    const int y = y0 + yy;
    const int r0 = p.bounds_v[2 * y] - ymin0;
    const int ycnt = p.bounds_v[2 * y + 1];
    const int* kv = p.coef_v + y * p.ksize_v;
    const uint8_t* src = s_h + r0 * row_bytes + x * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    for (int r = 0; r < ycnt; ++r) {
      const int k = kv[r];
      a0 += (int)src[r * row_bytes] * k;
      a1 += (int)src[r * row_bytes + 1] * k;
      a2 += (int)src[r * row_bytes + 2] * k;
    }
    const int v0 = clip8_q22(a0), v1 = clip8_q22(a1), v2 = clip8_q22(a2);
    if (stem) {
      const uint32_t cr = __bfloat16_as_ushort(lut_h[v0]);
      const uint32_t cg = __bfloat16_as_ushort(lut_h[v1]);
      const uint32_t cb = __bfloat16_as_ushort(lut_h[v2]);
      s_px[yy * p.rw + x] = make_uint2(cr | (cg << 16), cb);
    } else {
      uint8_t* dst = s_row + yy * row_bytes + x * 3;
      dst[0] = (uint8_t)v0; dst[1] = (uint8_t)v1; dst[2] = (uint8_t)v2;
    }
  }
  __syncthreads();

  const int crop = p.crop;
  if (stem) {
    // stem layout [clipcrop, t, crop, crop + 8, 4] bf16 ; one work item = two pixels (16 B)
    const int Wp = crop + 8;
    const int per_crop = Wp / 2;
    const int half = threadIdx.x >> 7;   // two crops in flight
    const int i0 = threadIdx.x & 127;
    for (int yy = 0; yy < ny; ++yy) {
      const int y = y0 + yy;
      const uint2* row = s_px + yy * p.rw;
      for (int k = half; k < p.ncrops; k += 2) {
        const int yo = y - p.tops[k];
        if (yo < 0 || yo >= crop) continue;
        const int left = p.lefts[k];
        const int flip = p.flips[k];
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                              ((((long long)clip_local * p.ncrops + k) * p.fpc + t) * crop + yo) * Wp * 4);
        for (int i = i0; i < per_crop; i += 128) {
          const int xa = 2 * i - p.pad_left, xb = xa + 1;
          uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
          if (xa >= 0 && xa < crop) a = row[flip ? left + crop - 1 - xa : left + xa];
          if (xb >= 0 && xb < crop) b = row[flip ? left + crop - 1 - xb : left + xb];
          dst[i] = make_uint4(a.x, a.y, b.x, b.y);
        }
      }
    }
  } else {
    // dataset layout [clip, crop_idx, t, 3, crop, crop] fp32
    for (int yy = 0; yy < ny; ++yy) {
      const int y = y0 + yy;
      const uint8_t* row = s_row + yy * row_bytes;
      for (int k = 0; k < p.ncrops; ++k) {
        const int yo = y - p.tops[k];
        if (yo < 0 || yo >= crop) continue;
        const int left = p.lefts[k];
        const int flip = p.flips[k];
        float* dst = reinterpret_cast<float*>(p.out) +
                     ((((long long)clip_local * p.ncrops + k) * p.fpc + t) * 3) * crop * crop + (long long)yo * crop;
        for (int i = threadIdx.x; i < 3 * crop; i += blockDim.x) {
          const int c = i / crop;
          const int xo = i - c * crop;
          const int xs = flip ? left + crop - 1 - xo : left + xo;
          dst[(long long)c * crop * crop + xo] = lut_f[row[xs * 3 + c]];
        }
      }
    }
  }
}

// K1, stem layout only, resampling kernels of at most KH x KV taps (Pillow's bilinear filter has 3 taps when it
// enlarges -- the 240 x 320 UCF-Crime frames --, 5 up to a 2 x reduction).  Same arithmetic as preprocess_kernel, far
// fewer instructions (that kernel is issue bound at 0.38 of the HBM write rate: 5.6 k instructions per thread):
//   * one thread owns one resized COLUMN: its horizontal taps and coefficients stay in registers for all source rows
//     (no index division, no per-pixel table loads);
//   * the horizontal result never goes to shared memory: the thread walks the staged source rows once and keeps the
//     last KV horizontally resampled pixels in a register window.  Row y's vertical taps are source rows
//     [ymin, ymin + ycnt) and ymin + ycnt never decreases with y, so when the window has just absorbed row
//     ymin + ycnt - 1 the taps are its LAST ycnt entries: the coefficients are stored right-aligned (zeros in front)
//     and every output is a fixed KV-term sum over statically indexed registers;
//   * crops: groups of per_crop threads (one 16-byte store = two pixels each) take (row, crop) pairs; the thread's
//     pad / range predicates and its mirrored source index are loop invariants.
// blockDim.x = the resized width rounded up to a warp (352 for 341 columns).
template <int KH, int KV, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 384 ? 4 : 2) preprocess_stem_kernel(const PreprocParams p) {
  extern __shared__ __align__(16) uint8_t s_mem[];
  const int R = p.rows_per_block;
  const int src_row_bytes = p.W * 3;
  const int src_pitch = (src_row_bytes + 15) & ~15;
  unsigned short* lut_h = reinterpret_cast<unsigned short*>(s_mem);          // 256 bf16 bit patterns
  int* s_kv = reinterpret_cast<int*>(s_mem + 512);                           // [R][KV] right-aligned vertical coefficients
  int* s_rend = s_kv + R * KV;                                               // [R] one past the last staged source row of output row yy
  uint2* s_px = reinterpret_cast<uint2*>(s_mem + p.off_px);                  // R x rw standardised bf16 pixels (r, g, b, 0)
  uint8_t* s_src = s_mem + p.off_src;                                        // staged source rows, pitch src_pitch (+ tap slack)

  const int y0 = blockIdx.x * R;
  const int ny = (p.rh - y0) < R ? (p.rh - y0) : R;
  const int slot = blockIdx.y;
  const int clip_local = slot / p.fpc;
  const int t = slot - clip_local * p.fpc;
  const int clip = p.clip_start + clip_local;
  int L = p.n_frames - clip * p.fpc;      // real frames in this clip (LoopPad, gtransforms.py:115-132)
  L = L > p.fpc ? p.fpc : L;
  const int src_frame = clip * p.fpc + (t % L);
  const uint8_t* frame = p.frames + (long long)src_frame * p.H * p.W * 3;
  const int tid = threadIdx.x, nthr = blockDim.x;

  for (int i = tid; i < 256; i += nthr) {   // (narrow resized frames run fewer than 256 threads)
    // GroupStandardizationTenCrop: t.sub_(114.75).div_(57.375), two fp32 roundings, then the stem's bf16
    const float v = __fdiv_rn(__fsub_rn((float)i, 114.75f), 57.375f);
    lut_h[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  }
  const int ymin0 = p.bounds_v[2 * y0];
  if (tid < ny) {
    const int y = y0 + tid;
    const int ymin = p.bounds_v[2 * y], ycnt = p.bounds_v[2 * y + 1];
    s_rend[tid] = ymin + ycnt - ymin0;
#pragma unroll
    for (int q = 0; q < KV; ++q) {
      const int j = q - (KV - ycnt);
      s_kv[tid * KV + q] = j >= 0 ? p.coef_v[y * p.ksize_v + j] : 0;
    }
  }
  const int nsrc = p.bounds_v[2 * (y0 + ny - 1)] + p.bounds_v[2 * (y0 + ny - 1) + 1] - ymin0;
  {
    const uint8_t* g0 = frame + (long long)ymin0 * src_row_bytes;
    const bool vec = ((src_row_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(g0) & 15) == 0);
    const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    if (vec) {
      const int per_row = src_row_bytes >> 4;
      for (int r = warp; r < nsrc; r += nwarp)
        for (int c = lane; c < per_row; c += 32)
          reinterpret_cast<uint4*>(s_src + r * src_pitch)[c] = __ldg(reinterpret_cast<const uint4*>(g0 + (long long)r * src_row_bytes) + c);
    } else {
      for (int r = warp; r < nsrc; r += nwarp)
        for (int c = lane; c < src_row_bytes; c += 32) s_src[r * src_pitch + c] = g0[(long long)r * src_row_bytes + c];
    }
  }
  __syncthreads();

  for (int x = tid; x < p.rw; x += nthr) {
    const int xmin = p.bounds_h[2 * x];
    int kh[KH];
#pragma unroll
    for (int j = 0; j < KH; ++j) kh[j] = j < p.ksize_h ? p.coef_h[x * p.ksize_h + j] : 0;   // zero beyond the tap count
    const uint8_t* col = s_src + xmin * 3;
    int win[KV][3];
#pragma unroll
    for (int q = 0; q < KV; ++q) win[q][0] = win[q][1] = win[q][2] = 0;
    int next_r = 0;
    for (int yy = 0; yy < ny; ++yy) {
      const int rend = s_rend[yy];
      while (next_r < rend) {
#pragma unroll
        for (int q = 0; q + 1 < KV; ++q) { win[q][0] = win[q + 1][0]; win[q][1] = win[q + 1][1]; win[q][2] = win[q + 1][2]; }
        const uint8_t* src = col + next_r * src_pitch;
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
        for (int j = 0; j < KH; ++j) {
          a0 += (int)src[3 * j] * kh[j];
          a1 += (int)src[3 * j + 1] * kh[j];
          a2 += (int)src[3 * j + 2] * kh[j];
        }
        win[KV - 1][0] = clip8_q22(a0); win[KV - 1][1] = clip8_q22(a1); win[KV - 1][2] = clip8_q22(a2);
        ++next_r;
      }
      int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
      for (int q = 0; q < KV; ++q) {
        const int k = s_kv[yy * KV + q];
        a0 += win[q][0] * k; a1 += win[q][1] * k; a2 += win[q][2] * k;
      }
      const uint32_t cr = lut_h[clip8_q22(a0)], cg = lut_h[clip8_q22(a1)], cb = lut_h[clip8_q22(a2)];
      s_px[yy * p.rw + x] = make_uint2(cr | (cg << 16), cb);
    }
  }
  __syncthreads();

  // stem layout [clipcrop, t, crop, crop + 8, 4] bf16; one work item = two pixels (16 B).  (row, crop) pairs are walked
  // row-major in steps of `groups` without a division; everything that depends only on the thread is hoisted.
  const int crop = p.crop;
  const int per_crop = (crop + 8) >> 1;
  const int groups = nthr / per_crop;
  const int g = tid / per_crop;
  if (g >= groups) return;
  const int i = tid - g * per_crop;
  const int xa = 2 * i - p.pad_left, xb = xa + 1;
  const bool pa = xa >= 0 && xa < crop, pb = xb >= 0 && xb < crop;
  const int fa = crop - 1 - xa, fb = crop - 1 - xb;                 // source columns in a mirrored crop
  const int ncrops = p.ncrops;
  const unsigned k_stride = (unsigned)(p.fpc * crop * per_crop);    // uint4 between two crops of a clip
  uint4* const outp = reinterpret_cast<uint4*>(p.out) + ((long long)clip_local * ncrops * p.fpc + t) * (long long)(crop * per_crop) + i;
  int yy = 0, k = g;
  for (;;) {
    while (k >= ncrops) { k -= ncrops; ++yy; }
    if (yy >= ny) break;
    const int yo = y0 + yy - p.tops[k];
    if ((unsigned)yo < (unsigned)crop) {
      const uint2* row = s_px + yy * p.rw + p.lefts[k];
      const bool flip = p.flips[k] != 0;
      uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
      if (pa) a = row[flip ? fa : xa];
      if (pb) b = row[flip ? fb : xb];
      outp[(unsigned)k * k_stride + (unsigned)(yo * per_crop)] = make_uint4(a.x, a.y, b.x, b.y);
    }
    k += groups;
  }
}

// ------------------------------------------------------------------------------------------- K3
struct PoolParams {
  const __nv_bfloat16* in;
  __nv_bfloat16* out;  // already offset by dst_c_off
  int B, Ti, Hi, Wi, C;
  int To, Ho, Wo;
  int kt, kh, kw, st, sh, sw;
  int pt, ph, pw;   // front padding
  int pad_zero;     // 1: out-of-range taps contribute 0 (SAME-padding port), 0: they are ignored
  int ldo;          // dst row pitch (elements)
};

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// one thread per (output pixel, 8-channel vector): 128-bit loads/stores, channel vectors of a pixel
// are adjacent threads so every tap is a contiguous row segment
__global__ void __launch_bounds__(256) maxpool3d_kernel(const PoolParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.B * p.To * p.Ho * p.Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const long long m_out = m;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const int to = (int)(m % p.To); m /= p.To;
    const long long b = m;
    const uint32_t neg_inf2 = 0xFF80FF80u;
    uint4 acc = make_uint4(neg_inf2, neg_inf2, neg_inf2, neg_inf2);
    bool any_oob = false;
    for (int dt = 0; dt < p.kt; ++dt) {
      const int ti = to * p.st - p.pt + dt;
      for (int dh = 0; dh < p.kh; ++dh) {
        const int hi = ho * p.sh - p.ph + dh;
        for (int dw = 0; dw < p.kw; ++dw) {
          const int wi = wo * p.sw - p.pw + dw;
          if ((unsigned)ti < (unsigned)p.Ti && (unsigned)hi < (unsigned)p.Hi && (unsigned)wi < (unsigned)p.Wi) {
            const uint4 x = *reinterpret_cast<const uint4*>(
                p.in + ((((b * p.Ti + ti) * p.Hi + hi) * p.Wi + wi) * (long long)p.C) + v * 8);
            acc.x = bf16x2_max(acc.x, x.x);
            acc.y = bf16x2_max(acc.y, x.y);
            acc.z = bf16x2_max(acc.z, x.z);
            acc.w = bf16x2_max(acc.w, x.w);
          } else {
            any_oob = true;
          }
        }
      }
    }
    if (any_oob && p.pad_zero) {
      acc.x = bf16x2_max(acc.x, 0u);
      acc.y = bf16x2_max(acc.y, 0u);
      acc.z = bf16x2_max(acc.z, 0u);
      acc.w = bf16x2_max(acc.w, 0u);
    }
    *reinterpret_cast<uint4*>(p.out + m_out * p.ldo + v * 8) = acc;
  }
}

// Fast path for un-padded windows with compile-time extents (every tap in range by construction: floor
// output size, pad 0): the KT*KH*KW 128-bit loads of a thread are all issued before the first max, and
// the streaming loads bypass L1 allocation, so enough bytes are in flight to approach HBM bandwidth.
__device__ __forceinline__ uint4 ld_stream_16(const void* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr));
  return r;
}

template <int KT, int KH, int KW>
__global__ void __launch_bounds__(256) maxpool3d_fixed_kernel(const PoolParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.B * p.To * p.Ho * p.Wo * cv;
  const long long sW = p.C, sH = (long long)p.Wi * p.C, sT = sH * p.Hi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const long long m_out = m;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const int to = (int)(m % p.To); m /= p.To;
    // m is now the clip index; pad is 0, so the window origin is (to*st, ho*sh, wo*sw)
    const __nv_bfloat16* base = p.in + (m * p.Ti + (long long)to * p.st) * sT + (long long)ho * p.sh * sH +
                                (long long)wo * p.sw * sW + v * 8;
    uint4 x[KT * KH * KW];
#pragma unroll
    for (int dt = 0; dt < KT; ++dt)
#pragma unroll
      for (int dh = 0; dh < KH; ++dh)
#pragma unroll
        for (int dw = 0; dw < KW; ++dw) x[(dt * KH + dh) * KW + dw] = ld_stream_16(base + dt * sT + dh * sH + dw * sW);
    uint4 acc = x[0];
#pragma unroll
    for (int k = 1; k < KT * KH * KW; ++k) {
      acc.x = bf16x2_max(acc.x, x[k].x);
      acc.y = bf16x2_max(acc.y, x[k].y);
      acc.z = bf16x2_max(acc.z, x[k].z);
      acc.w = bf16x2_max(acc.w, x[k].w);
    }
    *reinterpret_cast<uint4*>(p.out + m_out * p.ldo + v * 8) = acc;
  }
}

// Padded windows with compile-time extents (the SAME-padding pools of the Inception port between stages: (1,3,3)/(1,2,2),
// (3,3,3)/(2,2,2), (2,2,2)/(2,2,2)): every tap's 128-bit load is issued (out-of-frame taps clamped) before the first max,
// instead of the dependent load -> max chain of the general kernel.
// read-only 128-bit load that DOES allocate in L1: neighbouring threads of the sliding-window pool re-read the same lines
__device__ __forceinline__ uint4 ld_nc_16(const void* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
}

template <int KT, int KH, int KW>
__global__ void __launch_bounds__(256) maxpool3d_checked_kernel(const PoolParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.B * p.To * p.Ho * p.Wo * cv;
  const long long sW = p.C, sH = (long long)p.Wi * p.C, sT = sH * p.Hi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const long long m_out = m;
    const int wo = (int)(m % p.Wo); m /= p.Wo;
    const int ho = (int)(m % p.Ho); m /= p.Ho;
    const int to = (int)(m % p.To); m /= p.To;
    const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph, w0 = wo * p.sw - p.pw;
    // A tap outside the frame is clamped onto the nearest tap inside it: the padding is narrower than the window, so
    // that tap belongs to the same window and the maximum is unchanged -- and every load is unconditional.
    const __nv_bfloat16* base = p.in + m * p.Ti * sT + v * 8;
    long long ot[KT], oh[KH], ow[KW];
#pragma unroll
    for (int dt = 0; dt < KT; ++dt) ot[dt] = (long long)min(max(t0 + dt, 0), p.Ti - 1) * sT;
#pragma unroll
    for (int dh = 0; dh < KH; ++dh) oh[dh] = (long long)min(max(h0 + dh, 0), p.Hi - 1) * sH;
#pragma unroll
    for (int dw = 0; dw < KW; ++dw) ow[dw] = (long long)min(max(w0 + dw, 0), p.Wi - 1) * sW;
    const bool any_oob = t0 < 0 || h0 < 0 || w0 < 0 || t0 + KT > p.Ti || h0 + KH > p.Hi || w0 + KW > p.Wi;
    uint4 x[KT * KH * KW];
#pragma unroll
    for (int dt = 0; dt < KT; ++dt)
#pragma unroll
      for (int dh = 0; dh < KH; ++dh)
#pragma unroll
        for (int dw = 0; dw < KW; ++dw) x[(dt * KH + dh) * KW + dw] = ld_stream_16(base + ot[dt] + oh[dh] + ow[dw]);
    uint4 acc = x[0];
#pragma unroll
    for (int k = 1; k < KT * KH * KW; ++k) {
      acc.x = bf16x2_max(acc.x, x[k].x);
      acc.y = bf16x2_max(acc.y, x[k].y);
      acc.z = bf16x2_max(acc.z, x[k].z);
      acc.w = bf16x2_max(acc.w, x[k].w);
    }
    if (any_oob && p.pad_zero) {
      acc.x = bf16x2_max(acc.x, 0u);
      acc.y = bf16x2_max(acc.y, 0u);
      acc.z = bf16x2_max(acc.z, 0u);
      acc.w = bf16x2_max(acc.w, 0u);
    }
    *reinterpret_cast<uint4*>(p.out + m_out * p.ldo + v * 8) = acc;
  }
}

// 3x3x3 / stride 1 / pad 1 (the pooling branch of every Inception module): one thread walks all T frames of one
// (clip, h, w, 8-channel vector) column, computes each frame's 3x3 spatial max once (9 loads, not 27 per output) and
// slides a 3-frame window over those.  Out-of-range taps are ignored, or count as 0 when pad_zero (SAME-padding port).
__global__ void __launch_bounds__(256) maxpool3d_k3s1_kernel(const PoolParams p) {
  const int cv = p.C >> 3;
  const long long total = (long long)p.B * p.Hi * p.Wi * cv;
  const long long sW = p.C, sH = (long long)p.Wi * p.C, sT = sH * p.Hi;
  const uint32_t neg_inf2 = 0xFF80FF80u;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long m = i / cv;
    const int w = (int)(m % p.Wi); m /= p.Wi;
    const int h = (int)(m % p.Hi); m /= p.Hi;
    const long long b = m;
    const bool border = h == 0 || w == 0 || h == p.Hi - 1 || w == p.Wi - 1;
    const __nv_bfloat16* col = p.in + (b * p.Ti) * sT + (long long)h * sH + (long long)w * sW + v * 8;
    auto spatial = [&](int t) {
      const __nv_bfloat16* c = col + t * sT;
      uint4 x[9];
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const bool ok = (unsigned)(h + dh) < (unsigned)p.Hi && (unsigned)(w + dw) < (unsigned)p.Wi;
          x[(dh + 1) * 3 + dw + 1] = ok ? ld_nc_16(c + dh * sH + dw * sW) : make_uint4(neg_inf2, neg_inf2, neg_inf2, neg_inf2);
        }
      uint4 a = x[0];
#pragma unroll
      for (int k = 1; k < 9; ++k) {
        a.x = bf16x2_max(a.x, x[k].x); a.y = bf16x2_max(a.y, x[k].y); a.z = bf16x2_max(a.z, x[k].z); a.w = bf16x2_max(a.w, x[k].w);
      }
      return a;
    };
    const uint4 none = make_uint4(neg_inf2, neg_inf2, neg_inf2, neg_inf2);
    uint4 prev = none, cur = spatial(0);
    __nv_bfloat16* out = p.out + ((b * p.Ti) * (long long)p.Hi * p.Wi + (long long)h * p.Wi + w) * p.ldo + v * 8;
    const long long oT = (long long)p.Hi * p.Wi * p.ldo;
    for (int t = 0; t < p.Ti; ++t) {
      const uint4 next = (t + 1 < p.Ti) ? spatial(t + 1) : none;
      uint4 a;
      a.x = bf16x2_max(bf16x2_max(prev.x, cur.x), next.x);
      a.y = bf16x2_max(bf16x2_max(prev.y, cur.y), next.y);
      a.z = bf16x2_max(bf16x2_max(prev.z, cur.z), next.z);
      a.w = bf16x2_max(bf16x2_max(prev.w, cur.w), next.w);
      if (p.pad_zero && (border || t == 0 || t == p.Ti - 1)) {
        a.x = bf16x2_max(a.x, 0u); a.y = bf16x2_max(a.y, 0u); a.z = bf16x2_max(a.z, 0u); a.w = bf16x2_max(a.w, 0u);
      }
      *reinterpret_cast<uint4*>(out + t * oT) = a;
      prev = cur;
      cur = next;
    }
  }
}

// Row-per-block forms of the two padded pool families above (round 2; ncu on the thread-per-output kernels: 350 warp
// instructions per output vector, most of them 64-bit index divisions, 0 % L1 hits on 9 - 27 window loads per output).
//
// maxpool3d_rows_kernel<KT, KH, KW>: a block owns one output row (b, to, ho); a thread one (wo, 8-channel vector) item.
// The (dt, dh) bounds are block-uniform, every index is 32-bit, the window loads allocate in L1 (neighbouring wo share
// columns), and all loads of a thread are issued before the first max.  Measured per 160 clip-crops against the
// thread-per-output kernel with predicated taps: MaxPool3d_3a (192 ch) 0.46 -> 0.38 ms, 4a (480 ch) 0.43 -> 0.32 ms; 2a (64 ch:
// 448 one-item threads per block) 0.59 -> 0.62 ms.  Once that kernel clamped its taps instead (maxpool3d_checked_kernel above) it
// took 4a to 0.27 and 5a to 0.038 ms, so the plan now uses this kernel for the wide (1,3,3) pool only (3a: 0.38 vs 0.41 ms;
// VAD_POOL_ROWS forces either, both stay under test).  Also measured and rejected: a
// persistent grid-strided form (3a: 0.55 ms) and a separable row-per-block form of the 3x3x3 / 1 branch pools (column max
// through shared memory, one barrier per frame: 0.40 vs 0.30 ms on Mixed_3c).
template <int KT, int KH, int KW>
__global__ void __launch_bounds__(512) maxpool3d_rows_kernel(const PoolParams p) {
  const int cv = p.C >> 3;
  const int sH = p.Wi * p.C;                      // elements per input row (< 2^31 by construction)
  const long long sT = (long long)sH * p.Hi;
  const int items = p.Wo * cv;
  int r = blockIdx.x;
  const int ho = r % p.Ho; r /= p.Ho;
  const int to = r % p.To;
  const int b = r / p.To;
  const int t0 = to * p.st - p.pt, h0 = ho * p.sh - p.ph;
  const __nv_bfloat16* in_b = p.in + (long long)b * p.Ti * sT;
  __nv_bfloat16* out_row = p.out + (((long long)b * p.To + to) * p.Ho + ho) * (long long)p.Wo * p.ldo;
  const uint32_t neg_inf2 = 0xFF80FF80u;
  const uint4 none = make_uint4(neg_inf2, neg_inf2, neg_inf2, neg_inf2);
  bool row_ok[KT * KH];
  bool row_oob = false;
#pragma unroll
  for (int dt = 0; dt < KT; ++dt)
#pragma unroll
    for (int dh = 0; dh < KH; ++dh) {
      row_ok[dt * KH + dh] = (unsigned)(t0 + dt) < (unsigned)p.Ti && (unsigned)(h0 + dh) < (unsigned)p.Hi;
      row_oob |= !row_ok[dt * KH + dh];
    }
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int wo = it / cv, v = it - wo * cv;
    const int w0 = wo * p.sw - p.pw;
    bool col_ok[KW];
    bool any_oob = row_oob;
#pragma unroll
    for (int dw = 0; dw < KW; ++dw) { col_ok[dw] = (unsigned)(w0 + dw) < (unsigned)p.Wi; any_oob |= !col_ok[dw]; }
    const __nv_bfloat16* base = in_b + (long long)t0 * sT + (long long)h0 * sH + w0 * p.C + v * 8;
    uint4 x[KT * KH * KW];
#pragma unroll
    for (int dt = 0; dt < KT; ++dt)
#pragma unroll
      for (int dh = 0; dh < KH; ++dh)
#pragma unroll
        for (int dw = 0; dw < KW; ++dw)
          x[(dt * KH + dh) * KW + dw] = (row_ok[dt * KH + dh] && col_ok[dw]) ? ld_nc_16(base + dt * sT + dh * sH + dw * p.C) : none;
    uint4 acc = x[0];
#pragma unroll
    for (int k = 1; k < KT * KH * KW; ++k) {
      acc.x = bf16x2_max(acc.x, x[k].x); acc.y = bf16x2_max(acc.y, x[k].y); acc.z = bf16x2_max(acc.z, x[k].z); acc.w = bf16x2_max(acc.w, x[k].w);
    }
    if (any_oob && p.pad_zero) {
      acc.x = bf16x2_max(acc.x, 0u); acc.y = bf16x2_max(acc.y, 0u); acc.z = bf16x2_max(acc.z, 0u); acc.w = bf16x2_max(acc.w, 0u);
    }
    *reinterpret_cast<uint4*>(out_row + (long long)wo * p.ldo + v * 8) = acc;
  }
}

// ------------------------------------------------------------------------------------------- K4
// global average pool: in [B, P, C] bf16 -> out [B, C] fp32.  A warp covers 64 channels: lane =
// pg*8 + cv reads 16 B of channel vector cv at positions pg, pg+4, ... (4 x 128 B contiguous per
// step), then the four position groups are combined with warp shuffles.
// kt_win > 0: AvgPool3d((kt_win, H, W), stride 1) followed by a global mean over the T - kt_win + 1 windows (the head the
// reference puts on pytorchvideo's I3D-R50, src/i3d.py:21-57): frame t of the P = T * HW positions is weighted by the number
// of windows that cover it.
__global__ void __launch_bounds__(256) avgpool_kernel(const __nv_bfloat16* __restrict__ in, int B, int P, int C,
                                                      float* __restrict__ out, int HW = 0, int kt_win = 0) {
  const int warps_per_clip = C >> 6;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)B * warps_per_clip) return;
  const int b = (int)(gw / warps_per_clip);
  const int c0 = (int)(gw % warps_per_clip) * 64 + (lane & 7) * 8;
  const int pg = lane >> 3;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const __nv_bfloat16* base = in + (long long)b * P * C + c0;
  if (kt_win > 0) {
    const int T = P / HW, last = T - kt_win;
    for (int pos = pg; pos < P; pos += 4) {
      const int t = pos / HW;
      const int lo = t - kt_win + 1 > 0 ? t - kt_win + 1 : 0, hi = t < last ? t : last;
      const float w = (float)(hi - lo + 1);
      const uint4 x = *reinterpret_cast<const uint4*>(base + (long long)pos * C);
      acc[0] = fmaf(w, bf16_lo(x.x), acc[0]); acc[1] = fmaf(w, bf16_hi(x.x), acc[1]);
      acc[2] = fmaf(w, bf16_lo(x.y), acc[2]); acc[3] = fmaf(w, bf16_hi(x.y), acc[3]);
      acc[4] = fmaf(w, bf16_lo(x.z), acc[4]); acc[5] = fmaf(w, bf16_hi(x.z), acc[5]);
      acc[6] = fmaf(w, bf16_lo(x.w), acc[6]); acc[7] = fmaf(w, bf16_hi(x.w), acc[7]);
    }
  } else
  for (int pos = pg; pos < P; pos += 4) {
    const uint4 x = *reinterpret_cast<const uint4*>(base + (long long)pos * C);
    acc[0] += bf16_lo(x.x); acc[1] += bf16_hi(x.x);
    acc[2] += bf16_lo(x.y); acc[3] += bf16_hi(x.y);
    acc[4] += bf16_lo(x.z); acc[5] += bf16_hi(x.z);
    acc[6] += bf16_lo(x.w); acc[7] += bf16_hi(x.w);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  }
  if (pg == 0) {
    const float inv = kt_win > 0 ? 1.f / ((float)HW * (float)kt_win * (float)(P / HW - kt_win + 1)) : 1.f / (float)P;
    float4* o = reinterpret_cast<float4*>(out + (long long)b * C + c0);
    o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

// segment mean: feats [n_clips, ncrops, C] -> out [ncrops, seg, C].  Bin edges i*n/seg are the
// integer form of np.linspace(0, n, seg+1, dtype=int); rows are added in clip order and divided
// once (IEEE), which is what np.mean over axis 0 does, so the result is bit-identical.
// Lanes span channels (coalesced); the clip axis is a serial loop by construction.
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ feats, int n_clips, int ncrops,
                                                           int C, int seg, float* __restrict__ out) {
  const long long total = (long long)ncrops * seg * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int s = (int)((i / C) % seg);
    const int k = (int)(i / ((long long)C * seg));
    const int a = (int)(((long long)s * n_clips) / seg);
    const int b = (int)(((long long)(s + 1) * n_clips) / seg);
    const float* col = feats + (long long)k * C + c;
    const long long stride = (long long)ncrops * C;
    float r;
    if (a != b) {
      float sum = col[a * stride];
      for (int j = a + 1; j < b; ++j) sum = __fadd_rn(sum, col[j * stride]);
      r = __fdiv_rn(sum, (float)(b - a));
    } else {
      r = col[a * stride];
    }
    out[i] = r;
  }
}

// add_magnitude: one warp per row; out[row] = concat(feat[row], ||feat[row]||_2)
__global__ void __launch_bounds__(256) add_magnitude_kernel(const float* __restrict__ feats, long long rows, int C,
                                                            float* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = feats + row * C;
  float* dst = out + row * (C + 1);
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = src[c];
    dst[c] = v;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) dst[C] = sqrtf(ss);
}

}  // namespace vad
