// K5t: backward pass, BatchNorm batch statistics and Adam for the MGFN scoring head -- the training step of
// src/runner.py:29-39,53-59 over src/models/mgfn/modeling_mgfn.py:36-427 and src/loss/*.py.
//
// Contractions reuse the forward tcgen05 kind::tf32 GEMM (head_kernels.cuh) with re-arranged operands:
//   dgrad  dA = conv_taps(dOut) against WT[cin][tap'][n] = W[n][taps - 1 - tap'][cin]  (same kernel, transformed weights;
//          the residual branch's gradient rides in as the epilogue's `res`)
//   wgrad  dW[:, tap, :] = dOutT [N, tokens] . (A shifted by tap)T [Cin, tokens]       (same kernel: "activation" = dOutT,
//          "weights" = the transposed, tap-shifted activation; K = tokens)
// so the only new data movement is the tiled transpose below (which also produces the bias gradients as column sums).
// Everything else is small fp32 SIMT: the two LayerNorm flavours, GELU, Glance attention, the Focus depth-wise relation
// conv and its BatchNorm1d (batch statistics in train mode), selection scatter, all loss terms, Adam.
#pragma once

#include "head_kernels.cuh"

namespace vad {

// out[c * ld_out + tok] = in[(tok + shift) * ld_in + c] if token tok + shift lies in the same T-token sequence, else 0.
// colsum (optional, shift must be 0): colsum[c] += sum over tokens of in[tok, c]   (bias gradient).
// grid (ceil(ntok / 32), ceil(C / 32)), block (32, 8).
__global__ void __launch_bounds__(256) head_transpose_kernel(const float* __restrict__ in, long long ld_in, long long ntok, int C, int T, int shift,
                                                             float* __restrict__ out, long long ld_out, float* __restrict__ colsum) {
  __shared__ float tile[32][33];
  const long long tok0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const long long tok = tok0 + i;
    float v = 0.f;
    if (tok < ntok && c0 + tx < C) {
      const int t = (int)(tok % T) + shift;
      if (t >= 0 && t < T) v = in[(tok + shift) * ld_in + c0 + tx];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long tok = tok0 + tx;
    if (c < C && tok < ntok) out[(long long)c * ld_out + tok] = tile[tx][i];
  }
  if (colsum != nullptr && ty == 0 && c0 + tx < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += tile[i][tx];
    atomicAdd(colsum + c0 + tx, s);
  }
}

// z = gelu(zp): the train-mode forward keeps the pre-activation for the backward pass
__global__ void head_gelu_fwd_kernel(const float* __restrict__ zp, float* __restrict__ z, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) z[i] = gelu_erf(zp[i]);
}

// dzp = dz * gelu'(zp), exact (erf) GELU; in place on dz
__global__ void head_gelu_bwd_kernel(float* __restrict__ dz, const float* __restrict__ zp, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = zp[i];
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    dz[i] *= cdf + x * pdf;
  }
}

// MGFNLayerNorm backward (y = (x - mean) / (std + eps) * g + b, population std over channels):
//   dx = r * (dxh - mean(dxh)) - r^2 / (C * std) * (x - mean) * sum(dxh * (x - mean)),  dxh = dy * g,  r = 1 / (std + eps)
// dx_out = dx (+ dres);  dg += sum_tok dy * xh, db += sum_tok dy.  One warp per token (grid-stride), C <= 1024.
__global__ void __launch_bounds__(256) head_mgfn_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ g,
                                                               const float* __restrict__ dres, float eps, long long ntok, int C,
                                                               float* __restrict__ dx, float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float sg[1024], sb[1024];
  for (int c = threadIdx.x; c < C; c += blockDim.x) { sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float pg[32], pb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  for (long long tok = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tok < ntok; tok += warps) {
    const float* xr = x + tok * C;
    const float* dyr = dy + tok * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f, a = 0.f, d = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xc = xr[c] - mean, dxh = dyr[c] * g[c];
      v = fmaf(xc, xc, v);
      a += dxh;
      d = fmaf(dxh, xc, d);
    }
    const float sd = sqrtf(warp_sum(v) / (float)C);
    const float r = 1.f / (sd + eps);
    const float mdxh = warp_sum(a) / (float)C;
    const float k2 = sd > 0.f ? r * r * warp_sum(d) / ((float)C * sd) : 0.f;
    float* o = dx + tok * C;
    const float* rs = dres ? dres + tok * C : nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float xc = xr[c] - mean, dyc = dyr[c];
        float val = r * (dyc * g[c] - mdxh) - k2 * xc;
        if (rs) val += rs[c];
        o[c] = val;
        pg[i] = fmaf(dyc, xc * r, pg[i]);
        pb[i] += dyc;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(&sg[c], pg[i]); atomicAdd(&sb[c], pb[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(dg + c, sg[c]); atomicAdd(db + c, sb[c]); }
}

// Glance attention backward (modeling_mgfn.py:109-127), dim_head 64, T <= 64.  One block per (sequence, head):
//   S = (q * scale) k^T, P = softmax(S), O = P v;   dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dP o P)),
//   dq = dS k * scale, dk = dS^T (q * scale).   qkv / dqkv: [ntok, 3 * heads * 64] (q | k | v), dout: [ntok, heads * 64].
__global__ void __launch_bounds__(256) head_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dqkv,
                                                                 int T, int heads, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;                 // [T][65]  q * scale
  float* sk = sq + T * 65;        // [T][65]
  float* sv = sk + T * 65;        // [T][65]
  float* sdo = sv + T * 65;       // [T][65]
  float* sp = sdo + T * 65;       // [T][T + 1]  P, then dS
  float* sdp = sp + T * (T + 1);  // [T][T + 1]  dP
  const int s = blockIdx.x, h = blockIdx.y;
  const int inner = heads * 64;
  const long long base = (long long)s * T;
  for (int e = threadIdx.x; e < T * 64; e += blockDim.x) {
    const int i = e >> 6, d = e & 63;
    const float* row = qkv + (base + i) * 3 * inner + h * 64 + d;
    sq[i * 65 + d] = row[0] * scale;
    sk[i * 65 + d] = row[inner];
    sv[i * 65 + d] = row[2 * inner];
    sdo[i * 65 + d] = dout[(base + i) * inner + h * 64 + d];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int i = e / T, j = e - i * T;
    float a = 0.f, b = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) {
      a = fmaf(sq[i * 65 + d], sk[j * 65 + d], a);
      b = fmaf(sdo[i * 65 + d], sv[j * 65 + d], b);
    }
    sp[i * (T + 1) + j] = a;
    sdp[i * (T + 1) + j] = b;
  }
  __syncthreads();
  // row softmax and dS, one warp per row
  for (int i = threadIdx.x >> 5; i < T; i += blockDim.x >> 5) {
    const int lane = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sp[i * (T + 1) + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
    for (int j = lane; j < T; j += 32) { const float e = __expf(sp[i * (T + 1) + j] - mx); sp[i * (T + 1) + j] = e; l += e; }
    l = warp_sum(l);
    const float inv = 1.f / l;
    float dot = 0.f;
    for (int j = lane; j < T; j += 32) { const float pj = sp[i * (T + 1) + j] * inv; sp[i * (T + 1) + j] = pj; dot = fmaf(pj, sdp[i * (T + 1) + j], dot); }
    dot = warp_sum(dot);
    // keep P in sdp's place? no: dV needs P, dq / dk need dS -> write dS into sdp, P stays in sp
    for (int j = lane; j < T; j += 32) sdp[i * (T + 1) + j] = sp[i * (T + 1) + j] * (sdp[i * (T + 1) + j] - dot);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < T * 64; e += blockDim.x) {
    const int i = e >> 6, d = e & 63;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < T; ++j) {
      dq = fmaf(sdp[i * (T + 1) + j], sk[j * 65 + d], dq);   // dS[i, j] k[j]
      dk = fmaf(sdp[j * (T + 1) + i], sq[j * 65 + d], dk);   // dS[j, i] (q[j] * scale)
      dv = fmaf(sp[j * (T + 1) + i], sdo[j * 65 + d], dv);   // P[j, i] dO[j]
    }
    float* row = dqkv + (base + i) * 3 * inner + h * 64 + d;
    row[0] = dq * scale;
    row[inner] = dk;
    row[2 * inner] = dv;
  }
}

// Focus relation conv backward, input gradient: dv[tok, c] = sum_j w[c % heads][j] * du[tok - (j - half), c]
__global__ void head_relpos_bwd_data_kernel(const float* __restrict__ du, const float* __restrict__ w, float* __restrict__ dv, int S, int T, int C,
                                            int heads, int k) {
  const long long total = (long long)S * T * C;
  const int half = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long tok = i / C;
    const int t = (int)(tok % T);
    const int h = c % heads;
    float a = 0.f;
    for (int j = 0; j < k; ++j) {
      const int ts = t - (j - half);
      if (ts >= 0 && ts < T) a = fmaf(w[h * k + j], du[(tok - (j - half)) * C + c], a);
    }
    dv[i] = a;
  }
}
// ... weight / bias gradient:  dw[h][j] += sum_{tok, c % heads == h} du[tok, c] * v[tok + j - half, c],  db[h] += sum du[tok, c]
// (dw / db zeroed by the caller).  grid (C / 32, token chunks), block (32, 8): a thread owns one channel over a slice of
// the tokens and keeps k + 1 partial sums; block reduction per channel, then one atomicAdd per (channel, tap).  k <= 7.
__global__ void __launch_bounds__(256) head_relpos_bwd_weight_kernel(const float* __restrict__ du, const float* __restrict__ v, float* __restrict__ dw,
                                                                     float* __restrict__ db, int S, int T, int C, int heads, int k) {
  __shared__ float red[8][8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const int half = k / 2;
  const long long ntok = (long long)S * T;
  const long long per = (ntok + gridDim.y - 1) / gridDim.y;
  const long long lo = (long long)blockIdx.y * per, hi = lo + per < ntok ? lo + per : ntok;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (c < C) {
    for (long long tok = lo + ty; tok < hi; tok += 8) {
      const int t = (int)(tok % T);
      const float g = du[tok * C + c];
      a[7] += g;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int ts = t + j - half;
        if (j < k && ts >= 0 && ts < T) a[j] = fmaf(g, v[(tok + j - half) * C + c], a[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j][ty][tx] = a[j];
  __syncthreads();
  if (ty == 0 && c < C) {
    const int h = c % heads;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += red[j][i][tx];
      if (j == 7) atomicAdd(db + h, sum);
      else if (j < k) atomicAdd(dw + h * k + j, sum);
    }
  }
}

// BatchNorm1d in train mode (FocusAttention.norm, modeling_mgfn.py:178): per-channel statistics over all tokens.
// grid (C / 32, token chunks), block (32, 8); every block writes its two partial sums to part[chunk][2][C] and a second
// kernel adds the chunks in a fixed order in double precision -- no atomics: the batch statistics (and with them the whole
// train-mode forward and the loss) are bit-reproducible run to run, and E[x^2] - E[x]^2 does not amplify summation-order
// noise (with fp32 atomics the loss terms moved by 1e-3 between two runs on the same input).
//   mode 0: sum x,  sum x^2          (head_bn_finalize_kernel turns them into mean / invstd)
//   mode 1: sum dy, sum dy * xhat    (d beta, d gamma: head_bn_grad_finalize_kernel)
__global__ void __launch_bounds__(256) head_bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long ntok, int C, int mode,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             float* __restrict__ part) {
  __shared__ float ra[8][33], rb[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  const long long per = (ntok + gridDim.y - 1) / gridDim.y;
  const long long lo = (long long)blockIdx.y * per, hi = lo + per < ntok ? lo + per : ntok;
  float a = 0.f, b = 0.f;
  if (c < C) {
    if (mode == 0) {
      for (long long r = lo + ty; r < hi; r += 8) { const float v = x[r * C + c]; a += v; b = fmaf(v, v, b); }
    } else {
      const float mu = mean[c], is = invstd[c];
      for (long long r = lo + ty; r < hi; r += 8) {
        const float g = dy[r * C + c];
        a += g;
        b = fmaf(g, (x[r * C + c] - mu) * is, b);
      }
    }
  }
  ra[ty][tx] = a; rb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa += ra[i][tx]; sb += rb[i][tx]; }
    part[((size_t)blockIdx.y * 2 + 0) * C + c] = sa;
    part[((size_t)blockIdx.y * 2 + 1) * C + c] = sb;
  }
}
// chunk partials -> batch mean / invstd (biased variance) and the running-statistics update (momentum, unbiased variance, as torch)
__global__ void head_bn_finalize_kernel(const float* __restrict__ part, int chunks, long long ntok, int C, float eps, float momentum,
                                        float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean, float* __restrict__ run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  for (int k = 0; k < chunks; ++k) { s += (double)part[((size_t)k * 2) * C + c]; ss += (double)part[((size_t)k * 2 + 1) * C + c]; }
  const double mud = s / (double)ntok;
  double vard = ss / (double)ntok - mud * mud;
  vard = vard > 0.0 ? vard : 0.0;
  const float mu = (float)mud, var = (float)vard;
  mean[c] = mu;
  invstd[c] = rsqrtf(var + eps);
  if (run_mean) {
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mu;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (ntok > 1 ? var * (float)ntok / (float)(ntok - 1) : var);
  }
}
// chunk partials of mode 1 -> s1 = sum dy (= d beta), s2 = sum dy * xhat (= d gamma), fixed summation order
__global__ void head_bn_grad_finalize_kernel(const float* __restrict__ part, int chunks, int C, float* __restrict__ s1, float* __restrict__ s2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < chunks; ++k) { a += part[((size_t)k * 2) * C + c]; b += part[((size_t)k * 2 + 1) * C + c]; }
  s1[c] = a;
  s2[c] = b;
}
// y = (x - mean) * invstd * g + b
__global__ void head_bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ g, const float* __restrict__ b, long long ntok, int C, float* __restrict__ y) {
  const long long total = ntok * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = (x[i] - mean[c]) * invstd[c] * g[c] + b[c];
  }
}
// dx = g * invstd * (dy - s1 / N - xhat * s2 / N) (+ dres);  dgamma = s2, dbeta = s1 are already in place
__global__ void head_bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ g, const float* __restrict__ s1,
                                         const float* __restrict__ s2, const float* __restrict__ dres, long long ntok, int C, float* __restrict__ dx) {
  const long long total = ntok * C;
  const float invn = 1.f / (float)ntok;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float xh = (x[i] - mean[c]) * invstd[c];
    float v = g[c] * invstd[c] * (dy[i] - s1[c] * invn - xh * s2[c] * invn);
    if (dres) v += dres[i];
    dx[i] = v;
  }
}

// Feature amplifier (modeling_mgfn.py:84-93), magnitude branch: x += ratio * (Conv1d(1 -> C, k 3)(mag) + b)
//   dw[c][j] = ratio * sum_tok dx[tok, c] * mag[tok + j - 1],  db[c] = ratio * sum_tok dx[tok, c].  One block per channel.
__global__ void __launch_bounds__(256) head_amplify_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ mag, float ratio, int S, int T,
                                                               int C, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[4][256];
  const int c = blockIdx.x;
  const long long ntok = (long long)S * T;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long tok = threadIdx.x; tok < ntok; tok += blockDim.x) {
    const int t = (int)(tok % T);
    const float g = dx[tok * C + c];
    if (t > 0) a[0] = fmaf(g, mag[tok - 1], a[0]);
    a[1] = fmaf(g, mag[tok], a[1]);
    if (t + 1 < T) a[2] = fmaf(g, mag[tok + 1], a[2]);
    a[3] += g;
  }
  for (int i = 0; i < 4; ++i) red[i][threadIdx.x] = a[i];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int i = 0; i < 4; ++i) red[i][threadIdx.x] += red[i][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dw[c * 3 + 0] = ratio * red[0][0];
    dw[c * 3 + 1] = ratio * red[1][0];
    dw[c * 3 + 2] = ratio * red[2][0];
    db[c] = ratio * red[3][0];
  }
}

// Selection in train mode (modeling_mgfn.py:341-346): like head_select_kernel, with the dropout mask of the selection
// (values 0 or 1 / (1 - p), one row per video of the batch) multiplied onto the crop-mean magnitudes before the top-k.
__global__ void __launch_bounds__(256) head_select_train_kernel(const float* __restrict__ score_tok, const float* __restrict__ fmag_tok,
                                                                const float* __restrict__ xln, const float* __restrict__ mask, int n_videos,
                                                                int ncrops, int T, int C, int k, float* __restrict__ scores,
                                                                float* __restrict__ vid_score, int* __restrict__ idx_out, float* __restrict__ sel,
                                                                int video_off) {
  extern __shared__ float sh[];  // [T] magnitudes
  __shared__ int s_idx[8];
  const int bl = blockIdx.x;
  const int b = video_off + bl;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float m = 0.f, sc = 0.f;
    for (int c = 0; c < ncrops; ++c) {
      const long long tok = ((long long)b * ncrops + c) * T + t;
      m += fmag_tok[tok];
      sc += score_tok[tok];
    }
    m /= (float)ncrops;
    sh[t] = mask ? m * mask[(long long)b * T + t] : m;
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

// Loss backward (src/loss/base.py:7-48, src/loss/mgfn.py:7-47): one block.  Inputs as head_loss_kernel (l1 = the L1 norms it
// left in its scratch buffer).  Outputs: dscores[bs, T] = d loss / d scores (smoothness, sparsity and, through the mean of the
// selected snippets, BCE), dl1[2 R k] = d loss / d (L1 norm of each selected feature).
__global__ void __launch_bounds__(256) head_loss_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ vid_score,
                                                            const float* __restrict__ labels, const int* __restrict__ idx, const float* __restrict__ l1,
                                                            int nn, int ncrops, int T, int k, float* __restrict__ dscores, float* __restrict__ dl1,
                                                            float w_smooth, float w_sparse, float alpha, float margin) {
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
  const int R = ncrops * nn;
  float v = 0.f;
  for (int e = threadIdx.x; e < nn * T; e += blockDim.x) v = fmaf(scores[e], scores[e], v);
  const float nrm = sqrtf(block_sum(v));
  for (int e = threadIdx.x; e < bs * T; e += blockDim.x) {
    const int b = e / T, t = e - b * T;
    const float s = scores[e];
    float g = 0.f;
    if (t > 0) g += s - scores[e - 1];
    if (t + 1 < T) g -= scores[e + 1] - s;
    g *= 2.f * w_smooth;
    if (b < nn && nrm > 0.f) g += w_sparse * s / nrm;
    dscores[e] = g;
  }
  __syncthreads();
  // BCE through vid_score[b] = mean_j scores[b, idx[b, j]]   (nn.BCELoss clamps the logs at -100: no gradient there)
  for (int b = threadIdx.x; b < bs; b += blockDim.x) {
    const float pr = vid_score[b], y = labels[b];
    float dp = 0.f;
    if (logf(pr) > -100.f) dp -= y / pr;
    if (logf(1.f - pr) > -100.f) dp += (1.f - y) / (1.f - pr);
    dp /= (float)bs;
    for (int j = 0; j < k; ++j) dscores[b * T + idx[b * k + j]] += dp / (float)k;   // indices of one video are distinct
  }
  // contrastive terms on the L1 norms: rows e of ln / la (k entries each)
  const float* ln = l1;
  const float* la = l1 + R * k;
  float* dn = dl1;
  float* da = dl1 + R * k;
  for (int e = threadIdx.x; e < 2 * R * k; e += blockDim.x) dl1[e] = 0.f;
  __syncthreads();
  const int sep = R / 2, rest = R - sep, cnt = sep < rest ? sep : rest;
  // every term touches row e (and row sep + e) from exactly one thread per loop; loops are separated by barriers
  for (int e = threadIdx.x; e < R; e += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; ++j) { const float d = la[e * k + j] - ln[e * k + j] + 1e-6f; s = fmaf(d, d, s); }
    const float dist = sqrtf(s);
    const float m = fmaxf(margin - dist, 0.f);
    if (m > 0.f && dist > 0.f) {
      const float coef = alpha * alpha * (-2.f * m) / ((float)R * dist);
      for (int j = 0; j < k; ++j) {
        const float d = la[e * k + j] - ln[e * k + j] + 1e-6f;
        da[e * k + j] += coef * d;
        dn[e * k + j] -= coef * d;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    for (int j = 0; j < k; ++j) {
      const float d = ln[(sep + e) * k + j] - ln[e * k + j] + 1e-6f;
      const float g = alpha * 2.f * d / (float)cnt;
      dn[(sep + e) * k + j] += g;
      dn[e * k + j] -= g;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    for (int j = 0; j < k; ++j) {
      const float d = la[(sep + e) * k + j] - la[e * k + j] + 1e-6f;
      const float g = alpha * 2.f * d / (float)cnt;
      da[(sep + e) * k + j] += g;
      da[e * k + j] -= g;
    }
  }
}

// d xln[token of (video, crop, idx[j])] += sign(sel) * dl1: the gradient of the L1 norms of the selected features scattered
// back to the tokens they were gathered from (crop-major sel, as head_select_kernel writes it).  grid = n_sel videos.
__global__ void __launch_bounds__(256) head_select_bwd_kernel(const float* __restrict__ sel, const float* __restrict__ dl1, const int* __restrict__ idx,
                                                              int n_videos, int ncrops, int T, int C, int k, int video_off, float* __restrict__ dxln) {
  const int bl = blockIdx.x;
  const int b = video_off + bl;
  for (int e = threadIdx.x; e < ncrops * k * C; e += blockDim.x) {
    const int c = e % C;
    const int j = (e / C) % k;
    const int crop = e / (C * k);
    const long long row = ((long long)crop * n_videos + bl) * k + j;
    const float s = sel[row * C + c];
    const float g = dl1[row] * (s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f));
    dxln[(((long long)b * ncrops + crop) * T + idx[b * k + j]) * C + c] += g;
  }
}

// Final nn.LayerNorm + Linear(C, 1) + sigmoid backward (modeling_mgfn.py:404-409).  Per token:
//   dlogit = dscores[video, t] / ncrops * s (1 - s);   dy = dxln + dlogit * fcw;   standard LayerNorm backward to dx;
//   dg += dy * xhat, db += dy, dfcw += dlogit * y, dfcb += dlogit.   One warp per token (grid-stride), C <= 1024.
__global__ void __launch_bounds__(256) head_final_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xln, const float* __restrict__ score,
                                                             const float* __restrict__ dxln, const float* __restrict__ dscores, const float* __restrict__ g,
                                                             const float* __restrict__ fcw, float eps, long long ntok, int C, int ncrops, int T,
                                                             float* __restrict__ dx, float* __restrict__ dg, float* __restrict__ db,
                                                             float* __restrict__ dfcw, float* __restrict__ dfcb) {
  __shared__ float sg[1024], sb[1024], sw[1024];
  __shared__ float sfb;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { sg[c] = 0.f; sb[c] = 0.f; sw[c] = 0.f; }
  if (threadIdx.x == 0) sfb = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float pg[32], pb[32], pw[32], pfb = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { pg[i] = 0.f; pb[i] = 0.f; pw[i] = 0.f; }
  for (long long tok = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tok < ntok; tok += warps) {
    const long long seq = tok / T;
    const int t = (int)(tok - seq * T);
    const long long vid = seq / ncrops;
    const float sc = score[tok];
    const float dlogit = dscores[vid * T + t] / (float)ncrops * sc * (1.f - sc);
    const float* xr = x + tok * C;
    const float* yr = xln + tok * C;
    const float* dr = dxln + tok * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f, a = 0.f, d = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xc = xr[c] - mean;
      const float dxh = (dr[c] + dlogit * fcw[c]) * g[c];
      v = fmaf(xc, xc, v);
      a += dxh;
      d = fmaf(dxh, xc, d);
    }
    const float r = rsqrtf(warp_sum(v) / (float)C + eps);
    const float m1 = warp_sum(a) / (float)C;
    const float m2 = warp_sum(d) * r * r / (float)C;   // mean(dxh * xhat) * r  (xhat = xc * r)
    float* o = dx + tok * C;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float xc = xr[c] - mean;
        const float dy = dr[c] + dlogit * fcw[c];
        o[c] = r * (dy * g[c] - m1 - xc * m2);
        pg[i] = fmaf(dy, xc * r, pg[i]);
        pb[i] += dy;
        pw[i] = fmaf(dlogit, yr[c], pw[i]);
      }
    }
    if (lane == 0) pfb += dlogit;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(&sg[c], pg[i]); atomicAdd(&sb[c], pb[i]); atomicAdd(&sw[c], pw[i]); }
  }
  if (lane == 0) atomicAdd(&sfb, pfb);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { atomicAdd(dg + c, sg[c]); atomicAdd(db + c, sb[c]); atomicAdd(dfcw + c, sw[c]); }
  if (threadIdx.x == 0) atomicAdd(dfcb, sfb);
}

// torch.optim.Adam (src/runner.py:53-59: lr 1e-3, weight_decay 5e-4 as L2 added to the gradient), one fused pass over the
// flat parameter blob.  bc1 = 1 - beta1^step, bc2 = 1 - beta2^step (computed on the host in double).
__global__ void head_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                 float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float gi = fmaf(wd, w, g[i] * grad_scale);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mi / denom);
  }
}

}  // namespace vad
