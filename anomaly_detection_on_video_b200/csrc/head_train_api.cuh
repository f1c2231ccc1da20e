// Host side of the MGFN training step (included at the end of vad_api.cu after head_api.cuh).
#pragma once

#include "head_train_kernels.cuh"

struct TrainBlock {
  int type = 0, dim = 0, heads = 0, inner = 0;
  size_t scc_w = 0, scc_b = 0;
  size_t ln_g = 0, ln_b = 0, qkv_w = 0;                        // glance
  size_t v_w = 0, bn_g = 0, bn_b = 0, rp_w = 0, rp_b = 0;      // focus (raw to_v weight + its BatchNorm1d)
  size_t bn_stat = 0;                                          // focus: offset of running_mean | running_var in the stats buffer
  size_t out_w = 0, out_b = 0;
  size_t fln_g = 0, fln_b = 0, in_w = 0, in_b = 0, o2_w = 0, o2_b = 0;
};

struct vad_head_train {
  vad_head_config cfg;
  int device = 0;
  vad_head gemm_ctx;  // tensor-map encoder + launch counter for head_gemm
  size_t amp_w = 0, amp_b = 0, mag_w = 0, mag_b = 0;
  std::vector<TrainBlock> blocks;
  std::vector<HeadInter> inter;
  size_t fin_g = 0, fin_b = 0, fc_w = 0, fc_b = 0;
  size_t total_floats = 0, bn_floats = 0;
  int max_dim = 0, max_wide = 0;
  size_t max_weight = 0;
  int n_launches = 0;
};

extern "C" int32_t vad_head_train_create(vad_head_train_t** out, const vad_head_config* cfg, int32_t device) {
  if (!out || !cfg) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: null pointer");
  if (cfg->n_stages < 1 || cfg->n_stages > 4) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: n_stages must be 1..4");
  if (cfg->dim_head != 64) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: dim_head must be 64");
  if (cfg->channels % 64 || cfg->k < 1 || cfg->k > 8 || cfg->local_aggr_kernel < 1 || cfg->local_aggr_kernel > 7 || !(cfg->local_aggr_kernel & 1) || cfg->ff_repe < 1)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: bad config (channels %% 64, 1 <= k <= 8, odd local_aggr_kernel <= 7)");
  int32_t rc = require_sm100(device);
  if (rc != VAD_OK) return rc;
  vad_head_train* h = new vad_head_train();
  h->cfg = *cfg;
  h->device = device;
  size_t cur = 0, bn = 0;
  const int d0 = cfg->dims[0];
  h->amp_w = head_take(cur, (size_t)d0 * 3 * cfg->channels);
  h->amp_b = head_take(cur, d0);
  h->mag_w = head_take(cur, (size_t)d0 * 3);
  h->mag_b = head_take(cur, d0);
  h->max_weight = (size_t)d0 * 3 * cfg->channels;
  for (int st = 0; st < cfg->n_stages; ++st) {
    const int d = cfg->dims[st];
    if (d % 64 || d <= 0 || d > 1024) { delete h; return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: dims must be multiples of 64, at most 1024"); }
    if (cfg->types[st] != VAD_HEAD_GLANCE && cfg->types[st] != VAD_HEAD_FOCUS) { delete h; return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_create: bad stage type"); }
    const int heads = d / cfg->dim_head, inner = heads * cfg->dim_head, wide = cfg->ff_repe * d;
    if (d > h->max_dim) h->max_dim = d;
    if (wide > h->max_wide) h->max_wide = wide;
    if (3 * inner > h->max_wide) h->max_wide = 3 * inner;
    if ((size_t)wide * d > h->max_weight) h->max_weight = (size_t)wide * d;
    if ((size_t)3 * d * d > h->max_weight) h->max_weight = (size_t)3 * d * d;
    for (int b = 0; b < cfg->depths[st]; ++b) {
      TrainBlock k;
      k.type = cfg->types[st]; k.dim = d; k.heads = heads; k.inner = inner;
      k.scc_w = head_take(cur, (size_t)d * 3 * d);
      k.scc_b = head_take(cur, d);
      if (k.type == VAD_HEAD_GLANCE) {
        k.ln_g = head_take(cur, d);
        k.ln_b = head_take(cur, d);
        k.qkv_w = head_take(cur, (size_t)3 * inner * d);
      } else {
        k.v_w = head_take(cur, (size_t)inner * d);
        k.bn_g = head_take(cur, d);
        k.bn_b = head_take(cur, d);
        k.rp_w = head_take(cur, (size_t)heads * cfg->local_aggr_kernel);
        k.rp_b = head_take(cur, heads);
        k.bn_stat = bn;
        bn += 2 * (size_t)d;
      }
      k.out_w = head_take(cur, (size_t)d * inner);
      k.out_b = head_take(cur, d);
      k.fln_g = head_take(cur, d);
      k.fln_b = head_take(cur, d);
      k.in_w = head_take(cur, (size_t)wide * d);
      k.in_b = head_take(cur, wide);
      k.o2_w = head_take(cur, (size_t)d * wide);
      k.o2_b = head_take(cur, d);
      h->blocks.push_back(k);
    }
    HeadInter it;
    if (st + 1 < cfg->n_stages) {
      it.present = true; it.din = d; it.dout = cfg->dims[st + 1];
      it.ln_g = head_take(cur, d);
      it.ln_b = head_take(cur, d);
      it.w = head_take(cur, (size_t)it.dout * d);
      it.b = head_take(cur, it.dout);
      if ((size_t)it.dout * d > h->max_weight) h->max_weight = (size_t)it.dout * d;
    }
    h->inter.push_back(it);
  }
  const int dl = cfg->dims[cfg->n_stages - 1];
  h->fin_g = head_take(cur, dl);
  h->fin_b = head_take(cur, dl);
  h->fc_w = head_take(cur, dl);
  h->fc_b = head_take(cur, 1);
  h->total_floats = cur;
  h->bn_floats = bn;
  void* fn = nullptr;
  rc = driver_symbol("cuTensorMapEncodeTiled", &fn);
  if (rc != VAD_OK) { delete h; return rc; }
  h->gemm_ctx.encode_tiled = reinterpret_cast<EncodeTiledFnH>(fn);
  *out = h;
  return VAD_OK;
}

extern "C" void vad_head_train_destroy(vad_head_train_t* h) { delete h; }
extern "C" uint64_t vad_head_train_param_floats(const vad_head_train_t* h) { return h ? h->total_floats : 0; }
extern "C" uint64_t vad_head_train_bn_floats(const vad_head_train_t* h) { return h ? h->bn_floats : 0; }
extern "C" int32_t vad_head_train_num_launches(const vad_head_train_t* h) { return h ? h->n_launches : 0; }

// Bump allocator over the caller's workspace (1024-byte granules); base == nullptr just sizes.
struct TrainArena {
  uint8_t* base;
  size_t cur = 0;
  explicit TrainArena(uint8_t* b) : base(b) {}
  float* take(size_t floats) {
    const size_t at = cur;
    cur += (floats * 4 + 1023) / 1024 * 1024;
    return base ? reinterpret_cast<float*>(base + at) : nullptr;
  }
};

struct TrainBlockBufs {
  float *x1, *y1, *qkv, *u, *xb, *v, *bn_mean, *bn_invstd, *x2, *y2, *zp, *z, *x3;
};
struct TrainBufs {
  float *feat, *mag, *x0;
  std::vector<TrainBlockBufs> blk;
  std::vector<float*> inter_y, inter_x;
  float *xln, *score_tok, *fmag_tok, *scores, *vid_score, *sel_n, *sel_a, *l1, *dl1, *dscores, *dxln;
  int* idx;
  float *g0, *g1, *g2, *gw0, *gw1;     // gradient ping-pong [ntok, max_dim] x3, [ntok, max_wide] x2
  float *t1, *t2, *wt;                 // transposes: [max_wide or max rows, ntok], [max cols, ntok]; transformed weights
  float *bn_part;                      // BatchNorm column-reduction partials [chunks][2][C] (chunks * C <= 592 * 32)
};

static void head_train_layout(const vad_head_train* h, int n_videos, int ncrops, int T, TrainArena& a, TrainBufs& b) {
  const vad_head_config& c = h->cfg;
  const long long ntok = (long long)n_videos * ncrops * T;
  const int dl = c.dims[c.n_stages - 1];
  b.feat = a.take((size_t)ntok * c.channels);
  b.mag = a.take((size_t)ntok);
  b.x0 = a.take((size_t)ntok * c.dims[0]);
  b.blk.clear();
  for (const TrainBlock& k : h->blocks) {
    TrainBlockBufs q;
    memset(&q, 0, sizeof(q));
    const size_t d = (size_t)k.dim, wide = (size_t)c.ff_repe * k.dim;
    q.x1 = a.take(ntok * d);
    if (k.type == VAD_HEAD_GLANCE) {
      q.y1 = a.take(ntok * d);
      q.qkv = a.take(ntok * 3 * k.inner);
    } else {
      q.xb = a.take(ntok * d);
      q.v = a.take(ntok * k.inner);
      q.bn_mean = a.take(d); q.bn_invstd = a.take(d);
    }
    q.u = a.take(ntok * k.inner);
    q.x2 = a.take(ntok * d);
    q.y2 = a.take(ntok * d);
    q.zp = a.take(ntok * wide);
    q.z = a.take(ntok * wide);
    q.x3 = a.take(ntok * d);
    b.blk.push_back(q);
  }
  b.inter_y.clear(); b.inter_x.clear();
  for (const HeadInter& it : h->inter) {
    b.inter_y.push_back(it.present ? a.take((size_t)ntok * it.din) : nullptr);
    b.inter_x.push_back(it.present ? a.take((size_t)ntok * it.dout) : nullptr);
  }
  const int half = n_videos / 2;
  b.xln = a.take((size_t)ntok * dl);
  b.score_tok = a.take((size_t)ntok);
  b.fmag_tok = a.take((size_t)ntok);
  b.scores = a.take((size_t)n_videos * T);
  b.vid_score = a.take((size_t)n_videos);
  b.sel_n = a.take((size_t)ncrops * half * c.k * dl);
  b.sel_a = a.take((size_t)ncrops * half * c.k * dl);
  b.l1 = a.take((size_t)2 * ncrops * half * c.k);
  b.dl1 = a.take((size_t)2 * ncrops * half * c.k);
  b.dscores = a.take((size_t)n_videos * T);
  b.dxln = a.take((size_t)ntok * dl);
  b.idx = reinterpret_cast<int*>(a.take((size_t)n_videos * c.k));
  b.g0 = a.take((size_t)ntok * h->max_dim);
  b.g1 = a.take((size_t)ntok * h->max_dim);
  b.g2 = a.take((size_t)ntok * h->max_dim);
  b.gw0 = a.take((size_t)ntok * h->max_wide);
  b.gw1 = a.take((size_t)ntok * h->max_wide);
  const size_t rows = (size_t)(h->max_wide > c.channels ? h->max_wide : c.channels);
  b.t1 = a.take(rows * (size_t)ntok);
  b.t2 = a.take(rows * (size_t)ntok);
  b.wt = a.take(h->max_weight);
  b.bn_part = a.take((size_t)2 * 592 * 32);
}

extern "C" int32_t vad_head_train_workspace_bytes(const vad_head_train_t* h, int32_t n_videos, int32_t ncrops, int32_t t, uint64_t* bytes) {
  if (!h || !bytes || n_videos <= 0 || ncrops <= 0 || t <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_workspace_bytes: bad argument");
  TrainArena a(nullptr);
  TrainBufs b;
  head_train_layout(h, n_videos, ncrops, t, a, b);
  *bytes = a.cur;
  return VAD_OK;
}

#define TRAIN_LAUNCHED() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return fail(VAD_ERR_CUDA, "head training kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); ++h->gemm_ctx.n_launches; } while (0)

// One training step's forward + loss + backward.  grads_dev is OVERWRITTEN with d loss / d params (same layout as params).
extern "C" int32_t vad_head_train_step(vad_head_train_t* h, const float* params_dev, float* grads_dev, float* bn_stats_dev,
                                       const float* video_dev, int32_t n_videos, int32_t ncrops, int32_t t, const float* labels_dev,
                                       const float* mask_dev, const float* loss_cfg, void* workspace_dev, uint64_t workspace_bytes,
                                       float* loss_terms_dev, float* scores_out_dev, int32_t* idx_out_dev, void* stream) {
  if (!h || !params_dev || !grads_dev || !video_dev || !labels_dev || !workspace_dev || !loss_terms_dev)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: null pointer");
  if (h->bn_floats && !bn_stats_dev) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: the BatchNorm statistics buffer is required");
  if (n_videos < 2 || (n_videos & 1) || ncrops <= 0 || t < 2) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: need an even batch (normal half, abnormal half) and t >= 2");
  const vad_head_config& c = h->cfg;
  // loss constants of the reference (src/loss/base.py:9,24, src/loss/mgfn.py:9-11) unless the caller overrides them
  const float w_smooth = loss_cfg ? loss_cfg[0] : 8e-4f, w_sparse = loss_cfg ? loss_cfg[1] : 8e-3f;
  const float alpha = loss_cfg ? loss_cfg[2] : 0.001f, margin = loss_cfg ? loss_cfg[3] : 200.f;
  const int S = n_videos * ncrops, T = t, half_v = n_videos / 2;
  const long long ntok = (long long)S * T;
  if (ntok % 32 || T % 4 || T > 64) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: tokens %% 32 == 0, t %% 4 == 0 and t <= 64 (the training bags are 32 segments)");
  if (c.k > T) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: top-k exceeds t");
  if ((half_v * ncrops) & 1) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: need an even number of crop-sequences per class");
  if ((uintptr_t)workspace_dev & 1023) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_train_step: workspace must be 1024 B aligned");
  TrainArena arena(static_cast<uint8_t*>(workspace_dev));
  TrainBufs B;
  head_train_layout(h, n_videos, ncrops, T, arena, B);
  if (workspace_bytes < arena.cur) return fail(VAD_ERR_WORKSPACE_TOO_SMALL, "head training workspace %llu < required %llu", (unsigned long long)workspace_bytes, (unsigned long long)arena.cur);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vad_head* G = &h->gemm_ctx;
  G->n_launches = 0;
  const float* P = params_dev;
  float* D = grads_dev;
  const int dl = c.dims[c.n_stages - 1];
  auto ew_grid = [](long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); };
  auto warp_grid = [](long long tokens) { return (int)((tokens * 32 + 255) / 256); };
  // column reductions over all tokens: enough token chunks for ~4 blocks per SM whatever the channel count
  auto red_chunks = [&](int channels) { int g = (148 * 4) / ((channels + 31) / 32); long long mx = (ntok + 63) / 64; if (g > mx) g = (int)mx; return g < 1 ? 1 : g; };
  auto ln_grid = [](long long tokens) { long long g = (tokens + 7) / 8; return (int)(g > 148 * 4 ? 148 * 4 : (g < 1 ? 1 : g)); };
  VAD_CUDA_CHECK(cudaMemsetAsync(D, 0, h->total_floats * sizeof(float), st));

  // ------------------------------------------------------------------------------------------ forward (train mode)
  head_split_kernel<<<ew_grid(ntok * (c.channels / 4)), 256, 0, st>>>(video_dev, ntok, c.channels, B.feat, B.mag);
  TRAIN_LAUNCHED();
  const int d0 = c.dims[0];
  HEAD_TRY(head_gemm(G, B.feat, S, T, c.channels, 3, P + h->amp_w, d0, P + h->amp_b, false, nullptr, B.x0, st));
  head_amplify_kernel<<<ew_grid(ntok * d0), 256, 0, st>>>(B.x0, B.mag, P + h->mag_w, P + h->mag_b, c.mag_ratio, S, T, d0);
  TRAIN_LAUNCHED();
  const float* x = B.x0;
  std::vector<const float*> block_in(h->blocks.size());
  size_t bi = 0;
  for (int sidx = 0; sidx < c.n_stages; ++sidx) {
    for (int b = 0; b < c.depths[sidx]; ++b, ++bi) {
      const TrainBlock& k = h->blocks[bi];
      const TrainBlockBufs& q = B.blk[bi];
      const int d = k.dim, wide = c.ff_repe * d;
      block_in[bi] = x;
      HEAD_TRY(head_gemm(G, x, S, T, d, 3, P + k.scc_w, d, P + k.scc_b, false, x, q.x1, st));
      if (k.type == VAD_HEAD_GLANCE) {
        head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(q.x1, P + k.ln_g, P + k.ln_b, c.ln_eps, ntok, d, q.y1);
        TRAIN_LAUNCHED();
        HEAD_TRY(head_gemm(G, q.y1, S, T, d, 1, P + k.qkv_w, 3 * k.inner, nullptr, false, nullptr, q.qkv, st));
        dim3 ag((T + 127) / 128, k.heads, S);
        head_attention_kernel<<<ag, 128, 0, st>>>(q.qkv, q.u, S, T, k.heads, 1.0f / sqrtf((float)c.dim_head));
        TRAIN_LAUNCHED();
      } else {
        float* rm = bn_stats_dev + k.bn_stat;
        const int bn_chunks = red_chunks(d);
        head_bn_reduce_kernel<<<dim3((d + 31) / 32, bn_chunks), dim3(32, 8), 0, st>>>(q.x1, nullptr, ntok, d, 0, nullptr, nullptr, B.bn_part);
        TRAIN_LAUNCHED();
        head_bn_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(B.bn_part, bn_chunks, ntok, d, 1e-5f, 0.1f, q.bn_mean, q.bn_invstd, rm, rm + d);
        TRAIN_LAUNCHED();
        head_bn_apply_kernel<<<ew_grid(ntok * d), 256, 0, st>>>(q.x1, q.bn_mean, q.bn_invstd, P + k.bn_g, P + k.bn_b, ntok, d, q.xb);
        TRAIN_LAUNCHED();
        HEAD_TRY(head_gemm(G, q.xb, S, T, d, 1, P + k.v_w, k.inner, nullptr, false, nullptr, q.v, st));
        head_relpos_kernel<<<ew_grid(ntok * k.inner), 256, 0, st>>>(q.v, P + k.rp_w, P + k.rp_b, q.u, S, T, k.inner, k.heads, c.local_aggr_kernel);
        TRAIN_LAUNCHED();
      }
      HEAD_TRY(head_gemm(G, q.u, S, T, k.inner, 1, P + k.out_w, d, P + k.out_b, false, q.x1, q.x2, st));
      head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(q.x2, P + k.fln_g, P + k.fln_b, c.ln_eps, ntok, d, q.y2);
      TRAIN_LAUNCHED();
      HEAD_TRY(head_gemm(G, q.y2, S, T, d, 1, P + k.in_w, wide, P + k.in_b, false, nullptr, q.zp, st));
      head_gelu_fwd_kernel<<<ew_grid(ntok * wide), 256, 0, st>>>(q.zp, q.z, ntok * wide);
      TRAIN_LAUNCHED();
      HEAD_TRY(head_gemm(G, q.z, S, T, wide, 1, P + k.o2_w, d, P + k.o2_b, false, q.x2, q.x3, st));
      x = q.x3;
    }
    const HeadInter& it = h->inter[sidx];
    if (it.present) {
      head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(x, P + it.ln_g, P + it.ln_b, c.ln_eps, ntok, it.din, B.inter_y[sidx]);
      TRAIN_LAUNCHED();
      HEAD_TRY(head_gemm(G, B.inter_y[sidx], S, T, it.din, 1, P + it.w, it.dout, P + it.b, false, nullptr, B.inter_x[sidx], st));
      x = B.inter_x[sidx];
    }
  }
  const float* x_final = x;
  head_final_kernel<<<warp_grid(ntok), 256, 0, st>>>(x_final, P + h->fin_g, P + h->fin_b, P + h->fc_w, P + h->fc_b, c.ln_eps, ntok, dl, B.xln,
                                                    B.score_tok, B.fmag_tok);
  TRAIN_LAUNCHED();
  // selection (dropout mask on the magnitudes), normal half then abnormal half, and the loss
  const size_t sel_smem = (size_t)T * sizeof(float);
  head_select_train_kernel<<<half_v, 256, sel_smem, st>>>(B.score_tok, B.fmag_tok, B.xln, mask_dev, half_v, ncrops, T, dl, c.k, B.scores, B.vid_score,
                                                          B.idx, B.sel_n, 0);
  TRAIN_LAUNCHED();
  head_select_train_kernel<<<half_v, 256, sel_smem, st>>>(B.score_tok, B.fmag_tok, B.xln, mask_dev, half_v, ncrops, T, dl, c.k, B.scores, B.vid_score,
                                                          B.idx, B.sel_a, half_v);
  TRAIN_LAUNCHED();
  head_loss_kernel<<<1, 256, 0, st>>>(B.scores, B.vid_score, labels_dev, B.sel_n, B.sel_a, half_v, ncrops, T, dl, c.k, B.l1, loss_terms_dev, w_smooth,
                                      w_sparse, alpha, margin);
  TRAIN_LAUNCHED();
  if (scores_out_dev) VAD_CUDA_CHECK(cudaMemcpyAsync(scores_out_dev, B.scores, (size_t)n_videos * T * 4, cudaMemcpyDeviceToDevice, st));
  if (idx_out_dev) VAD_CUDA_CHECK(cudaMemcpyAsync(idx_out_dev, B.idx, (size_t)n_videos * c.k * 4, cudaMemcpyDeviceToDevice, st));

  // ------------------------------------------------------------------------------------------ backward
  auto transpose = [&](const float* in, long long ld_in, long long rows, int C, int Tseq, int shift, float* out, long long ld_out, float* colsum) -> int32_t {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((C + 31) / 32));
    head_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld_in, rows, C, Tseq, shift, out, ld_out, colsum);
    TRAIN_LAUNCHED();
    return VAD_OK;
  };
  // out = conv_taps(A) . W^T (+ bias) backward: dW (+ db) always; dA = dgrad (+ dA_res) when dA_out != nullptr
  auto gemm_bwd = [&](const float* dOut, const float* A, size_t w_off, size_t b_off, bool has_bias, int taps, int Cin, int N, float* dA_out,
                      const float* dA_res) -> int32_t {
    HEAD_TRY(transpose(dOut, N, ntok, N, T, 0, B.t1, ntok, has_bias ? D + b_off : nullptr));               // dOutT [N, ntok] (+ bias gradient)
    for (int tap = 0; tap < taps; ++tap) {
      HEAD_TRY(transpose(A, Cin, ntok, Cin, T, tap - taps / 2, B.t2, ntok, nullptr));                       // (A shifted by the tap)T [Cin, ntok]
      HEAD_TRY(head_gemm(G, B.t1, 1, N, (int)ntok, 1, B.t2, Cin, nullptr, false, nullptr, D + w_off + (size_t)tap * Cin, st, taps * Cin, 0, true));
    }
    if (dA_out) {
      for (int tp = 0; tp < taps; ++tp)                                                                      // WT[cin][tp][n] = W[n][taps - 1 - tp][cin]
        HEAD_TRY(transpose(P + w_off + (size_t)(taps - 1 - tp) * Cin, (long long)taps * Cin, N, Cin, N, 0, B.wt + (size_t)tp * N, (long long)taps * N, nullptr));
      HEAD_TRY(head_gemm(G, dOut, S, T, N, taps, B.wt, Cin, nullptr, false, dA_res, dA_out, st));
    }
    return VAD_OK;
  };

  // loss -> dscores, dl1 -> dxln (scatter) -> final LayerNorm / fc
  head_loss_bwd_kernel<<<1, 256, 0, st>>>(B.scores, B.vid_score, labels_dev, B.idx, B.l1, half_v, ncrops, T, c.k, B.dscores, B.dl1, w_smooth, w_sparse,
                                          alpha, margin);
  TRAIN_LAUNCHED();
  VAD_CUDA_CHECK(cudaMemsetAsync(B.dxln, 0, (size_t)ntok * dl * 4, st));
  const size_t Rk = (size_t)ncrops * half_v * c.k;
  head_select_bwd_kernel<<<half_v, 256, 0, st>>>(B.sel_n, B.dl1, B.idx, half_v, ncrops, T, dl, c.k, 0, B.dxln);
  TRAIN_LAUNCHED();
  head_select_bwd_kernel<<<half_v, 256, 0, st>>>(B.sel_a, B.dl1 + Rk, B.idx, half_v, ncrops, T, dl, c.k, half_v, B.dxln);
  TRAIN_LAUNCHED();
  float* gx = B.g0;      // gradient w.r.t. the current block output
  float* ga = B.g1;
  float* gb = B.g2;
  head_final_bwd_kernel<<<ln_grid(ntok), 256, 0, st>>>(x_final, B.xln, B.score_tok, B.dxln, B.dscores, P + h->fin_g, P + h->fc_w, c.ln_eps, ntok, dl, ncrops,
                                                      T, gx, D + h->fin_g, D + h->fin_b, D + h->fc_w, D + h->fc_b);
  TRAIN_LAUNCHED();
  bi = h->blocks.size();
  for (int sidx = c.n_stages - 1; sidx >= 0; --sidx) {
    const HeadInter& it = h->inter[sidx];
    if (it.present) {
      // x_next = conv(LN(x3)) : gx is d x_next [ntok, dout]
      const float* x3 = B.blk[bi - 1].x3;
      HEAD_TRY(gemm_bwd(gx, B.inter_y[sidx], it.w, it.b, true, 1, it.din, it.dout, ga, nullptr));            // ga = d y
      head_mgfn_ln_bwd_kernel<<<ln_grid(ntok), 256, 0, st>>>(x3, ga, P + it.ln_g, nullptr, c.ln_eps, ntok, it.din, gb, D + it.ln_g, D + it.ln_b);
      TRAIN_LAUNCHED();
      { float* tmp = gx; gx = gb; gb = tmp; }
    }
    for (int b = c.depths[sidx] - 1; b >= 0; --b) {
      --bi;
      const TrainBlock& k = h->blocks[bi];
      const TrainBlockBufs& q = B.blk[bi];
      const int d = k.dim, wide = c.ff_repe * d;
      // ffn: x3 = out_conv(gelu(in_conv(LN(x2)))) + x2
      HEAD_TRY(gemm_bwd(gx, q.z, k.o2_w, k.o2_b, true, 1, wide, d, B.gw0, nullptr));                          // gw0 = d z
      head_gelu_bwd_kernel<<<ew_grid(ntok * wide), 256, 0, st>>>(B.gw0, q.zp, ntok * wide);                    // gw0 = d zp
      TRAIN_LAUNCHED();
      HEAD_TRY(gemm_bwd(B.gw0, q.y2, k.in_w, k.in_b, true, 1, d, wide, ga, nullptr));                          // ga = d y2
      head_mgfn_ln_bwd_kernel<<<ln_grid(ntok), 256, 0, st>>>(q.x2, ga, P + k.fln_g, gx, c.ln_eps, ntok, d, gb, D + k.fln_g, D + k.fln_b);
      TRAIN_LAUNCHED();                                                                                      // gb = d x2
      // attention: x2 = to_out(u) + x1
      HEAD_TRY(gemm_bwd(gb, q.u, k.out_w, k.out_b, true, 1, k.inner, d, B.gw0, nullptr));                     // gw0 = d u [ntok, inner]
      if (k.type == VAD_HEAD_GLANCE) {
        const size_t smem = (size_t)(4 * T * 65 + 2 * T * (T + 1)) * sizeof(float);
        if (smem > 48 * 1024) VAD_CUDA_CHECK(cudaFuncSetAttribute(head_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        head_attention_bwd_kernel<<<dim3(S, k.heads), 256, smem, st>>>(q.qkv, B.gw0, B.gw1, T, k.heads, 1.0f / sqrtf((float)c.dim_head));
        TRAIN_LAUNCHED();                                                                                    // gw1 = d qkv
        HEAD_TRY(gemm_bwd(B.gw1, q.y1, k.qkv_w, 0, false, 1, d, 3 * k.inner, ga, nullptr));                   // ga = d y1
        head_mgfn_ln_bwd_kernel<<<ln_grid(ntok), 256, 0, st>>>(q.x1, ga, P + k.ln_g, gb, c.ln_eps, ntok, d, gx, D + k.ln_g, D + k.ln_b);
        TRAIN_LAUNCHED();                                                                                    // gx = d x1
      } else {
        head_relpos_bwd_data_kernel<<<ew_grid(ntok * k.inner), 256, 0, st>>>(B.gw0, P + k.rp_w, B.gw1, S, T, k.inner, k.heads, c.local_aggr_kernel);
        TRAIN_LAUNCHED();                                                                                    // gw1 = d v
        head_relpos_bwd_weight_kernel<<<dim3((k.inner + 31) / 32, red_chunks(k.inner)), dim3(32, 8), 0, st>>>(B.gw0, q.v, D + k.rp_w, D + k.rp_b, S, T, k.inner,
                                                                                                              k.heads, c.local_aggr_kernel);
        TRAIN_LAUNCHED();
        HEAD_TRY(gemm_bwd(B.gw1, q.xb, k.v_w, 0, false, 1, d, k.inner, ga, nullptr));                         // ga = d xb
        const int bn_chunks = red_chunks(d);
        head_bn_reduce_kernel<<<dim3((d + 31) / 32, bn_chunks), dim3(32, 8), 0, st>>>(q.x1, ga, ntok, d, 1, q.bn_mean, q.bn_invstd, B.bn_part);
        TRAIN_LAUNCHED();
        head_bn_grad_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(B.bn_part, bn_chunks, d, D + k.bn_b, D + k.bn_g);
        TRAIN_LAUNCHED();                                                                                    // d beta = sum dy, d gamma = sum dy xhat
        head_bn_bwd_apply_kernel<<<ew_grid(ntok * d), 256, 0, st>>>(q.x1, ga, q.bn_mean, q.bn_invstd, P + k.bn_g, D + k.bn_b, D + k.bn_g, gb, ntok, d, gx);
        TRAIN_LAUNCHED();                                                                                    // gx = d x1
      }
      // scc: x1 = scc(x_in) + x_in
      HEAD_TRY(gemm_bwd(gx, block_in[bi], k.scc_w, k.scc_b, true, 3, d, d, ga, gx));                          // ga = d x_in
      { float* tmp = gx; gx = ga; ga = tmp; }
    }
  }
  // amplifier: x0 = to_tokens(feat) + ratio * to_mag(mag); gx = d x0 (no gradient for the input features)
  HEAD_TRY(gemm_bwd(gx, B.feat, h->amp_w, h->amp_b, true, 3, c.channels, d0, nullptr, nullptr));
  head_amplify_bwd_kernel<<<d0, 256, 0, st>>>(gx, B.mag, c.mag_ratio, S, T, d0, D + h->mag_w, D + h->mag_b);
  TRAIN_LAUNCHED();
  h->n_launches = G->n_launches;
  return VAD_OK;
}

extern "C" int32_t vad_adam_step(float* params_dev, const float* grads_dev, float* m_dev, float* v_dev, uint64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream) {
  if (!params_dev || !grads_dev || !m_dev || !v_dev || n == 0 || step < 1) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  long long g = ((long long)n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  head_adam_kernel<<<(int)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(params_dev, grads_dev, m_dev, v_dev, (long long)n, lr, beta1, beta2, eps, weight_decay,
                                                                           (float)bc1, (float)sqrt(bc2), grad_scale);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}
