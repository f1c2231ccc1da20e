// Host side of the MGFN scoring head (included at the end of vad_api.cu: same translation unit, so it
// shares the error / driver-symbol helpers).
#pragma once

#include "head_kernels.cuh"

typedef CUresult (*EncodeTiledFnH)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct HeadBlock {
  int type = 0, dim = 0, heads = 0, inner = 0;
  size_t scc_w = 0, scc_b = 0;
  size_t ln_g = 0, ln_b = 0, qkv_w = 0;         // glance
  size_t v_w = 0, v_b = 0, rp_w = 0, rp_b = 0;  // focus
  size_t out_w = 0, out_b = 0;
  size_t fln_g = 0, fln_b = 0, in_w = 0, in_b = 0, o2_w = 0, o2_b = 0;
};
struct HeadInter {
  bool present = false;
  int din = 0, dout = 0;
  size_t ln_g = 0, ln_b = 0, w = 0, b = 0;
};

struct vad_head {
  vad_head_config cfg;
  const float* params = nullptr;
  int device = 0;
  size_t amp_w = 0, amp_b = 0, mag_w = 0, mag_b = 0;
  std::vector<HeadBlock> blocks;
  std::vector<HeadInter> inter;  // one per stage (after its blocks)
  std::vector<int> stage_of_block;
  size_t fin_g = 0, fin_b = 0, fc_w = 0, fc_b = 0;
  size_t total_floats = 0;
  int max_dim = 0, max_wide = 0;
  EncodeTiledFnH encode_tiled = nullptr;
  int n_launches = 0;
};

static size_t head_take(size_t& cur, size_t n) {
  const size_t at = cur;
  cur += (n + 63) / 64 * 64;
  return at;
}

extern "C" int32_t vad_head_create(vad_head_t** out, const vad_head_config* cfg, const float* params_dev,
                                   uint64_t params_bytes, int32_t device) {
  if (!out || !cfg || !params_dev) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: null pointer");
  if (cfg->n_stages < 1 || cfg->n_stages > 4) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: n_stages must be 1..4");
  if (cfg->dim_head != 64) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: dim_head must be 64");
  if (cfg->channels % 32 || cfg->k < 1 || cfg->k > 8 || cfg->local_aggr_kernel < 1 || !(cfg->local_aggr_kernel & 1) || cfg->ff_repe < 1)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: bad config (channels %% 32, 1 <= k <= 8, odd local_aggr_kernel)");
  int32_t rc = require_sm100(device);
  if (rc != VAD_OK) return rc;
  vad_head* h = new vad_head();
  h->cfg = *cfg;
  h->params = params_dev;
  h->device = device;
  size_t cur = 0;
  const int d0 = cfg->dims[0];
  h->amp_w = head_take(cur, (size_t)d0 * 3 * cfg->channels);
  h->amp_b = head_take(cur, d0);
  h->mag_w = head_take(cur, (size_t)d0 * 3);
  h->mag_b = head_take(cur, d0);
  for (int st = 0; st < cfg->n_stages; ++st) {
    const int d = cfg->dims[st];
    if (d % 64 || d <= 0) { delete h; return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: dims must be multiples of 64"); }
    if (cfg->types[st] != VAD_HEAD_GLANCE && cfg->types[st] != VAD_HEAD_FOCUS) { delete h; return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: bad stage type"); }
    const int heads = d / cfg->dim_head;
    const int inner = heads * cfg->dim_head;
    const int wide = cfg->ff_repe * d;
    if (d > h->max_dim) h->max_dim = d;
    if (wide > h->max_wide) h->max_wide = wide;
    if (3 * inner > h->max_wide) h->max_wide = 3 * inner;
    for (int b = 0; b < cfg->depths[st]; ++b) {
      HeadBlock k;
      k.type = cfg->types[st]; k.dim = d; k.heads = heads; k.inner = inner;
      k.scc_w = head_take(cur, (size_t)d * 3 * d);
      k.scc_b = head_take(cur, d);
      if (k.type == VAD_HEAD_GLANCE) {
        k.ln_g = head_take(cur, d);
        k.ln_b = head_take(cur, d);
        k.qkv_w = head_take(cur, (size_t)3 * inner * d);
      } else {
        k.v_w = head_take(cur, (size_t)inner * d);
        k.v_b = head_take(cur, inner);
        k.rp_w = head_take(cur, (size_t)heads * cfg->local_aggr_kernel);
        k.rp_b = head_take(cur, heads);
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
      h->stage_of_block.push_back(st);
    }
    HeadInter it;
    if (st + 1 < cfg->n_stages) {
      it.present = true; it.din = d; it.dout = cfg->dims[st + 1];
      it.ln_g = head_take(cur, d);
      it.ln_b = head_take(cur, d);
      it.w = head_take(cur, (size_t)it.dout * d);
      it.b = head_take(cur, it.dout);
    }
    h->inter.push_back(it);
  }
  const int dl = cfg->dims[cfg->n_stages - 1];
  h->fin_g = head_take(cur, dl);
  h->fin_b = head_take(cur, dl);
  h->fc_w = head_take(cur, dl);
  h->fc_b = head_take(cur, 1);
  h->total_floats = cur;
  if (params_bytes != cur * sizeof(float)) {
    const unsigned long long want = cur * sizeof(float);
    delete h;
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_create: parameter blob is %llu bytes, the config implies %llu",
                (unsigned long long)params_bytes, want);
  }
  void* fn = nullptr;
  rc = driver_symbol("cuTensorMapEncodeTiled", &fn);
  if (rc != VAD_OK) { delete h; return rc; }
  h->encode_tiled = reinterpret_cast<EncodeTiledFnH>(fn);
  *out = h;
  return VAD_OK;
}

extern "C" void vad_head_destroy(vad_head_t* h) { delete h; }
extern "C" int32_t vad_head_num_launches(const vad_head_t* h) { return h ? h->n_launches : 0; }

// workspace: feat [ntok, channels] | mag [ntok] | xa, xb, y, u [ntok, max_dim] | z [ntok, max_wide]
static void head_ws_layout(const vad_head* h, long long ntok, size_t off[8], size_t* total) {
  size_t cur = 0;
  auto take = [&](size_t floats) { const size_t at = cur; cur += (floats * 4 + 1023) / 1024 * 1024; return at; };
  off[0] = take((size_t)ntok * h->cfg.channels);
  off[1] = take((size_t)ntok);
  off[2] = take((size_t)ntok * h->max_dim);
  off[3] = take((size_t)ntok * h->max_dim);
  off[4] = take((size_t)ntok * h->max_dim);
  off[5] = take((size_t)ntok * h->max_dim);
  off[6] = take((size_t)ntok * h->max_wide);
  *total = cur;
}

extern "C" int32_t vad_head_workspace_bytes(const vad_head_t* h, int32_t n_seq, int32_t t, uint64_t* bytes) {
  if (!h || !bytes || n_seq <= 0 || t <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_workspace_bytes: bad argument");
  size_t off[8], total;
  head_ws_layout(h, (long long)n_seq * t, off, &total);
  *bytes = total;
  return VAD_OK;
}

extern "C" double vad_head_flops(const vad_head_t* h, int32_t n_seq, int32_t t) {
  if (!h) return 0.0;
  const double tok = (double)n_seq * t;
  double f = 2.0 * tok * h->cfg.dims[0] * 3.0 * h->cfg.channels;
  for (const HeadBlock& k : h->blocks) {
    const double d = k.dim, wide = (double)h->cfg.ff_repe * k.dim;
    f += 2.0 * tok * d * 3.0 * d;                                              // scc
    f += 2.0 * tok * (k.type == VAD_HEAD_GLANCE ? 3.0 : 1.0) * k.inner * d;   // to_qkv / to_v
    if (k.type == VAD_HEAD_GLANCE) f += 4.0 * tok * t * k.inner;              // QK^T and PV
    f += 2.0 * tok * d * k.inner;                                              // to_out
    f += 4.0 * tok * d * wide;                                                 // ffn
  }
  for (const HeadInter& it : h->inter)
    if (it.present) f += 2.0 * tok * it.din * it.dout;
  return f;
}

// out[S*T, N] = act(conv1d_taps(A)[S*T, taps*Cin] . W[N, taps*Cin]^T + bias) (+ res)
// (ldo / ldr: row pitches of out / res in floats, 0 = N;  the weight matrix is [N][taps * Cin])
static int32_t head_gemm(vad_head* h, const float* A, int S, int T, int Cin, int taps, const float* W, int N, const float* bias,
                         bool gelu, const float* res, float* out, cudaStream_t st, int ldo = 0, int ldr = 0, bool split_k = false) {
  if (Cin % kHeadBK || N % 64) return fail(VAD_ERR_INVALID_ARGUMENT, "head gemm: Cin %% 32 / N %% 64 (Cin=%d, N=%d)", Cin, N);
  HeadGemmParams q;
  memset(&q, 0, sizeof(q));
  q.S = S; q.T = T;
  q.Tb = T <= 32 ? 32 : (T <= 64 ? 64 : 128);
  q.Sb = 128 / q.Tb;
  q.t_tiles = (T + q.Tb - 1) / q.Tb;
  q.N = N; q.Cin = Cin; q.taps = taps; q.gelu = gelu ? 1 : 0;
  q.ldo = ldo ? ldo : N; q.ldr = ldr ? ldr : N; q.bias = bias; q.res = res; q.out = out;
  const int bn = (N % 128 == 0) ? 128 : 64;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)T, (cuuint64_t)S};
    cuuint64_t gstr[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cin * 4 * (cuuint64_t)T};
    cuuint32_t box[3] = {(cuuint32_t)kHeadBK, (cuuint32_t)q.Tb, (cuuint32_t)q.Sb};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = h->encode_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)A, gdim, gstr, box, es,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "head gemm: cuTensorMapEncodeTiled(A) failed: %d", (int)cr);
    const cuuint64_t K = (cuuint64_t)taps * Cin;
    cuuint64_t wdim[2] = {K, (cuuint64_t)N};
    cuuint64_t wstr[1] = {K * 4};
    cuuint32_t wbox[2] = {(cuuint32_t)kHeadBK, (cuuint32_t)bn};
    cuuint32_t wes[2] = {1, 1};
    cr = h->encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "head gemm: cuTensorMapEncodeTiled(W) failed: %d", (int)cr);
  }
  const int s_tiles = (S + q.Sb - 1) / q.Sb;
  dim3 grid(s_tiles * q.t_tiles, N / bn);
  if (split_k) {
    // enough CTAs for two waves of the 148 SMs, at least 8 k-blocks each; `out` must be zero and there is no epilogue math
    if (bias || gelu || res) return fail(VAD_ERR_INVALID_ARGUMENT, "head gemm: split-K takes no bias / GELU / residual");
    const int total_kb = taps * Cin / kHeadBK;
    int splits = (2 * 148 + (int)(grid.x * grid.y) - 1) / (int)(grid.x * grid.y);
    if (splits > total_kb / 8) splits = total_kb / 8;
    if (splits > 1) {
      q.kb_per_split = (total_kb + splits - 1) / splits;
      grid.z = (unsigned)((total_kb + q.kb_per_split - 1) / q.kb_per_split);
    }
  }
  cudaError_t e;
  if (bn == 128) {
    static bool attr = false;
    if (!attr) { VAD_CUDA_CHECK(cudaFuncSetAttribute(head_gemm_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadGemmCfg<128>::kSmemBytes)); attr = true; }
    head_gemm_tf32_kernel<128><<<grid, 192, HeadGemmCfg<128>::kSmemBytes, st>>>(tmA, tmB, q);
  } else {
    static bool attr = false;
    if (!attr) { VAD_CUDA_CHECK(cudaFuncSetAttribute(head_gemm_tf32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadGemmCfg<64>::kSmemBytes)); attr = true; }
    head_gemm_tf32_kernel<64><<<grid, 192, HeadGemmCfg<64>::kSmemBytes, st>>>(tmA, tmB, q);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(VAD_ERR_CUDA, "head gemm launch failed: %s", cudaGetErrorString(e));
  ++h->n_launches;
  return VAD_OK;
}

#define HEAD_TRY(expr) do { int32_t rc__ = (expr); if (rc__ != VAD_OK) return rc__; } while (0)
#define HEAD_LAUNCHED() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return fail(VAD_ERR_CUDA, "head kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); ++h->n_launches; } while (0)

extern "C" int32_t vad_head_forward(vad_head_t* h, const float* video_dev, int32_t n_videos, int32_t ncrops, int32_t t,
                                    void* workspace_dev, uint64_t workspace_bytes, float* xln_dev, float* score_dev,
                                    float* fmag_dev, void* stream) {
  if (!h || !video_dev || !workspace_dev || !xln_dev || !score_dev || !fmag_dev) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_forward: null pointer");
  if (n_videos <= 0 || ncrops <= 0 || t <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_forward: bad size");
  const int S = n_videos * ncrops, T = t;
  const long long ntok = (long long)S * T;
  size_t off[8], total;
  head_ws_layout(h, ntok, off, &total);
  if (workspace_bytes < total) return fail(VAD_ERR_WORKSPACE_TOO_SMALL, "head workspace %llu < required %llu", (unsigned long long)workspace_bytes, (unsigned long long)total);
  if ((uintptr_t)workspace_dev & 1023) return fail(VAD_ERR_INVALID_ARGUMENT, "head workspace must be 1024 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  float* feat = reinterpret_cast<float*>(ws + off[0]);
  float* mag = reinterpret_cast<float*>(ws + off[1]);
  float* xa = reinterpret_cast<float*>(ws + off[2]);
  float* xb = reinterpret_cast<float*>(ws + off[3]);
  float* y = reinterpret_cast<float*>(ws + off[4]);
  float* u = reinterpret_cast<float*>(ws + off[5]);
  float* z = reinterpret_cast<float*>(ws + off[6]);
  const float* P = h->params;
  const vad_head_config& c = h->cfg;
  h->n_launches = 0;
  auto ew_grid = [](long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); };
  auto warp_grid = [](long long tokens) { return (int)((tokens * 32 + 255) / 256); };

  head_split_kernel<<<ew_grid(ntok * (c.channels / 4)), 256, 0, st>>>(video_dev, ntok, c.channels, feat, mag);
  HEAD_LAUNCHED();
  // amplifier: x = Conv1d(channels -> d0, k3)(feat) + mag_ratio * Conv1d(1 -> d0, k3)(mag)
  const int d0 = c.dims[0];
  HEAD_TRY(head_gemm(h, feat, S, T, c.channels, 3, P + h->amp_w, d0, P + h->amp_b, false, nullptr, xa, st));
  head_amplify_kernel<<<ew_grid(ntok * d0), 256, 0, st>>>(xa, mag, P + h->mag_w, P + h->mag_b, c.mag_ratio, S, T, d0);
  HEAD_LAUNCHED();
  float* x = xa;
  float* xo = xb;
  size_t bi = 0;
  for (int sidx = 0; sidx < c.n_stages; ++sidx) {
    for (int b = 0; b < c.depths[sidx]; ++b, ++bi) {
      const HeadBlock& k = h->blocks[bi];
      const int d = k.dim, wide = c.ff_repe * d;
      // x = scc(x) + x
      HEAD_TRY(head_gemm(h, x, S, T, d, 3, P + k.scc_w, d, P + k.scc_b, false, x, xo, st));
      { float* tmp = x; x = xo; xo = tmp; }
      // x = attention(x) + x
      if (k.type == VAD_HEAD_GLANCE) {
        head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(x, P + k.ln_g, P + k.ln_b, c.ln_eps, ntok, d, y);
        HEAD_LAUNCHED();
        HEAD_TRY(head_gemm(h, y, S, T, d, 1, P + k.qkv_w, 3 * k.inner, nullptr, false, nullptr, z, st));
        dim3 ag((T + 127) / 128, k.heads, S);
        head_attention_kernel<<<ag, 128, 0, st>>>(z, u, S, T, k.heads, 1.0f / sqrtf((float)c.dim_head));
        HEAD_LAUNCHED();
      } else {
        HEAD_TRY(head_gemm(h, x, S, T, d, 1, P + k.v_w, k.inner, P + k.v_b, false, nullptr, z, st));
        head_relpos_kernel<<<ew_grid(ntok * k.inner), 256, 0, st>>>(z, P + k.rp_w, P + k.rp_b, u, S, T, k.inner, k.heads, c.local_aggr_kernel);
        HEAD_LAUNCHED();
      }
      HEAD_TRY(head_gemm(h, u, S, T, k.inner, 1, P + k.out_w, d, P + k.out_b, false, x, xo, st));
      { float* tmp = x; x = xo; xo = tmp; }
      // x = ffn(x) + x
      head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(x, P + k.fln_g, P + k.fln_b, c.ln_eps, ntok, d, y);
      HEAD_LAUNCHED();
      HEAD_TRY(head_gemm(h, y, S, T, d, 1, P + k.in_w, wide, P + k.in_b, true, nullptr, z, st));
      HEAD_TRY(head_gemm(h, z, S, T, wide, 1, P + k.o2_w, d, P + k.o2_b, false, x, xo, st));
      { float* tmp = x; x = xo; xo = tmp; }
    }
    const HeadInter& it = h->inter[sidx];
    if (it.present) {
      head_mgfn_layernorm_kernel<<<warp_grid(ntok), 256, 0, st>>>(x, P + it.ln_g, P + it.ln_b, c.ln_eps, ntok, it.din, y);
      HEAD_LAUNCHED();
      HEAD_TRY(head_gemm(h, y, S, T, it.din, 1, P + it.w, it.dout, P + it.b, false, nullptr, xo, st));
      { float* tmp = x; x = xo; xo = tmp; }
    }
  }
  const int dl = c.dims[c.n_stages - 1];
  head_final_kernel<<<warp_grid(ntok), 256, 0, st>>>(x, P + h->fin_g, P + h->fin_b, P + h->fc_w, P + h->fc_b, c.ln_eps, ntok, dl,
                                                    xln_dev, score_dev, fmag_dev);
  HEAD_LAUNCHED();
  return VAD_OK;
}

extern "C" int32_t vad_head_select(const vad_head_t* hc, const float* xln_dev, const float* score_dev, const float* fmag_dev,
                                   int32_t n_videos, int32_t ncrops, int32_t t, int32_t video_off, int32_t n_sel,
                                   float* scores_dev, float* vid_score_dev, int32_t* idx_dev, float* sel_dev, void* stream) {
  vad_head* h = const_cast<vad_head*>(hc);
  if (!h || !xln_dev || !score_dev || !fmag_dev || !scores_dev || !vid_score_dev || !idx_dev || !sel_dev)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_select: null pointer");
  if (n_videos <= 0 || ncrops <= 0 || t <= 0 || video_off < 0 || n_sel <= 0 || video_off + n_sel > n_videos)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_select: bad range");
  if (h->cfg.k > t) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_select: top-k %d exceeds the %d snippets", h->cfg.k, t);
  const int dl = h->cfg.dims[h->cfg.n_stages - 1];
  const size_t smem = (size_t)t * sizeof(float);
  if (smem > 48 * 1024) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_select: more than 12288 snippets per video");
  head_select_kernel<<<n_sel, 256, smem, static_cast<cudaStream_t>(stream)>>>(score_dev, fmag_dev, xln_dev, n_sel, ncrops, t, dl, h->cfg.k,
                                                                              scores_dev, vid_score_dev, idx_dev, sel_dev, video_off);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}

extern "C" int32_t vad_head_loss(const vad_head_t* h, const float* scores_dev, const float* vid_score_dev, const float* labels_dev,
                                 const float* sel_normal_dev, const float* sel_abnormal_dev, int32_t n_half, int32_t ncrops,
                                 int32_t t, float* scratch_dev, float* out_dev, void* stream) {
  if (!h || !scores_dev || !vid_score_dev || !labels_dev || !sel_normal_dev || !sel_abnormal_dev || !scratch_dev || !out_dev)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_loss: null pointer");
  if (n_half <= 0 || ncrops <= 0 || t < 2 || ((n_half * ncrops) & 1))
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_head_loss: need t >= 2 and an even number of crop-sequences per class");
  const int dl = h->cfg.dims[h->cfg.n_stages - 1];
  head_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores_dev, vid_score_dev, labels_dev, sel_normal_dev,
                                                                     sel_abnormal_dev, n_half, ncrops, t, dl, h->cfg.k, scratch_dev, out_dev);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}
