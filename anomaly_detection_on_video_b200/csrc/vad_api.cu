// Host side of libvad_b200.so: the C ABI declared in include/vad_b200.h.
// Shape inference, tensor-map encoding and kernel launches.  No device allocations happen here
// except the few KB of resampling tables a preprocessing handle owns.
#include "../../include/vad_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "aux_kernels.cuh"
#include "conv_umma.cuh"
#include "stem_umma.cuh"
#include "conv_thalo.cuh"
#include "conv_s3x3.cuh"
#include "conv_pair.cuh"
#include "stem_pair.cuh"
#include "conv_tail.cuh"

using namespace vad;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define VAD_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(VAD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                       \
  } while (0)

static bool g_pdl = true;  // VAD_NO_PDL=1 switches programmatic dependent launch off (read at plan creation)

extern "C" const char* vad_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t vad_abi_version(void) { return VAD_ABI_VERSION; }

// ------------------------------------------------------------------------------------ driver API
// cuTensorMapEncode* live in libcuda; resolve them through the runtime so the library has no
// link-time dependency on the driver (it must load on a GPU-less build box).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int32_t driver_symbol(const char* name, void** fn) {
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || *fn == nullptr)
    return fail(VAD_ERR_DRIVER_SYMBOL, "cannot resolve driver symbol %s (%s)", name, cudaGetErrorString(e));
  return VAD_OK;
}

static int32_t require_sm100(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(VAD_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(VAD_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  VAD_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(VAD_ERR_UNSUPPORTED_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a only", device,
                prop.major, prop.minor);
  return VAD_OK;
}

// ------------------------------------------------------------------------------------ plan
struct SlotInfo {
  int T = 0, H = 0, W = 0, C = 0;
  uint64_t offset = UINT64_MAX, bytes = 0;
  bool defined = false;
};

struct OpRuntime {
  ConvParams cp;
  PoolParams pp;
  CUtensorMap tmA, tmB, tmR, tmO;
  CUtensorMap tmO2;         // third output of a fused sibling conv (staged epilogue)
  bool epi = false;   // residual prefetched / output stored by TMA through shared memory
  int res_c = 0, dst_c = 0;
  int bn = 0;
  int pf[3] = {0, 0, 0};  // front padding (t, h, w) after resolving VAD_FLAG_CONV_SAME
  int pb[3] = {0, 0, 0};  // back padding
  int bk = 64;
  int kps = 1;        // k-blocks per pipeline stage (2: eight MMAs per barrier round trip for BN <= 128)
  int grid = 0;
  bool fold = false;
  bool stem = false;  // dedicated stem kernel (spatial tiles, resident weights)
  StemParams sp;
  CUtensorMap tmE, tmOdd, tmW, tmSO;
  CUtensorMap tmWh;         // stem_pair: the weight matrix with 32-row boxes (one channel half of a tap)
  bool stem_pair = false;   // stem on CTA pairs with half of the weights resident per CTA (taps that do not fit one CTA: 7x7x7)
  int stem_smem = 0;
  bool pair_epi = false; // CTA-pair kernel with the staged residual epilogue (conv3 of layer 4: cout % 256 == 0, >= 8 k-blocks)
  CUtensorMap tmBh;      // pair kernels: weight map with (64 x BN/2)-row boxes
  bool pair = false;     // CTA-pair kernel (tcgen05 cta_group::2, 256 x BN tiles): BN = 256 layers with TMA operands and the direct epilogue
  bool pool_tp = false;  // 1x1x1 residual conv with maxpool2 (2,1,1)/(2,1,1) fused into its staged epilogue
  bool s3 = false;       // spatial (1,3,3) 64 -> 64 kernel (conv_s3x3.cuh)
  S3x3Params s3p;
  bool thalo = false;    // temporal-halo kernel for (3,1,1) convs (conv_thalo.cuh)
  ThaloParams tp;
  bool stem_mf = false;  // multi-frame (input-frame stationary) stem kernel
  int stem_ti = 0, stem_ti_max = 0;
  int a_mode = 0;
  // bottleneck-tail fusion (conv_tail.cuh): this (1,3,3) 64 -> 64 op also runs the following 1x1x1 64 -> 256 conv3
  // (+ residual: tail = 1; + the block's 1x1x1 downsample of X folded into conv3's contraction: tail = 2)
  int tail = 0;
  int tail_c3 = -1, tail_ds = -1;  // op indices of the fused conv3 / downsample
  bool skip = false;               // this op runs inside an earlier op's launch
  TailParams tlp;
  CUtensorMap tmW3, tmX;
  int avg_P = 0, avg_C = 0, avg_HW = 0, avg_kt = 0;
  // for tensor-map encoding
  int Ci = 0, Ti = 0, Hi = 0, Wi = 0;
  int K_pad = 0;
};

struct vad_plan {
  std::vector<vad_op_desc> ops;
  int n_slots = 0;
  const uint8_t* params = nullptr;
  uint64_t params_bytes = 0;
  int in_pad_left = 0;
  int in_channels = 0;
  int device = 0;
  int sm_count = 148;
  bool stem_v3 = false;      // VAD_STEM_V3=1: one-output-frame-per-tile stem kernel instead of the multi-frame one
  bool stem_generic = false; // VAD_STEM_GENERIC=1: run the stem through the generic implicit-GEMM kernel
  bool stem_no_pair = false; // VAD_STEM_NO_PAIR=1: a stem whose taps exceed one CTA streams its weights instead of running on CTA pairs
  bool no_epi = false;       // VAD_NO_EPI=1: residual layers use the direct (register) epilogue
  bool no_tail = false;      // VAD_NO_TAIL=1: no conv2 -> conv3 (+ downsample) fusion in layer1 (A/B and bit-identity tests)
  std::vector<std::pair<int, void*>> fold_bufs;  // (op index, 64 KB device buffer): BN-scaled [W3 | Wd] of a tail = 2 op
  bool fold_pending = false;                     // fold_bufs must be (re)filled on the next bind
  bool no_s3 = false;        // VAD_NO_S3X3=1: layer1's (1,3,3) convs through the generic im2col kernel
  bool no_thalo = false;     // VAD_NO_THALO=1: (3,1,1) convs through the generic im2col kernel
  bool no_pair_split = false; // VAD_NO_PAIR_SPLIT=1: no half-width tail items in the CTA-pair kernel
  bool no_bk32 = false;      // VAD_NO_BK32=1: Cin % 64 == 32 layers through the gather producer (as before the 32-wide TMA path)
  int pair_mode = 1;         // VAD_PAIR=0: never use the CTA-pair kernel; 1 (default): for the long-K layers without residual
  int pair_epi_min_kb = 8;   // VAD_PAIR_EPI_MIN_KB: same for residual layers (staged epilogue; measured: K = 512 gains, K = 256 loses); 0 = off
  int pair_min_kb = 12;      // VAD_PAIR_MIN_KB: fewest 64-wide k-blocks for which a layer goes to the CTA-pair kernel
  int kps_override = 0;      // VAD_KPS=1|2: force k-blocks per stage (tuning only)
  int batch = 0, T = 0, H = 0, W = 0;
  bool configured = false;
  std::vector<SlotInfo> slots;
  uint64_t ws_bytes = 0;
  std::vector<OpRuntime> rt;
  const void* bound_x = nullptr;
  void* bound_ws = nullptr;
  // CUDA graph of one forward (small batches: ~55 launches of 5-15 us each are fixed cost; VAD_GRAPH=0|1 overrides the
  // batch <= 32 default).  Captured on an internal stream at the second forward of a binding (the first one sets the
  // kernels' shared-memory attributes), launched into the caller's stream; a re-bind or re-configure drops it.
  int direct_runs = 0;            // forwards launched op by op since the last bind
  cudaGraphExec_t graph_exec = nullptr;
  float* graph_feat = nullptr;
  int feat_misses = 0;            // forwards whose feature pointer differed from the captured one
  cudaStream_t cap_stream = nullptr;
  bool graph_failed = false;
  double flops = 0.0;
  int feat_c = 0;
  EncodeTiledFn encode_tiled = nullptr;
  EncodeIm2colFn encode_im2col = nullptr;
  int driver_version = 0;
  // optional per-op timing (vad_plan_profile_begin/end): events bracket every launch
  bool profiling = false;
  int prof_first = 0, prof_count = -1;  // ops bracketed by events (vad_plan_profile_select); -1: all
  std::vector<cudaEvent_t> ev_pool;   // recycled events
  std::vector<cudaEvent_t> ev_used;   // (n_ops + 1) events per profiled forward, in order
  std::vector<double> op_flops;       // useful FLOPs per op at the configured size
  std::vector<double> op_bytes;       // algorithmic bytes per op (inputs read once + outputs written once)
  std::vector<double> prof_flops, prof_bytes;  // totals over the profiled launches
  ~vad_plan() {
    for (auto& fb : fold_bufs) cudaFree(fb.second);
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_used) cudaEventDestroy(e);
  }
};

static inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

extern "C" int32_t vad_plan_create(vad_plan_t** plan, const vad_op_desc* ops, int32_t n_ops, int32_t n_slots,
                                   const void* params_dev, uint64_t params_bytes, int32_t in_channels,
                                   int32_t in_pad_left, int32_t device) {
  if (!plan || !ops || n_ops <= 0 || n_slots <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_create: null/empty argument");
  if (in_pad_left < 0 || in_pad_left > 8) return fail(VAD_ERR_INVALID_ARGUMENT, "in_pad_left must be in [0,8]");
  if (in_channels < 0 || in_channels % 8) return fail(VAD_ERR_INVALID_ARGUMENT, "in_channels must be 0 (stem layout) or a multiple of 8");
  int32_t rc = require_sm100(device);
  if (rc != VAD_OK) return rc;
  for (int i = 0; i < n_ops; ++i) {
    const vad_op_desc& d = ops[i];
    if (d.kind < VAD_OP_CONV || d.kind > VAD_OP_AVGPOOL) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: bad kind %d", i, d.kind);
    if (d.src < 0 || d.src >= n_slots) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: src slot %d out of range", i, d.src);
    if (d.kind != VAD_OP_AVGPOOL && (d.dst <= 0 || d.dst >= n_slots))
      return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: dst slot %d out of range (slot 0 is the read-only input)", i, d.dst);
    if (d.kind != VAD_OP_AVGPOOL && d.dst == d.src) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: in-place ops are not supported", i);
    if (d.kind == VAD_OP_CONV) {
      const bool fold = d.flags & VAD_FLAG_STEM_FOLD_W;
      if (fold ? (d.cin != 4) : (d.cin % 8 != 0 || d.cin <= 0))
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: cin=%d must be a multiple of 8 (4 with STEM_FOLD_W)", i, d.cin);
      if (d.cout <= 0 || d.cout % 8) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: cout=%d must be a multiple of 8", i, d.cout);
      if (d.dst_c_off % 8) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: dst_c_off must be a multiple of 8", i);
      if (d.res >= n_slots) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: res slot out of range", i);
      if (d.w_off % 128 || d.scale_off % 16 || d.shift_off % 16)
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: parameter offsets must be aligned (weights 128 B, scale/shift 16 B)", i);
      if (d.kt < 1 || d.kh < 1 || d.kw < 1 || d.st < 1 || d.sh < 1 || d.sw < 1 || d.pt < 0 || d.ph < 0 || d.pw < 0)
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: bad kernel/stride/pad", i);
      if ((d.flags & VAD_FLAG_POOL_T2) && !fold &&
          !(d.kt == 1 && d.kh == 1 && d.kw == 1 && d.st == 1 && d.sh == 1 && d.sw == 1 && d.res >= 0 && d.cout % 128 == 0 && d.cin % 64 == 0))
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: POOL_T2 needs the stem (STEM_FOLD_W) or a 1x1x1 residual conv with cout %% 128 == 0", i);
      if (fold && in_channels != 0) return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: STEM_FOLD_W needs the stem input layout (in_channels == 0)", i);
      if (fold && (d.kw > 8 || d.src != 0 || (d.sw & 1)))
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: STEM_FOLD_W needs src=0, kw<=8, even sw", i);
      if (fold && !(d.flags & VAD_FLAG_CONV_SAME) && (in_pad_left < d.pw || ((in_pad_left - d.pw) & 1)))
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %d: STEM_FOLD_W needs in_pad_left - pw even and >= 0", i);
    }
  }
  vad_plan* p = new vad_plan();
  p->ops.assign(ops, ops + n_ops);
  p->n_slots = n_slots;
  p->params = static_cast<const uint8_t*>(params_dev);
  p->params_bytes = params_bytes;
  p->in_pad_left = in_pad_left;
  p->in_channels = in_channels;
  p->device = device;
  void* fn = nullptr;
  rc = driver_symbol("cuTensorMapEncodeTiled", &fn);
  if (rc != VAD_OK) { delete p; return rc; }
  p->encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  rc = driver_symbol("cuTensorMapEncodeIm2col", &fn);
  if (rc != VAD_OK) { delete p; return rc; }
  p->encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&p->driver_version);
  cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (p->sm_count <= 0) p->sm_count = 148;
  { const char* k = getenv("VAD_STEM_V3"); p->stem_v3 = k && k[0] == '1'; }
  const char* sgen = getenv("VAD_STEM_GENERIC");
  p->stem_generic = sgen && sgen[0] == '1';
  { const char* k = getenv("VAD_NO_S3X3"); p->no_s3 = k && k[0] == '1'; }
  { const char* k = getenv("VAD_NO_TAIL"); p->no_tail = k && k[0] == '1'; }
  { const char* k = getenv("VAD_NO_THALO"); p->no_thalo = k && k[0] == '1'; }
  { const char* k = getenv("VAD_PAIR"); p->pair_mode = k ? atoi(k) : 1; }
  { const char* k = getenv("VAD_NO_BK32"); p->no_bk32 = k && k[0] == '1'; }
  { const char* k = getenv("VAD_STEM_NO_PAIR"); p->stem_no_pair = k && k[0] == '1'; }
  { const char* k = getenv("VAD_NO_PAIR_SPLIT"); p->no_pair_split = k && k[0] == '1'; }
  { const char* k = getenv("VAD_PAIR_MIN_KB"); p->pair_min_kb = k ? atoi(k) : 12; }
  { const char* k = getenv("VAD_PAIR_EPI_MIN_KB"); p->pair_epi_min_kb = k ? atoi(k) : 8; }
  { const char* k = getenv("VAD_NO_PDL"); g_pdl = !(k && k[0] == '1'); }
  { const char* k = getenv("VAD_KPS"); p->kps_override = k ? atoi(k) : 0; }
  const char* ne = getenv("VAD_NO_EPI");
  p->no_epi = ne && ne[0] == '1';
  *plan = p;
  return VAD_OK;
}

extern "C" void vad_plan_destroy(vad_plan_t* plan) {
  if (!plan) return;
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
  delete plan;
}

#include "plan_configure.cuh"

extern "C" int32_t vad_plan_slot_info(const vad_plan_t* p, int32_t slot, int32_t dims[4], uint64_t* offset,
                                      uint64_t* bytes) {
  if (!p || !p->configured) return fail(VAD_ERR_NOT_CONFIGURED, "plan is not configured");
  if (slot < 0 || slot >= p->n_slots || !p->slots[slot].defined) return fail(VAD_ERR_INVALID_ARGUMENT, "slot %d undefined", slot);
  const SlotInfo& s = p->slots[slot];
  if (dims) { dims[0] = s.T; dims[1] = s.H; dims[2] = s.W; dims[3] = s.C; }
  if (offset) *offset = s.offset;
  if (bytes) *bytes = s.bytes;
  return VAD_OK;
}

extern "C" int32_t vad_plan_num_launches(const vad_plan_t* p) {
  if (!p) return 0;
  int32_t n = (int32_t)p->ops.size();
  if (p->configured)
    for (const OpRuntime& r : p->rt) n -= r.skip ? 1 : 0;  // ops that run inside a fused launch
  return n;
}
extern "C" double vad_plan_flops(const vad_plan_t* p) { return (p && p->configured) ? p->flops : 0.0; }

// Slot shapes change while the op list runs (slots are reused), so shapes are re-derived here in
// op order; only pointers and tensor maps are (re)bound.
#include "plan_bind.cuh"
#include "plan_run.cuh"

extern "C" int32_t vad_plan_profile_select(vad_plan_t* p, int32_t first_op, int32_t n_ops) {
  if (!p) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_select: null plan");
  if (p->profiling) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_select: profiling is running");
  if (n_ops < 0) { p->prof_first = 0; p->prof_count = -1; return VAD_OK; }
  if (first_op < 0 || n_ops == 0 || first_op + n_ops > (int32_t)p->ops.size())
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_select: ops [%d, %d) out of range", first_op, first_op + n_ops);
  p->prof_first = first_op; p->prof_count = n_ops;
  return VAD_OK;
}

extern "C" int32_t vad_plan_profile_begin(vad_plan_t* p) {
  if (!p) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_begin: null plan");
  for (cudaEvent_t e : p->ev_used) p->ev_pool.push_back(e);
  p->ev_used.clear();
  p->prof_flops.assign(p->ops.size(), 0.0);
  p->prof_bytes.assign(p->ops.size(), 0.0);
  p->profiling = true;
  return VAD_OK;
}

extern "C" int32_t vad_plan_profile_end(vad_plan_t* p, int32_t n_ops, double* op_ms_sum, int32_t* op_calls,
                                        double* op_flops, double* op_bytes) {
  if (!p || !p->profiling) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_end: profiling was not started");
  p->profiling = false;
  const size_t n = p->ops.size();
  if (n_ops != (int32_t)n) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_profile_end: expected %zu ops", n);
  for (size_t i = 0; i < n; ++i) {
    if (op_ms_sum) op_ms_sum[i] = 0.0;
    if (op_calls) op_calls[i] = 0;
    if (op_flops) op_flops[i] = p->prof_flops[i];
    if (op_bytes) op_bytes[i] = p->prof_bytes[i];
  }
  if (p->ev_used.empty()) return VAD_OK;
  VAD_CUDA_CHECK(cudaEventSynchronize(p->ev_used.back()));
  const size_t first = p->prof_count < 0 ? 0 : (size_t)p->prof_first;
  const size_t cnt = p->prof_count < 0 ? n : (size_t)p->prof_count;
  const size_t per = cnt + 1;
  for (size_t f = 0; f + per <= p->ev_used.size(); f += per) {
    for (size_t i = 0; i < cnt; ++i) {
      float ms = 0.f;
      VAD_CUDA_CHECK(cudaEventElapsedTime(&ms, p->ev_used[f + i], p->ev_used[f + i + 1]));
      if (op_ms_sum) op_ms_sum[first + i] += ms;
      if (op_calls) op_calls[first + i] += 1;
    }
  }
  for (cudaEvent_t e : p->ev_used) p->ev_pool.push_back(e);
  p->ev_used.clear();
  return VAD_OK;
}

extern "C" int32_t vad_ingest_ncthw_f32(const float* x_dev, int32_t batch, int32_t t, int32_t h, int32_t w,
                                        int32_t pad_left, void* out_dev, void* stream) {
  if (!x_dev || !out_dev || batch <= 0 || t <= 0 || h <= 0 || w <= 0 || pad_left < 0 || pad_left > 8)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_ingest_ncthw_f32: bad argument");
  const long long total = (long long)batch * t * h * (w + 8);
  ingest_ncthw_f32_kernel<<<grid_for(total, 256, 148 * 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_dev, batch, t, h, w, pad_left, static_cast<uint2*>(out_dev));
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}

// ------------------------------------------------------------------------------------ preprocessing, segment mean, magnitude
#include "aux_api.cuh"

// ------------------------------------------------------------------------------------ MGFN scoring head
#include "head_api.cuh"
#include "head_train_api.cuh"

// ------------------------------------------------------------------------------------ TF32 precision mode
#include "tf32_api.cuh"
