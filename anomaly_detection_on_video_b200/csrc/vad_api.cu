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

static int pool_out_same(int in, int s) { return (in + s - 1) / s; }

extern "C" int32_t vad_plan_configure(vad_plan_t* p, int32_t batch, int32_t t, int32_t h, int32_t w,
                                      uint64_t* workspace_bytes) {
  if (!p) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_configure: null plan");
  if (batch <= 0 || t <= 0 || h <= 0 || w <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_configure: bad size");
  p->configured = false;
  p->slots.assign(p->n_slots, SlotInfo());
  p->rt.assign(p->ops.size(), OpRuntime());
  p->op_flops.assign(p->ops.size(), 0.0);
  p->op_bytes.assign(p->ops.size(), 0.0);
  p->flops = 0.0;
  p->feat_c = 0;
  SlotInfo& s0 = p->slots[0];
  s0.T = t; s0.H = h; s0.defined = true;
  if (p->in_channels == 0) { s0.W = w + 8; s0.C = 4; } else { s0.W = w; s0.C = p->in_channels; }
  s0.bytes = (uint64_t)batch * t * h * s0.W * s0.C * 2;

  for (size_t i = 0; i < p->ops.size(); ++i) {
    const vad_op_desc& d = p->ops[i];
    const SlotInfo src = p->slots[d.src];
    if (!src.defined) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu reads slot %d before it is written", i, d.src);
    OpRuntime& r = p->rt[i];
    int To, Ho, Wo, Cdst;
    if (d.kind == VAD_OP_CONV) {
      const bool fold = d.flags & VAD_FLAG_STEM_FOLD_W;
      const int Wi = fold ? src.W - 8 : src.W;
      if (src.C != d.cin) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: cin=%d but slot %d has C=%d", i, d.cin, d.src, src.C);
      if (d.flags & VAD_FLAG_CONV_SAME) {
        // TF "SAME" (Unit3D.compute_pad of the public I3D port): out = ceil(in / stride), the padding that needs is
        // split front = total / 2, back = total - front (asymmetric for the 7x7x7 / 2 stem: 2 in front, 3 behind)
        const int in3[3] = {src.T, src.H, Wi}, k3[3] = {d.kt, d.kh, d.kw}, s3[3] = {d.st, d.sh, d.sw};
        int out3[3];
        for (int a = 0; a < 3; ++a) {
          out3[a] = (in3[a] + s3[a] - 1) / s3[a];
          int tot = (out3[a] - 1) * s3[a] + k3[a] - in3[a];
          if (tot < 0) tot = 0;
          r.pf[a] = tot / 2;
          r.pb[a] = tot - tot / 2;
        }
        To = out3[0]; Ho = out3[1]; Wo = out3[2];
        if (fold && (p->in_pad_left < r.pf[2] || ((p->in_pad_left - r.pf[2]) & 1)))
          return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: STEM_FOLD_W needs in_pad_left - (SAME front pad %d) even and >= 0", i, r.pf[2]);
      } else {
        r.pf[0] = r.pb[0] = d.pt; r.pf[1] = r.pb[1] = d.ph; r.pf[2] = r.pb[2] = d.pw;
        To = (src.T + 2 * d.pt - d.kt) / d.st + 1;
        Ho = (src.H + 2 * d.ph - d.kh) / d.sh + 1;
        Wo = (Wi + 2 * d.pw - d.kw) / d.sw + 1;
      }
      const int pt = r.pf[0], ph = r.pf[1], pw = r.pf[2];
      const bool sym_pad = r.pf[0] == r.pb[0] && r.pf[1] == r.pb[1] && r.pf[2] == r.pb[2];
      if (To <= 0 || Ho <= 0 || Wo <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: empty output", i);
      Cdst = d.dst_c_total ? d.dst_c_total : d.cout;
      if (d.dst_c_off + d.cout > Cdst) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: channel slice exceeds dst_c_total", i);
      const long long M = (long long)batch * To * Ho * Wo;
      if (M > 0x7fffffffLL - 256) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: too many output pixels (%lld)", i, M);
      if (fold && d.sw * (Wo - 1) - pw + p->in_pad_left + 7 > src.W - 1)
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: folded stem window overruns the padded row", i);
      ConvParams& c = r.cp;
      memset(&c, 0, sizeof(c));
      c.M = (int)M; c.N = d.cout;
      c.To = To; c.Ho = Ho; c.Wo = Wo;
      c.Ti = src.T; c.Hi = src.H;
      c.kt = d.kt; c.kh = d.kh; c.st = d.st; c.sh = d.sh; c.pt = pt; c.ph = ph;
      if (fold) {
        c.Wi = Wo; c.kw = 1; c.sw = 1; c.pw = 0;
        c.cin_eff = 32; c.ntaps = d.kt * d.kh;
        c.sW = d.sw * 4; c.sH = (long long)src.W * 4; c.sT = c.sH * src.H; c.sN = c.sT * src.T;
      } else {
        c.Wi = Wi; c.kw = d.kw; c.sw = d.sw; c.pw = pw;
        c.cin_eff = d.cin; c.ntaps = d.kt * d.kh * d.kw;
        c.sW = d.cin; c.sH = (long long)Wi * d.cin; c.sT = c.sH * src.H; c.sN = c.sT * src.T;
      }
      const int K = c.ntaps * c.cin_eff;
      r.K_pad = (int)align_up(K, 64);  // packed weight rows are padded to 64 whatever BK the kernel uses
      c.relu = (d.flags & VAD_FLAG_RELU) ? 1 : 0;
      c.ldo = Cdst;
      r.dst_c = Cdst;
      const bool unit = d.kt == 1 && d.kh == 1 && d.kw == 1 && d.st == 1 && d.sh == 1 && d.sw == 1 && !pt && !ph && !pw;
      const bool tma_geom_ok = r.pb[0] <= 15 && r.pb[1] <= 15 && r.pb[2] <= 15 && d.kt <= 16 && d.kh <= 16 && d.kw <= 16 &&
                               d.st <= 8 && d.sh <= 8 && d.sw <= 8;
      r.bk = 64;
      // Cin % 64 == 32 (Inception's 96 / 160 / 480-channel inputs, 32-channel 5x5 branches): TMA operands with 32-wide
      // k-blocks (64-byte rows, SWIZZLE_64B) -- direct epilogue only; anything else that is not a multiple of 64: gather
      const bool epi_wanted = !p->no_epi && (d.res >= 0 || (d.cout >= 128 && K <= 256 && d.cout >= 2 * K));
      // ... and Cin % 32 == 16 (16 / 48 / 112 / 144 / 528 channels) with 16-wide ones (32-byte rows, SWIZZLE_32B, one MMA each)
      const int sub_k = (fold || epi_wanted || p->no_bk32) ? 0 : ((d.cin % 64) == 32 ? 32 : ((d.cin % 32) == 16 ? 16 : 0));
      const bool half_k = sub_k != 0;
      if ((d.flags & VAD_FLAG_FORCE_GATHER) || !tma_geom_ok || (!fold && (d.cin % 64) && !half_k))
        r.a_mode = A_GATHER;
      else if (half_k) {
        r.a_mode = unit ? A_TMA_2D : A_TMA_IM2COL;
        r.bk = sub_k;
      } else if (fold) {
        r.a_mode = A_TMA_IM2COL;  // im2col over the overlapping 8-pixel window view: 32 bf16 = 64-byte rows
        r.bk = 32;
      } else if (unit)
        r.a_mode = A_TMA_2D;
      else
        r.a_mode = A_TMA_IM2COL;
      c.a_mode = r.a_mode;
      c.num_kb = r.bk == 64 ? r.K_pad / 64 : (K + r.bk - 1) / r.bk;
      // staged epilogue (two 128 x BN tiles in smem, TMA store; residual prefetched by TMA): residual layers, and
      // output-dominated small-K layers without one (K <= 256, cout >= 2K: the first downsample projections)
      r.epi = !p->no_epi && (d.res >= 0 || (d.cout >= 128 && K <= 256 && d.cout >= 2 * K));
      r.bn = (d.cout > 128 && !r.epi && r.bk == 64) ? 256 : (d.cout > 64 ? 128 : 64);
      r.pair_epi = r.epi && p->pair_mode > 0 && p->pair_epi_min_kb > 0 && d.res >= 0 && r.a_mode != A_GATHER && r.bk == 64 && d.cout % 256 == 0 &&
                   !(d.flags & VAD_FLAG_POOL_T2) && (p->sm_count % 2) == 0 && c.num_kb >= p->pair_epi_min_kb && M > kBlockM;
      if (r.pair_epi) r.bn = 256;
      r.kps = (r.a_mode != A_GATHER && r.bk == 64 && r.bn <= 128 && c.num_kb >= 2 && !r.epi) ? 2 : 1;
      // folded stem through the generic kernel (InceptionI3d's 7x7x7): a 32-wide k-block is only two N = 64 MMAs, far below
      // the ~300 cycles a barrier round trip costs the issuing thread; four of them per stage
      if (r.a_mode != A_GATHER && r.bk == 32 && c.num_kb >= 4 && !r.epi) r.kps = 4;
      if (r.a_mode != A_GATHER && r.bk == 16 && !r.epi) r.kps = 8;
      if (p->kps_override == 1) r.kps = 1;
      if (p->kps_override == 2 && r.a_mode != A_GATHER && r.bk == 64 && r.bn <= 128 && c.num_kb >= 2) r.kps = 2;
      long long m_tiles = (M + kBlockM - 1) / kBlockM;
      r.thalo = !p->no_thalo && r.a_mode == A_TMA_IM2COL && !fold && !r.epi && d.res < 0 && r.bn <= 64 && d.kt == 3 && d.kh == 1 &&
                d.kw == 1 && d.st == 1 && d.sh == 1 && d.sw == 1 && pt == 1 && sym_pad && ph == 0 && pw == 0 &&
                (src.T == 2 || src.T == 4) && d.cin % 64 == 0;
      if (r.thalo) {
        ThaloParams& q = r.tp;
        memset(&q, 0, sizeof(q));
        q.B = batch; q.T = src.T; q.HW = src.H * src.W;
        q.P = 128 / src.T; q.logP = src.T == 2 ? 6 : 5;
        q.tiles_per_clip = (q.HW + q.P - 1) / q.P;
        q.N = d.cout; q.Cin = d.cin;
        q.relu = c.relu; q.ldo = Cdst;
        m_tiles = (long long)batch * q.tiles_per_clip;
        // weights resident in shared memory when one n tile covers cout and they leave room for >= 3 A-only stages
        const int kb_bytes = r.bn * 128, w_all = 3 * (d.cin / 64) * kb_bytes;
        const int budget = ThaloCfg<64>::kBudget;
        q.a_region = 16384 + 256 * q.P;
        q.resident = (d.cout <= r.bn && w_all + 3 * q.a_region <= budget) ? 1 : 0;
        q.stage_bytes = q.a_region + (q.resident ? 0 : 3 * kb_bytes);
        q.n_stages = (budget - (q.resident ? w_all : 0)) / q.stage_bytes;
        if (q.n_stages > ThaloCfg<64>::kMaxStages) q.n_stages = ThaloCfg<64>::kMaxStages;
        if (q.n_stages < 2) r.thalo = false;
      }
      const bool pool_t2 = (d.flags & VAD_FLAG_POOL_T2) != 0;
      r.pool_tp = false;
      if (pool_t2 && !fold) {
        // maxpool2 fused into a 1x1x1 residual conv: (all 4 frames x 32 pixels) tiles through the staged epilogue
        if (!(r.epi && r.bn == 128 && r.a_mode == A_TMA_2D && src.T == 4 && r.kps == 1))
          return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: POOL_T2 on a 1x1x1 conv needs 4 input frames, the staged epilogue and TMA operands", i);
        r.pool_tp = true;
        m_tiles = (long long)batch * ((src.H * Wi + 31) / 32);
      }
      r.s3 = !p->no_s3 && r.a_mode == A_TMA_IM2COL && !fold && !r.epi && d.res < 0 && d.cin == 64 && d.cout == 64 && d.kt == 1 &&
             d.kh == 3 && d.kw == 3 && d.st == 1 && d.sh == 1 && d.sw == 1 && pt == 0 && ph == 1 && pw == 1 && sym_pad;
      if (r.s3) {
        S3x3Params& q = r.s3p;
        memset(&q, 0, sizeof(q));
        q.F = batch * src.T; q.H = src.H; q.W = Wi;
        q.tiles_w = (Wi + 7) / 8; q.tiles_h = (src.H + 15) / 16;
        q.relu = c.relu;
        m_tiles = (long long)q.F * q.tiles_w * q.tiles_h;
      }
      const long long n_tiles = (d.cout + r.bn - 1) / r.bn;
      if (m_tiles * n_tiles > 0x7fffffffLL) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: grid too large", i);
      c.n_tiles = (int)n_tiles;
      c.num_tiles = (int)(m_tiles * n_tiles);
      r.grid = c.num_tiles < p->sm_count ? c.num_tiles : p->sm_count;  // persistent: one CTA per SM
      // CTA pairs pay off where the L2 -> shared-memory path is the limit (long K); short-K layers are output bound
      r.pair = p->pair_mode > 0 && (r.bn == 256 || r.bn == 128) && !r.epi && r.a_mode != A_GATHER && r.bk == 64 && !r.thalo && !r.s3 &&
               !r.pool_tp && d.cout % r.bn == 0 && (p->sm_count % 2) == 0 && m_tiles >= 2 && c.num_kb >= p->pair_min_kb;
      if (r.pair || r.pair_epi) {
        c.mc_items = (int)(((m_tiles + 1) / 2) * n_tiles);
        r.grid = 2 * c.mc_items < p->sm_count ? 2 * c.mc_items : p->sm_count;  // whole clusters, each with at least one item
        c.pair_split = c.pair_total = c.mc_items;
        c.pair_box_rows = r.bn / 2;
        // wave quantisation: when the last round of pair tiles fills at most half of the CTA pairs, run its items as two
        // 128-column halves each (one n tile only: cout == 256), e.g. layer3: 245 tiles on 74 pairs = 3.31 -> 3.5 rounds, not 4
        const int clusters = r.grid / 2, rem = clusters > 0 ? c.mc_items % clusters : 0;
        if (r.pair && r.bn == 256 && n_tiles == 1 && !p->no_pair_split && c.mc_items > clusters && rem > 0 && 2 * rem <= clusters) {
          c.pair_split = c.mc_items - rem;
          c.pair_total = c.pair_split + 2 * rem;
          c.pair_box_rows = 64;
        }
      }
      if (r.thalo) { r.tp.n_tiles = c.n_tiles; r.tp.num_tiles = c.num_tiles; }
      if (r.pool_tp) { c.pool_tp = 1; c.tp_tiles_per_clip = (src.H * Wi + 31) / 32; }
      if (r.s3) r.s3p.num_tiles = c.num_tiles;
      r.Ci = d.cin; r.Ti = src.T; r.Hi = src.H; r.Wi = Wi; r.fold = fold;
      r.stem = false;
      // (only the front padding enters the stem kernels: out-of-range rows / frames / columns behind the data are
      // zero-filled by TMA, so the asymmetric SAME padding of the Inception port needs nothing extra)
      if (fold && r.a_mode == A_TMA_IM2COL && !p->stem_generic && d.sh == 2 && d.sw == 2 && d.cout == 64 && d.res < 0 &&
          r.pf[2] <= p->in_pad_left) {
        StemParams& q = r.sp;
        memset(&q, 0, sizeof(q));
        q.clk_out = nullptr;
        q.B = batch; q.To = To; q.Ho = Ho; q.Wo = Wo;
        q.pool_t = pool_t2 ? 2 : 1;
        q.To_out = To / q.pool_t;
        q.kt = d.kt; q.kh = d.kh; q.st = d.st; q.pt = pt; q.ph = ph;
        const int th = 16, tw = 8;  // output tile: 8 (w) x 16 (h)
        q.tiles_w = (Wo + tw - 1) / tw; q.tiles_h = (Ho + th - 1) / th;
        long long nu = (long long)batch * q.To_out * q.tiles_h * q.tiles_w;
        q.rows_even = th + (d.kh + 1) / 2 - 1;
        q.rows_odd = th + d.kh / 2 - 1;
        q.seg_bytes = ((tw - 1) * d.sw * 4 + 32) * 2;  // bytes per raw input-row segment in smem (176)
        q.off_odd = (int)align_up((uint64_t)q.rows_even * q.seg_bytes, 128);
        q.stage_bytes = (int)align_up((uint64_t)q.off_odd + (uint64_t)q.rows_odd * q.seg_bytes, 128);
        int w_bytes = d.kt * d.kh * kStemTapBytes;
        r.stem_pair = w_bytes > 150 * 1024 && w_bytes <= 300 * 1024 && q.pool_t == 1 && (p->sm_count & 1) == 0 && !p->stem_no_pair;
        if (r.stem_pair) {
          // too many taps for one CTA (7x7x7: 196 KB): a CTA pair, each CTA keeping half of the output channels' weights resident
          w_bytes = (int)align_up((uint64_t)w_bytes / 2, 1024);
        } else if (w_bytes > 150 * 1024) {
          // too many taps to keep resident (7x7x7: 196 KB): the kh taps of one dt ride in that dt's stage
          q.w_stream = 1;
          q.off_w = (int)align_up((uint64_t)q.stage_bytes, 1024);
          q.stage_bytes = q.off_w + d.kh * kStemTapBytes;
          w_bytes = 0;
        }
        const int fixed = w_bytes + 2 * kStemStagingBytes + 2 * 64 * 4 + (2 * kStemMaxStages + 17) * 8 + 16 + 32 * 32 + 1024;
        int ns = (227 * 1024 - fixed) / q.stage_bytes;
        if (ns > kStemMaxStages) ns = kStemMaxStages;
        // multi-frame variant: all output frames of a spatial tile live in TMEM (8 x 64 columns), input frames are
        // walked once; needs the I3D temporal geometry (kt 5, stride 2, pad 2) and at most 8 output frames
        r.stem_mf = !r.stem_pair && !p->stem_v3 && d.kt == 5 && d.st == 2 && pt == 2 && To <= 8 && To >= 1 && d.kh <= 7;
        if (r.stem_mf) {
          nu = (long long)batch * q.tiles_h * q.tiles_w;
          r.stem_ti = src.T;
          r.stem_ti_max = src.T - 1 < 2 * (To - 1) + 2 ? src.T - 1 : 2 * (To - 1) + 2;
        }
        if (ns >= 2 && nu > 0 && nu <= 0x7fffffffLL && d.kh > 1) {
          q.n_stages = ns;
          q.num_units = (int)nu;
          q.relu = c.relu;
          { const char* sd = getenv("VAD_STEM_DEBUG"); q.dbg = sd ? atoi(sd) : 0; }
          r.stem_smem = fixed + ns * q.stage_bytes;
          r.stem = true;
          r.grid = q.num_units < p->sm_count ? q.num_units : p->sm_count;
          if (r.stem_pair) {   // one item = two tiles
            const int items = (q.num_units + 1) / 2, pairs = p->sm_count / 2;
            r.grid = 2 * (items < pairs ? items : pairs);
          }
        }
      }
      if (pool_t2 && fold) {
        if (!r.stem)
          return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: POOL_T2 needs the dedicated stem kernel (stride 2, cout 64, TMA input, "
                      "no residual)", i);
        if (To < 2) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: POOL_T2 needs at least two output frames", i);
        To = To / 2;  // shape of the dst slot
      }
      const uint64_t need_w = d.w_off + (uint64_t)d.cout * r.K_pad * 2;
      if (need_w > p->params_bytes || d.scale_off + 4ull * d.cout > p->params_bytes || d.shift_off + 4ull * d.cout > p->params_bytes)
        return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: parameters exceed the blob (%llu > %llu)", i,
                    (unsigned long long)need_w, (unsigned long long)p->params_bytes);
      if (d.res >= 0) {
        const SlotInfo& rs = p->slots[d.res];
        if (!rs.defined || rs.T != To || rs.H != Ho || rs.W != Wo || rs.C < d.cout)
          return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: residual slot %d shape mismatch", i, d.res);
        c.ldr = rs.C;
        r.res_c = rs.C;
      }
      if (r.pool_tp) To = To / 2;  // shape of the dst slot (the residual above has the unpooled shape)
      const int cin_real = fold ? 3 : d.cin;
      p->op_flops[i] = 2.0 * (double)M * d.cout * d.kt * d.kh * d.kw * cin_real;  // frames the reference conv produces
      p->flops += p->op_flops[i];
      // activations read once, weights once, output written once (+ residual read)
      p->op_bytes[i] = 2.0 * ((double)batch * src.T * src.H * src.W * src.C + (double)d.cout * r.K_pad +
                              (double)M * d.cout * (d.res >= 0 ? 2.0 : (pool_t2 ? 0.5 : 1.0)));
    } else if (d.kind == VAD_OP_MAXPOOL) {
      PoolParams& q = r.pp;
      memset(&q, 0, sizeof(q));
      if (src.C % 8) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: max-pool needs C %% 8 == 0", i);
      if (d.flags & VAD_FLAG_POOL_SAME) {
        To = pool_out_same(src.T, d.st); Ho = pool_out_same(src.H, d.sh); Wo = pool_out_same(src.W, d.sw);
        auto front = [](int in, int out, int k, int s) { int tot = (out - 1) * s + k - in; if (tot < 0) tot = 0; return tot / 2; };
        q.pt = front(src.T, To, d.kt, d.st); q.ph = front(src.H, Ho, d.kh, d.sh); q.pw = front(src.W, Wo, d.kw, d.sw);
        q.pad_zero = 1;
      } else {
        To = (src.T + 2 * d.pt - d.kt) / d.st + 1;
        Ho = (src.H + 2 * d.ph - d.kh) / d.sh + 1;
        Wo = (src.W + 2 * d.pw - d.kw) / d.sw + 1;
        q.pt = d.pt; q.ph = d.ph; q.pw = d.pw; q.pad_zero = 0;
      }
      if (To <= 0 || Ho <= 0 || Wo <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: empty output", i);
      Cdst = d.dst_c_total ? d.dst_c_total : src.C;
      if (d.dst_c_off % 8 || d.dst_c_off + src.C > Cdst) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: bad channel slice", i);
      q.B = batch; q.Ti = src.T; q.Hi = src.H; q.Wi = src.W; q.C = src.C;
      q.To = To; q.Ho = Ho; q.Wo = Wo;
      q.kt = d.kt; q.kh = d.kh; q.kw = d.kw; q.st = d.st; q.sh = d.sh; q.sw = d.sw;
      q.ldo = Cdst;
      p->op_bytes[i] = 2.0 * ((double)batch * src.T * src.H * src.W * src.C + (double)batch * To * Ho * Wo * src.C);
    } else {  // AVGPOOL
      if (src.C % 64) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: avg-pool needs C %% 64 == 0", i);
      r.avg_P = src.T * src.H * src.W;
      r.avg_C = src.C;
      // kt > 0: AvgPool3d((kt, H, W), stride 1) + global mean over the windows (kh / kw, when given, must cover the map)
      if (d.kt > 1 && d.kt < src.T) {
        if ((d.kh && d.kh != src.H) || (d.kw && d.kw != src.W)) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: windowed avg-pool must span the whole %dx%d map", i, src.H, src.W);
        r.avg_HW = src.H * src.W;
        r.avg_kt = d.kt;
      }
      p->feat_c = src.C;
      p->op_bytes[i] = 2.0 * batch * (double)r.avg_P * src.C + 4.0 * batch * src.C;
      continue;
    }
    SlotInfo& dst = p->slots[d.dst];
    if (dst.defined && (dst.T != To || dst.H != Ho || dst.W != Wo || dst.C != Cdst)) {
      // a slot may be reused with a new shape once its previous contents are dead
      dst.T = To; dst.H = Ho; dst.W = Wo; dst.C = Cdst;
    } else if (!dst.defined) {
      dst.T = To; dst.H = Ho; dst.W = Wo; dst.C = Cdst; dst.defined = true;
    }
    const uint64_t bytes = (uint64_t)batch * To * Ho * Wo * Cdst * 2;
    if (bytes > dst.bytes) dst.bytes = bytes;
  }
  // ---- bottleneck-tail fusion: a (1,3,3) 64 -> 64 halo-tile conv whose output feeds only the next 1x1x1 64 -> 256
  // residual conv runs both in one launch (conv_tail.cuh); when the residual is the block's own 1x1x1 downsample of a
  // 64-channel X, that conv joins the contraction as well.  The intermediate slots must be dead afterwards.
  if (!p->no_tail) {
    const int n_ops = (int)p->ops.size();
    auto dead_after = [&](int slot, int last_reader) {
      for (int j = last_reader + 1; j < n_ops; ++j) {
        const vad_op_desc& e = p->ops[j];
        if (e.src == slot || (e.kind == VAD_OP_CONV && e.res == slot)) return false;
        if (e.kind != VAD_OP_AVGPOOL && e.dst == slot) return true;
      }
      return true;
    };
    auto is_proj = [&](int j) {  // 1x1x1, stride 1, 64 -> 256 into a whole 256-channel slot, TMA operands
      if (j >= n_ops) return false;
      const vad_op_desc& e = p->ops[j];
      return e.kind == VAD_OP_CONV && e.kt == 1 && e.kh == 1 && e.kw == 1 && e.st == 1 && e.sh == 1 && e.sw == 1 && !e.pt && !e.ph && !e.pw &&
             !(e.flags & ~VAD_FLAG_RELU) && e.cin == 64 && e.cout == 256 && e.dst_c_off == 0 && (e.dst_c_total == 0 || e.dst_c_total == 256) &&
             p->rt[j].K_pad == 64;
    };
    for (int i = 0; i + 1 < n_ops; ++i) {
      OpRuntime& r = p->rt[i];
      const vad_op_desc& d = p->ops[i];
      if (!r.s3 || r.skip) continue;
      int c3 = -1, ds = -1;
      if (is_proj(i + 1) && p->ops[i + 1].src == d.dst && p->ops[i + 1].res >= 0 && p->ops[i + 1].res != d.dst &&
          p->rt[i + 1].res_c == 256 && p->ops[i + 1].dst != p->ops[i + 1].res && dead_after(d.dst, i + 1)) {
        c3 = i + 1;
      } else if (is_proj(i + 1) && is_proj(i + 2) && p->ops[i + 1].res < 0 && p->ops[i + 1].src != d.dst && p->ops[i + 1].dst != d.dst &&
                 p->rt[i + 1].cp.M == r.cp.M && p->ops[i + 2].src == d.dst && p->ops[i + 2].res == p->ops[i + 1].dst &&
                 p->ops[i + 2].dst != p->ops[i + 1].src && p->ops[i + 2].dst != d.dst && dead_after(d.dst, i + 2) &&
                 dead_after(p->ops[i + 1].dst, i + 2)) {
        ds = i + 1; c3 = i + 2;
      }
      if (c3 < 0) continue;
      r.tail = ds >= 0 ? 2 : 1;
      r.tail_c3 = c3; r.tail_ds = ds;
      p->rt[c3].skip = true;
      if (ds >= 0) p->rt[ds].skip = true;
      TailParams& q = r.tlp;
      memset(&q, 0, sizeof(q));
      q.F = r.s3p.F; q.H = r.s3p.H; q.W = r.s3p.W;
      q.tiles_w = r.s3p.tiles_w; q.tiles_h = r.s3p.tiles_h; q.num_tiles = r.s3p.num_tiles;
      q.relu2 = r.cp.relu; q.relu3 = p->rt[c3].cp.relu;
      {
        // BN scale / shift travel in the kernel parameter block (constant bank): fetch them from the parameter blob once
        // per configure (a few KB, synchronous like the rest of configure; the blob is immutable for the life of the plan)
        const vad_op_desc& d3 = p->ops[c3];
        float sd_shift[256];
        VAD_CUDA_CHECK(cudaMemcpy(q.s2, p->params + d.scale_off, 64 * 4, cudaMemcpyDeviceToHost));
        VAD_CUDA_CHECK(cudaMemcpy(q.b2, p->params + d.shift_off, 64 * 4, cudaMemcpyDeviceToHost));
        VAD_CUDA_CHECK(cudaMemcpy(q.s3, p->params + d3.scale_off, 256 * 4, cudaMemcpyDeviceToHost));
        VAD_CUDA_CHECK(cudaMemcpy(q.b3, p->params + d3.shift_off, 256 * 4, cudaMemcpyDeviceToHost));
        if (ds >= 0) {
          VAD_CUDA_CHECK(cudaMemcpy(sd_shift, p->params + p->ops[ds].shift_off, 256 * 4, cudaMemcpyDeviceToHost));
          for (int k = 0; k < 256; ++k) q.b3[k] += sd_shift[k];
        }
      }
      // per-op accounting: the fused launch carries the FLOPs of its parts; bytes = each tensor touched once
      const double Md = (double)r.cp.M;
      p->op_flops[i] += p->op_flops[c3] + (ds >= 0 ? p->op_flops[ds] : 0.0);
      p->op_flops[c3] = 0.0;
      p->op_bytes[i] = 2.0 * (Md * 64 + Md * 256 + Md * (ds >= 0 ? 64 : 256) + 9.0 * 64 * 64 + 256.0 * 64 * (ds >= 0 ? 2 : 1));
      p->op_bytes[c3] = 0.0;
      if (ds >= 0) {
        p->op_flops[ds] = 0.0; p->op_bytes[ds] = 0.0;
        bool have = false;
        for (auto& fb : p->fold_bufs) have = have || fb.first == i;
        if (!have) {
          void* buf = nullptr;
          VAD_CUDA_CHECK(cudaMalloc(&buf, 256 * 128 * 2));
          p->fold_bufs.emplace_back(i, buf);
        }
        p->fold_pending = true;
      }
    }
  }
  uint64_t off = 0;
  for (int s = 1; s < p->n_slots; ++s) {
    if (!p->slots[s].defined) continue;
    p->slots[s].offset = off;
    // + one tile of slack: TMA boxes of the last (partial) tile never leave the allocation
    off += align_up(p->slots[s].bytes + 1024, 1024);
  }
  p->ws_bytes = off;
  p->batch = batch; p->T = t; p->H = h; p->W = w;
  p->bound_x = nullptr; p->bound_ws = nullptr;
  p->configured = true;
  if (workspace_bytes) *workspace_bytes = off;
  return VAD_OK;
}

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
static void drop_graph(vad_plan* p) {
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  p->direct_runs = 0;
}

static int32_t bind_plan(vad_plan* p, const void* x, void* ws, cudaStream_t st) {
  auto slot_ptr = [&](int s) -> uint8_t* {
    return s == 0 ? const_cast<uint8_t*>(static_cast<const uint8_t*>(x)) : static_cast<uint8_t*>(ws) + p->slots[s].offset;
  };
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const vad_op_desc& d = p->ops[i];
    OpRuntime& r = p->rt[i];
    if (d.kind == VAD_OP_CONV) {
      ConvParams& c = r.cp;
      const bool fold = d.flags & VAD_FLAG_STEM_FOLD_W;
      c.in = reinterpret_cast<const __nv_bfloat16*>(slot_ptr(d.src)) + (fold ? (p->in_pad_left - r.pf[2]) * 4 : 0);
      c.out = reinterpret_cast<__nv_bfloat16*>(slot_ptr(d.dst)) + d.dst_c_off;
      c.res = d.res >= 0 ? reinterpret_cast<const __nv_bfloat16*>(slot_ptr(d.res)) : nullptr;
      c.scale = reinterpret_cast<const float*>(p->params + d.scale_off);
      c.shift = reinterpret_cast<const float*>(p->params + d.shift_off);
      // weights: [cout][K_pad] bf16, box = 64 (K) x BN (rows), 128B swizzle
      {
        cuuint64_t gdim[2] = {(cuuint64_t)r.K_pad, (cuuint64_t)d.cout};
        cuuint64_t gstr[1] = {(cuuint64_t)r.K_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)r.bk, (cuuint32_t)r.bn};
        cuuint32_t es[2] = {1, 1};
        CUresult cr = p->encode_tiled(&r.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(p->params + d.w_off), gdim,
                                      gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      r.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (r.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(weights) failed: %d", i, (int)cr);
      }
      memset(&r.tmA, 0, sizeof(r.tmA));
      memset(&r.tmR, 0, sizeof(r.tmR));
      memset(&r.tmO, 0, sizeof(r.tmO));
      memset(&r.tmBh, 0, sizeof(r.tmBh));
      if (r.pair || r.pair_epi) {
        // each CTA of a pair loads half of the BN weight rows
        cuuint64_t gdim[2] = {(cuuint64_t)r.K_pad, (cuuint64_t)d.cout};
        cuuint64_t gstr[1] = {(cuuint64_t)r.K_pad * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)c.pair_box_rows};
        cuuint32_t es[2] = {1, 1};
        CUresult cr = p->encode_tiled(&r.tmBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(p->params + d.w_off), gdim, gstr, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(weight halves) failed: %d", i, (int)cr);
      }
      if (r.epi) {
        // residual [M, res_c] -> 128-row x 64-channel boxes; output slice [M, cout] (row pitch dst_c) <- 32-row boxes
        cuuint32_t es2[2] = {1, 1};
        CUresult cr;
        if (r.pool_tp) {
          // (channels, H*W, T, clips) views: residual boxes are 64 ch x 32 px x 4 frames, the pooled output 64 x 32 x 1
          const uint64_t hw = (uint64_t)c.Ho * c.Wo;
          cuuint32_t es4[4] = {1, 1, 1, 1};
          cuuint64_t rdim[4] = {(cuuint64_t)r.res_c, hw, 4, (cuuint64_t)p->batch};
          cuuint64_t rstr[3] = {(cuuint64_t)r.res_c * 2, (cuuint64_t)r.res_c * 2 * hw, (cuuint64_t)r.res_c * 2 * hw * 4};
          cuuint32_t rbox[4] = {64, 32, 4, 1};
          cr = p->encode_tiled(&r.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c.res, rdim, rstr, rbox, es4,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr == CUDA_SUCCESS) {
            cuuint64_t odim[4] = {(cuuint64_t)d.cout, hw, 2, (cuuint64_t)p->batch};
            cuuint64_t ostr[3] = {(cuuint64_t)r.dst_c * 2, (cuuint64_t)r.dst_c * 2 * hw, (cuuint64_t)r.dst_c * 2 * hw * 2};
            cuuint32_t obox[4] = {64, 32, 1, 1};
            cr = p->encode_tiled(&r.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c.out, odim, ostr, obox, es4,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          }
          if (cr == CUDA_SUCCESS) {
            cuuint64_t adim[4] = {(cuuint64_t)r.Ci, hw, 4, (cuuint64_t)p->batch};
            cuuint64_t astr[3] = {(cuuint64_t)r.Ci * 2, (cuuint64_t)r.Ci * 2 * hw, (cuuint64_t)r.Ci * 2 * hw * 4};
            cuuint32_t abox[4] = {64, 32, 4, 1};
            cr = p->encode_tiled(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(d.src), adim, astr, abox, es4,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          }
          if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(fused temporal pool) failed: %d", i, (int)cr);
        } else {
        if (d.res >= 0) {
          cuuint64_t rdim[2] = {(cuuint64_t)r.res_c, (cuuint64_t)c.M};
          cuuint64_t rstr[1] = {(cuuint64_t)r.res_c * 2};
          cuuint32_t rbox[2] = {64, (cuuint32_t)kBlockM};
          cr = p->encode_tiled(&r.tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)c.res, rdim, rstr, rbox, es2,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(residual) failed: %d", i, (int)cr);
        }
        cuuint64_t odim[2] = {(cuuint64_t)d.cout, (cuuint64_t)c.M};
        cuuint64_t ostr[1] = {(cuuint64_t)r.dst_c * 2};
        cuuint32_t obox[2] = {64, 32};
        cr = p->encode_tiled(&r.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)c.out, odim, ostr, obox, es2,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(output) failed: %d", i, (int)cr);
        }
      }
      if (r.stem) {
        // raw padded rows viewed as (x = Wp * 4 elements, H, T, N): a box is 88 contiguous elements (the union of
        // 8 overlapping windows) x rows with stride 2, no swizzle; even / odd input rows are two boxes
        StemParams& q = r.sp;
        q.scale = c.scale; q.shift = c.shift;
        const uint64_t wp = (uint64_t)p->slots[0].W;
        cuuint64_t rdim[4] = {(cuuint64_t)wp * 4, (cuuint64_t)r.Hi, (cuuint64_t)r.Ti, (cuuint64_t)p->batch};
        cuuint64_t rstr[3] = {wp * 4 * 2, wp * 4 * 2 * r.Hi, wp * 4 * 2 * r.Hi * r.Ti};
        cuuint32_t res4[4] = {1, 2, 1, 1};
        cuuint32_t bE[4] = {(cuuint32_t)(q.seg_bytes / 2), (cuuint32_t)(2 * q.rows_even - 1), 1, 1};
        cuuint32_t bO[4] = {(cuuint32_t)(q.seg_bytes / 2), (cuuint32_t)(2 * q.rows_odd - 1), 1, 1};
        CUresult cr = p->encode_tiled(&r.tmE, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c.in, rdim, rstr, bE, res4,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr == CUDA_SUCCESS)
          cr = p->encode_tiled(&r.tmOdd, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c.in, rdim, rstr, bO, res4,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr == CUDA_SUCCESS) {
          cuuint64_t wdim[2] = {(cuuint64_t)r.K_pad, (cuuint64_t)d.cout};
          cuuint64_t wstr[1] = {(cuuint64_t)r.K_pad * 2};
          cuuint32_t wbox[2] = {32, 64};
          cuuint32_t wes[2] = {1, 1};
          cr = p->encode_tiled(&r.tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(p->params + d.w_off), wdim, wstr, wbox,
                               wes, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr == CUDA_SUCCESS && r.stem_pair) {
            cuuint32_t hbox[2] = {32, 32};
            cr = p->encode_tiled(&r.tmWh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(p->params + d.w_off), wdim, wstr, hbox,
                                 wes, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          }
        }
        if (cr == CUDA_SUCCESS) {
          // output [N, To_out, Ho, Wo, Cdst] (channel slice at c.out): one store per epilogue warp = 64 channels x
          // 8 columns x 4 rows out of the 128B-swizzled staging tile
          const uint64_t cb = (uint64_t)r.dst_c * 2;
          cuuint64_t odim[5] = {(cuuint64_t)d.cout, (cuuint64_t)q.Wo, (cuuint64_t)q.Ho, (cuuint64_t)q.To_out, (cuuint64_t)p->batch};
          cuuint64_t ostr[4] = {cb, cb * q.Wo, cb * q.Wo * q.Ho, cb * q.Wo * q.Ho * q.To_out};
          cuuint32_t obox[5] = {64, 8, 4, 1, 1};
          cuuint32_t oes[5] = {1, 1, 1, 1, 1};
          cr = p->encode_tiled(&r.tmSO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)c.out, odim, ostr, obox, oes,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (cr != CUDA_SUCCESS)
          return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(stem) failed: %d; set VAD_STEM_GENERIC=1", i, (int)cr);
      } else if (r.s3) {
        // input (C, W, H, F): one box = 64 channels x 10 columns x 18 rows (tile + halo); output slice (cout, W, H, F): 8 x 4 per store
        S3x3Params& q = r.s3p;
        q.scale = c.scale; q.shift = c.shift;
        cuuint64_t gdim[4] = {64, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.F};
        cuuint64_t gstr[3] = {128, (cuuint64_t)128 * q.W, (cuuint64_t)128 * q.W * q.H};
        cuuint32_t box[4] = {64, 10, 18, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult cr = p->encode_tiled(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(d.src), gdim, gstr, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr == CUDA_SUCCESS) {
          const uint64_t cb = (uint64_t)r.dst_c * 2;
          cuuint64_t odim[4] = {64, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.F};
          cuuint64_t ostr[3] = {cb, cb * q.W, cb * q.W * q.H};
          cuuint32_t obox[4] = {64, 8, 4, 1};
          cr = p->encode_tiled(&r.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c.out, odim, ostr, obox, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(3x3 halo tile) failed: %d", i, (int)cr);
        if (r.tail) {
          // fused conv3 (+ downsample): resident 256 x 64 weight tiles; 64 ch x 8 w x 16 h boxes over the 256-channel
          // residual (tail = 1) or the 64-channel block input X (tail = 2), and over the 256-channel output
          const vad_op_desc& d3 = p->ops[r.tail_c3];
          cuuint32_t es2[2] = {1, 1};
          cuuint32_t wbox[2] = {64, 256};
          cuuint32_t cbox[4] = {64, 8, 16, 1};
          cuuint64_t wide[4] = {256, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.F};
          cuuint64_t wstr4[3] = {512, (cuuint64_t)512 * q.W, (cuuint64_t)512 * q.W * q.H};
          cr = p->encode_tiled(&r.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(d3.dst), wide, wstr4, cbox, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (cr == CUDA_SUCCESS && r.tail == 1) {
            cuuint64_t wdim[2] = {64, 256};
            cuuint64_t wstr[1] = {128};
            cr = p->encode_tiled(&r.tmW3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(p->params + d3.w_off), wdim, wstr, wbox, es2,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr == CUDA_SUCCESS)
              cr = p->encode_tiled(&r.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(d3.res), wide, wstr4, cbox, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          } else if (cr == CUDA_SUCCESS) {
            const vad_op_desc& dd = p->ops[r.tail_ds];
            void* fb = nullptr;
            for (auto& e : p->fold_bufs) if (e.first == (int)i) fb = e.second;
            if (!fb) return fail(VAD_ERR_CUDA, "op %zu: folded tail weights were not allocated", i);
            if (p->fold_pending) {
              fold_tail_weights_kernel<<<(256 * 128 + 255) / 256, 256, 0, st>>>(
                  reinterpret_cast<const __nv_bfloat16*>(p->params + d3.w_off), reinterpret_cast<const __nv_bfloat16*>(p->params + dd.w_off),
                  reinterpret_cast<const float*>(p->params + d3.scale_off), reinterpret_cast<const float*>(p->params + dd.scale_off), 64, 64,
                  static_cast<__nv_bfloat16*>(fb));
              if (cudaGetLastError() != cudaSuccess) return fail(VAD_ERR_CUDA, "op %zu: fold_tail_weights_kernel launch failed", i);
            }
            cuuint64_t wdim[2] = {128, 256};
            cuuint64_t wstr[1] = {256};
            cr = p->encode_tiled(&r.tmW3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, fb, wdim, wstr, wbox, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr == CUDA_SUCCESS)
              cr = p->encode_tiled(&r.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(dd.src), gdim, gstr, cbox, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          }
          if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(bottleneck tail) failed: %d", i, (int)cr);
        }
      } else if (r.thalo) {
        // (C, HW, T, N): one box = 64 channels x P pixels x all T frames of one clip
        ThaloParams& q = r.tp;
        q.scale = c.scale; q.shift = c.shift; q.out = c.out;
        cuuint64_t gdim[4] = {(cuuint64_t)r.Ci, (cuuint64_t)q.HW, (cuuint64_t)q.T, (cuuint64_t)p->batch};
        cuuint64_t gstr[3] = {(cuuint64_t)r.Ci * 2, (cuuint64_t)r.Ci * 2 * q.HW, (cuuint64_t)r.Ci * 2 * q.HW * q.T};
        cuuint32_t box[4] = {64, (cuuint32_t)q.P, (cuuint32_t)q.T, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult cr = p->encode_tiled(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)slot_ptr(d.src), gdim, gstr, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(temporal halo A) failed: %d", i, (int)cr);
      } else if (r.pool_tp) {
        // operand map encoded with the epilogue maps above
      } else if (r.a_mode == A_TMA_2D) {
        cuuint64_t gdim[2] = {(cuuint64_t)r.Ci, (cuuint64_t)c.M};
        cuuint64_t gstr[1] = {(cuuint64_t)r.Ci * 2};
        cuuint32_t box[2] = {(cuuint32_t)r.bk, (cuuint32_t)kBlockM};
        cuuint32_t es[2] = {1, 1};
        CUresult cr = p->encode_tiled(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)slot_ptr(d.src), gdim, gstr,
                                      box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      r.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (r.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(A) failed: %d", i, (int)cr);
      } else if (r.a_mode == A_TMA_IM2COL && r.fold) {
        // Stem: view the padded [N, T, H, Wp, 4] input as (C' = 32, W' = Wo, H, T, N) where pixel w' is the
        // 8-pixel x 4-channel window starting at padded column sw * w' -- consecutive windows overlap, so
        // the W' stride (sw * 8 B = 16 B) is smaller than the row extent (64 B).  kw is folded into C', so
        // only (dh, dt) remain as im2col offsets.
        const uint64_t wp = (uint64_t)p->slots[0].W;
        cuuint64_t gdim[5] = {32, (cuuint64_t)c.Wo, (cuuint64_t)r.Hi, (cuuint64_t)r.Ti, (cuuint64_t)p->batch};
        cuuint64_t gstr[4];
        gstr[0] = (cuuint64_t)d.sw * 4 * 2;
        gstr[1] = wp * 4 * 2;
        gstr[2] = gstr[1] * r.Hi;
        gstr[3] = gstr[2] * r.Ti;
        int lower[3] = {0, -r.pf[1], -r.pf[0]};
        int upper[3] = {0, r.pb[1] - (d.kh - 1), r.pb[0] - (d.kt - 1)};
        cuuint32_t es[5] = {1, 1, (cuuint32_t)d.sh, (cuuint32_t)d.st, 1};
        CUresult cr = p->encode_im2col(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)c.in, gdim, gstr, lower, upper,
                                       32, (cuuint32_t)kBlockM, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS)
          return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeIm2col(stem window view) failed: %d", i, (int)cr);
        const uint64_t tensor_bytes = gstr[3] * (uint64_t)p->batch;
        if (p->driver_version <= 13010 && tensor_bytes < 131072)
          reinterpret_cast<uint64_t*>(&r.tmA)[1] &= ~(1ull << 21);
      } else if (r.a_mode == A_TMA_IM2COL) {
        // (C, W, H, D, N); the bounding box of base pixels runs from -pad to (extent - 1 + pad - (k-1))
        cuuint64_t gdim[5] = {(cuuint64_t)r.Ci, (cuuint64_t)r.Wi, (cuuint64_t)r.Hi, (cuuint64_t)r.Ti, (cuuint64_t)p->batch};
        cuuint64_t gstr[4];
        gstr[0] = (cuuint64_t)r.Ci * 2;
        gstr[1] = gstr[0] * r.Wi;
        gstr[2] = gstr[1] * r.Hi;
        gstr[3] = gstr[2] * r.Ti;
        int lower[3] = {-r.pf[2], -r.pf[1], -r.pf[0]};
        int upper[3] = {r.pb[2] - (d.kw - 1), r.pb[1] - (d.kh - 1), r.pb[0] - (d.kt - 1)};
        cuuint32_t es[5] = {1, (cuuint32_t)d.sw, (cuuint32_t)d.sh, (cuuint32_t)d.st, 1};
        CUresult cr = p->encode_im2col(&r.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)slot_ptr(d.src), gdim, gstr,
                                       lower, upper, (cuuint32_t)r.bk, (cuuint32_t)kBlockM, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       r.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (r.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeIm2col failed: %d", i, (int)cr);
        // Drivers up to 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (bit 21 of the
        // second descriptor word must be cleared); public CUTLASS applies the same fix-up.
        const uint64_t tensor_bytes = gstr[3] * (uint64_t)p->batch;
        if (p->driver_version <= 13010 && tensor_bytes < 131072)
          reinterpret_cast<uint64_t*>(&r.tmA)[1] &= ~(1ull << 21);
      }
    } else if (d.kind == VAD_OP_MAXPOOL) {
      r.pp.in = reinterpret_cast<const __nv_bfloat16*>(slot_ptr(d.src));
      r.pp.out = reinterpret_cast<__nv_bfloat16*>(slot_ptr(d.dst)) + d.dst_c_off;
    }
  }
  p->fold_pending = false;
  p->bound_x = x;
  p->bound_ws = ws;
  drop_graph(p);
  return VAD_OK;
}

// Launch with the programmatic-stream-serialization attribute: the kernel may begin (barrier init, TMEM allocation,
// loads of constant weights) while its predecessor in the stream is still draining; every kernel launched this way
// executes griddepcontrol.wait before it touches anything a predecessor wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, int cluster,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int BN, int BK, int KPS, bool GATHER, bool EPI>
static cudaError_t launch_conv(const OpRuntime& r, cudaStream_t st) {
  using Cfg = ConvCfg<BN, BK, KPS, GATHER, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<BN, BK, KPS, GATHER, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_k(conv_umma_kernel<BN, BK, KPS, GATHER, EPI>, r.grid, Cfg::kThreads, Cfg::kSmemBytes, st, g_pdl, 1, r.tmA, r.tmB, r.tmR, r.tmO, r.cp);
}

template <int BN, bool EPI>
static cudaError_t launch_conv_bn(const OpRuntime& r, cudaStream_t st) {
  if (r.a_mode == A_GATHER) return launch_conv<BN, 64, 1, true, EPI>(r, st);
  if (r.kps == 2) return launch_conv<BN, 64, 2, false, EPI>(r, st);
  return launch_conv<BN, 64, 1, false, EPI>(r, st);
}

template <int BN, int KPS, bool EPI>
static cudaError_t launch_conv_pair_t(const OpRuntime& r, cudaStream_t st) {
  using Cfg = PairCfg<BN, KPS, EPI>;
  auto kern = conv_pair_kernel<BN, KPS, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_k(kern, r.grid, Cfg::kThreads, Cfg::kSmemBytes, st, g_pdl, 2, r.tmA, r.tmBh, r.tmR, r.tmO, r.cp);
}
static cudaError_t launch_conv_pair(const OpRuntime& r, cudaStream_t st) {
  if (r.pair_epi) return launch_conv_pair_t<256, 1, true>(r, st);
  return r.bn == 256 ? launch_conv_pair_t<256, 1, false>(r, st) : launch_conv_pair_t<128, 2, false>(r, st);
}

static cudaError_t launch_conv_any(const OpRuntime& r, cudaStream_t st) {
  if (r.pair || r.pair_epi) return launch_conv_pair(r, st);
  if (r.bk == 16)  // Cin % 32 == 16 layers: 16-wide k-blocks, eight per stage
    return r.bn == 128 ? launch_conv<128, 16, 8, false, false>(r, st) : launch_conv<64, 16, 8, false, false>(r, st);
  if (r.bk == 32) {  // folded stem (TMA window view) and Cin % 64 == 32 layers: 32-wide k-blocks, direct epilogue
    if (r.bn == 128) return r.kps == 4 ? launch_conv<128, 32, 4, false, false>(r, st) : launch_conv<128, 32, 1, false, false>(r, st);
    return r.kps == 4 ? launch_conv<64, 32, 4, false, false>(r, st) : launch_conv<64, 32, 1, false, false>(r, st);
  }
  if (r.epi) return r.bn == 128 ? launch_conv_bn<128, true>(r, st) : launch_conv_bn<64, true>(r, st);
  switch (r.bn) {
    case 256: return r.a_mode == A_GATHER ? launch_conv<256, 64, 1, true, false>(r, st) : launch_conv<256, 64, 1, false, false>(r, st);
    case 128: return launch_conv_bn<128, false>(r, st);
    default:  return launch_conv_bn<64, false>(r, st);
  }
}

static int grid_for(long long total, int threads, int cap = 148 * 32) {
  long long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// every op of the table, in order, into stream st (directly, or inside a stream capture)
static int32_t run_ops(vad_plan* p, const void* x_dev, void* workspace_dev, float* feat_out_dev, cudaStream_t st) {
  auto mark = [&]() -> cudaError_t {
    if (!p->profiling) return cudaSuccess;
    cudaEvent_t ev;
    if (!p->ev_pool.empty()) { ev = p->ev_pool.back(); p->ev_pool.pop_back(); }
    else { cudaError_t ce = cudaEventCreate(&ev); if (ce != cudaSuccess) return ce; }
    p->ev_used.push_back(ev);
    return cudaEventRecord(ev, st);
  };
  const int pf0 = p->prof_count < 0 ? 0 : p->prof_first;
  const int pf1 = p->prof_count < 0 ? (int)p->ops.size() : p->prof_first + p->prof_count;  // events before ops pf0..pf1-1 and after op pf1-1
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if ((int)i >= pf0 && (int)i < pf1 && mark() != cudaSuccess) return fail(VAD_ERR_CUDA, "profiling event failed");
    const vad_op_desc& d = p->ops[i];
    const OpRuntime& r = p->rt[i];
    cudaError_t e = cudaSuccess;
    if (r.skip) {
      // ran inside the fused launch of an earlier op
    } else if (d.kind == VAD_OP_CONV) {
      if (r.tail) {
        // residual form: 2 halo stages + a ring of 3 staging tiles; downsample form: 1 halo stage + 2 staging tiles
        static long long* tail_dbg = nullptr;
        static const bool want_dbg = getenv("VAD_TAIL_DEBUG") != nullptr;
        if (want_dbg && !tail_dbg) { cudaMalloc(&tail_dbg, 2 * 4 * 32 * 8); cudaMemset(tail_dbg, 0, 2 * 4 * 32 * 8); }
        auto launch_tail = [&](auto kern, int smem, auto mode_tag) -> cudaError_t {
          static bool attr = false;   // one per instantiation of this generic lambda: mode_tag tells the two kernels (same pointer type) apart
          cudaError_t le = cudaSuccess;
          if (!attr) { le = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = (le == cudaSuccess); }
          if (le != cudaSuccess) return le;
          TailParams tp = r.tlp;
          tp.dbg = want_dbg ? tail_dbg + (r.tail == 2 ? 0 : 128) : nullptr;
          le = launch_k(kern, r.grid, kTailThreads, (size_t)smem, st, g_pdl, 1, r.tmA, r.tmB, r.tmW3, r.tmX, r.tmO, tp);
          if (want_dbg && le == cudaSuccess) {  // debug only: synchronises and prints CTA 0's timeline of its tiles 8..11
            long long h[128];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, tp.dbg, sizeof(h), cudaMemcpyDeviceToHost);
            const long long t0 = h[17];
            fprintf(stderr, "tail mode %d timeline (cycles since tile 8's acc2_full):\n", r.tail);
            for (int t = 0; t < 4; ++t) {
              fprintf(stderr, " tile %d:", 8 + t);
              for (int e = 0; e < 28; ++e) fprintf(stderr, " %lld", h[t * 32 + e] ? h[t * 32 + e] - t0 : -1);
              fprintf(stderr, "\n");
            }
          }
          return le;
        };
        if (r.tail == 2)            e = launch_tail(conv_tail_kernel<true, 1, 2>, TailCfg<true, 1, 2>::kSmemBytes, std::integral_constant<int, 2>{});
        else                        e = launch_tail(conv_tail_kernel<false, 2, 3>, TailCfg<false, 2, 3>::kSmemBytes, std::integral_constant<int, 1>{});
      } else if (r.stem) {
        static bool stem_attr = false;
        if (!stem_attr) {
          e = cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
          stem_attr = (e == cudaSuccess);
        }
        if (e == cudaSuccess && r.stem_pair) {
          static bool pair_attr = false;
          if (!pair_attr) {
            e = cudaFuncSetAttribute(stem_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            pair_attr = (e == cudaSuccess);
          }
          if (e == cudaSuccess)
            e = launch_k(stem_umma_pair_kernel, r.grid, kStemPairThreads, (size_t)r.stem_smem, st, g_pdl, 2, r.tmE, r.tmOdd, r.tmWh, r.tmSO, r.sp);
        } else if (e == cudaSuccess && r.stem_mf) {
          static bool mf_attr = false;
          if (!mf_attr) {
            e = cudaFuncSetAttribute(stem_umma_mf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            mf_attr = (e == cudaSuccess);
          }
          if (e == cudaSuccess) {
            StemMfParams mp;
            mp.s = r.sp;
            static long long* clk_dev = nullptr;
            static const bool want_clk = getenv("VAD_STEM_CLOCKS") != nullptr;
            if (want_clk && !clk_dev) cudaMalloc(&clk_dev, 32);
            mp.s.clk_out = want_clk ? clk_dev : nullptr;
            mp.Ti = r.stem_ti;
            mp.ti_max = r.stem_ti_max;
            e = launch_k(stem_umma_mf_kernel, r.grid, kStemMfThreads, (size_t)r.stem_smem, st, g_pdl, 1, r.tmE, r.tmOdd, r.tmW, r.tmSO, mp);
            if (want_clk && e == cudaSuccess) {  // debug only: synchronises
              long long hclk[3] = {0, 0, 0};
              cudaStreamSynchronize(st);
              cudaMemcpy(hclk, clk_dev, 24, cudaMemcpyDeviceToHost);
              fprintf(stderr, "stem mf: CTA 0 MMA thread %lld cycles in %lld ns = %.0f MHz, %lld units, %.0f cycles/unit\n", hclk[0], hclk[1],
                      hclk[1] ? 1e3 * (double)hclk[0] / (double)hclk[1] : 0.0, hclk[2], hclk[2] ? (double)hclk[0] / (double)hclk[2] : 0.0);
            }
          }
        } else if (e == cudaSuccess) {
          e = launch_k(stem_umma_kernel, r.grid, kStemThreads, (size_t)r.stem_smem, st, g_pdl, 1, r.tmE, r.tmOdd, r.tmW, r.tmSO, r.sp);
        }
      } else if (r.s3) {
        static bool attr = false;
        if (!attr) { e = cudaFuncSetAttribute(conv_s3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS3SmemBytes); attr = (e == cudaSuccess); }
        if (e == cudaSuccess) e = launch_k(conv_s3x3_kernel, r.grid, kS3Threads, (size_t)kS3SmemBytes, st, g_pdl, 1, r.tmA, r.tmB, r.tmO, r.s3p);
      } else if (r.thalo) {
        const int w_all = r.tp.resident ? 3 * (r.tp.Cin / 64) * r.bn * 128 : 0;
        const int smem = w_all + r.tp.n_stages * r.tp.stage_bytes + ThaloCfg<64>::kFixedBytes;
        static bool attr = false;
        if (!attr) { e = cudaFuncSetAttribute(conv_thalo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr = (e == cudaSuccess); }
        if (e == cudaSuccess) e = launch_k(conv_thalo_kernel<64>, r.grid, ThaloCfg<64>::kThreads, (size_t)smem, st, g_pdl, 1, r.tmA, r.tmB, r.tp);
        if (e == cudaSuccess) e = cudaGetLastError();
      } else {
        e = launch_conv_any(r, st);
      }
    } else if (d.kind == VAD_OP_MAXPOOL) {
      const long long total = (long long)r.pp.B * r.pp.To * r.pp.Ho * r.pp.Wo * (r.pp.C / 8);
      const PoolParams& q = r.pp;
      const bool inb = !q.pt && !q.ph && !q.pw && (q.To - 1) * q.st + q.kt <= q.Ti && (q.Ho - 1) * q.sh + q.kh <= q.Hi &&
                       (q.Wo - 1) * q.sw + q.kw <= q.Wi;
      const int g = grid_for(total, 256, 148 * 64);
      if (inb && q.kt == 2 && q.kh == 3 && q.kw == 3)
        maxpool3d_fixed_kernel<2, 3, 3><<<g, 256, 0, st>>>(q);   // I3Res50 maxpool1
      else if (inb && q.kt == 1 && q.kh == 3 && q.kw == 3)
        maxpool3d_fixed_kernel<1, 3, 3><<<g, 256, 0, st>>>(q);   // I3Res50 maxpool1 after the stem's fused temporal max
      else if (inb && q.kt == 2 && q.kh == 1 && q.kw == 1)
        maxpool3d_fixed_kernel<2, 1, 1><<<g, 256, 0, st>>>(q);   // I3Res50 maxpool2
      else if (inb && q.kt == 1 && q.kh == 1 && q.kw == 1)
        maxpool3d_fixed_kernel<1, 1, 1><<<g, 256, 0, st>>>(q);   // strided copy
      else if (q.kt == 3 && q.kh == 3 && q.kw == 3 && q.st == 1 && q.sh == 1 && q.sw == 1 && q.pt == 1 && q.ph == 1 && q.pw == 1 &&
               q.To == q.Ti && q.Ho == q.Hi && q.Wo == q.Wi) {          // Inception branch pools
        maxpool3d_k3s1_kernel<<<grid_for((long long)q.B * q.Hi * q.Wi * (q.C / 8), 256, 148 * 64), 256, 0, st>>>(q);
      } else if ((q.kt == 1 || q.kt == 3) && q.kh == 3 && q.kw == 3 || (q.kt == 2 && q.kh == 2 && q.kw == 2)) {
        // MaxPool3d_2a / 3a ((1,3,3) / (1,2,2)), 4a ((3,3,3) / 2), 5a ((2,2,2) / 2), SAME padding: one block per output row
        const int items = q.Wo * (q.C / 8);
        const int iters = (items + 511) / 512;
        const int threads = ((items + iters - 1) / iters + 31) / 32 * 32;
        const long long rows = (long long)q.B * q.To * q.Ho;
        if (rows > 0x7fffffffLL) return fail(VAD_ERR_INVALID_ARGUMENT, "max-pool: too many output rows for one launch");
        if (q.C < 128 || getenv("VAD_POOL_OLD")) {
          if (q.kt == 1)      maxpool3d_checked_kernel<1, 3, 3><<<g, 256, 0, st>>>(q);
          else if (q.kt == 3) maxpool3d_checked_kernel<3, 3, 3><<<g, 256, 0, st>>>(q);
          else                maxpool3d_checked_kernel<2, 2, 2><<<g, 256, 0, st>>>(q);
        } else if (q.kt == 1) maxpool3d_rows_kernel<1, 3, 3><<<(int)rows, threads, 0, st>>>(q);
        else if (q.kt == 3)   maxpool3d_rows_kernel<3, 3, 3><<<(int)rows, threads, 0, st>>>(q);
        else                  maxpool3d_rows_kernel<2, 2, 2><<<(int)rows, threads, 0, st>>>(q);
      }
      else
        maxpool3d_kernel<<<g, 256, 0, st>>>(q);
      e = cudaGetLastError();
    } else {
      if (!feat_out_dev) return fail(VAD_ERR_INVALID_ARGUMENT, "plan ends in AVGPOOL but feat_out_dev is null");
      const uint8_t* src = d.src == 0 ? static_cast<const uint8_t*>(x_dev)
                                      : static_cast<const uint8_t*>(workspace_dev) + p->slots[d.src].offset;
      const long long warps = (long long)p->batch * (r.avg_C / 64);
      avgpool_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), p->batch,
                                                                     r.avg_P, r.avg_C, feat_out_dev, r.avg_HW, r.avg_kt);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return fail(VAD_ERR_CUDA, "op %zu launch failed: %s", i, cudaGetErrorString(e));
    if ((int)i == pf1 - 1 && mark() != cudaSuccess) return fail(VAD_ERR_CUDA, "profiling event failed");
    if (p->profiling && (int)i >= pf0 && (int)i < pf1) { p->prof_flops[i] += p->op_flops[i]; p->prof_bytes[i] += p->op_bytes[i]; }
  }
  return VAD_OK;
}

extern "C" int32_t vad_plan_forward(vad_plan_t* p, const void* x_dev, void* workspace_dev, uint64_t workspace_bytes,
                                    float* feat_out_dev, void* stream) {
  if (!p || !p->configured) return fail(VAD_ERR_NOT_CONFIGURED, "vad_plan_forward: plan is not configured");
  if (!x_dev || (!workspace_dev && p->ws_bytes)) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_plan_forward: null pointer");
  if (workspace_bytes < p->ws_bytes)
    return fail(VAD_ERR_WORKSPACE_TOO_SMALL, "workspace %llu < required %llu", (unsigned long long)workspace_bytes,
                (unsigned long long)p->ws_bytes);
  if (((uintptr_t)x_dev & 15) || ((uintptr_t)workspace_dev & 1023))
    return fail(VAD_ERR_INVALID_ARGUMENT, "x must be 16 B aligned and the workspace 1024 B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->bound_x != x_dev || p->bound_ws != workspace_dev) {
    int32_t rc = bind_plan(p, x_dev, workspace_dev, st);
    if (rc != VAD_OK) return rc;
  }
  const char* graph_env = getenv("VAD_GRAPH");   // 0: never, 1: any batch; unset: batch <= 32
  const int gmode = graph_env ? atoi(graph_env) : -1;
  const bool want_graph = !p->profiling && !p->graph_failed && !getenv("VAD_TAIL_DEBUG") && !getenv("VAD_STEM_CLOCKS") &&
                          (gmode == 1 || (gmode < 0 && p->batch <= 32));
  if (want_graph && p->graph_exec && p->graph_feat == feat_out_dev) {
    VAD_CUDA_CHECK(cudaGraphLaunch(p->graph_exec, st));
    return VAD_OK;
  }
  if (want_graph && p->graph_exec && ++p->feat_misses > 2) {
    // the caller hands out a different feature pointer every forward: a graph bakes it in, so stop re-capturing
    drop_graph(p);
    p->graph_failed = true;
    ++p->direct_runs;
    return run_ops(p, x_dev, workspace_dev, feat_out_dev, st);
  }
  if (want_graph && p->direct_runs >= 1) {
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    if (!p->cap_stream) VAD_CUDA_CHECK(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
      const int32_t rc = run_ops(p, x_dev, workspace_dev, feat_out_dev, p->cap_stream);
      ce = cudaStreamEndCapture(p->cap_stream, &graph);
      if (rc != VAD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    }
    if (ce == cudaSuccess) ce = cudaGraphInstantiate(&p->graph_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce == cudaSuccess) {
      p->graph_feat = feat_out_dev;
      VAD_CUDA_CHECK(cudaGraphLaunch(p->graph_exec, st));
      return VAD_OK;
    }
    // capture not possible here (old driver, ...): remember and launch op by op from now on
    cudaGetLastError();
    p->graph_exec = nullptr;
    p->graph_failed = true;
  }
  ++p->direct_runs;
  return run_ops(p, x_dev, workspace_dev, feat_out_dev, st);
}

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
