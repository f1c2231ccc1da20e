/*
 * vad_b200.h -- C ABI of libvad_b200.so: the B200 (sm_100a) hot path of
 * jinmang2/anomaly_detection_on_video (I3D snippet-feature extraction).
 *
 * The reference has no FFI: its boundary is plain Python call sites.  Every entry point below
 * names the reference call site(s) it replaces (file:line relative to the reference root) so a
 * maintainer can bind it with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every function returns VAD_OK (0) or a negative vad_status; the message for the calling
 *     thread's last failure is available from vad_last_error()
 *   - pointers named *_dev are device pointers on the handle's device, owned by the caller
 *     (PyTorch in our host code); the library borrows them for the duration of the call
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it and no
 *     call synchronises the device
 *   - handles are bound to one device, are not thread-safe, and own only host memory plus (for the
 *     preprocessing handle) a few KB of resampling-coefficient tables
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute call fails
 *     with VAD_ERR_CUDA / VAD_ERR_UNSUPPORTED_DEVICE
 */
#ifndef VAD_B200_H_
#define VAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAD_ABI_VERSION 1

typedef enum vad_status {
  VAD_OK = 0,
  VAD_ERR_INVALID_ARGUMENT = -1,
  VAD_ERR_CUDA = -2,
  VAD_ERR_UNSUPPORTED_DEVICE = -3,
  VAD_ERR_WORKSPACE_TOO_SMALL = -4,
  VAD_ERR_NOT_CONFIGURED = -5,
  VAD_ERR_DRIVER_SYMBOL = -6
} vad_status;

/* Message describing the last failure on the calling thread ("" if none). Never NULL. */
const char* vad_last_error(void);
int32_t vad_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * Backbone plan: a list of ops over numbered activation slots.
 *
 * Replaces the torch.nn graph of I3Res50 (src/i3d.py:198-318: conv1/bn1/relu/maxpool1, Bottleneck
 * stacks src/i3d.py:60-121 incl. downsample src/i3d.py:262-272, maxpool2, avgpool) that
 * extract_features.py:88 calls as `model(crop)`.  The same op list expresses an InceptionV1-3D
 * (Unit3D / MaxPool3dSamePadding / channel-slice concat) -- only the table differs.
 *
 * Activations are channels-last bf16: slot tensor = [batch, T, H, W, C] with C a multiple of 8.
 * Slot 0 is the network input in *stem layout*: [batch, T, H, W + 8, 4] bf16 (RGB + one zero
 * channel; `in_pad_left` zero pixels on the left of every row, 8 - in_pad_left on the right), which
 * is what vad_preproc_run(out_mode = VAD_OUT_STEM_BF16) and vad_ingest_ncthw_f32() produce.
 * ---------------------------------------------------------------------------------------------- */

enum { VAD_OP_CONV = 0, VAD_OP_MAXPOOL = 1, VAD_OP_AVGPOOL = 2 };
/* AVGPOOL: global mean over (T, H, W) of the src slot -> fp32 features.  With 1 < kt < T it is AvgPool3d((kt, H, W), stride 1)
 * followed by a global mean over the T - kt + 1 windows (the head the reference puts on pytorchvideo's I3D-R50,
 * src/i3d.py:21-57): frame t is weighted by the number of windows covering it. */

/* vad_op_desc.flags */
enum {
  VAD_FLAG_RELU = 1,        /* ReLU after (scale,shift) and the optional residual add               */
  VAD_FLAG_STEM_FOLD_W = 2, /* conv reads slot 0 in stem layout; the kw taps of one row are folded
                               into the contraction dim (window of 8 px x 4 ch, 64-byte aligned)    */
  VAD_FLAG_POOL_SAME = 4,   /* TF "SAME" padding for max-pool (MaxPool3dSamePadding); pad value 0   */
  VAD_FLAG_FORCE_GATHER = 8,/* conv: feed the A operand with the cp.async gather producer instead
                               of TMA (debug / comparison; results are identical)                   */
  VAD_FLAG_POOL_T2 = 16,    /* stem conv (STEM_FOLD_W): max over output frames (2k, 2k+1) fused into
                               the epilogue, i.e. the temporal half of a following MaxPool3d with
                               kt = st = 2, pt = 0 (reference src/i3d.py:212-214); dst has To / 2    */
  VAD_FLAG_CONV_SAME = 32,  /* conv: TF "SAME" padding (InceptionI3d's Unit3D): out = ceil(in / stride),
                               front pad = total / 2, back pad = total - front; pt/ph/pw are ignored */
  VAD_FLAG_STEM_PLANES = 64 /* TF32 plans only, with STEM_FOLD_W: slot 0 is the column-parity plane layout of
                               vad_tf32_ingest_ncthw_planes and the stem runs on the dedicated tcgen05 kind::tf32
                               stem kernel (stride 2 in h and w, cout = 64, kt * kh <= 36, kw <= 8, pw <= 3)   */
};

typedef struct vad_op_desc {
  int32_t kind;           /* VAD_OP_*                                                               */
  int32_t src;            /* input slot                                                             */
  int32_t dst;            /* output slot (ignored by AVGPOOL: it writes the fp32 feature output)    */
  int32_t res;            /* CONV: slot added before ReLU (`out += residual`, src/i3d.py:115), or -1 */
  int32_t cin;            /* CONV: input channels as stored (multiple of 8; 4 with STEM_FOLD_W)     */
  int32_t cout;           /* CONV: output channels (multiple of 8)                                  */
  int32_t kt, kh, kw;     /* kernel                                                                 */
  int32_t st, sh, sw;     /* stride                                                                 */
  int32_t pt, ph, pw;     /* zero padding (front == back), ignored with VAD_FLAG_POOL_SAME          */
  int32_t flags;          /* VAD_FLAG_*                                                             */
  int32_t dst_c_off;      /* first channel written inside the dst slot (Inception concat), else 0   */
  int32_t dst_c_total;    /* channel count of the dst slot tensor; 0 means `cout` (or C of src)     */
  uint64_t w_off;         /* CONV: byte offset of bf16 weights [cout][K_pad] in the parameter blob; K
                             order is (kt, kh, kw, cin), K_pad = K rounded up to 64, zero padded    */
  uint64_t scale_off;     /* CONV: byte offset of fp32 scale[cout]  (gamma / sqrt(var + eps))       */
  uint64_t shift_off;     /* CONV: byte offset of fp32 shift[cout]  (beta - mean * scale)           */
  /* Sibling 1x1x1 convs over the same input run as ONE conv (bf16 plans; InceptionI3d's b0 | b1a | b2a): the weight matrix
   * stacks their output channels, each sibling starting on a multiple of 64 columns, and the output columns are routed to up
   * to three destinations.  dst1 > 0 switches it on:
   *   columns [0, split1)      -> slot dst (dst_c_off / dst_c_total as usual), only the first seg_w0 of them are stored
   *   columns [split1, split2) -> slot dst1 (a tensor of seg_w1 channels)
   *   columns [split2, cout)   -> slot dst2 (a tensor of seg_w2 channels); dst2 <= 0: no third part, split2 == cout
   * Columns between a sibling's width and the next start carry zero weights and are not stored.  All zero: ordinary conv. */
  int32_t dst1, dst2;
  int32_t split1, split2;
  int32_t seg_w0, seg_w1, seg_w2;
  int32_t reserved0;
} vad_op_desc;

typedef struct vad_plan vad_plan_t;

/* Build a plan.  `params_dev` (bf16 weights + fp32 scale/shift, laid out as the op offsets say) is
 * borrowed for the life of the plan.  Replaces I3Res50.__init__ / _make_layer (src/i3d.py:199-300)
 * plus load_state_dict (src/i3d.py:356-359) on the device side.
 * in_channels == 0: slot 0 is the stem layout described above (the production case).
 * in_channels  > 0: slot 0 is a plain [batch, T, H, W, in_channels] bf16 tensor (multiple of 8);
 *                   used to run any sub-graph (e.g. a single conv) on its own. */
int32_t vad_plan_create(vad_plan_t** plan, const vad_op_desc* ops, int32_t n_ops, int32_t n_slots,
                        const void* params_dev, uint64_t params_bytes, int32_t in_channels,
                        int32_t in_pad_left, int32_t device);

/* Fix the problem size (clips per forward, frames, crop height/width); infers every slot shape and
 * returns the workspace the caller must provide to vad_plan_forward. */
int32_t vad_plan_configure(vad_plan_t* plan, int32_t batch, int32_t t, int32_t h, int32_t w,
                           uint64_t* workspace_bytes);

/* Run the op list: x_dev (slot 0, stem layout) -> feat_out_dev fp32 [batch, C_last] when the plan
 * ends in AVGPOOL.  Replaces `model(crop)` at extract_features.py:86-89 (I3Res50.forward,
 * src/i3d.py:302-318) for `batch` clip-crops at once. */
int32_t vad_plan_forward(vad_plan_t* plan, const void* x_dev, void* workspace_dev,
                         uint64_t workspace_bytes, float* feat_out_dev, void* stream);

/* Shape / location of a slot after configure (for tests and for chaining): dims = {T,H,W,C},
 * byte offset inside the workspace (slot 0 lives outside the workspace: offset = UINT64_MAX). */
int32_t vad_plan_slot_info(const vad_plan_t* plan, int32_t slot, int32_t dims[4], uint64_t* offset,
                           uint64_t* bytes);

/* Number of kernels one vad_plan_forward launches (for the bench's gpu_launches count). */
int32_t vad_plan_num_launches(const vad_plan_t* plan);

/* Conv FLOPs (2 * MAC, useful taps only) of one forward at the configured size. */
double vad_plan_flops(const vad_plan_t* plan);

/* Per-op device timing for roofline reports.  Between begin and end every vad_plan_forward brackets
 * each launch with CUDA events on the launching stream (no synchronisation is added); end waits for
 * the last event and returns, per op, totals over the timed launches: milliseconds, launch count,
 * useful FLOPs and algorithmic bytes (each input / output tensor touched once). */
int32_t vad_plan_profile_begin(vad_plan_t* plan);
/* Restrict the events to ops [first_op, first_op + n_ops) (n_ops < 0: all ops again).  Event records between kernels
 * keep consecutive kernels from overlapping their prologues (programmatic dependent launch), so a timed run that only
 * needs the dominant kernel's duration brackets just that one. */
int32_t vad_plan_profile_select(vad_plan_t* plan, int32_t first_op, int32_t n_ops);
int32_t vad_plan_profile_end(vad_plan_t* plan, int32_t n_ops, double* op_ms_sum, int32_t* op_calls,
                             double* op_flops, double* op_bytes);

void vad_plan_destroy(vad_plan_t* plan);

/* fp32 NCTHW clips (what the reference feeds its model: extract_features.py:83-86) -> stem layout.
 * x_dev: [batch, 3, T, H, W] fp32.  out_dev: [batch, T, H, W + 8, 4] bf16. */
int32_t vad_ingest_ncthw_f32(const float* x_dev, int32_t batch, int32_t t, int32_t h, int32_t w,
                             int32_t pad_left, void* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused clip preprocessing.  Replaces src/gtransforms.py:9-73,115-132 (GroupResize, GroupTenCrop,
 * ToTensorTenCrop, GroupStandardizationTenCrop, LoopPad) and the permutes at src/dataset.py:188-195
 * / extract_features.py:83.  The resize is Pillow's two-pass fixed-point bilinear resample, so
 * the fp32 output is bit-identical to the reference's.
 * ---------------------------------------------------------------------------------------------- */
enum {
  VAD_OUT_DATASET_F32 = 0, /* [n_clips, ncrops, frames_per_clip, 3, crop, crop] fp32: the stacked
                              TenCropVideoFrameDataset.__getitem__ tensors (src/dataset.py:188-195) */
  VAD_OUT_STEM_BF16 = 1    /* [n_clips * ncrops, frames_per_clip, crop, crop + 8, 4] bf16 stem layout */
};

typedef struct vad_preproc vad_preproc_t;

/* ncrops is 10 (TenCrop order: tl, tr, bl, br, center, then the same five of the h-flipped image)
 * or 1 (center crop == TenCrop index 4). */
int32_t vad_preproc_create(vad_preproc_t** pp, int32_t src_h, int32_t src_w, int32_t resize,
                           int32_t crop, int32_t ncrops, int32_t device);
/* resized_hw = {H, W} after GroupResize; tops/lefts hold ncrops entries (offsets in the resized,
 * un-flipped image of the pixel block each crop reads; flipped crops read it right-to-left). */
int32_t vad_preproc_info(const vad_preproc_t* pp, int32_t resized_hw[2], int32_t* tops, int32_t* lefts,
                         int32_t* flips);
/* frames_dev: [n_frames, src_h, src_w, 3] uint8.  Clip i covers frames [i*fpc, (i+1)*fpc); a short
 * last clip is loop-padded (frame j <- frame j mod L).  Processes clips
 * [clip_start, clip_start + n_clips). */
int32_t vad_preproc_run(vad_preproc_t* pp, const uint8_t* frames_dev, int32_t n_frames,
                        int32_t clip_start, int32_t n_clips, int32_t frames_per_clip, int32_t out_mode,
                        int32_t pad_left, void* out_dev, void* stream);
void vad_preproc_destroy(vad_preproc_t* pp);

/* ------------------------------------------------------------------------------------------------
 * 32-segment averaging.  Replaces segment() (extract_features.py:159-185): for each crop,
 * r = linspace(0, n_clips, seg+1, dtype=int); out[i] = mean(f[r[i]:r[i+1]]) or f[r[i]] if empty.
 * feats_dev: [n_clips, ncrops, C] fp32 -> out_dev: [ncrops, seg_length, C] fp32.  Clips are added in
 * index order and divided once, which reproduces np.mean bit for bit.
 * ---------------------------------------------------------------------------------------------- */
int32_t vad_segment_mean(const float* feats_dev, int32_t n_clips, int32_t ncrops, int32_t c,
                         int32_t seg_length, float* out_dev, void* stream);

/* FeatureDataset.add_magnitude (src/dataset.py:121-124): [rows, C] fp32 -> [rows, C + 1] with the
 * L2 norm of each row appended. */
int32_t vad_add_magnitude(const float* feats_dev, int64_t rows, int32_t c, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MGFN scoring head.  Replaces MGFNForVideoAnomalyDetection.forward in eval mode
 * (src/models/mgfn/modeling_mgfn.py:36-427: feature amplifier, Glance / Focus blocks, final LayerNorm +
 * fc + sigmoid, magnitude selection and score prediction) and the losses of src/loss/base.py:7-48 and
 * src/loss/mgfn.py:7-47.  fp32 activations; every Conv1d is a tcgen05 kind::tf32 GEMM.  Training
 * (backward, dropout on the selection mask, the optimizer step of src/runner.py:29-59) is not built.
 * ---------------------------------------------------------------------------------------------- */
enum { VAD_HEAD_GLANCE = 0, VAD_HEAD_FOCUS = 1 };

typedef struct vad_head_config {
  int32_t channels;          /* feature width without the magnitude column (2048)                  */
  int32_t n_stages;          /* <= 4                                                               */
  int32_t dims[4];           /* stage widths (64, 128, 1024); multiples of 64                      */
  int32_t depths[4];         /* blocks per stage (3, 3, 2)                                         */
  int32_t types[4];          /* VAD_HEAD_GLANCE / VAD_HEAD_FOCUS per stage                         */
  int32_t dim_head;          /* 64 (the attention kernel is specialised for it)                    */
  int32_t ff_repe;           /* FFN expansion (4)                                                  */
  int32_t local_aggr_kernel; /* Focus relation kernel (5)                                          */
  int32_t k;                 /* top-k snippets (3)                                                 */
  float mag_ratio;           /* 0.1                                                                */
  float ln_eps;              /* 1e-5 for MGFNLayerNorm and nn.LayerNorm                            */
} vad_head_config;

typedef struct vad_head vad_head_t;

/* params_dev: fp32 blob, tensors in module order, each starting on a 64-float boundary:
 *   amplifier   to_tokens W[d0][3][channels] (tap-major contraction), b[d0]; to_mag w[d0][3], b[d0]
 *   per block   scc W[d][3][d], b[d]
 *     glance    norm g[d], b[d]; to_qkv W[3*inner][d]; to_out W[d][inner], b[d]
 *     focus     to_v W[inner][d], b[inner] (BatchNorm1d folded in); rel_pos w[heads][k], b[heads];
 *               to_out W[d][inner], b[d]
 *     ffn       layer_norm g[d], b[d]; in_conv W[r*d][d], b[r*d]; out_conv W[d][r*d], b[d]
 *   after every stage but the last: layer_norm g[d], b[d]; conv W[d_next][d], b[d_next]
 *   final       layer_norm w[d], b[d]; fc w[d], b[1]
 * vad_head_create fails unless params_bytes equals the size this layout implies. */
int32_t vad_head_create(vad_head_t** head, const vad_head_config* cfg, const float* params_dev,
                        uint64_t params_bytes, int32_t device);
/* workspace for n_seq = videos * ncrops sequences of t snippets */
int32_t vad_head_workspace_bytes(const vad_head_t* head, int32_t n_seq, int32_t t, uint64_t* bytes);
/* video_dev [n_videos, ncrops, t, channels + 1] fp32 ->
 *   xln_dev   [n_videos * ncrops, t, d_last]  layer-normed snippet features
 *   score_dev [n_videos * ncrops, t]          per-crop snippet scores (sigmoid)
 *   fmag_dev  [n_videos * ncrops, t]          L2 norms of xln                                      */
int32_t vad_head_forward(vad_head_t* head, const float* video_dev, int32_t n_videos, int32_t ncrops,
                         int32_t t, void* workspace_dev, uint64_t workspace_bytes, float* xln_dev,
                         float* score_dev, float* fmag_dev, void* stream);
/* magnitude selection + score prediction (eval: no dropout on the mask) for videos
 * [video_off, video_off + n_sel) of an n_videos batch:
 *   scores_dev [n_videos, t] crop-mean snippet scores (rows of the selected videos are written)
 *   vid_score_dev [n_videos], idx_dev [n_videos, k]
 *   sel_dev [ncrops, n_sel, k, d_last] gathered features, crop-major like the reference's torch.cat loop */
int32_t vad_head_select(const vad_head_t* head, const float* xln_dev, const float* score_dev,
                        const float* fmag_dev, int32_t n_videos, int32_t ncrops, int32_t t,
                        int32_t video_off, int32_t n_sel, float* scores_dev, float* vid_score_dev,
                        int32_t* idx_dev, float* sel_dev, void* stream);
/* loss = MGFNLoss + TemporalSmoothnessLoss + SparsityLoss for a batch whose first n_half videos are normal and
 * last n_half abnormal.  labels_dev [2 * n_half] (normal labels then abnormal labels); scratch_dev holds
 * 2 * ncrops * n_half * k floats; out_dev[7] = {total, smooth, sparsity, bce, con, con_n, con_a}. */
int32_t vad_head_loss(const vad_head_t* head, const float* scores_dev, const float* vid_score_dev,
                      const float* labels_dev, const float* sel_normal_dev, const float* sel_abnormal_dev,
                      int32_t n_half, int32_t ncrops, int32_t t, float* scratch_dev, float* out_dev, void* stream);
int32_t vad_head_num_launches(const vad_head_t* head);
double vad_head_flops(const vad_head_t* head, int32_t n_seq, int32_t t);
void vad_head_destroy(vad_head_t* head);

/* ------------------------------------------------------------------------------------------------
 * MGFN training step: replaces MGFNRunner.training_step + loss.backward() + the Adam step Lightning drives from
 * configure_optimizers (src/runner.py:29-39,53-59) over MGFNForVideoAnomalyDetection.forward in train() mode
 * (modeling_mgfn.py:302-427: BatchNorm1d on batch statistics, dropout on the magnitude-selection mask) and the losses
 * of src/loss/base.py:7-48, src/loss/mgfn.py:7-47.
 *
 * Parameters, gradients and the two Adam moments are flat fp32 blobs of vad_head_train_param_floats() floats in the
 * layout of vad_head_create, except that a Focus block stores its raw to_v weight and BatchNorm1d affine terms:
 *     focus     to_v W[inner][d]; norm weight[d], bias[d]; rel_pos w[heads][k], b[heads]; to_out ...
 * The BatchNorm running statistics (not trained) live in a separate buffer of vad_head_train_bn_floats() floats:
 * running_mean[d] | running_var[d] per Focus block, in block order; a step updates them with momentum 0.1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vad_head_train vad_head_train_t;
int32_t vad_head_train_create(vad_head_train_t** head, const vad_head_config* cfg, int32_t device);
uint64_t vad_head_train_param_floats(const vad_head_train_t* head);
uint64_t vad_head_train_bn_floats(const vad_head_train_t* head);
int32_t vad_head_train_workspace_bytes(const vad_head_train_t* head, int32_t n_videos, int32_t ncrops, int32_t t,
                                       uint64_t* bytes);
/* Forward (train mode) + loss + backward for a batch whose first n_videos / 2 bags are normal and the rest abnormal
 * (src/runner.py:31-33).  video_dev [n_videos, ncrops, t, channels + 1]; labels_dev [n_videos] (normal then abnormal);
 * mask_dev [n_videos, t]: the dropout mask of the magnitude selection (0 or 1 / (1 - p); modeling_mgfn.py:341-344), drawn
 * by the caller so that the step is reproducible, or NULL for none.  grads_dev is OVERWRITTEN with d loss / d params.
 * loss_cfg_host: NULL for the reference's constants, else HOST floats {smoothness weight 8e-4, sparsity weight 8e-3,
 * alpha 0.001, margin 200} (src/loss/base.py:9,24; src/loss/mgfn.py:9-11).
 * loss_terms_dev[7] as vad_head_loss; scores_out_dev [n_videos, t] and idx_out_dev [n_videos, k] are optional.
 * Every contraction (forward, input gradients, weight gradients) is a tcgen05 kind::tf32 GEMM; t <= 64, t % 4 == 0 and
 * n_videos * ncrops * t % 32 == 0 (the training bags are 32 segments, configs/data/default.yaml). */
int32_t vad_head_train_step(vad_head_train_t* head, const float* params_dev, float* grads_dev, float* bn_stats_dev,
                            const float* video_dev, int32_t n_videos, int32_t ncrops, int32_t t,
                            const float* labels_dev, const float* mask_dev, const float* loss_cfg_host,
                            void* workspace_dev, uint64_t workspace_bytes, float* loss_terms_dev,
                            float* scores_out_dev, int32_t* idx_out_dev, void* stream);
int32_t vad_head_train_num_launches(const vad_head_train_t* head);
void vad_head_train_destroy(vad_head_train_t* head);
/* torch.optim.Adam semantics (configure_optimizers, src/runner.py:53-59: lr 1e-3, weight_decay 5e-4 added to the
 * gradient) fused over flat blobs of n floats; `step` counts from 1; grads are multiplied by grad_scale first
 * (1 / world_size after a sum all-reduce). */
int32_t vad_adam_step(float* params_dev, const float* grads_dev, float* m_dev, float* v_dev, uint64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * TF32 precision mode of the backbone (BASELINE.json: "bf16 vs TF32 modes"; features within 1e-3 of the
 * reference's fp32 path).  Same op table and slot / channel-slice semantics as vad_plan_*, with
 *   - every slot a plain fp32 [batch, T, H, W, C] tensor (slot 0: the caller's input, C = in_channels,
 *     a multiple of 4 -- RGB padded with a zero channel, see vad_tf32_ingest_ncthw),
 *   - weights fp32 [cout][K_pad], K = (dt, dh, dw, cin) padded to a multiple of 32, scale / shift fp32,
 *   - no STEM_FOLD_W / POOL_T2 flags (the unfused op list).
 * Replaces the same reference code as vad_plan_* (src/i3d.py:199-318) in the higher-precision mode. */
typedef struct vad_tf32_plan vad_tf32_plan_t;
int32_t vad_tf32_plan_create(vad_tf32_plan_t** plan, const vad_op_desc* ops, int32_t n_ops, int32_t n_slots,
                             const void* params_dev, uint64_t params_bytes, int32_t in_channels, int32_t device);
int32_t vad_tf32_plan_configure(vad_tf32_plan_t* plan, int32_t batch, int32_t t, int32_t h, int32_t w,
                                uint64_t* workspace_bytes);
int32_t vad_tf32_plan_forward(vad_tf32_plan_t* plan, const void* x_dev, void* workspace_dev,
                              uint64_t workspace_bytes, float* feat_out_dev, void* stream);
int32_t vad_tf32_plan_slot_info(const vad_tf32_plan_t* plan, int32_t slot, int32_t* dims4, uint64_t* offset,
                                uint64_t* bytes);
int32_t vad_tf32_plan_num_launches(const vad_tf32_plan_t* plan);
double vad_tf32_plan_flops(const vad_tf32_plan_t* plan);
void vad_tf32_plan_destroy(vad_tf32_plan_t* plan);
/* fp32 [batch, 3, T, H, W] (the tensor the reference hands its model, extract_features.py:86) ->
 * fp32 [batch, T, H, W, 4] with a zero fourth channel: slot 0 of a TF32 plan with in_channels = 4. */
int32_t vad_tf32_ingest_ncthw(const float* x_dev, int32_t batch, int32_t t, int32_t h, int32_t w, float* out_dev,
                              void* stream);
/* The same hand-off for a plan whose stem carries VAD_FLAG_STEM_PLANES: fp32 [batch, 3, T, H, W] ->
 * fp32 [batch, T, H, 2, (W + 8) / 2, 4]: pixel x of a row sits at padded index xp = x + 3, in plane xp & 1 at position xp >> 1
 * (zero fourth channel, zero pad pixels; W even).  In each plane the 8-pixel windows of consecutive stride-2 output columns
 * start 16 bytes apart, which is what lets the tensor core read them straight out of raw row segments. */
int32_t vad_tf32_ingest_ncthw_planes(const float* x_dev, int32_t batch, int32_t t, int32_t h, int32_t w, float* out_dev,
                                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAD_B200_H_ */
