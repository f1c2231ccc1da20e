// vad_plan_configure: shapes, slot layout, kernel selection and launch geometry of every op for one (batch, T, H, W)
// (part of vad_api.cu: included there, after the plan structures; not a stand-alone translation unit)
#pragma once

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
      Cdst = d.dst_c_total ? d.dst_c_total : (d.dst1 > 0 ? d.seg_w0 : d.cout);
      if (d.dst1 <= 0 && d.dst_c_off + d.cout > Cdst) return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: channel slice exceeds dst_c_total", i);
      const bool multi = d.dst1 > 0;   // fused sibling 1x1x1 convs: output columns routed to up to three slots
      if (multi) {
        const bool unit1 = d.kt == 1 && d.kh == 1 && d.kw == 1 && d.st == 1 && d.sh == 1 && d.sw == 1;
        const int s2 = d.dst2 > 0 ? d.split2 : d.cout;
        if (!unit1 || d.res >= 0 || fold || d.dst1 >= p->n_slots || d.dst2 >= p->n_slots || d.dst1 == d.src || d.dst2 == d.src ||
            d.dst1 == d.dst || (d.dst2 > 0 && (d.dst2 == d.dst || d.dst2 == d.dst1)) || d.split1 <= 0 || d.split1 % 64 || s2 % 64 && d.dst2 > 0 ||
            d.split1 >= s2 || s2 > d.cout || (d.dst2 > 0 && s2 >= d.cout) || d.seg_w0 <= 0 || d.seg_w0 > d.split1 || d.seg_w0 % 8 ||
            d.seg_w1 <= 0 || d.seg_w1 > s2 - d.split1 || d.seg_w1 % 8 || (d.dst2 > 0 && (d.seg_w2 <= 0 || d.seg_w2 > d.cout - s2 || d.seg_w2 % 8)) ||
            d.dst_c_off + d.seg_w0 > Cdst || (d.flags & VAD_FLAG_POOL_T2))
          return fail(VAD_ERR_INVALID_ARGUMENT, "op %zu: bad fused-sibling description (1x1x1 stride 1, no residual; parts start on multiples of "
                      "64 columns, widths multiples of 8 inside their parts, distinct slots)", i);
      }
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
      // HBM-bound plain 1x1x1 convs with one k-block and few output channels (Inception's Conv3d_2b_1x1, 64 -> 64 over 4 M rows:
      // 0.247 -> 0.207 ms per 160 clip-crops through the staged epilogue).  Measured with larger limits as well: the pooling-branch
      // projections (K = 192 .. 832) gain nothing or lose (Mixed_3b.b3b 0.082 -> 0.090 ms), so the default stops at K = 64.
      const int epi_unit_max_k = getenv("VAD_EPI_UNIT_MAXK") ? atoi(getenv("VAD_EPI_UNIT_MAXK")) : 64;
      const bool epi_unit = unit && d.res < 0 && !multi && d.cout <= 128 && (d.cin % 64) == 0 && K <= epi_unit_max_k &&
                            !(d.flags & (VAD_FLAG_FORCE_GATHER | VAD_FLAG_POOL_T2));
      const bool epi_wanted = !multi && !p->no_epi && (d.res >= 0 || epi_unit || (d.cout >= 128 && K <= 256 && d.cout >= 2 * K));
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
      r.epi = epi_wanted;
      // fused siblings: the staged TMA-store epilogue (one tensor map per part) where the k-blocks are 64 wide; the direct epilogue
      // (parts routed per 32-column chunk; ncu: 7,900 cycles of epilogue latency per 128 x 256 tile) for the 32- / 16-wide k-block layers
      if (multi) r.epi = !p->no_epi && r.bk == 64 && r.a_mode != A_GATHER;
      r.bn = (d.cout > 128 && !r.epi && r.bk == 64 && !multi) ? 256 : (d.cout > 64 ? 128 : 64);   // fused siblings: 128-wide tiles waste less of a ragged N
      if (multi) {
        c.split1 = d.split1; c.split2 = d.dst2 > 0 ? d.split2 : d.cout;
        c.seg_w0 = d.seg_w0; c.seg_w1 = d.seg_w1; c.seg_w2 = d.dst2 > 0 ? d.seg_w2 : 0;
        c.ldo1 = d.seg_w1; c.ldo2 = d.seg_w2;
      }
      r.pair_epi = !multi && r.epi && p->pair_mode > 0 && p->pair_epi_min_kb > 0 && d.res >= 0 && r.a_mode != A_GATHER && r.bk == 64 && d.cout % 256 == 0 &&
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
      r.pair = !multi && p->pair_mode > 0 && (r.bn == 256 || r.bn == 128) && !r.epi && r.a_mode != A_GATHER && r.bk == 64 && !r.thalo && !r.s3 &&
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
          // pair kernel: both tiles of an item must share their output frame for the out-of-clip frame taps to be skipped
          q.Ti = (r.stem_pair && (q.tiles_w * q.tiles_h) % 2 == 0) ? src.T : 0;
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
      const int cout_real = d.dst1 > 0 ? d.seg_w0 + d.seg_w1 + (d.dst2 > 0 ? d.seg_w2 : 0) : d.cout;   // fused siblings: without the alignment gaps
      p->op_flops[i] = 2.0 * (double)M * cout_real * d.kt * d.kh * d.kw * cin_real;  // frames the reference conv produces
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
    if (d.kind == VAD_OP_CONV && d.dst1 > 0) {   // the sibling outputs of a fused 1x1x1 conv: tensors of exactly seg_w channels
      const int extra[2][2] = {{d.dst1, d.seg_w1}, {d.dst2, d.seg_w2}};
      for (int e = 0; e < 2; ++e) {
        if (extra[e][0] <= 0) continue;
        SlotInfo& ds = p->slots[extra[e][0]];
        ds.T = To; ds.H = Ho; ds.W = Wo; ds.C = extra[e][1]; ds.defined = true;
        const uint64_t b2 = (uint64_t)batch * To * Ho * Wo * extra[e][1] * 2;
        if (b2 > ds.bytes) ds.bytes = b2;
      }
    }
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

