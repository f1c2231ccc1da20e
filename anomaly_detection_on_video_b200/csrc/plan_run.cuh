// vad_plan_forward: per-kernel launch helpers, the op loop (run_ops) and the CUDA-graph replay of small batches
// (part of vad_api.cu: included there, after the plan structures; not a stand-alone translation unit)
#pragma once

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
  return launch_k(conv_umma_kernel<BN, BK, KPS, GATHER, EPI>, r.grid, Cfg::kThreads, Cfg::kSmemBytes, st, g_pdl, 1, r.tmA, r.tmB, r.tmR, r.tmO, r.tmO2, r.cp);
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
      } else if (((q.kt == 1 || q.kt == 3) && q.kh == 3 && q.kw == 3 || (q.kt == 2 && q.kh == 2 && q.kw == 2)) &&
                 q.pt < q.kt && q.ph < q.kh && q.pw < q.kw && (q.To - 1) * q.st - q.pt < q.Ti &&
                 (q.Ho - 1) * q.sh - q.ph < q.Hi && (q.Wo - 1) * q.sw - q.pw < q.Wi) {   // every window meets the frame
        // MaxPool3d_2a / 3a ((1,3,3) / (1,2,2)), 4a ((3,3,3) / 2), 5a ((2,2,2) / 2), SAME padding: one block per output row
        const int items = q.Wo * (q.C / 8);
        const int iters = (items + 511) / 512;
        const int threads = ((items + iters - 1) / iters + 31) / 32 * 32;
        const long long rows = (long long)q.B * q.To * q.Ho;
        if (rows > 0x7fffffffLL) return fail(VAD_ERR_INVALID_ARGUMENT, "max-pool: too many output rows for one launch");
        // measured on B200 (160 clip-crops): the row kernel wins only for the wide (1,3,3) pool (3a: 0.38 vs 0.41 ms);
        // the clamped grid-stride kernel wins for 2a (C = 64), 4a (0.27 vs 0.33 ms) and 5a (0.038 vs 0.055 ms)
        const bool by_rows = getenv("VAD_POOL_ROWS") ? atoi(getenv("VAD_POOL_ROWS")) != 0 : (q.kt == 1 && q.C >= 128);
        if (!by_rows) {
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

