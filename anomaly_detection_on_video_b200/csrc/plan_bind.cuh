// bind_plan: tensor maps and pointers of every op for one (input, workspace) pair; dropped CUDA graph on re-bind
// (part of vad_api.cu: included there, after the plan structures; not a stand-alone translation unit)
#pragma once

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
      c.out1 = d.dst1 > 0 ? reinterpret_cast<__nv_bfloat16*>(slot_ptr(d.dst1)) : nullptr;   // fused sibling 1x1x1 convs
      c.out2 = (d.dst1 > 0 && d.dst2 > 0) ? reinterpret_cast<__nv_bfloat16*>(slot_ptr(d.dst2)) : nullptr;
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
      memset(&r.tmO2, 0, sizeof(r.tmO2));
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
        cuuint64_t odim[2] = {(cuuint64_t)(d.dst1 > 0 ? d.seg_w0 : d.cout), (cuuint64_t)c.M};   // fused siblings: the map clips at the part's width
        cuuint64_t ostr[1] = {(cuuint64_t)r.dst_c * 2};
        cuuint32_t obox[2] = {64, 32};
        cr = p->encode_tiled(&r.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)c.out, odim, ostr, obox, es2,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(output) failed: %d", i, (int)cr);
        if (d.dst1 > 0) {   // the sibling outputs: tensors of exactly seg_w channels; the second one rides in the (unused) residual slot
          const int parts[2][2] = {{d.dst1, d.seg_w1}, {d.dst2, d.seg_w2}};
          CUtensorMap* maps[2] = {&r.tmR, &r.tmO2};
          for (int e = 0; e < 2 && cr == CUDA_SUCCESS; ++e) {
            if (parts[e][0] <= 0) continue;
            cuuint64_t pdim[2] = {(cuuint64_t)parts[e][1], (cuuint64_t)c.M};
            cuuint64_t pstr[1] = {(cuuint64_t)parts[e][1] * 2};
            cr = p->encode_tiled(maps[e], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, slot_ptr(parts[e][0]), pdim, pstr, obox, es2,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          }
          if (cr != CUDA_SUCCESS) return fail(VAD_ERR_CUDA, "op %zu: cuTensorMapEncodeTiled(sibling output) failed: %d", i, (int)cr);
        }
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

