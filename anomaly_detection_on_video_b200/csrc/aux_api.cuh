// Host side of preprocessing (K1), segment mean and magnitude (K4): included by vad_api.cu (same translation unit, shares
// the error / device helpers).
#pragma once

#include "aux_kernels.cuh"

// ------------------------------------------------------------------------------------ preprocessing
struct vad_preproc {
  int src_h = 0, src_w = 0, rh = 0, rw = 0, crop = 0, ncrops = 0, device = 0;
  int ksize_h = 0, ksize_v = 0;
  int tops[10] = {0}, lefts[10] = {0}, flips[10] = {0};
  int* tables_dev = nullptr;  // bounds_h | coef_h | bounds_v | coef_v
  size_t off_bh = 0, off_ch = 0, off_bv = 0, off_cv = 0;
  std::vector<int> bounds_v_host;  // (ymin, count) per resized row: sizes the kernel's source-row staging
};

// Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0):
// double-precision triangle weights, normalised, then quantised to 22 fractional bits.
static void resample_tables(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& coefs, int& ksize) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  coefs.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (double)(1 << 22);
      coefs[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

extern "C" int32_t vad_preproc_create(vad_preproc_t** out, int32_t src_h, int32_t src_w, int32_t resize, int32_t crop,
                                      int32_t ncrops, int32_t device) {
  if (!out || src_h <= 0 || src_w <= 0 || resize <= 0 || crop <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_preproc_create: bad size");
  if (ncrops != 1 && ncrops != 10) return fail(VAD_ERR_INVALID_ARGUMENT, "ncrops must be 1 or 10");
  if (crop & 1) return fail(VAD_ERR_INVALID_ARGUMENT, "crop must be even");
  int32_t rc = require_sm100(device);
  if (rc != VAD_OK) return rc;
  vad_preproc* pp = new vad_preproc();
  pp->src_h = src_h; pp->src_w = src_w; pp->crop = crop; pp->ncrops = ncrops; pp->device = device;
  // torchvision Resize(int): shorter side -> resize, longer -> int(resize * long / short)
  if (src_w <= src_h) { pp->rw = resize; pp->rh = (int)((double)((long long)resize * src_h) / (double)src_w); }
  else                { pp->rh = resize; pp->rw = (int)((double)((long long)resize * src_w) / (double)src_h); }
  if (pp->rh < crop || pp->rw < crop) { delete pp; return fail(VAD_ERR_INVALID_ARGUMENT, "crop %d larger than resized image %dx%d", crop, pp->rh, pp->rw); }
  // torchvision five_crop order tl, tr, bl, br, center(round-half-even); then the h-flipped image
  const int ct = (int)nearbyint((pp->rh - crop) / 2.0), cl = (int)nearbyint((pp->rw - crop) / 2.0);
  const int t5[5] = {0, 0, pp->rh - crop, pp->rh - crop, ct};
  const int l5[5] = {0, pp->rw - crop, 0, pp->rw - crop, cl};
  if (ncrops == 10) {
    for (int k = 0; k < 5; ++k) {
      pp->tops[k] = t5[k]; pp->lefts[k] = l5[k]; pp->flips[k] = 0;
      pp->tops[5 + k] = t5[k]; pp->lefts[5 + k] = pp->rw - crop - l5[k]; pp->flips[5 + k] = 1;
    }
  } else {
    pp->tops[0] = ct; pp->lefts[0] = cl; pp->flips[0] = 0;
  }
  std::vector<int> bh, ch, bv, cv;
  resample_tables(src_w, pp->rw, bh, ch, pp->ksize_h);
  resample_tables(src_h, pp->rh, bv, cv, pp->ksize_v);
  pp->bounds_v_host = bv;
  std::vector<int> all;
  pp->off_bh = 0;               all.insert(all.end(), bh.begin(), bh.end());
  pp->off_ch = all.size();      all.insert(all.end(), ch.begin(), ch.end());
  pp->off_bv = all.size();      all.insert(all.end(), bv.begin(), bv.end());
  pp->off_cv = all.size();      all.insert(all.end(), cv.begin(), cv.end());
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&pp->tables_dev, all.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(pp->tables_dev, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (pp->tables_dev) cudaFree(pp->tables_dev);
    delete pp;
    return fail(VAD_ERR_CUDA, "vad_preproc_create: %s", cudaGetErrorString(e));
  }
  *out = pp;
  return VAD_OK;
}

extern "C" int32_t vad_preproc_info(const vad_preproc_t* pp, int32_t resized_hw[2], int32_t* tops, int32_t* lefts,
                                    int32_t* flips) {
  if (!pp) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_preproc_info: null handle");
  if (resized_hw) { resized_hw[0] = pp->rh; resized_hw[1] = pp->rw; }
  for (int k = 0; k < pp->ncrops; ++k) {
    if (tops) tops[k] = pp->tops[k];
    if (lefts) lefts[k] = pp->lefts[k];
    if (flips) flips[k] = pp->flips[k];
  }
  return VAD_OK;
}

extern "C" int32_t vad_preproc_run(vad_preproc_t* pp, const uint8_t* frames_dev, int32_t n_frames, int32_t clip_start,
                                   int32_t n_clips, int32_t frames_per_clip, int32_t out_mode, int32_t pad_left,
                                   void* out_dev, void* stream) {
  if (!pp || !frames_dev || !out_dev) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_preproc_run: null pointer");
  if (n_frames <= 0 || frames_per_clip <= 0 || n_clips <= 0 || clip_start < 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_preproc_run: bad counts");
  const int total_clips = (n_frames - 1) / frames_per_clip + 1;  // src/dataset.py:171-173
  if (clip_start + n_clips > total_clips) return fail(VAD_ERR_INVALID_ARGUMENT, "clips [%d,%d) exceed the %d clips of %d frames", clip_start, clip_start + n_clips, total_clips, n_frames);
  if (out_mode != VAD_OUT_DATASET_F32 && out_mode != VAD_OUT_STEM_BF16) return fail(VAD_ERR_INVALID_ARGUMENT, "bad out_mode");
  if (pad_left < 0 || pad_left > 8) return fail(VAD_ERR_INVALID_ARGUMENT, "pad_left must be in [0,8]");
  if ((long long)n_clips * frames_per_clip > 65535) return fail(VAD_ERR_INVALID_ARGUMENT, "at most 65535 frame slots per call");
  PreprocParams q;
  memset(&q, 0, sizeof(q));
  q.frames = frames_dev; q.n_frames = n_frames; q.H = pp->src_h; q.W = pp->src_w;
  q.rh = pp->rh; q.rw = pp->rw; q.ksize_h = pp->ksize_h; q.ksize_v = pp->ksize_v;
  q.bounds_h = pp->tables_dev + pp->off_bh; q.coef_h = pp->tables_dev + pp->off_ch;
  q.bounds_v = pp->tables_dev + pp->off_bv; q.coef_v = pp->tables_dev + pp->off_cv;
  q.crop = pp->crop; q.ncrops = pp->ncrops;
  for (int k = 0; k < 10; ++k) { q.tops[k] = pp->tops[k]; q.lefts[k] = pp->lefts[k]; q.flips[k] = pp->flips[k]; }
  q.clip_start = clip_start; q.fpc = frames_per_clip; q.out_mode = out_mode; q.pad_left = pad_left; q.out = out_dev;
  if (out_mode == VAD_OUT_STEM_BF16 && pp->ksize_h <= 5 && pp->ksize_v <= 5 && !getenv("VAD_K1_GENERIC")) {
    // column-per-thread kernel (aux_kernels.cuh): bf16 LUT | right-aligned vertical coefficients | R rows of bf16 pixels |
    // staged source rows (+ 32 B so that zero-coefficient taps past the row end stay inside the allocation)
    const int KV = pp->ksize_v <= 3 ? 3 : 5;
    const size_t src_pitch = ((size_t)pp->src_w * 3 + 15) / 16 * 16;
    size_t smem = 0;
    int R = 8;
    for (;; R >>= 1) {
      int max_src = 0;
      for (int y0 = 0; y0 < pp->rh; y0 += R) {
        const int y1 = (y0 + R < pp->rh ? y0 + R : pp->rh) - 1;
        const int n = pp->bounds_v_host[2 * y1] + pp->bounds_v_host[2 * y1 + 1] - pp->bounds_v_host[2 * y0];
        if (n > max_src) max_src = n;
      }
      size_t off = 512 + (size_t)R * KV * 4 + (size_t)R * 4;
      off = (off + 15) / 16 * 16;
      q.off_px = (int)off;
      off += (size_t)R * pp->rw * 8;
      off = (off + 15) / 16 * 16;
      q.off_src = (int)off;
      off += (size_t)max_src * src_pitch + 32;
      smem = off;
      q.rows_per_block = R;
      q.max_src_rows = max_src;
      if (smem <= 64 * 1024 || R == 1) break;
    }
    if (smem > 200 * 1024) return fail(VAD_ERR_INVALID_ARGUMENT, "source frames too wide for the resampling kernel (%zu B of shared memory)", smem);
    int threads = (pp->rw + 31) / 32 * 32;
    const int per_crop = (pp->crop + 8) / 2;
    if (threads < (per_crop + 31) / 32 * 32) threads = (per_crop + 31) / 32 * 32;
    if (threads > 512) threads = 512;
    if (threads < per_crop) return fail(VAD_ERR_INVALID_ARGUMENT, "crop %d too wide for the preprocessing kernel", pp->crop);
    dim3 grid((pp->rh + R - 1) / R, n_clips * frames_per_clip);
    auto launch = [&](auto kern) -> int32_t {
      if (smem > 48 * 1024) VAD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(q);
      VAD_CUDA_CHECK(cudaGetLastError());
      return VAD_OK;
    };
    // up to 384 threads (resized rows of <= 384 pixels, every 4:3 source) the kernel is compiled for four blocks per SM
    if (threads <= 384) {
      if (pp->ksize_h <= 3) return KV == 3 ? launch(preprocess_stem_kernel<3, 3, 384>) : launch(preprocess_stem_kernel<3, 5, 384>);
      return KV == 3 ? launch(preprocess_stem_kernel<5, 3, 384>) : launch(preprocess_stem_kernel<5, 5, 384>);
    }
    if (pp->ksize_h <= 3) return KV == 3 ? launch(preprocess_stem_kernel<3, 3, 512>) : launch(preprocess_stem_kernel<3, 5, 512>);
    return KV == 3 ? launch(preprocess_stem_kernel<5, 3, 512>) : launch(preprocess_stem_kernel<5, 5, 512>);
  }
  // shared memory: LUTs | R resized u8 rows | R rows of bf16 pixels (stem mode) | horizontally resampled source rows |
  // staged source rows.  R (resized rows per block) is the largest of 8, 4, 2, 1 that fits.
  const size_t row_bytes = ((size_t)pp->rw * 3 + 15) / 16 * 16;
  const size_t src_pitch = ((size_t)pp->src_w * 3 + 15) / 16 * 16;
  size_t smem = 0;
  int R = 8;
  for (;; R >>= 1) {
    int max_src = 0;
    for (int y0 = 0; y0 < pp->rh; y0 += R) {
      const int y1 = (y0 + R < pp->rh ? y0 + R : pp->rh) - 1;
      const int n = pp->bounds_v_host[2 * y1] + pp->bounds_v_host[2 * y1 + 1] - pp->bounds_v_host[2 * y0];
      if (n > max_src) max_src = n;
    }
    size_t off = 256 * 4 + 256 * 2 + (size_t)R * row_bytes;
    q.off_px = (int)off;
    if (out_mode == VAD_OUT_STEM_BF16) off += (size_t)R * pp->rw * 8;
    off = (off + 15) / 16 * 16;
    q.off_h = (int)off;
    off += (size_t)max_src * row_bytes;
    q.off_src = (int)off;
    off += (size_t)max_src * src_pitch;
    smem = off;
    q.rows_per_block = R;
    q.max_src_rows = max_src;
    if (smem <= 96 * 1024 || R == 1) break;
  }
  if (smem > 200 * 1024) return fail(VAD_ERR_INVALID_ARGUMENT, "source frames too wide for the resampling kernel (%zu B of shared memory)", smem);
  if (smem > 48 * 1024) VAD_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((pp->rh + R - 1) / R, n_clips * frames_per_clip);
  preprocess_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(q);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}

extern "C" void vad_preproc_destroy(vad_preproc_t* pp) {
  if (!pp) return;
  if (pp->tables_dev) cudaFree(pp->tables_dev);
  delete pp;
}

// ------------------------------------------------------------------------------------ segment / magnitude
extern "C" int32_t vad_segment_mean(const float* feats_dev, int32_t n_clips, int32_t ncrops, int32_t c,
                                    int32_t seg_length, float* out_dev, void* stream) {
  if (!feats_dev || !out_dev || n_clips <= 0 || ncrops <= 0 || c <= 0 || seg_length <= 0)
    return fail(VAD_ERR_INVALID_ARGUMENT, "vad_segment_mean: bad argument");
  const long long total = (long long)ncrops * seg_length * c;
  segment_mean_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feats_dev, n_clips, ncrops, c,
                                                                                         seg_length, out_dev);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}

extern "C" int32_t vad_add_magnitude(const float* feats_dev, int64_t rows, int32_t c, float* out_dev, void* stream) {
  if (!feats_dev || !out_dev || rows <= 0 || c <= 0) return fail(VAD_ERR_INVALID_ARGUMENT, "vad_add_magnitude: bad argument");
  const long long threads = rows * 32;
  add_magnitude_kernel<<<(int)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feats_dev, rows, c, out_dev);
  VAD_CUDA_CHECK(cudaGetLastError());
  return VAD_OK;
}

