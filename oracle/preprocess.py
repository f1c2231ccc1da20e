"""Oracle for clip preprocessing (numpy, integer-exact).  TEST INFRASTRUCTURE ONLY.

Restates the transform chain the reference builds at src/dataset.py:175-183:

    GroupResize(256, BILINEAR)        src/gtransforms.py:9-18   -> torchvision Resize -> PIL resize
    GroupTenCrop(224)                 src/gtransforms.py:21-26  -> torchvision TenCrop
    ToTensorTenCrop                   src/gtransforms.py:29-38  (HWC u8 -> CHW f32)
    GroupStandardizationTenCrop       src/gtransforms.py:41-73  (x - 114.75) / 57.375, two roundings
    LoopPad(16)                       src/gtransforms.py:115-132
    permute(1, 0, 2, 3, 4)            src/dataset.py:195

PIL and torchvision are third-party dependencies of the reference (not vendored in it); their
algorithms are restated here from their published sources (Pillow ``src/libImaging/Resample.c``:
``precompute_coeffs`` / ``normalize_coeffs_8bpc`` / ``ImagingResampleHorizontal_8bpc`` /
``ImagingResampleVertical_8bpc``; torchvision ``transforms/functional.py``: ``resize`` /
``five_crop`` / ``ten_crop`` / ``center_crop``) and pinned by tests against the installed
PIL 12.2 / torchvision 0.26 and against fixtures produced by the reference's own classes.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Pillow: 8 bits of pixel, 2 bits of headroom


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for BILINEAR (support = 1.0).

    Returns (bounds [out,2] = (xmin, count), coeffs [out,ksize] int32 in 22-bit fixed point, ksize).
    """
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            v *= float(1 << PRECISION_BITS)
            coeffs[xx, x] = int(v - 0.5) if v < 0 else int(v + 0.5)
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs, ksize


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit resampling pass along ``axis`` (0 = vertical, 1 = horizontal) of [H, W, C] u8."""
    in_size = img.shape[axis]
    bounds, coeffs, _ = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)  # [in, other, C]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(cnt):
            acc += src[xmin + j] * int(coeffs[xx, j])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """``PIL.Image.resize((out_w, out_h), BILINEAR)`` on an [H, W, C] uint8 array.

    Pillow runs the horizontal pass first, rounds to uint8, then the vertical pass; a pass whose
    size does not change is skipped.
    """
    assert img.dtype == np.uint8 and img.ndim == 3
    out = img
    if out_w != img.shape[1]:
        out = _resample_axis(out, out_w, axis=1)
    if out_h != img.shape[0]:
        out = _resample_axis(out, out_h, axis=0)
    return out


def resized_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision ``Resize(int)``: shorter side -> size, longer -> int(size * long / short)."""
    if w <= h:
        return int(size * h / w), size
    return size, int(size * w / h)


def ten_crop_boxes(h: int, w: int, crop: int) -> List[Tuple[int, int, bool]]:
    """torchvision ``TenCrop``: (top, left, flipped) of [tl, tr, bl, br, center] on the image, then
    the same five on its horizontal flip.  For a flipped entry (top, left) index the flipped image."""
    ct = int(round((h - crop) / 2.0))  # python round: half to even, as torchvision's center_crop
    cl = int(round((w - crop) / 2.0))
    five = [(0, 0), (0, w - crop), (h - crop, 0), (h - crop, w - crop), (ct, cl)]
    return [(t, l, False) for t, l in five] + [(t, l, True) for t, l in five]


def standardize(u8: np.ndarray) -> np.ndarray:
    """``t.sub_(114.75).div_(57.375)`` in fp32 (src/gtransforms.py:60-61,71-72): two roundings."""
    x = u8.astype(np.float32)
    x = (x - np.float32(114.75)).astype(np.float32)
    return (x / np.float32(57.375)).astype(np.float32)


def clip_tensor(frames: np.ndarray, clip_idx: int, frames_per_clip: int = 16, resize: int = 256, crop: int = 224,
                ncrops: int = 10) -> np.ndarray:
    """``TenCropVideoFrameDataset(frames)[clip_idx]`` -> (ncrops, frames_per_clip, 3, crop, crop) fp32.

    frames: [n_frames, H, W, 3] uint8.  ``ncrops=1`` returns only the center crop (TenCrop index 4).
    """
    n, h, w, _ = frames.shape
    start = clip_idx * frames_per_clip
    chunk = frames[start:start + frames_per_clip]  # src/dataset.py:189-191
    length = chunk.shape[0]
    rh, rw = resized_size(h, w, resize)
    boxes = ten_crop_boxes(rh, rw, crop)
    if ncrops == 1:
        boxes = [boxes[4]]
    per_frame = []
    for f in range(length):
        img = chunk[f] if (rh, rw) == (h, w) else pil_resize_bilinear(chunk[f], rh, rw)
        flipped = img[:, ::-1]
        crops = []
        for top, left, flip in boxes:
            srcimg = flipped if flip else img
            c = srcimg[top:top + crop, left:left + crop]          # HWC u8
            crops.append(standardize(np.ascontiguousarray(c.transpose(2, 0, 1))))  # CHW f32
        per_frame.append(np.stack(crops))                         # (ncrops, 3, crop, crop)
    t = np.stack(per_frame)                                       # (L, ncrops, 3, crop, crop)
    if length != frames_per_clip:                                 # LoopPad: frame j <- frame j mod L
        idx = np.arange(frames_per_clip) % length
        t = t[idx]
    return np.ascontiguousarray(t.transpose(1, 0, 2, 3, 4))       # src/dataset.py:195


def n_clips(n_frames: int, frames_per_clip: int = 16) -> int:
    """src/dataset.py:171-173."""
    return (n_frames - 1) // frames_per_clip + 1


def to_stem_layout(clip: np.ndarray, pad_left: int = 3) -> np.ndarray:
    """(ncrops, T, 3, H, W) fp32 -> (ncrops, T, H, W + 8, 4) fp32 with zero padding: the layout the
    CUDA stem consumes (values still fp32; the kernel rounds them to bf16)."""
    k, t, c, h, w = clip.shape
    out = np.zeros((k, t, h, w + 8, 4), dtype=np.float32)
    out[:, :, :, pad_left:pad_left + w, :3] = clip.transpose(0, 1, 3, 4, 2)
    return out
