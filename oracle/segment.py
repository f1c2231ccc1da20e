"""Oracle for feature stacking, 32-segment averaging and magnitude append (numpy).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

from typing import List

import numpy as np


def stack_clip_features(per_batch: List[List[np.ndarray]], strict_compat: bool = False) -> np.ndarray:
    """``_extract``'s stacking (extract_features.py:93-100).

    per_batch: for each loader batch, the 10 per-crop outputs of shape (B, C, 1, 1, 1).
    Returns (n_clips, 10, C).  The reference ends with ``np.squeeze``, which also removes the clip
    axis when n_clips == 1; ``strict_compat=True`` reproduces that quirk.
    """
    outs = [np.stack(crops, axis=1) for crops in per_batch]  # (B, 10, C, 1, 1, 1)
    allc = np.vstack(outs)
    if strict_compat:
        return np.squeeze(allc)
    return allc.reshape(allc.shape[0], allc.shape[1], allc.shape[2])


def segment_edges(n: int, seg_length: int = 32) -> np.ndarray:
    """``np.linspace(0, n, seg_length + 1, dtype=int)`` (extract_features.py:175)."""
    return np.linspace(0, n, seg_length + 1, dtype=int)


def segment_features(features: np.ndarray, seg_length: int = 32) -> np.ndarray:
    """``segment()`` body for one file (extract_features.py:171-183).

    features: (n_clips, ncrops, C) fp32 as stored in ``<video>_i3d.npy`` -> (ncrops, seg_length, C).
    """
    feats = features.transpose(1, 0, 2)
    out = []
    for f in feats:
        new_feat = np.zeros((seg_length, f.shape[1])).astype(np.float32)
        r = segment_edges(len(f), seg_length)
        for i in range(seg_length):
            if r[i] != r[i + 1]:
                new_feat[i, :] = np.mean(f[r[i]:r[i + 1], :], 0)
            else:
                new_feat[i, :] = f[r[i], :]
        out.append(new_feat)
    return np.array(out, dtype=np.float32)


def segment_features_sequential(features: np.ndarray, seg_length: int = 32) -> np.ndarray:
    """The arithmetic the CUDA kernel performs, spelled out: integer bin edges ``i*n // seg``, rows
    added one at a time in clip order in fp32, one fp32 divide.  Tests assert this is bit-identical to
    ``segment_features`` (i.e. to numpy's mean over axis 0)."""
    n, k, c = features.shape
    out = np.zeros((k, seg_length, c), dtype=np.float32)
    for crop in range(k):
        for i in range(seg_length):
            a, b = (i * n) // seg_length, ((i + 1) * n) // seg_length
            if a != b:
                acc = features[a, crop].astype(np.float32).copy()
                for j in range(a + 1, b):
                    acc = (acc + features[j, crop]).astype(np.float32)
                out[crop, i] = acc / np.float32(b - a)
            else:
                out[crop, i] = features[a, crop]
    return out


def add_magnitude(feature: np.ndarray) -> np.ndarray:
    """``FeatureDataset.add_magnitude`` (src/dataset.py:121-124): (P, T, C) -> (P, T, C + 1)."""
    magnitude = np.linalg.norm(feature, axis=2)[:, :, np.newaxis]
    return np.concatenate((feature, magnitude), axis=2)
