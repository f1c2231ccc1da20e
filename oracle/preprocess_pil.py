"""The reference's clip transform chain run with the reference's own third-party calls -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

``oracle.preprocess`` restates Pillow's resampling arithmetic in numpy so that parity can be pinned bit for bit; it is far
slower than what the reference actually executes.  This module is the timing-faithful twin: the same sequence of PIL /
torchvision operations, with the same Python loop structure, as

    src/dataset.py:155-163,175-195          List[PIL.Image] -> transform -> permute(1, 0, 2, 3, 4)
    src/gtransforms.py:9-18                 transforms.Resize(256, BILINEAR) per frame
    src/gtransforms.py:21-26                transforms.TenCrop(224) per frame
    src/gtransforms.py:29-38                PILToTensor per crop, stack, stack, .float()
    src/gtransforms.py:41-73                per (frame, crop): per channel t.sub_(114.75).div_(57.375)
    src/gtransforms.py:115-132              LoopPad

PIL and torchvision are installed in the image (they are the reference's dependencies, not its source).  Only
``bench.py``'s CPU legs time it; ``tests/test_oracle.py`` holds it bit-equal to ``oracle.preprocess.clip_tensor``.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def to_pil_list(frames: np.ndarray) -> list:
    """[n, H, W, 3] uint8 -> List[PIL.Image] (what the decoder loop at src/dataset.py:156-159 produces)."""
    from PIL import Image
    return [Image.fromarray(f) for f in frames]


class RefClipTransform:
    """transform(images[start:end]).permute(1, 0, 2, 3, 4): (10, frames_per_clip, 3, crop, crop) fp32."""

    def __init__(self, frames_per_clip: int = 16, resize: int = 256, cropsize: int = 224) -> None:
        from PIL import Image
        from torchvision import transforms
        self.frames_per_clip = frames_per_clip
        self.resize = transforms.Resize(resize, interpolation=Image.BILINEAR)
        self.tencrop = transforms.TenCrop(cropsize)
        self.to_tensor = transforms.PILToTensor()
        self.mean = torch.FloatTensor([114.75] * 3)
        self.std = torch.FloatTensor([57.375] * 3)

    def __call__(self, images: List) -> torch.Tensor:
        resized = [self.resize(im) for im in images]
        cropped = [self.tencrop(im) for im in resized]
        t = torch.stack([torch.stack([self.to_tensor(c) for c in crops]) for crops in cropped], dim=0).float()
        for b in range(t.size(0)):            # frames
            for c in range(t.size(1)):        # crops
                for ch, m, s in zip(t[b, c], self.mean, self.std):
                    ch.sub_(m).div_(s)
        n = t.size(0)
        if n != self.frames_per_clip:         # cyclic repeat up to frames_per_clip
            pad = self.frames_per_clip - n
            parts = [t] * (1 + pad // n)
            if pad % n:
                parts.append(t[: pad % n])
            t = torch.cat(parts, dim=0)
        return t.permute(1, 0, 2, 3, 4)


def clip_tensor_pil(images: List, clip_idx: int, tf: RefClipTransform) -> torch.Tensor:
    """``TenCropVideoFrameDataset(images)[clip_idx]`` (src/dataset.py:188-195)."""
    k = tf.frames_per_clip
    return tf(images[clip_idx * k:(clip_idx + 1) * k])
