"""Smallest program that launches every kernel of the hot path once at the benchmark shapes
(16 clips x 10 crops = 160 clip-crops): the target of the `ncu --set full` captures under profiles/.

    python tools/ncu_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
from anomaly_detection_on_video_b200.engine import segment_mean
from anomaly_detection_on_video_b200.i3d import I3Res50

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = I3Res50()   # constructor initialisation: the kernels' time does not depend on the weights
m.eval().to(dev)
frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(16 * 16, 240, 320, 3), dtype=np.uint8)).to(dev)
ds = TenCropVideoFrameDataset(frames, device=dev)
x = ds.clips_stem(0, 16)                       # preprocess_kernel
f = m.forward_stem_layout(x)                   # stem + 52 conv + 2 maxpool + avgpool
s = segment_mean(f.view(16, 10, -1), 32)       # segment_mean_kernel
torch.cuda.synchronize()
print("ok", tuple(f.shape), tuple(s.shape), float(f.abs().mean()))
