"""Time the fused preprocessing kernel alone: 16 clips of 16 frames 240x320 -> 160 clip-crops in stem layout."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
dev = torch.device("cuda", 0)
frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(256, 240, 320, 3), dtype=np.uint8)).to(dev)
ds = TenCropVideoFrameDataset(frames, device=dev)
out = ds.clips_stem(0, 16)
for _ in range(3): ds.clips_stem(0, 16, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): ds.clips_stem(0, 16, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gb = (frames.numel() + out.numel() * 2) / 1e9
print(f"preprocess 16 clips: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic ({gb:.3f} GB)")
