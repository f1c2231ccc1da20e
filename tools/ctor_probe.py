"""Where does the host-side time of one end-to-end step go?  (debug tool for bench.py's e2e number)"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
from anomaly_detection_on_video_b200.engine import Preprocessor, segment_mean
from anomaly_detection_on_video_b200.extract_features import extract_clip_features
from anomaly_detection_on_video_b200.i3d import I3Res50
dev = torch.device("cuda", 0)
fh = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2000, 240, 320, 3), dtype=np.uint8)).pin_memory()
m = I3Res50().eval().to(dev)
def sync(): torch.cuda.synchronize(dev)
def ms(t0): return round((time.perf_counter() - t0) * 1e3, 2)
for it in range(6):
    sync(); t0 = time.perf_counter()
    pp = Preprocessor(240, 320, 256, 224, 10, dev); sync(); a = ms(t0); t0 = time.perf_counter()
    dst = torch.empty(fh.shape, dtype=torch.uint8, device=dev); sync(); b = ms(t0); t0 = time.perf_counter()
    dst.copy_(fh, non_blocking=True); sync(); c = ms(t0); t0 = time.perf_counter()
    del pp, dst
    d = TenCropVideoFrameDataset(fh, device=dev); sync(); e = ms(t0); t0 = time.perf_counter()
    feats = extract_clip_features(d, m, dev, clips_per_batch=16, strict_compat=False, as_numpy=False); sync(); f = ms(t0); t0 = time.perf_counter()
    seg = segment_mean(feats, 32); out = feats.cpu(), seg.cpu(); g = ms(t0)
    del d, feats, seg
    print(f"iter {it}: preproc ctor {a}  empty {b}  single copy {c}  dataset ctor {e}  extract {f}  d2h {g}", flush=True)
