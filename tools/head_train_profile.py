"""One MGFN training step at the bench shape (32 bags x 10 crops x 32 segments) for `ncu --metrics gpu__time_duration.sum`:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/head_train_launches.csv python tools/head_train_profile.py
and, without ncu, the CUDA-event time of a step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from anomaly_detection_on_video_b200.mgfn import MGFNConfig, MGFNForVideoAnomalyDetection, NativeAdam

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = MGFNForVideoAnomalyDetection(MGFNConfig()).to(dev).train()
opt = NativeAdam(m)
g = torch.Generator().manual_seed(1)
feat = torch.randn(32, 10, 32, 2048, generator=g).abs() * 0.5
video = torch.cat([feat, feat.norm(dim=3, keepdim=True)], dim=3).to(dev)
nl, al = torch.zeros(16, device=dev), torch.ones(16, device=dev)
steps = int(os.environ.get("STEPS", "3"))
for i in range(steps):
    if i == steps - 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
    out = m(video, abnormal_labels=al, normal_labels=nl)
    opt.step()
e1.record()
torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1), "loss", float(out.loss.detach()), "launches", m.train_launches)
