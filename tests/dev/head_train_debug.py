"""Per-parameter gradient error of the native MGFN training step against autograd over the CPU oracle (network order)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from anomaly_detection_on_video_b200.mgfn import MGFNConfig, MGFNForVideoAnomalyDetection
from oracle import mgfn as M

dev = torch.device("cuda", 0)
m = MGFNForVideoAnomalyDetection(MGFNConfig(dropout_rate=0.0))
m.load_state_dict(M.seeded_state_dict(0), strict=True)
m = m.to(dev).train()
video = M.synthetic_video(3, 4, 10, 32)
out = m(video.to(dev), abnormal_labels=torch.ones(2), normal_labels=torch.zeros(2))
out.loss.backward()
torch.cuda.synchronize()
sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running_" not in k and "num_batches" not in k) for k, v in M.seeded_state_dict(0).items()}
ref = M.forward(video, sd, normal_labels=torch.zeros(2), abnormal_labels=torch.ones(2), training=True)
names = [k for k, v in sd.items() if v.requires_grad]
grads = dict(zip(names, torch.autograd.grad(ref["loss"], [sd[n] for n in names])))
print("loss", float(out.loss), float(ref["loss"]))
for n, p in m.named_parameters():
    g, r = p.grad.detach().cpu().double(), grads[n].double()
    e = float((g - r).norm() / max(float(r.norm()), 1e-30))
    print(f"{e:10.3e}  |g| {float(g.norm()):10.3e}  |ref| {float(r.norm()):10.3e}  {n}")
