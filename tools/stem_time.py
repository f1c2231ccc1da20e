import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from anomaly_detection_on_video_b200.i3d import I3Res50
m = I3Res50().eval().cuda()
dev = torch.device("cuda", 0)
xs = torch.randn(160, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
for _ in range(2): m.forward_stem_layout(xs)
plan = m.plan(dev); plan.profile_begin()
for _ in range(3): m.forward_stem_layout(xs)
prof = plan.profile_end()
print(os.environ.get("VAD_STEM_DEBUG", "0"), "conv1 ms", round(prof[0]["ms"] / prof[0]["calls"], 3), "maxpool1", round(prof[1]["ms"] / prof[1]["calls"], 3))
