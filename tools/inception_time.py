"""Per-op timing of the InceptionV1-3D op table at B = 160 clip-crops (16 x 224 x 224)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from anomaly_detection_on_video_b200.inception import InceptionI3d
B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = torch.device("cuda", 0)
m = InceptionI3d().eval().to(dev)
xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
for _ in range(2): m.forward_stem_layout(xs)
plan = m.plan(dev); plan.profile_begin()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): m.forward_stem_layout(xs)
e1.record(); torch.cuda.synchronize()
prof = plan.profile_end()
ms = e0.elapsed_time(e1) / 3
print(f"InceptionI3d forward B={B}: {ms:.3f} ms, {B / ms * 1e3:.0f} clips/s, {B * 55.575e9 / ms / 1e9:.0f} TFLOP/s")
for p in sorted(prof, key=lambda p: -p["ms"])[:int(os.environ.get("TOP", "14"))]:
    print(f"{p['name']:24s} {p['ms'] / p['calls']:8.3f} ms  {p['flops'] / max(p['ms'], 1e-9) / 1e9:8.1f} TFLOP/s")
