"""Per-op CUDA-event time of the layer1 tail launches under different VAD_* settings, variants interleaved on one box
(whole-forward A/B drowns a 1-2 % kernel change in the power-cap clock drift).

    python tools/tail_ab.py "VAD_TAIL_CFG=23" "VAD_TAIL_CFG=15"
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.i3d import I3Res50
variants = sys.argv[1:] or ["", "VAD_NO_TAIL=1"]
dev = torch.device("cuda", 0)
xs = torch.randn(160, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
models = []
for v in variants:
    kv = [x.split("=", 1) for x in v.split(",") if x]
    for k, val in kv: os.environ[k] = val
    m = I3Res50().eval().to(dev)
    for _ in range(2): m.forward_stem_layout(xs)
    torch.cuda.synchronize()
    for k, _ in kv: del os.environ[k]
    models.append(m)
acc = {v: {} for v in variants}
for lap in range(4):
    for v, m in zip(variants, models):
        plan = m.plan(dev)
        plan.profile_begin()
        for _ in range(10): m.forward_stem_layout(xs)
        torch.cuda.synchronize()
        for p in plan.profile_end():
            if p["calls"] and p["name"].startswith("layer1."):
                a = acc[v].setdefault(p["name"], [0.0, 0])
                a[0] += p["ms"]; a[1] += p["calls"]
for v in variants:
    row = {n: round(a[0] / a[1], 4) for n, a in acc[v].items() if a[0] / a[1] > 0.01}
    print(v or "(default)", "layer1 total", round(sum(row.values()), 4), row, flush=True)
