"""Small-batch I3Res50 forwards, op-by-op launches (VAD_GRAPH=0) against CUDA-graph replay (VAD_GRAPH=1)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.i3d import I3Res50
dev = torch.device("cuda", 0)
m = I3Res50().eval().to(dev)
for B in (1, 2, 4, 8, 16, 32, 64):
    xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
    row = {"clip_crops": B}
    for mode in ("0", "1"):
        os.environ["VAD_GRAPH"] = mode
        m._plan = None
        for _ in range(4): m.forward_stem_layout(xs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): m.forward_stem_layout(xs)
        e1.record(); torch.cuda.synchronize()
        row["graph_ms" if mode == "1" else "direct_ms"] = round(e0.elapsed_time(e1) / 50, 4)
    row["clips_per_s_graph"] = round(B / row["graph_ms"] * 1e3)
    print(json.dumps(row), flush=True)
