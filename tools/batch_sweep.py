"""BASELINE config 3: batched I3Res50 throughput, 1..256 clip-crops per forward (bf16 operands, fp32 accumulate;
the TF32 backbone mode is not built), CUDA-event timing, >= 10 iterations after warm-up."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.i3d import I3Res50
dev = torch.device("cuda", 0)
m = I3Res50().eval().to(dev)
for B in (1, 2, 4, 8, 16, 32, 64, 128, 160, 256):
    xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
    for _ in range(3): m.forward_stem_layout(xs)
    n = 20 if B <= 64 else 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): m.forward_stem_layout(xs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"clip_crops": B, "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3), "tflops": round(B * 32.829 / ms, 1)}), flush=True)
