"""BASELINE config 3: batched I3Res50 throughput, 1..256 clip-crops per forward, in both precision modes:
bf16 (bf16 operands and stored activations, fp32 accumulate -- the production path) and tf32 (fp32 activations and
weights, tcgen05 kind::tf32, general kernels).  CUDA-event timing, >= 10 iterations after warm-up.

    python tools/batch_sweep.py [bf16|tf32|both]
"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomaly_detection_on_video_b200.engine import ingest_ncthw_tf32
from anomaly_detection_on_video_b200.i3d import I3Res50
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
m = I3Res50().eval().to(dev)
for mode in (("bf16", "tf32") if which == "both" else (which,)):
    m.precision = mode
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 160, 256):
        if mode == "tf32" and B > 128:
            continue  # fp32 activations of the unfused table: keep the sweep inside a few GB
        if mode == "bf16":
            xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
            run = lambda: m.forward_stem_layout(xs)
        else:
            plan = m.plan(dev)
            xs = ingest_ncthw_tf32(torch.randn(B, 3, 16, 224, 224, device=dev), planes=bool(plan.ops[0].flags & 64))
            run = lambda: plan.forward(xs)
        for _ in range(3): run()
        n = 20 if B <= 64 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"mode": mode, "clip_crops": B, "ms": round(ms, 3), "clips_per_s": round(B / ms * 1e3),
                          "tflops": round(B * 32.829 / ms, 1)}), flush=True)
        del xs
        torch.cuda.empty_cache()
