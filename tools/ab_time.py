"""A/B timing of the backbone forward under different VAD_* environment settings, on one box, interleaved.

    python tools/ab_time.py "VAD_PAIR=0" "VAD_PAIR=1" "VAD_PAIR=1,VAD_PAIR_MIN_KB=20"

Each argument is a comma-separated list of NAME=VALUE applied while that variant's plan is created (the knobs are read in
vad_plan_create); variants are then timed round-robin (B = 160 clip-crops, 20 forwards per lap, 4 laps).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from anomaly_detection_on_video_b200.i3d import I3Res50

    variants = sys.argv[1:] or ["", "VAD_PAIR=0"]
    dev = torch.device("cuda", 0)
    B = int(os.environ.get("AB_BATCH", "160"))
    xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
    torch.manual_seed(0)
    sd = I3Res50().state_dict()   # one constructor initialisation shared by every variant
    models = []
    for v in variants:
        kv = [x.split("=", 1) for x in v.split(",") if x]
        for k, val in kv:
            os.environ[k] = val
        m = I3Res50()
        m.load_state_dict(sd)
        m.eval().cuda()
        for _ in range(2):
            m.forward_stem_layout(xs)
        torch.cuda.synchronize()
        for k, _ in kv:
            del os.environ[k]
        models.append(m)
    laps = {v: [] for v in variants}
    for lap in range(4):
        for v, m in zip(variants, models):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                m.forward_stem_layout(xs)
            e1.record()
            torch.cuda.synchronize()
            laps[v].append(e0.elapsed_time(e1) / 20)
    for v in variants:
        ms = sorted(laps[v])
        print(json.dumps({"variant": v or "(default)", "ms_median": (ms[1] + ms[2]) / 2, "ms_min": ms[0], "laps": [round(x, 3) for x in laps[v]],
                          "clips_per_s": B / ((ms[1] + ms[2]) / 2) * 1e3}), flush=True)


if __name__ == "__main__":
    main()
