"""BASELINE config 4 in miniature: sharded ten-crop extraction of synthetic UCF-Crime-shaped videos through the CLI's
work queue, and the check that the output files do not depend on the number of GPUs.

    python tools/sharded_extract_check.py gen   --dir /tmp/vad_shard --videos 24
    python tools/sharded_extract_check.py run   --dir /tmp/vad_shard                      # 1 GPU  -> <dir>/out_w1
    python -m torch.distributed.run --nproc-per-node 2 ... tools/sharded_extract_check.py run --dir /tmp/vad_shard   # -> out_w2
    python tools/sharded_extract_check.py compare --dir /tmp/vad_shard

Video lengths: seeded log-normal (median 240 frames here; UCF-Crime's real median is ~2,000), clipped to [40, 1500],
240x320 uint8 noise frames stored as .npy (the decoder is host I/O and out of scope).
"""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("phase", choices=["gen", "run", "compare"])
    ap.add_argument("--dir", required=True)
    ap.add_argument("--videos", type=int, default=24)
    a = ap.parse_args()
    vdir = os.path.join(a.dir, "videos")
    if a.phase == "gen":
        os.makedirs(vdir, exist_ok=True)
        rng = np.random.default_rng(2024)
        lens = np.clip(np.exp(rng.normal(np.log(240.0), 0.8, a.videos)), 40, 1500).astype(int)
        for i, n in enumerate(lens):
            name = ("Normal_Videos_%03d_x264" if i % 2 else "Abuse%03d_x264") % i
            np.save(os.path.join(vdir, name + ".npy"), rng.integers(0, 256, size=(int(n), 240, 320, 3), dtype=np.uint8))
        print("generated", a.videos, "videos,", int(lens.sum()), "frames,", int(sum((n - 1) // 16 + 1 for n in lens)), "clips")
        return
    if a.phase == "run":
        import torch

        from anomaly_detection_on_video_b200.extract_features import _rows_from_dir, extract, segment
        from anomaly_detection_on_video_b200.i3d import I3Res50
        from anomaly_detection_on_video_b200.workqueue import WorkQueue

        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        queue = WorkQueue.from_env()
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dev = torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(0)   # every rank and every run: the same constructor initialisation
        model = I3Res50()
        model.eval().to(dev)
        out = os.path.join(a.dir, f"out_w{world}")
        rows = _rows_from_dir(vdir)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        written = extract(rows, model, dev, os.path.join(out, "train"), queue=queue)
        if queue is not None:
            queue.barrier()
        segment(os.path.join(out, "train"), os.path.join(out, "segment_features_32"), 32, queue=queue)
        if queue is not None:
            queue.barrier()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clips = sum((np.load(r["video_path"], mmap_mode="r").shape[0] - 1) // 16 + 1 for r in rows)
        print(json.dumps({"world": world, "rank": rank, "videos_here": len(written), "seconds": dt,
                          "clip_crops_per_s_all_ranks": clips * 10 / dt}), flush=True)
        return
    digests = {}
    for w in sorted(d for d in os.listdir(a.dir) if d.startswith("out_w")):
        h = {}
        for root, _, files in os.walk(os.path.join(a.dir, w)):
            for f in files:
                p = os.path.join(root, f)
                h[os.path.relpath(p, os.path.join(a.dir, w))] = hashlib.sha256(open(p, "rb").read()).hexdigest()
        digests[w] = h
        print(w, len(h), "files")
    names = list(digests)
    ok = all(digests[n] == digests[names[0]] for n in names[1:]) and len(names) >= 2 and len(digests[names[0]]) > 0
    print("IDENTICAL" if ok else "DIFFERENT", names)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
