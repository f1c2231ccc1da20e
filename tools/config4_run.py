"""BASELINE config 4 at its stated size: sharded 10-crop extraction of 1,900 synthetic UCF-Crime-shaped videos over the
GPUs of one box through the work queue (no collective on the data path), with byte-identity against a 1-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/config4_run.py --out gpurun_out/r02_config4_w8.json

Videos: 1,900 (UCF-Crime: 1,610 train + 290 test), frame counts log-normal with median 2,000 and sigma 0.8 under a fixed
seed, clipped to [64, 20,000]; 240 x 320 uint8 noise frames generated ON THE DEVICE from a per-video seed (the same bytes
whichever rank draws them; decoding is host I/O and out of scope).  Items are claimed longest-first from the
TCP-store counter (workqueue.WorkQueue); every video goes through the product path -- fused preprocessing, backbone in
16-clip batches, 32-segment mean -- its snippet features come back to the host like extract() returns them, its segment
features are written as .npy (atomic rename) like segment() does.  Identity digest per video: sha256 of the segment-feature
bytes + an exact, order-independent checksum of the snippet features' bit patterns computed on the device (int64 sum and xor of
the fp32 words; hashing 26 GB of snippet features on the host would cost more than extracting them).  Afterwards rank 0 alone recomputes every --check-stride-th video and
compares the hashes: the output of a video must not depend on which GPU, or how many, produced it.
Reports whole-job clip-crops/s (barrier to last rank done), per-rank items / clips / busy seconds and the idle tail.
"""
import argparse
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch


def video_lengths(n_videos: int, seed: int = 2024) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.clip(np.exp(rng.normal(np.log(2000.0), 0.8, n_videos)), 64, 20000).astype(np.int64)


def frames_of(video: int, n_frames: int, dev: torch.device) -> torch.Tensor:
    g = torch.Generator(device=dev)
    g.manual_seed(900_000 + video)
    return torch.randint(0, 256, (n_frames, 240, 320, 3), dtype=torch.uint8, device=dev, generator=g)


def _xor_reduce(words: torch.Tensor) -> torch.Tensor:
    """xor of all int32 words (exact, order independent), as int64."""
    w = words.reshape(-1)
    n = 1 << (int(w.numel()) - 1).bit_length()
    if n != w.numel():
        w = torch.cat([w, w.new_zeros(n - w.numel())])
    while w.numel() > 1:
        h = w.numel() // 2
        w = torch.bitwise_xor(w[:h], w[h:])
    return w[0].to(torch.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=1900)
    ap.add_argument("--out", default="gpurun_out/r02_config4.json")
    ap.add_argument("--workdir", default="/tmp/vad_config4")
    ap.add_argument("--check-stride", type=int, default=61)
    ap.add_argument("--static", action="store_true", help="static round-robin sharding instead of the dynamic queue")
    a = ap.parse_args()

    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
    from anomaly_detection_on_video_b200.engine import segment_mean
    from anomaly_detection_on_video_b200.extract_features import _atomic_save, extract_clip_features
    from anomaly_detection_on_video_b200.hostaffinity import bind_to_gpu
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from anomaly_detection_on_video_b200.workqueue import WorkQueue

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu(local)
    queue = WorkQueue.from_env(dynamic=not a.static) or WorkQueue()
    torch.manual_seed(0)
    model = I3Res50()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
    model.eval().to(dev)
    lens = video_lengths(a.videos)
    order = np.argsort(-lens, kind="stable")  # longest first: the tail of the schedule is made of short items
    outdir = os.path.join(a.workdir, f"w{world}", "segment_features_32")
    os.makedirs(outdir, exist_ok=True)

    def process(video: int):
        frames = frames_of(video, int(lens[video]), dev)
        ds = TenCropVideoFrameDataset(frames, device=dev)
        feats = extract_clip_features(ds, model, dev, strict_compat=False, as_numpy=False)
        seg = segment_mean(feats, 32)
        words = feats.view(torch.int32)
        digest = torch.stack([words.sum(dtype=torch.int64), _xor_reduce(words)])
        fh = feats.cpu()                                   # the product's output: (n_clips, 10, 2048) fp32 on the host
        sh = seg.cpu().numpy()
        _atomic_save(os.path.join(outdir, f"video_{video:04d}_i3d.npy"), sh)
        d = digest.cpu().tolist()
        return f"{hashlib.sha256(sh.tobytes()).hexdigest()}:{d[0]}:{d[1]}", fh.shape[0]

    process(int(order[-1]))  # warm-up: kernels loaded, plan built, workspace allocated
    torch.cuda.synchronize(dev)
    queue.barrier()
    t0 = time.perf_counter()
    mine, clips, busy = {}, 0, 0.0
    for i in queue.claim(a.videos, tag="config4"):
        v = int(order[i])
        t1 = time.perf_counter()
        mine[v], n = process(v)
        busy += time.perf_counter() - t1
        clips += n
    torch.cuda.synchronize(dev)
    done = time.perf_counter() - t0
    part = os.path.join(a.workdir, f"w{world}", f"rank{rank}.json")
    with open(part + ".tmp", "w") as f:
        json.dump({"rank": rank, "videos": len(mine), "clips": clips, "busy_s": busy, "done_s": done, "sha": mine}, f)
    os.replace(part + ".tmp", part)
    queue.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        parts = [json.load(open(os.path.join(a.workdir, f"w{world}", f"rank{r}.json"))) for r in range(world)]
        sha = {}
        for p in parts:
            for k, v in p["sha"].items():
                assert k not in sha, f"video {k} processed twice"
                sha[k] = v
        assert len(sha) == a.videos, (len(sha), a.videos)
        total_clips = int(sum(p["clips"] for p in parts))
        last = max(p["done_s"] for p in parts)
        first = min(p["done_s"] for p in parts)
        # byte-identity: this GPU alone recomputes a strided subset
        check = list(range(0, a.videos, a.check_stride))
        t2 = time.perf_counter()
        mism = [v for v in check if process(v)[0] != sha[str(v)]]
        res = {
            "config": "BASELINE config 4", "videos": a.videos, "world_size": world, "sharding": "static round-robin" if a.static else "dynamic queue (TCP store), longest first",
            "frames_total": int(lens.sum()), "frames_median": int(np.median(lens)), "frames_max": int(lens.max()),
            "clips_total": total_clips, "clip_crops_total": total_clips * 10,
            "seconds": last, "clip_crops_per_s": total_clips * 10 / last, "clip_crops_per_s_per_gpu": total_clips * 10 / last / world,
            "first_rank_done_s": first, "last_rank_done_s": last, "idle_tail_fraction": (last - first) / last,
            "mean_busy_fraction": float(np.mean([p["busy_s"] for p in parts]) / last),
            "per_rank": [{k: p[k] for k in ("rank", "videos", "clips", "busy_s", "done_s")} for p in parts],
            "identity_check": {"videos_recomputed_on_one_gpu": len(check), "mismatches": len(mism), "seconds": time.perf_counter() - t2},
            "wall_s_with_barriers": wall,
        }
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps({k: v for k, v in res.items() if k != "per_rank"}))
        for p in res["per_rank"]:
            print(p)
        assert not mism, f"outputs differ between the sharded run and one GPU: videos {mism}"
    queue.close()


if __name__ == "__main__":
    main()
