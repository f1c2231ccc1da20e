"""Drop-in for the reference CLI module ``extract_features.py``: same functions, same on-disk layout.

    main(outdir)                                   extract_features.py:43-52
    extract(dataset, model, device, outpath)       extract_features.py:55-156
    segment(feature_path, seg_outpath, seg_length) extract_features.py:159-185
    load_feature_extraction_model(model_name)      extract_features.py:34-40

Files written (extract_features.py:106-107,126-131,156,165,183-185):
    <outpath>/<stem>_i3d.npy              (n_clips, 10, C) float32   [(10, C) when n_clips == 1, see below]
    <outpath>/<stem>/<stem>_<k>.npy       per-chunk cache for videos larger than 1 GB (3008-frame chunks)
    <seg_outpath>/<stem>_i3d.npy          (10, seg_length, C) float32

What differs from the reference, on purpose:
  * the 10 crops of a batch run as ONE forward of 10x the batch (the reference loops over crops and
    synchronises on a D2H copy after each), features stay on the GPU until the video is done;
  * preprocessing is one fused kernel fed from a single uint8 upload of the frames;
  * files are written to a temp name and renamed, so a crash never leaves a truncated .npy that the
    skip-if-exists logic would then trust;
  * ``segment`` creates its output directory (the reference raises FileNotFoundError) and an
    exact multiple of 3008 frames does not produce an empty trailing chunk.
``strict_compat=True`` (default) keeps the reference's ``np.squeeze`` quirk: a one-clip video is saved
as (10, C) rather than (1, 10, C).
"""
from __future__ import annotations

import argparse
import os
from typing import Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .dataset import FrameSource, TenCropVideoFrameDataset
from .engine import segment_mean
from .i3d import _NativeBackbone, build_i3d_feature_extractor
from .workqueue import WorkQueue

DEFAULT_REPO_ID = "jinmang2/ucf_crime"
DEFAULT_DATASET_CONF_NAME = "anomaly"
DEFAULT_CACHE_DIR = "/content/drive/MyDrive/ucf_crime"
CHUNK_FRAMES = 16 * 188  # 3,008 (extract_features.py:121)
CLIPS_PER_BATCH = 16     # extract_features.py:79


def load_ucf_crime_dataset(repo_id: str = DEFAULT_REPO_ID, cache_dir: str = DEFAULT_CACHE_DIR,
                           config_name: str = DEFAULT_DATASET_CONF_NAME):
    """extract_features.py:26-31 (needs the HF hub; out of scope offline, kept for API parity)."""
    from datasets import load_dataset

    return load_dataset(repo_id, config_name, cache_dir=cache_dir)


def load_feature_extraction_model(model_name: str = "tushar-n-baseline", state_dict_path: Optional[str] = None,
                                  device: Optional[torch.device] = None, precision: str = "bf16") -> Tuple[torch.nn.Module, torch.device]:
    """extract_features.py:34-40.  The default is the reference *factory's* default backbone (I3Res50);
    the reference CLI's own default, pytorchvideo's ``i3d_8x8_r50``, is built from its published architecture (third-party,
    parity unpinned: ``ptv_resnet.I3D8x8R50``).
    ``model_name="inception-i3d"`` builds the InceptionV1-3D backbone BASELINE.json names (1024-d features; not part of
    the reference).  ``precision``: "bf16" (production) or "tf32" (fp32 activations, features within 1e-3)."""
    if precision not in ("bf16", "tf32"):
        raise ValueError(f"precision must be 'bf16' or 'tf32', not {precision!r}")
    if not torch.cuda.is_available():
        raise RuntimeError("feature extraction runs on sm_100a GPUs only; no CUDA device is visible")
    if model_name == "inception-i3d":
        from .inception import InceptionI3d

        model = InceptionI3d()
        if state_dict_path is not None:
            model.load_state_dict(torch.load(state_dict_path, map_location="cpu"), strict=True)
    else:
        model = build_i3d_feature_extractor(model_name=model_name, state_dict_path=state_dict_path)
    model.precision = precision
    model.eval()
    model.to(device if device is not None else "cuda")
    return model, next(model.parameters()).device


def _atomic_save(path: str, arr: np.ndarray) -> None:
    """Write-then-rename.  The temporary is a dot file that does NOT end in ``.npy`` (np.save gets an open handle, so it
    appends no suffix): a rank killed mid-write leaves nothing that ``segment()`` or a consumer globbing ``*.npy`` picks up."""
    d, base = os.path.split(path)
    tmp = os.path.join(d, f".{base}.{os.getpid()}.tmp")
    with open(tmp, "wb") as fh:
        np.save(fh, arr)
    os.replace(tmp, path)


def _stem_name(video_path: str) -> str:
    return video_path.split(os.sep)[-1].split(".")[0]  # extract_features.py:106


@torch.no_grad()
def extract_clip_features(video_dataset: TenCropVideoFrameDataset, model: torch.nn.Module, device: torch.device,
                          clips_per_batch: int = CLIPS_PER_BATCH, strict_compat: bool = True,
                          as_numpy: bool = True) -> Union[np.ndarray, torch.Tensor]:
    """The reference's inner ``_extract`` (extract_features.py:77-102): all clips of one frame container
    -> (n_clips, ncrops, C)."""
    n_clips, k = len(video_dataset), video_dataset.ncrops
    feats: Optional[torch.Tensor] = None
    if isinstance(model, _NativeBackbone) and model.precision == "bf16":
        stem_buf = None
        for start in range(0, n_clips, clips_per_batch):
            n = min(clips_per_batch, n_clips - start)
            if stem_buf is None or stem_buf.shape[0] != n * k:
                stem_buf = None
                stem_buf = video_dataset.clips_stem(start, n, pad_left=model.pad_left)
            else:
                video_dataset.clips_stem(start, n, pad_left=model.pad_left, out=stem_buf)
            f = model.forward_stem_layout(stem_buf)  # (n * k, C), clip-major then crop
            if feats is None:
                feats = torch.empty(n_clips, k, f.shape[1], dtype=torch.float32, device=f.device)
            feats[start:start + n] = f.view(n, k, -1)
    else:
        # any other nn.Module (and the native backbones in tf32 mode, which take the reference's fp32 NCTHW crops): same
        # call pattern as the reference, one forward per crop index
        for start in range(0, n_clips, clips_per_batch):
            n = min(clips_per_batch, n_clips - start)
            inputs = video_dataset.clips_f32(start, n).permute(0, 1, 3, 2, 4, 5)  # (B, 10, 3, 16, H, W), :83
            crops = [model(inputs[:, c].to(device)).detach().reshape(n, -1) for c in range(inputs.shape[1])]
            f = torch.stack(crops, dim=1)
            if feats is None:
                feats = torch.empty(n_clips, k, f.shape[2], dtype=torch.float32, device=f.device)
            feats[start:start + n] = f
    if strict_compat:
        feats = feats.squeeze()  # np.squeeze at extract_features.py:100
    return feats.cpu().numpy() if as_numpy else feats


def extract_stream(videos: Iterable, model: torch.nn.Module, device: torch.device, *, clips_per_batch: int = CLIPS_PER_BATCH,
                   seg_length: Optional[int] = 32, frames_per_clip: int = 16, ncrops: int = 10):
    """Features of a sequence of videos with the host <-> device traffic of neighbouring videos overlapped.

    ``videos`` yields frame containers (anything ``TenCropVideoFrameDataset`` accepts).  For each one, in order, yields
    ``(features, segments)`` as host tensors: ``(n_clips, ncrops, C)`` fp32 and ``(ncrops, seg_length, C)`` fp32 (``None``
    when ``seg_length`` is None) -- what ``extract`` + ``segment`` (extract_features.py:55-185) write per video.
    Software pipeline, one video deep: while video *i* runs on the compute stream, the frames of video *i+1* are already
    uploading (own stream) and the results of video *i-1* are copied back into pinned buffers; the host blocks only on
    the event of the video it is about to yield.  No collective, no CPU compute.
    """
    from .engine import segment_mean

    pending = None  # (features_host, segments_host, event, keep-alive) of the previous video

    def finish(p):
        p[2].synchronize()
        return p[0], p[1]

    for frames in videos:
        ds = TenCropVideoFrameDataset(frames, frames_per_clip=frames_per_clip, device=device, ncrops=ncrops)  # upload starts here
        feats = extract_clip_features(ds, model, device, clips_per_batch=clips_per_batch, strict_compat=False, as_numpy=False)
        seg = segment_mean(feats, seg_length) if seg_length else None
        f_host = torch.empty(feats.shape, dtype=feats.dtype, pin_memory=True)
        f_host.copy_(feats, non_blocking=True)
        s_host = None
        if seg is not None:
            s_host = torch.empty(seg.shape, dtype=seg.dtype, pin_memory=True)
            s_host.copy_(seg, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        if pending is not None:
            yield finish(pending)  # everything of video i is queued before the host waits for video i-1
        pending = (f_host, s_host, ev, (ds, feats, seg))
    if pending is not None:
        yield finish(pending)


def _iter_rows(dataset) -> Iterable[Mapping]:
    for row in dataset:
        yield row


def extract(dataset, model: torch.nn.Module, device: torch.device, outpath: str, *, strict_compat: bool = True,
            clips_per_batch: int = CLIPS_PER_BATCH, chunk_frames: int = CHUNK_FRAMES,
            queue: Optional[WorkQueue] = None) -> Optional[List[str]]:
    """Per-video extraction loop (extract_features.py:55-156).

    ``dataset``: a ``datasets.DatasetDict`` / dict of splits (recursed, one sub-directory per split) or
    an iterable of rows with ``video_path`` and ``size`` (KB) -- a ``datasets.Dataset`` or a plain list
    of dicts.  ``queue``: when several processes (one per GPU) run this on the same row list, the
    shared work queue hands every row to exactly one of them; there is no collective on this path.
    """
    if isinstance(dataset, Mapping):  # datasets.DatasetDict is a dict of splits
        for mode, dset in dataset.items():
            extract(dset, model, device, os.path.join(outpath, mode), strict_compat=strict_compat,
                    clips_per_batch=clips_per_batch, chunk_frames=chunk_frames, queue=queue)
        return None
    if not hasattr(dataset, "__iter__"):
        raise AssertionError(
            "The type of dataset argument must be `datasets.Dataset` or `datasets.DatasetDict` "
            f"(or a list of rows). Your input's type is {type(dataset)}.")
    os.makedirs(outpath, exist_ok=True)
    rows = list(_iter_rows(dataset))
    if queue is not None:
        # longest first (by the size column) limits the tail when videos are spread over GPUs
        order = sorted(range(len(rows)), key=lambda i: (-float(rows[i].get("size", 0)), i))
        todo = (order[j] for j in queue.claim(len(rows), tag=outpath))
    else:
        todo = iter(range(len(rows)))
    written: List[str] = []
    from collections import deque

    streamed: "deque[str]" = deque()  # save paths of the videos currently inside the pipeline, in order

    def small_videos():
        """Claims rows; videos above the reference's 1 GB threshold are chunked right here (sequentially, with the
        chunk cache), everything else is handed to extract_stream so that neighbouring videos overlap."""
        for ri in todo:
            sample = rows[ri]
            filename = _stem_name(sample["video_path"])
            savepath = os.path.join(outpath, filename + "_i3d.npy")
            if os.path.exists(savepath):  # idempotent resume, extract_features.py:109-110
                continue
            if sample.get("size", 0) > 1024 ** 2:  # > 1 GB (size is in KB), extract_features.py:116
                _extract_chunked(sample, filename, savepath)
                continue
            streamed.append(savepath)
            yield sample["video_path"]

    def _extract_chunked(sample, filename, savepath):
        src = FrameSource(sample["video_path"])
        n_seg = (len(src) + chunk_frames - 1) // chunk_frames
        seg_folder = os.path.join(outpath, filename)
        os.makedirs(seg_folder, exist_ok=True)
        segments = []
        for seg in range(n_seg):
            seg_savepath = os.path.join(seg_folder, filename + f"_{seg}.npy")
            if os.path.exists(seg_savepath):
                outputs = np.load(seg_savepath)
            else:
                frames = src.read(seg * chunk_frames, (seg + 1) * chunk_frames)
                outputs = extract_clip_features(TenCropVideoFrameDataset(frames, device=device), model, device,
                                                clips_per_batch, strict_compat)
                _atomic_save(seg_savepath, outputs)
            segments.append(outputs)
        outputs = np.vstack(segments)  # extract_features.py:148
        _atomic_save(savepath, outputs)
        written.append(savepath)

    for f_host, _ in extract_stream(small_videos(), model, device, clips_per_batch=clips_per_batch, seg_length=None):
        savepath = streamed.popleft()
        outputs = f_host.numpy()
        if strict_compat:
            outputs = np.squeeze(outputs)  # np.squeeze at extract_features.py:100
        _atomic_save(savepath, outputs)
        written.append(savepath)
    return written


def segment(feature_path: str, seg_outpath: str, seg_length: int = 32, queue: Optional[WorkQueue] = None) -> None:
    """32-segment averaging of every ``.npy`` in ``feature_path`` (extract_features.py:159-185)."""
    os.makedirs(seg_outpath, exist_ok=True)
    files = [f for f in sorted(os.listdir(feature_path))
             if f.endswith(".npy") and not f.startswith(".") and ".tmp." not in f and os.path.isfile(os.path.join(feature_path, f))]
    idxs: Iterable[int] = queue.claim(len(files), tag=seg_outpath) if queue is not None else range(len(files))
    for i in idxs:
        file = files[i]
        savepath = os.path.join(seg_outpath, file)
        if os.path.exists(savepath):
            continue
        features = np.load(os.path.join(feature_path, file))
        if features.ndim == 2:  # a one-clip video squeezed by the reference
            features = features[None]
        out = segment_mean(torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)).cuda(), seg_length)
        _atomic_save(savepath, out.cpu().numpy())


def main(outdir: str = "/content/drive/MyDrive/ucf_crime", dataset=None, model_name: str = "tushar-n-baseline",
         state_dict_path: Optional[str] = None, queue: Optional[WorkQueue] = None, precision: str = "bf16") -> None:
    """extract_features.py:43-52."""
    outpath = os.path.join(outdir, "anomaly_features")
    anomaly = dataset if dataset is not None else load_ucf_crime_dataset()
    model, device = load_feature_extraction_model(model_name, state_dict_path, precision=precision)
    extract(anomaly, model, device, outpath, queue=queue)
    seg_length = 32
    seg_outpath = os.path.join(outdir, f"segment_features_{seg_length}")
    train_dir = os.path.join(outpath, "train")
    if os.path.isdir(train_dir):  # segments only for the train split, extract_features.py:51-52
        if queue is not None:
            queue.barrier()
        segment(train_dir, seg_outpath, seg_length, queue=queue)


def _rows_from_dir(video_dir: str) -> List[dict]:
    rows = []
    for f in sorted(os.listdir(video_dir)):
        p = os.path.join(video_dir, f)
        if os.path.isfile(p) and f.rsplit(".", 1)[-1].lower() in ("npy", "mp4", "avi", "mkv", "mov"):
            rows.append({"video_path": p, "size": os.path.getsize(p) // 1024})
    return rows


def cli(argv: Optional[Sequence[str]] = None) -> None:
    ap = argparse.ArgumentParser(description="I3D ten-crop snippet-feature extraction on B200")
    ap.add_argument("--outdir", required=True)
    ap.add_argument("--videos", help="directory of videos (.npy frame arrays or containers); one 'train' split")
    ap.add_argument("--weights", help="state_dict for the backbone (reference checkpoint format)")
    ap.add_argument("--model-name", default="tushar-n-baseline", help="tushar-n-baseline (I3Res50, 2048-d) or inception-i3d (1024-d)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"],
                    help="bf16: bf16 operands / activations, fp32 accumulate (features within 1e-2 of the fp32 reference); "
                         "tf32: fp32 activations, TF32 tensor-core products (within 1e-3, ~5x slower)")
    args = ap.parse_args(argv)
    queue = WorkQueue.from_env()
    if queue is not None:
        torch.cuda.set_device(queue.local_rank)
    from .hostaffinity import bind_to_gpu

    bind_to_gpu(queue.local_rank if queue is not None else 0)  # pinned frame buffers land on the GPU's own socket
    dataset = {"train": _rows_from_dir(args.videos)} if args.videos else None
    try:
        main(args.outdir, dataset=dataset, model_name=args.model_name, state_dict_path=args.weights, queue=queue, precision=args.precision)
    finally:
        if queue is not None:
            queue.close()  # the rank that hosts the store outlives every other rank's last claim


if __name__ == "__main__":
    cli()
