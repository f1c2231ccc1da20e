"""Drop-in for the reference clip source ``TenCropVideoFrameDataset`` (src/dataset.py:145-195) and the
``add_magnitude`` step of its ``FeatureDataset`` (src/dataset.py:121-124).

Same constructor, ``len()`` and ``[i] -> (10, 16, 3, 224, 224) float32`` contract; the whole transform
chain of src/gtransforms.py runs as ONE fused sm_100a kernel (``vad_preproc_run``), bit-identical to
the reference's PIL/torchvision result.  Frames are uploaded once as uint8 and every clip tensor is
produced on the GPU; ``clips_stem`` hands the native backbone its bf16 input layout directly.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib
from .engine import Preprocessor, add_magnitude as _add_magnitude

try:  # PIL is only needed to accept List[PIL.Image] like the reference does
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None

BILINEAR = 2  # PIL.Image.BILINEAR


class FrameSource:
    """Random access to decoded frames: ``len()`` and ``[i] -> (H, W, 3) uint8``.

    Stands where ``decord.VideoReader`` stands in the reference (src/dataset.py:155-159,
    extract_features.py:123).  ``.npy`` files hold pre-decoded frames (synthetic videos); real
    containers are decoded with decord when installed, else OpenCV.  Decoding is host I/O and not
    part of the accelerated path.
    """

    def __init__(self, uri: str) -> None:
        self.uri = uri
        self._arr = None
        self._reader = None
        self._cv = None
        if uri.endswith(".npy"):
            self._arr = np.load(uri, mmap_mode="r")
            if self._arr.ndim != 4 or self._arr.shape[3] != 3 or self._arr.dtype != np.uint8:
                raise ValueError(f"{uri}: expected a uint8 [n_frames, H, W, 3] array")
            return
        try:
            import decord  # type: ignore

            self._reader = decord.VideoReader(uri=uri)
            return
        except ImportError:
            pass
        import cv2  # type: ignore

        cap = cv2.VideoCapture(uri)
        if not cap.isOpened():
            raise RuntimeError(f"cannot open video {uri}")
        frames = []
        while True:
            ok, bgr = cap.read()
            if not ok:
                break
            frames.append(bgr[:, :, ::-1])
        cap.release()
        if not frames:
            raise RuntimeError(f"{uri}: no frames decoded")
        self._arr = np.stack(frames)

    def __len__(self) -> int:
        return len(self._arr) if self._arr is not None else len(self._reader)

    def __getitem__(self, i: int) -> np.ndarray:
        if self._arr is not None:
            return np.asarray(self._arr[i])
        return self._reader[i].asnumpy()

    def read(self, start: int, stop: int) -> np.ndarray:
        stop = min(stop, len(self))
        if self._arr is not None:
            return np.ascontiguousarray(self._arr[start:stop])
        return np.stack([self._reader[i].asnumpy() for i in range(start, stop)])


def _from_numpy_readonly(a: np.ndarray) -> torch.Tensor:
    """``torch.from_numpy`` without the non-writable warning: memory-mapped / read-only frame arrays are only ever read
    (they are copied to pinned memory or to the GPU next)."""
    import warnings

    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
        return torch.from_numpy(a)


def _frames_to_tensor(src) -> torch.Tensor:
    """Any accepted frame container -> uint8 [n, H, W, 3] tensor (CPU or already on the GPU)."""
    if isinstance(src, torch.Tensor):
        t = src
    elif isinstance(src, np.ndarray):
        t = _from_numpy_readonly(np.ascontiguousarray(src))
    elif isinstance(src, FrameSource):
        t = _from_numpy_readonly(src.read(0, len(src)))
    elif isinstance(src, str):
        fs = FrameSource(src)
        t = _from_numpy_readonly(fs.read(0, len(fs)))
    elif isinstance(src, (list, tuple)) and len(src) > 0 and Image is not None and isinstance(src[0], Image.Image):
        t = torch.from_numpy(np.stack([np.asarray(im.convert("RGB")) for im in src]))
    else:
        raise ValueError(
            "The type of `video_path_or_images` must be either `str` or `List[PIL.Image.Image]` "
            "(a uint8 [n, H, W, 3] numpy array / torch tensor is accepted too). "
            f"The type of your input is {type(src)}."
        )
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
        raise ValueError("frames must be uint8 [n_frames, H, W, 3]")
    return t


# Preprocessor handles (resize tap tables + crop table on the device) depend only on the frame geometry, never on the
# frames: one per geometry and device is shared by every dataset of the process.  Creating one costs several
# cudaMalloc / synchronous upload / (on destruction) cudaFree calls -- 10-150 ms of host time once the context holds the
# backbone's multi-GB workspace -- which would otherwise sit in front of every video, where nothing can hide it.
_PP_CACHE: "OrderedDict[tuple, Preprocessor]" = OrderedDict()
_PP_CACHE_MAX = 32


_UPLOAD_STREAMS: dict = {}


def _upload_stream(device: torch.device) -> "torch.cuda.Stream":
    index = device.index if device.index is not None else torch.cuda.current_device()
    st = _UPLOAD_STREAMS.get(index)
    if st is None:
        st = torch.cuda.Stream(torch.device("cuda", index))
        _UPLOAD_STREAMS[index] = st
    return st


def _shared_preprocessor(src_h: int, src_w: int, resize: int, crop: int, ncrops: int, device: torch.device) -> Preprocessor:
    index = device.index if device.index is not None else torch.cuda.current_device()
    key = (int(src_h), int(src_w), int(resize), int(crop), int(ncrops), int(index))
    pp = _PP_CACHE.get(key)
    if pp is None:
        pp = Preprocessor(src_h, src_w, resize, crop, ncrops, device)
        _PP_CACHE[key] = pp
        while len(_PP_CACHE) > _PP_CACHE_MAX:
            _PP_CACHE.popitem(last=False)
    else:
        _PP_CACHE.move_to_end(key)
    return pp


class TenCropVideoFrameDataset(Dataset):
    """GPU clip source with the reference's interface (src/dataset.py:145-195)."""

    def __init__(self, video_path_or_images: Union[str, Sequence, np.ndarray, torch.Tensor], frames_per_clip: int = 16,
                 resize: int = 256, cropsize: int = 224, resample: int = BILINEAR,
                 device: Optional[torch.device] = None, ncrops: int = 10) -> None:
        if resample != BILINEAR:
            raise NotImplementedError("only PIL.Image.BILINEAR (the reference default) is implemented")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("TenCropVideoFrameDataset preprocesses on the GPU only (no CPU fallback)")
        frames = _frames_to_tensor(video_path_or_images)
        self._ready = []  # (frames uploaded so far, event) per chunk of the pipelined H2D copy
        if not frames.is_cuda:
            # One upload of the raw uint8 frames, pipelined: chunks go out on a copy stream and every
            # preprocessing launch waits (on the device, not the host) only for the chunk it needs, so the
            # PCIe transfer overlaps the backbone instead of preceding it.
            frames = frames.contiguous()
            if not frames.is_pinned():
                try:
                    frames = frames.pin_memory()
                except RuntimeError:
                    pass
            # The destination is ALLOCATED on the (per-device, long-lived) upload stream: the caching allocator then
            # orders its reuse against that stream, so the copies need not wait for whatever is still queued on the
            # compute stream -- the upload of the next video overlaps the backbone of the current one
            # (extract_features.extract_stream).  The compute stream reads it, hence record_stream.
            copy_stream = _upload_stream(self.device)
            chunk = max(frames_per_clip, (256 // frames_per_clip) * frames_per_clip)
            with torch.cuda.stream(copy_stream):
                dst = torch.empty(frames.shape, dtype=torch.uint8, device=self.device)
                for c0 in range(0, frames.shape[0], chunk):
                    c1 = min(frames.shape[0], c0 + chunk)
                    dst[c0:c1].copy_(frames[c0:c1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    self._ready.append((c1, ev))
            dst.record_stream(torch.cuda.current_stream(self.device))
            self._host_frames = frames  # keep the pinned source alive until the copies are done
            frames = dst
        self.frames = frames.contiguous()
        self.frames_per_clip = frames_per_clip
        self.ncrops = ncrops
        self.cropsize = cropsize
        n_frames = self.frames.shape[0]
        self.indices = list(range((n_frames - 1) // frames_per_clip + 1))  # src/dataset.py:171-173
        self._pp = _shared_preprocessor(self.frames.shape[1], self.frames.shape[2], resize, cropsize, ncrops, self.device)

    def __len__(self) -> int:
        return len(self.indices)

    def _wait_uploaded(self, clip_end: int) -> None:
        """Make the current stream wait for the H2D chunks covering clips [0, clip_end)."""
        if not self._ready:
            return
        need = min(self.frames.shape[0], clip_end * self.frames_per_clip)
        cur = torch.cuda.current_stream(self.device)
        while self._ready:
            upto, ev = self._ready[0]
            cur.wait_event(ev)
            if upto >= need:
                break
            self._ready.pop(0)  # earlier chunks are implied by later events on the same copy stream

    def __getitem__(self, idx: int) -> torch.Tensor:
        """(ncrops, clip_len, 3, H, W) float32 on the GPU -- values identical to the reference's."""
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        self._wait_uploaded(idx + 1)
        return self._pp.run(self.frames, idx, 1, self.frames_per_clip, _lib.VAD_OUT_DATASET_F32)[0]

    def clips_f32(self, start: int, n: int) -> torch.Tensor:
        """(n, ncrops, clip_len, 3, H, W) float32: ``torch.stack([self[i] for i in range(start, start+n)])``."""
        self._wait_uploaded(start + n)
        return self._pp.run(self.frames, start, n, self.frames_per_clip, _lib.VAD_OUT_DATASET_F32)

    def clips_stem(self, start: int, n: int, pad_left: int = 3, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bf16 stem layout (n * ncrops, clip_len, H, W + 8, 4), clip-major then crop: the direct input of
        ``I3Res50.forward_stem_layout`` (skips the fp32 NCTHW tensor the reference materialises)."""
        self._wait_uploaded(start + n)
        return self._pp.run(self.frames, start, n, self.frames_per_clip, _lib.VAD_OUT_STEM_BF16, pad_left, out=out)


def add_magnitude(feature: Union[np.ndarray, torch.Tensor]) -> Union[np.ndarray, torch.Tensor]:
    """``FeatureDataset.add_magnitude`` (src/dataset.py:121-124) on the GPU: (..., C) -> (..., C + 1)."""
    if isinstance(feature, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(feature, dtype=np.float32)).cuda()
        return _add_magnitude(t).cpu().numpy()
    return _add_magnitude(feature)


class FeatureDataset(Dataset):
    """Drop-in for the reference ``FeatureDataset`` (src/dataset.py:96-142): one item per ``<video>_i3d.npy`` feature
    file, ``{"feature": (crops, T, C + 1) fp32 with the L2 magnitude appended, "anomaly": 0.0 / 1.0 from the file name,
    ["label": per-frame ground truth]}``.  ``values`` maps a file name to an array or, with ``open_func``, to whatever
    ``np.load(open_func(value))`` can read (the reference's lazily opened zip members).  The magnitude column is
    computed by ``vad_add_magnitude`` on ``device`` (the reference uses ``np.linalg.norm``; they agree to fp32
    rounding, see tests/test_gpu_kernels.py)."""

    def __init__(self, filenames, values, labels=None, open_func=None, device: Optional[torch.device] = None) -> None:
        self.filenames = list(filenames)
        self.values = values
        self.labels = labels
        self.open_func = open_func
        self.device = torch.device(device if device is not None else "cuda")

    def __len__(self) -> int:
        return len(self.values)

    def open(self, value):
        if self.open_func is None:
            return value
        return np.load(self.open_func(value))  # dynamic loading (src/dataset.py:115-119)

    def add_magnitude(self, feature: np.ndarray) -> np.ndarray:
        if self.device.type != "cuda":
            raise RuntimeError("FeatureDataset.add_magnitude runs on the GPU only (no CPU fallback)")
        t = torch.from_numpy(np.ascontiguousarray(feature, dtype=np.float32)).to(self.device)
        return _add_magnitude(t).cpu().numpy()

    def get_filename(self, idx: int) -> str:
        return self.filenames[idx]

    def __getitem__(self, idx: int):
        fname = self.get_filename(idx)
        feature = self.open(self.values[fname])
        outputs = {"feature": self.add_magnitude(feature),
                   "anomaly": np.array(0.0 if "Normal" in fname else 1.0, dtype=np.float32)}
        if self.labels is not None:
            outputs["label"] = np.array(self.labels[fname], dtype=np.float32)
        return outputs


def build_feature_dataset(mode: str = "train", local_path: Optional[str] = None, filename: Optional[str] = None,
                          ground_truth: Optional[str] = None, dynamic_load: bool = True,
                          device: Optional[torch.device] = None):
    """``build_feature_dataset`` (src/dataset.py:20-93) for a LOCAL ``train.zip`` / ``test.zip`` of ``*_i3d.npy``
    feature files (the reference downloads them from the HF hub; there is no network here): ``train`` returns
    ``{"normal": FeatureDataset, "abnormal": FeatureDataset}`` split on "Normal" in the file name, ``test`` one dataset
    with the per-frame labels of ``ground_truth`` (a json: file name -> list)."""
    import json
    import zipfile

    assert mode in ("train", "test")
    if local_path is None or filename is None:
        raise RuntimeError("no network: pass local_path and filename of the feature archive "
                           "(the reference's default downloads jinmang2/ucf_crime_tencrop_i3d_seg32)")
    zipf = zipfile.ZipFile(os.path.join(local_path, filename))
    names, values = [], {}
    for member in zipf.infolist():
        if member.is_dir():
            continue
        name = member.filename.split("/")[-1]
        names.append(name)
        values[name] = member if dynamic_load else np.load(zipf.open(member))
    open_func = zipf.open if dynamic_load else None
    if mode == "test":
        labels = json.load(open(ground_truth)) if ground_truth is not None else None
        return FeatureDataset(names, values, labels=labels, open_func=open_func, device=device)
    normal = [n for n in names if "Normal" in n]
    abnormal = [n for n in names if "Normal" not in n]
    return {"normal": FeatureDataset(normal, {n: values[n] for n in normal}, open_func=open_func, device=device),
            "abnormal": FeatureDataset(abnormal, {n: values[n] for n in abnormal}, open_func=open_func, device=device)}


__all__ = ["TenCropVideoFrameDataset", "FrameSource", "FeatureDataset", "build_feature_dataset", "add_magnitude", "BILINEAR"]
_ = List
