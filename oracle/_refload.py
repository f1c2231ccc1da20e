"""Import the UNMODIFIED reference (``/root/reference``) for pinning the oracle.  TEST INFRASTRUCTURE ONLY.

The reference's top-level imports need ``pytorchvideo`` and ``decord``, which are not installed
(and not needed for the classes we pin: ``I3Res50``, ``TenCropVideoFrameDataset``, ``segment``,
``MGFNForVideoAnomalyDetection``).  Empty stub modules are injected for exactly those names; no
reference source is copied or patched.  Only available in the build container -- the GPU box has no
``/root/reference``, so nothing that runs there may call this.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VAD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def load():
    """Returns a namespace with the reference modules: .i3d, .dataset, .extract_features, .models, .loss."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("pytorchvideo", "pytorchvideo.models", "pytorchvideo.models.resnet", "decord"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorchvideo.models.resnet"].create_resnet = lambda *a, **k: (_ for _ in ()).throw(
        RuntimeError("pytorchvideo is not installed"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.i3d = importlib.import_module("src.i3d")
    ns.dataset = importlib.import_module("src.dataset")
    ns.gtransforms = importlib.import_module("src.gtransforms")
    ns.extract_features = importlib.import_module("extract_features")
    ns.models = importlib.import_module("src.models")
    ns.loss = importlib.import_module("src.loss")
    return ns
