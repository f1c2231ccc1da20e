"""Keep the host side of one GPU's extraction on the CPU socket that GPU hangs off.

The end-to-end path uploads 230 KB per frame from pinned host memory while the backbone runs (dataset.py); when the
pinned pages sit on the other socket every byte crosses the inter-socket link first and the upload no longer hides
behind the compute.  ``bind_to_gpu`` restricts the calling process to the CPUs NVML reports as local to the GPU
(intersected with the cpuset the container allows) BEFORE the pinned buffers are allocated, so first touch places them
on the right node.  Best effort: without NVML, or when the allowed cpuset has no CPU local to the GPU, nothing changes.

No reference counterpart: the reference loads frames in DataLoader workers (extract_features.py:79-84) and leaves
placement to the OS.
"""
from __future__ import annotations

import os
from typing import Dict


def bind_to_gpu(device_index: int) -> Dict[str, object]:
    info: Dict[str, object] = {"bound": False, "allowed_cpus": len(os.sched_getaffinity(0))}
    if os.environ.get("VAD_NO_NUMA_BIND") == "1":
        info["why"] = "disabled by VAD_NO_NUMA_BIND"
        return info
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        nvml_index = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                nvml_index = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(nvml_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        local = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        try:
            info["gpu_numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
    except Exception as exc:  # NVML missing or refusing: leave placement to the OS
        info["why"] = f"{type(exc).__name__}: {exc}"
        return info
    allowed = os.sched_getaffinity(0)
    target = allowed & local
    info["gpu_local_cpus"] = len(local)
    if not target:
        info["why"] = "no allowed CPU is local to the GPU"
        return info
    if target != allowed:
        os.sched_setaffinity(0, target)
    info["bound"] = True
    info["cpus"] = len(target)
    return info
