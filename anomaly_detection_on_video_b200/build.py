"""Build ``libvad_b200.so`` in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

    python -m anomaly_detection_on_video_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libvad_b200.so")
SOURCES = ["vad_api.cu"]
HEADERS = ["ptx_sm100.cuh", "conv_umma.cuh", "stem_umma.cuh", "conv_thalo.cuh", "conv_s3x3.cuh", "conv_pair.cuh", "stem_pair.cuh", "conv_tail.cuh", "aux_kernels.cuh", "aux_api.cuh", "plan_configure.cuh", "plan_bind.cuh", "plan_run.cuh", "head_kernels.cuh", "head_api.cuh", "head_train_kernels.cuh", "head_train_api.cuh", "tf32_kernels.cuh", "tf32_api.cuh", os.path.join("..", "..", "include", "vad_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libvad_b200.so")
    return nvcc


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    out_m = os.path.getmtime(OUT)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > out_m:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [
        os.path.join(CSRC, s) for s in SOURCES
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
