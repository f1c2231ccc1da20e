"""The C-ABI library builds, loads without a GPU, exports exactly what include/vad_b200.h declares,
and refuses to compute when there is no sm_100 device (no silent fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vad_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vad_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_match_binding_table():
    from anomaly_detection_on_video_b200 import _lib

    assert _declared_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(native_lib):
    for name in _declared_functions():
        assert hasattr(native_lib, name), f"{name} declared in vad_b200.h but not exported"
    assert native_lib.vad_abi_version() == 1


def test_library_has_no_driver_link_dependency():
    from anomaly_detection_on_video_b200 import _lib, build

    build.build()
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out, "libcuda must be resolved at run time so the library loads on a build box"
    assert "libtorch" not in out and "libc10" not in out, "no torch types may cross the C ABI"


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from anomaly_detection_on_video_b200 import _lib, build

    build.build()
    res = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = res.stdout
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass and "IM2COL" in sass, "TMA (tiled + im2col) loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    assert "HMMA.16816" not in sass, "legacy mma.sync path must not be present"


def test_compute_calls_fail_loudly_without_gpu(native_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is for GPU-less hosts")
    h = ctypes.c_void_p()
    rc = native_lib.vad_preproc_create(ctypes.byref(h), 240, 320, 256, 224, 10, 0)
    assert rc != 0
    assert b"no CPU fallback" in native_lib.vad_last_error()


def test_python_api_refuses_cpu_tensors(native_lib):
    import torch

    from anomaly_detection_on_video_b200 import engine
    from anomaly_detection_on_video_b200.i3d import I3Res50

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.segment_mean(torch.zeros(4, 10, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        I3Res50().eval()(torch.zeros(1, 3, 8, 32, 32))


def test_tf32_and_pair_paths_are_in_the_binary():
    """The TF32 mode and the CTA-pair kernels are compiled into the same library: kind::tf32 and cta_group::2 MMAs in SASS."""
    from anomaly_detection_on_video_b200 import _lib, build

    build.build()
    res = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = res.stdout
    assert "conv_tf32_kernel" in sass and "conv_pair_kernel" in sass
    assert "2CTA" in sass, "tcgen05.mma.cta_group::2 missing from SASS"


def test_tf32_mode_fails_loudly_without_gpu(native_lib):
    import torch

    from anomaly_detection_on_video_b200 import engine
    from anomaly_detection_on_video_b200.i3d import I3Res50

    m = I3Res50().eval()
    m.precision = "tf32"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 8, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.ingest_ncthw_tf32(torch.zeros(1, 3, 8, 32, 32))
    m.precision = "fp16"
    if torch.cuda.is_available():
        with pytest.raises(ValueError):
            m(torch.zeros(1, 3, 8, 32, 32, device="cuda"))
