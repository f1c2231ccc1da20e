"""ctypes binding of ``libvad_b200.so`` (the C ABI in ``include/vad_b200.h``).

The shared library is built in-tree by ``anomaly_detection_on_video_b200.build`` (``nvcc`` for
sm_100a) and must sit next to this file.  There is deliberately no fallback: if the library is
missing, or a compute entry point is called without an sm_100 GPU, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_uint8, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libvad_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

VAD_OP_CONV, VAD_OP_MAXPOOL, VAD_OP_AVGPOOL = 0, 1, 2
VAD_FLAG_RELU, VAD_FLAG_STEM_FOLD_W, VAD_FLAG_POOL_SAME, VAD_FLAG_FORCE_GATHER, VAD_FLAG_POOL_T2, VAD_FLAG_CONV_SAME = 1, 2, 4, 8, 16, 32
VAD_FLAG_STEM_PLANES = 64
VAD_OUT_DATASET_F32, VAD_OUT_STEM_BF16 = 0, 1

# every symbol include/vad_b200.h declares (tests check the .so exports exactly these)
EXPORTED_SYMBOLS = (
    "vad_last_error",
    "vad_abi_version",
    "vad_plan_create",
    "vad_plan_configure",
    "vad_plan_forward",
    "vad_plan_slot_info",
    "vad_plan_num_launches",
    "vad_plan_flops",
    "vad_plan_profile_begin",
    "vad_plan_profile_select",
    "vad_plan_profile_end",
    "vad_plan_destroy",
    "vad_head_train_create",
    "vad_head_train_param_floats",
    "vad_head_train_bn_floats",
    "vad_head_train_workspace_bytes",
    "vad_head_train_step",
    "vad_head_train_num_launches",
    "vad_head_train_destroy",
    "vad_adam_step",
    "vad_ingest_ncthw_f32",
    "vad_preproc_create",
    "vad_preproc_info",
    "vad_preproc_run",
    "vad_preproc_destroy",
    "vad_segment_mean",
    "vad_add_magnitude",
    "vad_head_create",
    "vad_head_workspace_bytes",
    "vad_head_forward",
    "vad_head_select",
    "vad_head_loss",
    "vad_head_num_launches",
    "vad_head_flops",
    "vad_head_destroy",
    "vad_tf32_plan_create",
    "vad_tf32_plan_configure",
    "vad_tf32_plan_forward",
    "vad_tf32_plan_slot_info",
    "vad_tf32_plan_num_launches",
    "vad_tf32_plan_flops",
    "vad_tf32_plan_destroy",
    "vad_tf32_ingest_ncthw",
    "vad_tf32_ingest_ncthw_planes",
)


class OpDesc(ctypes.Structure):
    """Mirror of ``vad_op_desc``."""

    _fields_ = [
        ("kind", c_int32),
        ("src", c_int32),
        ("dst", c_int32),
        ("res", c_int32),
        ("cin", c_int32),
        ("cout", c_int32),
        ("kt", c_int32),
        ("kh", c_int32),
        ("kw", c_int32),
        ("st", c_int32),
        ("sh", c_int32),
        ("sw", c_int32),
        ("pt", c_int32),
        ("ph", c_int32),
        ("pw", c_int32),
        ("flags", c_int32),
        ("dst_c_off", c_int32),
        ("dst_c_total", c_int32),
        ("w_off", c_uint64),
        ("scale_off", c_uint64),
        ("shift_off", c_uint64),
        ("dst1", c_int32),
        ("dst2", c_int32),
        ("split1", c_int32),
        ("split2", c_int32),
        ("seg_w0", c_int32),
        ("seg_w1", c_int32),
        ("seg_w2", c_int32),
        ("reserved0", c_int32),
    ]


class HeadConfig(ctypes.Structure):
    """Mirror of ``vad_head_config``."""

    _fields_ = [
        ("channels", c_int32),
        ("n_stages", c_int32),
        ("dims", c_int32 * 4),
        ("depths", c_int32 * 4),
        ("types", c_int32 * 4),
        ("dim_head", c_int32),
        ("ff_repe", c_int32),
        ("local_aggr_kernel", c_int32),
        ("k", c_int32),
        ("mag_ratio", c_float),
        ("ln_eps", c_float),
    ]


VAD_HEAD_GLANCE, VAD_HEAD_FOCUS = 0, 1

_lib = None


def load() -> ctypes.CDLL:
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m anomaly_detection_on_video_b200.build` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.vad_last_error.restype = c_char_p
    lib.vad_last_error.argtypes = []
    lib.vad_abi_version.restype = c_int32
    lib.vad_plan_create.restype = c_int32
    lib.vad_plan_create.argtypes = [POINTER(c_void_p), POINTER(OpDesc), c_int32, c_int32, c_void_p, c_uint64, c_int32, c_int32, c_int32]
    lib.vad_plan_configure.restype = c_int32
    lib.vad_plan_configure.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_uint64)]
    lib.vad_plan_forward.restype = c_int32
    lib.vad_plan_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]
    lib.vad_plan_slot_info.restype = c_int32
    lib.vad_plan_slot_info.argtypes = [c_void_p, c_int32, POINTER(c_int32), POINTER(c_uint64), POINTER(c_uint64)]
    lib.vad_plan_num_launches.restype = c_int32
    lib.vad_plan_num_launches.argtypes = [c_void_p]
    lib.vad_plan_flops.restype = c_double
    lib.vad_plan_flops.argtypes = [c_void_p]
    lib.vad_plan_profile_begin.restype = c_int32
    lib.vad_plan_profile_begin.argtypes = [c_void_p]
    lib.vad_plan_profile_select.restype = c_int32
    lib.vad_plan_profile_select.argtypes = [c_void_p, c_int32, c_int32]
    lib.vad_plan_profile_end.restype = c_int32
    lib.vad_plan_profile_end.argtypes = [c_void_p, c_int32, POINTER(c_double), POINTER(c_int32), POINTER(c_double), POINTER(c_double)]
    lib.vad_plan_destroy.restype = None
    lib.vad_plan_destroy.argtypes = [c_void_p]
    lib.vad_ingest_ncthw_f32.restype = c_int32
    lib.vad_ingest_ncthw_f32.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.vad_preproc_create.restype = c_int32
    lib.vad_preproc_create.argtypes = [POINTER(c_void_p), c_int32, c_int32, c_int32, c_int32, c_int32, c_int32]
    lib.vad_preproc_info.restype = c_int32
    lib.vad_preproc_info.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
    lib.vad_preproc_run.restype = c_int32
    lib.vad_preproc_run.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.vad_preproc_destroy.restype = None
    lib.vad_preproc_destroy.argtypes = [c_void_p]
    lib.vad_segment_mean.restype = c_int32
    lib.vad_segment_mean.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.vad_add_magnitude.restype = c_int32
    lib.vad_add_magnitude.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.vad_head_create.restype = c_int32
    lib.vad_head_create.argtypes = [POINTER(c_void_p), POINTER(HeadConfig), c_void_p, c_uint64, c_int32]
    lib.vad_head_workspace_bytes.restype = c_int32
    lib.vad_head_workspace_bytes.argtypes = [c_void_p, c_int32, c_int32, POINTER(c_uint64)]
    lib.vad_head_forward.restype = c_int32
    lib.vad_head_forward.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_uint64, c_void_p, c_void_p,
                                     c_void_p, c_void_p]
    lib.vad_head_select.restype = c_int32
    lib.vad_head_select.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.vad_head_loss.restype = c_int32
    lib.vad_head_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                  c_void_p, c_void_p, c_void_p]
    lib.vad_tf32_plan_create.restype = c_int32
    lib.vad_tf32_plan_create.argtypes = [POINTER(c_void_p), POINTER(OpDesc), c_int32, c_int32, c_void_p, c_uint64, c_int32, c_int32]
    lib.vad_tf32_plan_configure.restype = c_int32
    lib.vad_tf32_plan_configure.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, POINTER(c_uint64)]
    lib.vad_tf32_plan_forward.restype = c_int32
    lib.vad_tf32_plan_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]
    lib.vad_tf32_plan_slot_info.restype = c_int32
    lib.vad_tf32_plan_slot_info.argtypes = [c_void_p, c_int32, POINTER(c_int32), POINTER(c_uint64), POINTER(c_uint64)]
    lib.vad_tf32_plan_num_launches.restype = c_int32
    lib.vad_tf32_plan_num_launches.argtypes = [c_void_p]
    lib.vad_tf32_plan_flops.restype = c_double
    lib.vad_tf32_plan_flops.argtypes = [c_void_p]
    lib.vad_tf32_plan_destroy.restype = None
    lib.vad_tf32_plan_destroy.argtypes = [c_void_p]
    lib.vad_tf32_ingest_ncthw.restype = c_int32
    lib.vad_tf32_ingest_ncthw.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.vad_tf32_ingest_ncthw_planes.restype = c_int32
    lib.vad_tf32_ingest_ncthw_planes.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.vad_head_num_launches.restype = c_int32
    lib.vad_head_num_launches.argtypes = [c_void_p]
    lib.vad_head_flops.restype = c_double
    lib.vad_head_flops.argtypes = [c_void_p, c_int32, c_int32]
    lib.vad_head_destroy.restype = None
    lib.vad_head_destroy.argtypes = [c_void_p]
    lib.vad_head_train_create.restype = c_int32
    lib.vad_head_train_create.argtypes = [POINTER(c_void_p), POINTER(HeadConfig), c_int32]
    lib.vad_head_train_param_floats.restype = c_uint64
    lib.vad_head_train_param_floats.argtypes = [c_void_p]
    lib.vad_head_train_bn_floats.restype = c_uint64
    lib.vad_head_train_bn_floats.argtypes = [c_void_p]
    lib.vad_head_train_workspace_bytes.restype = c_int32
    lib.vad_head_train_workspace_bytes.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(c_uint64)]
    lib.vad_head_train_step.restype = c_int32
    lib.vad_head_train_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                        c_void_p, POINTER(c_float), c_void_p, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.vad_head_train_num_launches.restype = c_int32
    lib.vad_head_train_num_launches.argtypes = [c_void_p]
    lib.vad_head_train_destroy.restype = None
    lib.vad_head_train_destroy.argtypes = [c_void_p]
    lib.vad_adam_step.restype = c_int32
    lib.vad_adam_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_float, c_float, c_float, c_float, c_float,
                                  c_int32, c_float, c_void_p]
    if lib.vad_abi_version() != 1:
        raise RuntimeError(f"{LIB_PATH}: ABI version {lib.vad_abi_version()} != 1; rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Raise ``RuntimeError`` carrying ``vad_last_error()`` when a call failed."""
    if rc != 0:
        msg = load().vad_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libvad_b200 {what} failed (status {rc}): {msg}")


__all__ = ["OpDesc", "HeadConfig", "load", "check", "LIB_PATH", "EXPORTED_SYMBOLS"]
# keep the ctypes scalar types importable from here for the thin wrappers
_ = (c_float, c_uint8)
