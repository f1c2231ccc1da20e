"""-m gpu: stand-ins for compute-sanitizer's memcheck / initcheck, which is closed on this GPU pool (profiles/r02_sanitizer.md).

Every workspace the library is handed (backbone plan, scoring head, training step) is embedded in a larger buffer with guard
bands of a sentinel byte on both sides and poisoned with a NaN bit pattern before the launches:
  * the guard bands must be untouched afterwards (no write outside the workspace the call was given), and
  * the results must be bit-identical to a run on a zero-filled workspace (nothing reads memory it did not write first)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 1 << 20
SENTINEL = 0xA5


def _guarded(nbytes: int, device, poison: bool):
    """(buffer, aligned pointer of the usable region, usable-region view)."""
    buf = torch.full((nbytes + 2 * GUARD + 2048,), SENTINEL, dtype=torch.uint8, device=device)
    ptr = (buf.data_ptr() + GUARD + 1023) // 1024 * 1024
    off = ptr - buf.data_ptr()
    region = buf[off:off + nbytes]
    if poison:
        region.view(torch.int16).fill_(-64)  # 0xFFC0: a bf16 NaN; read as fp32, 0xFFC0FFC0 is a NaN too
    else:
        region.zero_()
    return buf, ptr, off


def _bands_intact(buf, off, nbytes):
    return bool((buf[:off] == SENTINEL).all()) and bool((buf[off + nbytes:] == SENTINEL).all())


def test_backbone_workspace_guard_bands_and_poison(cuda_device):
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from oracle import i3res50 as O

    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0), strict=True)
    m = m.eval().to(cuda_device)
    x = torch.randn(3, 3, 16, 112, 112, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444).to(cuda_device)
    ref = m(x).clone()                       # configures the plan, allocates its own workspace
    plan = m.plan(torch.device(cuda_device))
    need = plan._ws_bytes
    outs = []
    for poison in (False, True):
        buf, ptr, off = _guarded(need, cuda_device, poison)
        plan._ws, plan._ws_ptr = buf, ptr    # the next forward binds the plan to this workspace
        outs.append(m(x).clone())
        torch.cuda.synchronize()
        assert _bands_intact(buf, off, need), "a kernel wrote outside the workspace it was given"
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref), "features depend on what the workspace held before the forward"


def test_head_eval_and_training_workspaces(cuda_device):
    from anomaly_detection_on_video_b200 import _lib
    from anomaly_detection_on_video_b200.mgfn import MGFNConfig, MGFNForVideoAnomalyDetection
    from oracle import mgfn as M

    lib = _lib.load()
    video = M.synthetic_video(3, 4, 10, 32).to(cuda_device)
    # ---- eval
    m = MGFNForVideoAnomalyDetection(MGFNConfig(dropout_rate=0.0))
    m.load_state_dict(M.seeded_state_dict(0), strict=True)
    m = m.eval().to(cuda_device)
    ref = m(video).scores.clone()
    need = ctypes.c_uint64()
    _lib.check(lib.vad_head_workspace_bytes(m._handle, 40, 32, ctypes.byref(need)))
    for poison in (False, True):
        buf, ptr, off = _guarded(int(need.value), cuda_device, poison)
        m._ws = buf[off - 1023:]  # forward() aligns (data_ptr + 1023) // 1024 * 1024: exactly our region
        got = m(video).scores.clone()
        torch.cuda.synchronize()
        assert _bands_intact(buf, off, int(need.value))
        assert torch.equal(got, ref)
    # ---- training step
    m.train()
    labels = dict(abnormal_labels=torch.ones(2, device=cuda_device), normal_labels=torch.zeros(2, device=cuda_device))
    out = m(video, **labels)
    st = m._train
    ref_loss, ref_grad = out.loss_terms.clone(), st["grad"].clone()
    _lib.check(lib.vad_head_train_workspace_bytes(st["handle"], 4, 10, 32, ctypes.byref(need)))
    m.load_state_dict(M.seeded_state_dict(0), strict=True)   # the running statistics moved; start the repeat from the same state
    for poison in (False, True):
        buf, ptr, off = _guarded(int(need.value), cuda_device, poison)
        st["ws"] = buf[off - 1023:]
        m.load_state_dict(M.seeded_state_dict(0), strict=True)
        out = m(video, **labels)
        torch.cuda.synchronize()
        assert _bands_intact(buf, off, int(need.value))
        # the BatchNorm batch sums and the split-K weight gradients are reduced with fp32 atomics: equal up to summation order
        # (1e-6 relative), and never NaN from the poison
        torch.testing.assert_close(out.loss_terms, ref_loss, rtol=1e-5, atol=1e-7)
        assert torch.isfinite(st["grad"]).all()
        torch.testing.assert_close(st["grad"], ref_grad, rtol=1e-4, atol=1e-6)
