"""Thin Python objects over the C ABI: device memory and streams come from PyTorch, every kernel
comes from ``libvad_b200.so``.  Nothing in here computes on the CPU or through ``torch`` ops.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import OpDesc, check

KBLOCK = 64  # contraction tile of the conv kernel; packed weight rows are padded to it


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device: this path has no CPU fallback")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# --------------------------------------------------------------------------------------- op table
@dataclass
class Op:
    """One row of the backbone op table (mirrors ``vad_op_desc``)."""

    kind: int
    src: int
    dst: int = 0
    res: int = -1
    cin: int = 0
    cout: int = 0
    kernel: Tuple[int, int, int] = (1, 1, 1)
    stride: Tuple[int, int, int] = (1, 1, 1)
    pad: Tuple[int, int, int] = (0, 0, 0)
    flags: int = 0
    dst_c_off: int = 0
    dst_c_total: int = 0
    w_off: int = 0
    scale_off: int = 0
    shift_off: int = 0
    name: str = ""
    # sibling 1x1x1 convs fused into one launch (vad_op_desc.dst1 ...): columns [split1, split2) -> slot dst1, [split2, cout) -> dst2
    dst1: int = 0
    dst2: int = 0
    split1: int = 0
    split2: int = 0
    seg_w: Tuple[int, int, int] = (0, 0, 0)

    def to_c(self) -> OpDesc:
        d = OpDesc()
        d.kind, d.src, d.dst, d.res = self.kind, self.src, self.dst, self.res
        d.cin, d.cout = self.cin, self.cout
        d.kt, d.kh, d.kw = self.kernel
        d.st, d.sh, d.sw = self.stride
        d.pt, d.ph, d.pw = self.pad
        d.flags = self.flags
        d.dst_c_off, d.dst_c_total = self.dst_c_off, self.dst_c_total
        d.w_off, d.scale_off, d.shift_off = self.w_off, self.scale_off, self.shift_off
        d.dst1, d.dst2, d.split1, d.split2 = self.dst1, self.dst2, self.split1, self.split2
        d.seg_w0, d.seg_w1, d.seg_w2 = self.seg_w
        return d


class ParamPacker:
    """Packs conv weights (bf16, K-major, K = (kt, kh, kw, cin) padded to 64) and the folded
    BatchNorm scale / shift (fp32) into one blob; returns the byte offsets the op table needs."""

    def __init__(self) -> None:
        self._chunks: List[torch.Tensor] = []
        self._size = 0

    def _append(self, t: torch.Tensor, align: int) -> int:
        pad = (-self._size) % align
        if pad:
            self._chunks.append(torch.zeros(pad, dtype=torch.uint8))
            self._size += pad
        off = self._size
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        self._chunks.append(raw)
        self._size += raw.numel()
        return off

    def add_conv(self, weight: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, fold_w: bool = False,
                 cin_pad: Optional[int] = None, tf32: bool = False, planes: bool = False) -> Tuple[int, int, int]:
        """weight: [cout, cin, kt, kh, kw] fp32 (torch Conv3d layout).  ``tf32``: keep the weights in fp32 with K padded
        to 32 (the layout of the TF32 precision mode, ``Tf32Plan``)."""
        w = weight.detach().to(torch.float32).cpu()
        cout, cin, kt, kh, kw = w.shape
        w = w.permute(0, 2, 3, 4, 1)  # [cout, kt, kh, kw, cin]
        if fold_w:
            # stem: one (kt, kh) tap contracts a window of 8 pixels x 4 channels; pixels >= kw and
            # channel 3 carry zero weights
            if cin > 4 or kw > 8:
                raise ValueError("fold_w packing needs cin <= 4 and kw <= 8")
            wf = torch.zeros(cout, kt, kh, 8, 4, dtype=torch.float32)
            wf[:, :, :, :kw, :cin] = w
            if planes:
                # VAD_FLAG_STEM_PLANES: window pixel j = 2 i + p sits at position i of column-parity plane p, so one
                # (dt, dh) tap is two 16-float virtual taps [plane][4 px][4 ch]
                wf = wf.reshape(cout, kt, kh, 4, 2, 4).permute(0, 1, 2, 4, 3, 5).contiguous()
            w2 = wf.reshape(cout, kt * kh * 32)
        else:
            cp = cin_pad or cin
            if cp != cin:
                wp = torch.zeros(cout, kt, kh, kw, cp, dtype=torch.float32)
                wp[..., :cin] = w
                w = wp
            w2 = w.reshape(cout, -1)
        k = w2.shape[1]
        kblock = 32 if tf32 else KBLOCK
        k_pad = (k + kblock - 1) // kblock * kblock
        if k_pad != k:
            w2 = torch.cat([w2, torch.zeros(cout, k_pad - k)], dim=1)
        if tf32:
            # round to the nearest TF32 value (ties away from zero, = cvt.rna.tf32.f32): the tensor core truncates
            w2 = ((w2.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        w_off = self._append(w2 if tf32 else w2.to(torch.bfloat16), 128)
        s_off = self._append(scale.detach().to(torch.float32).cpu(), 16)
        b_off = self._append(shift.detach().to(torch.float32).cpu(), 16)
        return w_off, s_off, b_off

    def blob(self) -> torch.Tensor:
        if not self._chunks:
            return torch.zeros(0, dtype=torch.uint8)
        return torch.cat(self._chunks)


def fold_bn(gamma: torch.Tensor, beta: torch.Tensor, mean: torch.Tensor, var: torch.Tensor, eps: float):
    """BatchNorm (eval) as a per-channel affine: y = x * scale + shift, computed in fp32."""
    scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
    shift = beta.detach().float() - mean.detach().float() * scale
    return scale, shift


# --------------------------------------------------------------------------------------- plan
class BackbonePlan:
    """Owns a ``vad_plan_t``; PyTorch owns the parameter blob and the workspace."""

    def __init__(self, ops: Sequence[Op], params: torch.Tensor, n_slots: int, in_pad_left: int,
                 device: torch.device, in_channels: int = 0) -> None:
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BackbonePlan needs a CUDA device: this path has no CPU fallback")
        self.ops = list(ops)
        self.n_slots = n_slots
        self.in_pad_left = in_pad_left
        self.in_channels = in_channels
        self.params = params.to(self.device) if not params.is_cuda else params
        self._c_ops = (OpDesc * len(self.ops))(*[o.to_c() for o in self.ops])
        self._h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self.lib.vad_plan_create(ctypes.byref(self._h), self._c_ops, len(self.ops), n_slots,
                                       ctypes.c_void_p(self.params.data_ptr()), self.params.numel(), in_channels,
                                       in_pad_left, dev_index), "vad_plan_create")
        self._cfg: Optional[Tuple[int, int, int, int]] = None
        self._feat_buf: Optional[torch.Tensor] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_ptr = 0
        self._ws_bytes = 0

    def __del__(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.vad_plan_destroy(h)
            self._h = ctypes.c_void_p()

    def configure(self, batch: int, t: int, h: int, w: int) -> None:
        if self._cfg == (batch, t, h, w):
            return
        need = ctypes.c_uint64()
        check(self.lib.vad_plan_configure(self._h, batch, t, h, w, ctypes.byref(need)), "vad_plan_configure")
        self._cfg = (batch, t, h, w)
        self._ws_bytes = int(need.value)
        if self._ws is None or self._ws.numel() < self._ws_bytes + 1024:
            self._ws = None
            self._ws = torch.empty(self._ws_bytes + 1024, dtype=torch.uint8, device=self.device)
        base = self._ws.data_ptr()
        self._ws_ptr = (base + 1023) // 1024 * 1024

    @property
    def flops(self) -> float:
        return float(self.lib.vad_plan_flops(self._h))

    @property
    def num_launches(self) -> int:
        return int(self.lib.vad_plan_num_launches(self._h))

    def profile_begin(self, first_op: int = 0, n_ops: int = -1) -> None:
        """Bracket every launch (default) or only ops [first_op, first_op + n_ops) with CUDA events."""
        check(self.lib.vad_plan_profile_select(self._h, first_op, n_ops), "vad_plan_profile_select")
        check(self.lib.vad_plan_profile_begin(self._h), "vad_plan_profile_begin")

    def profile_end(self) -> List[Dict[str, object]]:
        """Per-op totals since ``profile_begin``: name, kind, ms, launches, useful FLOPs, algorithmic bytes."""
        n = len(self.ops)
        ms = (ctypes.c_double * n)()
        calls = (ctypes.c_int32 * n)()
        flops = (ctypes.c_double * n)()
        nbytes = (ctypes.c_double * n)()
        check(self.lib.vad_plan_profile_end(self._h, n, ms, calls, flops, nbytes), "vad_plan_profile_end")
        return [{"name": op.name, "kind": op.kind, "ms": float(ms[i]), "calls": int(calls[i]), "flops": float(flops[i]),
                 "bytes": float(nbytes[i])} for i, op in enumerate(self.ops)]

    def feature_dim(self) -> int:
        for op in reversed(self.ops):
            if op.kind == _lib.VAD_OP_AVGPOOL:
                return self.slot_shape(op.src)[3]
        raise RuntimeError("plan has no AVGPOOL op")

    def slot_shape(self, slot: int) -> Tuple[int, int, int, int]:
        dims = (ctypes.c_int32 * 4)()
        check(self.lib.vad_plan_slot_info(self._h, slot, dims, None, None), "vad_plan_slot_info")
        return tuple(int(v) for v in dims)

    def slot_tensor(self, slot: int) -> torch.Tensor:
        """bf16 view [batch, T, H, W, C] of a workspace slot (valid after ``forward``)."""
        dims = (ctypes.c_int32 * 4)()
        off = ctypes.c_uint64()
        nbytes = ctypes.c_uint64()
        check(self.lib.vad_plan_slot_info(self._h, slot, dims, ctypes.byref(off), ctypes.byref(nbytes)), "vad_plan_slot_info")
        batch = self._cfg[0]
        t, h, w, c = (int(v) for v in dims)
        start = self._ws_ptr - self._ws.data_ptr() + int(off.value)
        n = batch * t * h * w * c
        return self._ws[start:start + 2 * n].view(torch.bfloat16).view(batch, t, h, w, c)

    def forward(self, x_stem: torch.Tensor, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """x_stem: [batch, T, H, W + 8, 4] bf16 (stem layout; or plain [batch, T, H, W, in_channels] for a
        plan created with ``in_channels > 0``).  Returns fp32 [batch, C] features when the op table ends
        in AVGPOOL, else None (read results with ``slot_tensor``)."""
        _require_cuda(x_stem, "x_stem")
        c_in = self.in_channels or 4
        if x_stem.dtype != torch.bfloat16 or x_stem.dim() != 5 or x_stem.shape[-1] != c_in or not x_stem.is_contiguous():
            raise ValueError(f"input must be a contiguous bf16 [batch, T, H, W(+8), {c_in}] tensor")
        b, t, h, wp, _ = x_stem.shape
        self.configure(b, t, h, wp - 8 if self.in_channels == 0 else wp)
        has_feat = any(op.kind == _lib.VAD_OP_AVGPOOL for op in self.ops)
        clone = False
        if has_feat and out is None:
            # a plan-owned result buffer: a small-batch forward is replayed as a CUDA graph (vad_plan_forward), which bakes
            # the feature pointer in; the caller gets a copy
            if self._feat_buf is None or self._feat_buf.shape[0] != b:
                self._feat_buf = torch.empty(b, self.feature_dim(), dtype=torch.float32, device=self.device)
            out, clone = self._feat_buf, True
        check(self.lib.vad_plan_forward(self._h, ctypes.c_void_p(x_stem.data_ptr()), ctypes.c_void_p(self._ws_ptr),
                                        self._ws_bytes, ctypes.c_void_p(out.data_ptr()) if has_feat else None,
                                        ctypes.c_void_p(_stream_ptr(self.device))), "vad_plan_forward")
        if not has_feat:
            return None
        return out.clone() if clone else out


class Tf32Plan:
    """Owns a ``vad_tf32_plan_t``: the same op table run with fp32 activations / weights and tcgen05 kind::tf32 MMAs
    (features within 1e-3 of the reference's fp32 path; ``BackbonePlan`` is the bf16 production mode)."""

    def __init__(self, ops: Sequence[Op], params: torch.Tensor, n_slots: int, device: torch.device, in_channels: int = 4) -> None:
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Tf32Plan needs a CUDA device: this path has no CPU fallback")
        self.ops = list(ops)
        self.n_slots = n_slots
        self.in_channels = in_channels
        self.params = params.to(self.device) if not params.is_cuda else params
        self._c_ops = (OpDesc * len(self.ops))(*[o.to_c() for o in self.ops])
        self._h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self.lib.vad_tf32_plan_create(ctypes.byref(self._h), self._c_ops, len(self.ops), n_slots,
                                            ctypes.c_void_p(self.params.data_ptr()), self.params.numel(), in_channels, dev_index),
              "vad_tf32_plan_create")
        self._cfg: Optional[Tuple[int, int, int, int]] = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_ptr = 0
        self._ws_bytes = 0

    def __del__(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.vad_tf32_plan_destroy(h)
            self._h = ctypes.c_void_p()

    def configure(self, batch: int, t: int, h: int, w: int) -> None:
        if self._cfg == (batch, t, h, w):
            return
        need = ctypes.c_uint64()
        check(self.lib.vad_tf32_plan_configure(self._h, batch, t, h, w, ctypes.byref(need)), "vad_tf32_plan_configure")
        self._cfg = (batch, t, h, w)
        self._ws_bytes = int(need.value)
        if self._ws is None or self._ws.numel() < self._ws_bytes + 1024:
            self._ws = None
            self._ws = torch.empty(self._ws_bytes + 1024, dtype=torch.uint8, device=self.device)
        self._ws_ptr = (self._ws.data_ptr() + 1023) // 1024 * 1024

    @property
    def flops(self) -> float:
        return float(self.lib.vad_tf32_plan_flops(self._h))

    @property
    def num_launches(self) -> int:
        return int(self.lib.vad_tf32_plan_num_launches(self._h))

    def feature_dim(self) -> int:
        for op in reversed(self.ops):
            if op.kind == _lib.VAD_OP_AVGPOOL:
                return self.slot_shape(op.src)[3]
        raise RuntimeError("plan has no AVGPOOL op")

    def slot_shape(self, slot: int) -> Tuple[int, int, int, int]:
        dims = (ctypes.c_int32 * 4)()
        check(self.lib.vad_tf32_plan_slot_info(self._h, slot, dims, None, None), "vad_tf32_plan_slot_info")
        return tuple(int(v) for v in dims)

    def slot_tensor(self, slot: int) -> torch.Tensor:
        """fp32 view [batch, T, H, W, C] of a workspace slot (valid after ``forward``)."""
        dims = (ctypes.c_int32 * 4)()
        off = ctypes.c_uint64()
        check(self.lib.vad_tf32_plan_slot_info(self._h, slot, dims, ctypes.byref(off), None), "vad_tf32_plan_slot_info")
        batch = self._cfg[0]
        t, h, w, c = (int(v) for v in dims)
        start = self._ws_ptr - self._ws.data_ptr() + int(off.value)
        n = batch * t * h * w * c
        return self._ws[start:start + 4 * n].view(torch.float32).view(batch, t, h, w, c)

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """x: contiguous fp32 [batch, T, H, W, in_channels] on the device; [batch, T, H, 2, (W + 8) / 2, 4] (``ingest_ncthw_tf32(...,
        planes=True)``) when the stem op carries VAD_FLAG_STEM_PLANES."""
        _require_cuda(x, "x")
        planes = bool(self.ops and self.ops[0].flags & _lib.VAD_FLAG_STEM_PLANES)
        if planes:
            if x.dtype != torch.float32 or x.dim() != 6 or x.shape[3] != 2 or x.shape[-1] != 4 or not x.is_contiguous():
                raise ValueError("input must be a contiguous fp32 [batch, T, H, 2, (W + 8) / 2, 4] plane-layout tensor")
            b, t, h, _, wh, _ = x.shape
            w = 2 * wh - 8
        else:
            if x.dtype != torch.float32 or x.dim() != 5 or x.shape[-1] != self.in_channels or not x.is_contiguous():
                raise ValueError(f"input must be a contiguous fp32 [batch, T, H, W, {self.in_channels}] tensor")
            b, t, h, w, _ = x.shape
        self.configure(b, t, h, w)
        has_feat = any(op.kind == _lib.VAD_OP_AVGPOOL for op in self.ops)
        if has_feat and out is None:
            out = torch.empty(b, self.feature_dim(), dtype=torch.float32, device=self.device)
        check(self.lib.vad_tf32_plan_forward(self._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(self._ws_ptr), self._ws_bytes,
                                             ctypes.c_void_p(out.data_ptr()) if has_feat else None,
                                             ctypes.c_void_p(_stream_ptr(self.device))), "vad_tf32_plan_forward")
        return out if has_feat else None


def ingest_ncthw_tf32(x: torch.Tensor, planes: bool = False) -> torch.Tensor:
    """fp32 [B, 3, T, H, W] clips -> fp32 [B, T, H, W, 4] (zero fourth channel): the input of a ``Tf32Plan``.
    ``planes``: the column-parity plane layout [B, T, H, 2, (W + 8) / 2, 4] of a plan whose stem has VAD_FLAG_STEM_PLANES."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 5 or x.shape[1] != 3:
        raise ValueError("expected a float32 [B, 3, T, H, W] tensor")
    x = x.contiguous()
    b, _, t, h, w = x.shape
    lib = _lib.load()
    if planes:
        if w % 2:
            raise ValueError("the plane layout needs an even frame width")
        out = torch.empty(b, t, h, 2, (w + 8) // 2, 4, dtype=torch.float32, device=x.device)
        check(lib.vad_tf32_ingest_ncthw_planes(ctypes.c_void_p(x.data_ptr()), b, t, h, w, ctypes.c_void_p(out.data_ptr()),
                                               ctypes.c_void_p(_stream_ptr(x.device))), "vad_tf32_ingest_ncthw_planes")
        return out
    out = torch.empty(b, t, h, w, 4, dtype=torch.float32, device=x.device)
    check(lib.vad_tf32_ingest_ncthw(ctypes.c_void_p(x.data_ptr()), b, t, h, w, ctypes.c_void_p(out.data_ptr()),
                                    ctypes.c_void_p(_stream_ptr(x.device))), "vad_tf32_ingest_ncthw")
    return out


def ingest_ncthw(x: torch.Tensor, pad_left: int) -> torch.Tensor:
    """fp32 [B, 3, T, H, W] clips (the tensor the reference hands its model) -> stem layout."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 5 or x.shape[1] != 3:
        raise ValueError("expected a float32 [B, 3, T, H, W] tensor")
    x = x.contiguous()
    b, _, t, h, w = x.shape
    out = torch.empty(b, t, h, w + 8, 4, dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    check(lib.vad_ingest_ncthw_f32(ctypes.c_void_p(x.data_ptr()), b, t, h, w, pad_left, ctypes.c_void_p(out.data_ptr()),
                                   ctypes.c_void_p(_stream_ptr(x.device))), "vad_ingest_ncthw_f32")
    return out


# --------------------------------------------------------------------------------------- preprocessing
class Preprocessor:
    """Fused resize / ten-crop / standardise / loop-pad for frames of one size (``vad_preproc_t``)."""

    def __init__(self, src_h: int, src_w: int, resize: int = 256, crop: int = 224, ncrops: int = 10,
                 device: Optional[torch.device] = None) -> None:
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("Preprocessor needs a CUDA device: this path has no CPU fallback")
        self.src_h, self.src_w, self.resize, self.crop, self.ncrops = src_h, src_w, resize, crop, ncrops
        self._h = ctypes.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self.lib.vad_preproc_create(ctypes.byref(self._h), src_h, src_w, resize, crop, ncrops, dev_index),
              "vad_preproc_create")

    def __del__(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self.lib.vad_preproc_destroy(h)
            self._h = ctypes.c_void_p()

    def info(self) -> Dict[str, object]:
        hw = (ctypes.c_int32 * 2)()
        tops = (ctypes.c_int32 * 10)()
        lefts = (ctypes.c_int32 * 10)()
        flips = (ctypes.c_int32 * 10)()
        check(self.lib.vad_preproc_info(self._h, hw, tops, lefts, flips), "vad_preproc_info")
        n = self.ncrops
        return {"resized_hw": (int(hw[0]), int(hw[1])), "tops": list(tops)[:n], "lefts": list(lefts)[:n],
                "flips": list(flips)[:n]}

    def run(self, frames: torch.Tensor, clip_start: int, n_clips: int, frames_per_clip: int = 16,
            out_mode: int = _lib.VAD_OUT_STEM_BF16, pad_left: int = 3, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8 [n_frames, H, W, 3] on the device."""
        _require_cuda(frames, "frames")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_contiguous():
            raise ValueError("frames must be a contiguous uint8 [n_frames, H, W, 3] tensor")
        if frames.shape[1] != self.src_h or frames.shape[2] != self.src_w:
            raise ValueError("frame size does not match this Preprocessor")
        c, k, f = self.crop, self.ncrops, frames_per_clip
        if out_mode == _lib.VAD_OUT_STEM_BF16:
            shape, dtype = (n_clips * k, f, c, c + 8, 4), torch.bfloat16
        else:
            shape, dtype = (n_clips, k, f, 3, c, c), torch.float32
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=frames.device)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {dtype} tensor of shape {shape}")
        # 65535 frame slots per launch (gridDim.y)
        max_clips = max(1, 65535 // f)
        done = 0
        while done < n_clips:
            n = min(max_clips, n_clips - done)
            check(self.lib.vad_preproc_run(self._h, ctypes.c_void_p(frames.data_ptr()), frames.shape[0], clip_start + done,
                                           n, f, out_mode, pad_left, ctypes.c_void_p(out[done * (k if out_mode else 1):].data_ptr()),
                                           ctypes.c_void_p(_stream_ptr(frames.device))), "vad_preproc_run")
            done += n
        return out


# --------------------------------------------------------------------------------------- reductions
def segment_mean(feats: torch.Tensor, seg_length: int = 32) -> torch.Tensor:
    """[n_clips, ncrops, C] fp32 -> [ncrops, seg_length, C] (extract_features.py:159-185 semantics)."""
    _require_cuda(feats, "feats")
    if feats.dtype != torch.float32 or feats.dim() != 3:
        raise ValueError("expected a float32 [n_clips, ncrops, C] tensor")
    feats = feats.contiguous()
    n, k, c = feats.shape
    out = torch.empty(k, seg_length, c, dtype=torch.float32, device=feats.device)
    lib = _lib.load()
    check(lib.vad_segment_mean(ctypes.c_void_p(feats.data_ptr()), n, k, c, seg_length, ctypes.c_void_p(out.data_ptr()),
                               ctypes.c_void_p(_stream_ptr(feats.device))), "vad_segment_mean")
    return out


def add_magnitude(feats: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> [..., C + 1] with the L2 norm appended (src/dataset.py:121-124)."""
    _require_cuda(feats, "feats")
    if feats.dtype != torch.float32:
        raise ValueError("expected float32 features")
    feats = feats.contiguous()
    c = feats.shape[-1]
    rows = feats.numel() // c
    out = torch.empty(*feats.shape[:-1], c + 1, dtype=torch.float32, device=feats.device)
    lib = _lib.load()
    check(lib.vad_add_magnitude(ctypes.c_void_p(feats.data_ptr()), rows, c, ctypes.c_void_p(out.data_ptr()),
                                ctypes.c_void_p(_stream_ptr(feats.device))), "vad_add_magnitude")
    return out


__all__ = ["Op", "ParamPacker", "fold_bn", "BackbonePlan", "Preprocessor", "ingest_ncthw", "segment_mean",
           "add_magnitude"]
_ = field  # dataclasses.field kept for subclasses
