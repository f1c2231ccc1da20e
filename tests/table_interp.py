"""CPU interpreter of a backbone op table (``engine.Op`` rows + the ``ParamPacker`` blob), in plain fp32 torch.

Test infrastructure only: it gives the op-table builders (`i3d.I3Res50`, `inception.InceptionI3d`) a semantic check
that needs no GPU -- every slot routing, channel slice, zero-widened temporary, fused-sibling split and folded BatchNorm
is executed exactly as the table states it and the result is compared with the oracle's forward.  What it does NOT
model is the kernels' arithmetic (bf16 activations, MMA summation order): activations stay fp32 here, only the weights
carry their packed bf16 rounding.

Semantics follow include/vad_b200.h (`vad_op_desc`):
  CONV     y = conv(src) * scale + shift (+ res) (ReLU); padding = `pad`, or TF-SAME when VAD_FLAG_CONV_SAME; output columns
           [0, split1) -> channel slice [dst_c_off, ..) of `dst`, [split1, split2) -> `dst1`, [split2, cout) -> `dst2` (only
           seg_w* columns of each part exist); VAD_FLAG_POOL_T2 = max over output frame pairs after the activation;
           VAD_FLAG_STEM_FOLD_W = weights packed as (kt, kh) taps over 8-pixel x 4-channel windows.
  MAXPOOL  `pad` with -inf, or TF-SAME with zeros when VAD_FLAG_POOL_SAME.
  AVGPOOL  mean over (T, H, W) (a temporal window when kernel[0] > 1: sliding (kt, H, W) average, then the mean of those).
"""
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from anomaly_detection_on_video_b200 import _lib
from anomaly_detection_on_video_b200.engine import KBLOCK, Op


def _same_pads(size: Tuple[int, int, int], k, s) -> List[int]:
    pads = []
    for dim in (2, 1, 0):  # F.pad wants the last dimension first
        tot = max(k[dim] - s[dim], 0) if size[dim] % s[dim] == 0 else max(k[dim] - (size[dim] % s[dim]), 0)
        pads += [tot // 2, tot - tot // 2]
    return pads


def _weights(blob: torch.Tensor, op: Op, tf32: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> ([cout, cin, kt, kh, kw] fp32, scale [cout], shift [cout]) decoded from the packed blob (bf16 rows padded to 64, or --
    TF32 precision mode -- fp32 rows padded to 32)."""
    kt, kh, kw = op.kernel
    fold = bool(op.flags & _lib.VAD_FLAG_STEM_FOLD_W)
    k = kt * kh * 32 if fold else kt * kh * kw * op.cin
    kblock = 32 if tf32 else KBLOCK
    k_pad = (k + kblock - 1) // kblock * kblock
    if tf32:
        raw = blob[op.w_off:op.w_off + op.cout * k_pad * 4].view(torch.float32).reshape(op.cout, k_pad)
        assert not (raw.view(torch.int32) & 0x1FFF).any(), f"{op.name}: fp32 weights are not TF32 values"
    else:
        raw = blob[op.w_off:op.w_off + op.cout * k_pad * 2].view(torch.bfloat16).float().reshape(op.cout, k_pad)
    assert not raw[:, k:].any(), f"{op.name}: K padding of the packed weights is not zero"
    if fold:
        w = raw[:, :k]
        if op.flags & _lib.VAD_FLAG_STEM_PLANES:   # [plane][4 positions][4 channels] per (dt, dh) tap: pixel j = 2 i + plane
            w = w.reshape(op.cout, kt, kh, 2, 4, 4).permute(0, 1, 2, 4, 3, 5)
        w = w.reshape(op.cout, kt, kh, 8, 4)
        assert not w[:, :, :, kw:].any() and not w[..., 3].any(), f"{op.name}: folded-window padding taps are not zero"
        w = w[:, :, :, :kw, :3].permute(0, 4, 1, 2, 3)
    else:
        w = raw[:, :k].reshape(op.cout, kt, kh, kw, op.cin).permute(0, 4, 1, 2, 3)
    scale = blob[op.scale_off:op.scale_off + 4 * op.cout].view(torch.float32)
    shift = blob[op.shift_off:op.shift_off + 4 * op.cout].view(torch.float32)
    return w.contiguous(), scale.clone(), shift.clone()


@torch.no_grad()
def run_table(ops: List[Op], blob: torch.Tensor, x: torch.Tensor, tf32: bool = False) -> torch.Tensor:
    """x: [B, 3, T, H, W] fp32 clips -> [B, C] features, executing `ops` row by row on NCDHW fp32 slot tensors.
    ``tf32``: the table and blob of the TF32 precision mode (fp32 weights; an RGB conv that is not folded reads 4 channels)."""
    slots: Dict[int, torch.Tensor] = {0: x}
    concat_owner: Dict[int, str] = {}
    out = None
    for op in ops:
        src = slots[op.src]
        if op.kind == _lib.VAD_OP_CONV:
            w, scale, shift = _weights(blob, op, tf32)
            fold = bool(op.flags & _lib.VAD_FLAG_STEM_FOLD_W)
            cin_have = src.shape[1]
            if fold:
                assert op.cin == 4 and cin_have == 3, op.name
            elif tf32 and op.src == 0 and op.cin == 4 and cin_have == 3:
                assert not w[:, 3].any(), f"{op.name}: weights of the zero input channel are not zero"
                w = w[:, :3].contiguous()
            else:
                assert op.cin == cin_have, f"{op.name}: cin {op.cin} but slot {op.src} holds {cin_have} channels"
            if op.flags & _lib.VAD_FLAG_CONV_SAME:
                xin = F.pad(src, _same_pads(tuple(src.shape[2:]), op.kernel, op.stride))
                y = F.conv3d(xin, w, None, op.stride)
            else:
                y = F.conv3d(src, w, None, op.stride, op.pad)
            y = y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
            if op.res >= 0:
                assert not op.dst1, op.name
                y = y + slots[op.res]
            if op.flags & _lib.VAD_FLAG_RELU:
                y = F.relu(y)
            if op.flags & _lib.VAD_FLAG_POOL_T2:
                assert y.shape[2] % 2 == 0, op.name
                y = torch.maximum(y[:, :, 0::2], y[:, :, 1::2])
            if op.dst1:
                w0, w1, w2 = op.seg_w
                end2 = op.split2 if op.dst2 else op.cout
                assert w0 <= op.split1 and w1 <= end2 - op.split1 and (not op.dst2 or w2 == op.cout - op.split2), op.name
                # columns between a part's width and the next split are layout padding: the kernel never stores them
                parts = [(op.dst, y[:, :w0], op.dst_c_off, op.dst_c_total), (op.dst1, y[:, op.split1:op.split1 + w1], 0, 0)]
                if op.dst2:
                    parts.append((op.dst2, y[:, op.split2:op.split2 + w2], 0, 0))
            else:
                parts = [(op.dst, y, op.dst_c_off, op.dst_c_total)]
            block = op.name.split(".")[0]   # the branches of one Inception block share a concat destination
            for dst, val, off, total in parts:
                if total:
                    if concat_owner.get(dst) != block:
                        # a concat destination: NaN until every slice is written, so a gap in the slices poisons the features
                        slots[dst] = torch.full((val.shape[0], total) + tuple(val.shape[2:]), float("nan"))
                        concat_owner[dst] = block
                    slots[dst][:, off:off + val.shape[1]] = val
                else:
                    slots[dst] = val
                    concat_owner.pop(dst, None)
        elif op.kind == _lib.VAD_OP_MAXPOOL:
            if op.flags & _lib.VAD_FLAG_POOL_SAME:
                slots[op.dst] = F.max_pool3d(F.pad(src, _same_pads(tuple(src.shape[2:]), op.kernel, op.stride)), op.kernel, op.stride)
            else:
                slots[op.dst] = F.max_pool3d(src, op.kernel, op.stride, op.pad)
            concat_owner.pop(op.dst, None)
        elif op.kind == _lib.VAD_OP_AVGPOOL:
            if op.kernel[0] > 1:
                src = F.avg_pool3d(src, (min(op.kernel[0], src.shape[2]),) + tuple(src.shape[3:]), 1)
            out = src.mean(dim=(2, 3, 4))
        else:
            raise AssertionError(f"unknown op kind {op.kind}")
    assert out is not None, "the table does not end in AVGPOOL"
    return out
