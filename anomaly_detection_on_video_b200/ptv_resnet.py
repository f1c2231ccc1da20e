"""``i3d_8x8_r50``: the backbone the reference's CLI runs by default (extract_features.py:34,46) -- pytorchvideo's
``create_resnet`` with the arguments at src/i3d.py:339-350 under the reference's ``create_res_pooler`` head (src/i3d.py:21-57)
-- as a third op table over the native kernels.

pytorchvideo is third-party, un-vendored and unpinned in the reference, and absent from this image, so there is nothing to
execute as an oracle: **parity is unpinned** (SURVEY 8(c)); the table follows the published architecture restated in
``oracle/i3d_r50_ptv.py`` and the module tree keeps pytorchvideo's parameter names (``blocks.0.conv``, ``blocks.N.res_blocks.B.
branch1_conv / branch2.conv_a|b|c / norm_*``), so an ``I3D_8x8_R50.pyth`` checkpoint loads with ``load_state_dict(strict=False)``
exactly as the reference loads it (the classification projection of the hub checkpoint is dropped there too).

Differences from ``I3Res50`` that matter to the kernels: no temporal stride in the stem (16 frames stay 16; the one-frame-per-tile
stem kernel runs it), a padded spatial max-pool, a (2,1,1) max-pool after res2, 8 frames through res3-5, and a head that is
AvgPool3d((4,7,7), stride 1) followed by a global mean -- frame t weighted by the number of windows covering it (K4 with a
temporal window).  113.6 GFLOP per 16 x 224 x 224 clip (3.5x I3Res50).
"""
from __future__ import annotations

from typing import List, Tuple

from torch import nn

from . import _lib
from .engine import Op, ParamPacker
from .i3d import STEM_PAD_LEFT, _NativeBackbone

DEPTHS = (3, 4, 6, 3)
CONV_A_KT: Tuple[Tuple[int, ...], ...] = ((3,), (3, 1), (3, 1), (1, 3))   # src/i3d.py:342-347, cycled over the blocks of a stage
STAGE_BLOCK_INDEX = (1, 3, 4, 5)                                           # res2..res5 inside ``blocks`` (2: stage1 pool, 6: head)


class _Branch2(nn.Module):
    def __init__(self, dim_in: int, inner: int, out: int, kt: int, stride: int) -> None:
        super().__init__()
        self.conv_a = nn.Conv3d(dim_in, inner, (kt, 1, 1), stride=1, padding=(kt // 2, 0, 0), bias=False)
        self.norm_a = nn.BatchNorm3d(inner)
        self.conv_b = nn.Conv3d(inner, inner, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
        self.norm_b = nn.BatchNorm3d(inner)
        self.conv_c = nn.Conv3d(inner, out, 1, bias=False)
        self.norm_c = nn.BatchNorm3d(out)


class _ResBlock(nn.Module):
    def __init__(self, dim_in: int, inner: int, out: int, kt: int, stride: int, shortcut: bool) -> None:
        super().__init__()
        if shortcut:
            self.branch1_conv = nn.Conv3d(dim_in, out, 1, stride=(1, stride, stride), bias=False)
            self.branch1_norm = nn.BatchNorm3d(out)
        self.branch2 = _Branch2(dim_in, inner, out, kt, stride)


class _ResStage(nn.Module):
    def __init__(self, blocks: List[_ResBlock]) -> None:
        super().__init__()
        self.res_blocks = nn.ModuleList(blocks)


class _Stem(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.conv = nn.Conv3d(3, 64, (5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False)
        self.norm = nn.BatchNorm3d(64)


class I3D8x8R50(_NativeBackbone):
    """pytorchvideo I3D-R50 feature extractor (reference src/i3d.py:339-350), 2048-d output.  Parity unpinned."""

    feature_dim = 2048
    pad_left = STEM_PAD_LEFT

    def __init__(self) -> None:
        super().__init__()
        blocks: List[nn.Module] = [_Stem()]
        dim_in = 64
        for s, depth in enumerate(DEPTHS):
            inner, out = 64 * 2 ** s, 256 * 2 ** s
            stage = []
            for b in range(depth):
                kt = CONV_A_KT[s][b % len(CONV_A_KT[s])]
                stage.append(_ResBlock(dim_in if b == 0 else out, inner, out, kt, 2 if (b == 0 and s > 0) else 1, shortcut=b == 0))
            blocks.append(_ResStage(stage))
            if s == 0:
                blocks.append(nn.Identity())  # blocks.2: MaxPool3d((2,1,1)) -- no parameters, keeps pytorchvideo's numbering
            dim_in = out
        blocks.append(nn.Identity())          # blocks.6: the pooling head (no parameters)
        self.blocks = nn.ModuleList(blocks)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")

    def _build_table(self) -> Tuple[List[Op], ParamPacker, int]:
        if self.precision != "bf16":
            raise NotImplementedError("i3d_8x8_r50 is built for the bf16 mode only (the TF32 plan has no windowed average pool)")
        pk = ParamPacker()
        ops: List[Op] = []
        T1, T2, DS = 3, 4, 5
        stem = self.blocks[0]
        ops.append(self._conv_op(pk, stem.conv, stem.norm, src=0, dst=1, relu=True, fold_w=self.precision == "bf16", name="blocks.0.conv"))
        ops.append(Op(kind=_lib.VAD_OP_MAXPOOL, src=1, dst=2, kernel=(1, 3, 3), stride=(1, 2, 2), pad=(0, 1, 1), name="blocks.0.pool"))
        cur = 2
        for s in range(4):
            stage = self.blocks[STAGE_BLOCK_INDEX[s]]
            for b, blk in enumerate(stage.res_blocks):
                nxt = 1 if cur == 2 else 2
                n = f"blocks.{STAGE_BLOCK_INDEX[s]}.res_blocks.{b}"
                br = blk.branch2
                ops.append(self._conv_op(pk, br.conv_a, br.norm_a, cur, T1, relu=True, name=n + ".conv_a"))
                ops.append(self._conv_op(pk, br.conv_b, br.norm_b, T1, T2, relu=True, name=n + ".conv_b"))
                res = cur
                if hasattr(blk, "branch1_conv"):
                    ops.append(self._conv_op(pk, blk.branch1_conv, blk.branch1_norm, cur, DS, relu=False, name=n + ".branch1"))
                    res = DS
                ops.append(self._conv_op(pk, br.conv_c, br.norm_c, T2, nxt, relu=True, res=res, name=n + ".conv_c"))
                cur = nxt
            if s == 0:
                nxt = 1 if cur == 2 else 2
                ops.append(Op(kind=_lib.VAD_OP_MAXPOOL, src=cur, dst=nxt, kernel=(2, 1, 1), stride=(2, 1, 1), name="blocks.2.pool"))
                cur = nxt
        # AvgPool3d((4, 7, 7), stride 1) + AdaptiveAvgPool3d(1): K4 with a temporal window of 4 frames
        ops.append(Op(kind=_lib.VAD_OP_AVGPOOL, src=cur, kernel=(4, 0, 0), name="head"))
        return ops, pk, 6


__all__ = ["I3D8x8R50"]
