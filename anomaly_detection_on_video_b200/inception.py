"""InceptionV1-3D ("I3D", Carreira & Zisserman 2017) feature extractor on the same native kernels.

`north_star` names this backbone (``InceptionI3d`` / ``Unit3D`` / ``MaxPool3dSamePadding``, 1024-d features); the
reference repository does NOT contain it (it ships I3D-ResNet50), so there is no reference
code to be a drop-in for.  The module tree and parameter names follow the widely used public PyTorch port of the
Kinetics I3D checkpoint (``Conv3d_1a_7x7.conv3d.weight``, ``Conv3d_1a_7x7.bn.*``, ``Mixed_3b.b1a.conv3d.weight`` ...),
so those checkpoints load with ``load_state_dict``; the layer table is SURVEY.md Appendix B.  PARITY UNPINNED BY THE
REFERENCE: the oracle (``oracle/inception.py``) is our own fp32 restatement of the public architecture.

What maps onto which kernel:
  Unit3D (conv3d, TF-"SAME" padding, no bias + BatchNorm3d(eps 1e-3) + ReLU)  -> K2 with ``VAD_FLAG_CONV_SAME``;
      Cin % 64 == 32 inputs take TMA operands with 32-wide k-blocks; the narrow branch temporaries (16 .. 160 channels) and
      Mixed_4e's 528-channel output are widened with zero channels so that they do too, or take 64-wide ones (BRANCH_PAD,
      CONCAT_PAD below); b0 | b1a | b2a of a block are one launch with routed output columns (``_siblings``)
  MaxPool3dSamePadding                                                        -> K3 with ``VAD_FLAG_POOL_SAME``
  Inception branch concat                                                     -> every branch conv writes its
      channel slice of the concatenated tensor directly (``dst_c_off`` / ``dst_c_total``): no copy kernel
  AvgPool3d([2, 7, 7]) on the final 2 x 7 x 7 map                             -> K4 (global mean)
The 7x7x7 / 2 stem (49 taps x 4 KB of weights: too many for one CTA's shared memory) runs on CTA pairs, each CTA keeping the
taps of half of the output channels resident (csrc/stem_pair.cuh).  12.5 k clips/s per B200 at 160 clip-crops per forward (0.49-0.50 of
the sustained bf16 peak), 11.5 k for a whole video end to end.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from .engine import Op, ParamPacker, fold_bn
from .i3d import _NativeBackbone

INCEPTION_PAD_LEFT = 2  # == the SAME front pad of the 7-wide stride-2 stem, so the folded window starts on a 16 B boundary

# (name, in_channels, [b0, b1a, b1b, b2a, b2b, b3b]) -- SURVEY.md Appendix B
MIXED: Tuple[Tuple[str, int, Tuple[int, int, int, int, int, int]], ...] = (
    ("Mixed_3b", 192, (64, 96, 128, 16, 32, 32)),
    ("Mixed_3c", 256, (128, 128, 192, 32, 96, 64)),
    ("Mixed_4b", 480, (192, 96, 208, 16, 48, 64)),
    ("Mixed_4c", 512, (160, 112, 224, 24, 64, 64)),
    ("Mixed_4d", 512, (128, 128, 256, 24, 64, 64)),
    ("Mixed_4e", 512, (112, 144, 288, 32, 64, 64)),
    ("Mixed_4f", 528, (256, 160, 320, 32, 128, 128)),
    ("Mixed_5b", 832, (256, 160, 320, 32, 128, 128)),
    ("Mixed_5c", 832, (384, 192, 384, 48, 128, 128)),
)
# max-pools that sit in front of a Mixed block: name -> (kernel, stride)
# Branch temporaries whose channel count would put the following 3x3x3 conv on a narrow k-block path are widened with zero
# channels (zero weight rows in the producing 1x1x1 conv: relu(0 * x + 0) = 0; zero weight columns in the consumer): the
# consumer then contracts whole 64-channel k-blocks through the im2col TMA path.  Measured per 160 clip-crops (round 2):
# Cin = 144 (16-wide k-blocks, one 32-byte sector per TMA row) ran at 0.26 of the tensor peak, 112 at 0.28, 16 at 0.05, 24 on
# the cp.async gather at 0.08; a 64-multiple Cin reaches 0.8; 96 -> 128 takes
# Mixed_3b.b1b off the 32-wide k-block path (0.78 -> 0.65 ms); 32 -> 64 was measured and LOSES (Mixed_3c.b2b 0.33 -> 0.40 ms: that
# small-N layer is bound by its im2col traffic through L2, which doubles).  Internal to a Mixed block: results are unchanged (the
# extra terms are exact zeros).
# 16 and 24 channels go to 32, not 64 (32-wide k-blocks): Mixed_3b.b2b 0.366 -> 0.256 ms, 4b / 4c / 4d.b2b 0.049 -> 0.043 ms -- the
# same im2col-traffic argument as for 32 -> 64; 48 -> 64 stays (one 64-wide k-block per tap instead of a 32- and a 16-wide one).
BRANCH_PAD = {16: 32, 24: 32, 48: 64, 96: 128, 112: 128, 144: 192, 160: 192}
# The same for a whole block output: Mixed_4e's 528-channel concat (8 x 64 + 16) puts all of Mixed_4f's 1x1x1 convs on the 16-wide
# k-block path (ncu, 160 clip-crops: the fused sibling launch 180 us against 72-74 us for its 512-channel peers, b3b 58 against
# 28 us).  The last branch conv of 4e writes 48 zero channels behind the concat and 4f's convs carry zero weight columns for them;
# 4f's branch pool moves 9 % more bytes.  480 (Mixed_3c -> MaxPool3d_4a -> 4b, 32-wide k-blocks) is left alone: the pools in
# between would pay more than the convs gain.
CONCAT_PAD = {528: 576}
if os.environ.get("VAD_BRANCH_PAD"):   # A/B: "16:32,24:32"
    BRANCH_PAD.update({int(a): int(b) for a, b in (kv.split(":") for kv in os.environ["VAD_BRANCH_PAD"].split(","))})

POOL_BEFORE = {"Mixed_4b": ("MaxPool3d_4a_3x3", (3, 3, 3), (2, 2, 2)), "Mixed_5b": ("MaxPool3d_5a_2x2", (2, 2, 2), (2, 2, 2))}


class Unit3D(nn.Module):
    """Parameter container: Conv3d (no bias) + BatchNorm3d(eps=1e-3) + ReLU with TF-SAME padding."""

    def __init__(self, in_channels: int, output_channels: int, kernel_shape=(1, 1, 1), stride=(1, 1, 1)) -> None:
        super().__init__()
        self.conv3d = nn.Conv3d(in_channels, output_channels, kernel_shape, stride=stride, padding=0, bias=False)
        self.bn = nn.BatchNorm3d(output_channels, eps=0.001, momentum=0.01)


class MaxPool3dSamePadding(nn.Module):
    def __init__(self, kernel_size, stride) -> None:
        super().__init__()
        self.kernel_size, self.stride = tuple(kernel_size), tuple(stride)


class InceptionModule(nn.Module):
    def __init__(self, in_channels: int, out_channels: Sequence[int]) -> None:
        super().__init__()
        self.b0 = Unit3D(in_channels, out_channels[0])
        self.b1a = Unit3D(in_channels, out_channels[1])
        self.b1b = Unit3D(out_channels[1], out_channels[2], (3, 3, 3))
        self.b2a = Unit3D(in_channels, out_channels[3])
        self.b2b = Unit3D(out_channels[3], out_channels[4], (3, 3, 3))
        self.b3a = MaxPool3dSamePadding((3, 3, 3), (1, 1, 1))
        self.b3b = Unit3D(in_channels, out_channels[5])
        self.out_channels = tuple(out_channels)


class InceptionI3d(_NativeBackbone):
    """``model(x [B,3,16,224,224] fp32 cuda) -> [B,1024,1,1,1]`` (== ``extract_features``); the logits layer of the
    public port is kept as parameters for checkpoint compatibility and never evaluated (feature extraction stops at
    the average pool)."""

    feature_dim = 1024
    pad_left = INCEPTION_PAD_LEFT

    def __init__(self, num_classes: int = 400, in_channels: int = 3) -> None:
        super().__init__()
        if in_channels != 3:
            raise NotImplementedError("only the RGB stream is built (the flow stream has 2 input channels)")
        self.fuse_stem_pool = False
        self.pad_branches = True   # BRANCH_PAD; False keeps every temporary at its nominal width (A/B and tests)
        self.fuse_siblings = True  # b0 | b1a | b2a of a Mixed block as one launch (bf16 mode); False: three launches
        self.Conv3d_1a_7x7 = Unit3D(3, 64, (7, 7, 7), (2, 2, 2))
        self.MaxPool3d_2a_3x3 = MaxPool3dSamePadding((1, 3, 3), (1, 2, 2))
        self.Conv3d_2b_1x1 = Unit3D(64, 64)
        self.Conv3d_2c_3x3 = Unit3D(64, 192, (3, 3, 3))
        self.MaxPool3d_3a_3x3 = MaxPool3dSamePadding((1, 3, 3), (1, 2, 2))
        for name, cin, outs in MIXED:
            if name in POOL_BEFORE:
                pname, k, s = POOL_BEFORE[name]
                setattr(self, pname, MaxPool3dSamePadding(k, s))
            setattr(self, name, InceptionModule(cin, outs))
        self.logits = nn.Conv3d(1024, num_classes, 1)  # unused by extract_features; kept for state_dict compatibility
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _unit(self, pk: ParamPacker, u: Unit3D, src: int, dst: int, name: str, fold_w: bool = False, off: int = 0,
              total: int = 0, cin_pad: int = 0, cout_pad: int = 0) -> Op:
        scale, shift = fold_bn(u.bn.weight, u.bn.bias, u.bn.running_mean, u.bn.running_var, u.bn.eps)
        if self.precision != "tf32" and (cin_pad or cout_pad):
            # zero-widened branch temporary (BRANCH_PAD): extra output channels of the producer / input channels of the consumer
            conv = u.conv3d
            w = conv.weight.detach().float()
            co, ci = cout_pad or conv.out_channels, cin_pad or conv.in_channels
            wp = torch.zeros(co, ci, *conv.kernel_size, dtype=torch.float32, device=w.device)
            wp[:conv.out_channels, :conv.in_channels] = w
            sc = torch.zeros(co, dtype=torch.float32, device=w.device); sc[:conv.out_channels] = scale
            sh = torch.zeros(co, dtype=torch.float32, device=w.device); sh[:conv.out_channels] = shift
            w_off, s_off, b_off = pk.add_conv(wp, sc, sh)
            flags = _lib.VAD_FLAG_RELU | _lib.VAD_FLAG_CONV_SAME | (_lib.VAD_FLAG_FORCE_GATHER if self.force_gather else 0)
            return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, cin=ci, cout=co, kernel=tuple(conv.kernel_size), stride=tuple(conv.stride),
                      pad=(0, 0, 0), flags=flags, dst_c_off=off, dst_c_total=total, w_off=w_off, scale_off=s_off, shift_off=b_off, name=name)
        if self.precision == "tf32":
            conv = u.conv3d
            cin = (conv.in_channels + 3) // 4 * 4
            fold = fold_w and cin == 4 and conv.kernel_size[2] <= 8
            w_off, s_off, b_off = pk.add_conv(conv.weight, scale, shift, cin_pad=cin, tf32=True, fold_w=fold)
            return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, cin=cin, cout=conv.out_channels, kernel=tuple(conv.kernel_size),
                      stride=tuple(conv.stride), pad=(0, 0, 0),
                      flags=_lib.VAD_FLAG_RELU | _lib.VAD_FLAG_CONV_SAME | (_lib.VAD_FLAG_STEM_FOLD_W if fold else 0), dst_c_off=off,
                      dst_c_total=total, w_off=w_off, scale_off=s_off, shift_off=b_off, name=name)
        w_off, s_off, b_off = pk.add_conv(u.conv3d.weight, scale, shift, fold_w=fold_w)
        flags = _lib.VAD_FLAG_RELU | _lib.VAD_FLAG_CONV_SAME | (_lib.VAD_FLAG_STEM_FOLD_W if fold_w else 0)
        if self.force_gather:
            flags |= _lib.VAD_FLAG_FORCE_GATHER
        conv = u.conv3d
        return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, cin=4 if fold_w else conv.in_channels, cout=conv.out_channels,
                  kernel=tuple(conv.kernel_size), stride=tuple(conv.stride), pad=(0, 0, 0), flags=flags, dst_c_off=off,
                  dst_c_total=total, w_off=w_off, scale_off=s_off, shift_off=b_off, name=name)

    def _siblings(self, pk: ParamPacker, m: "InceptionModule", src: int, dst: int, t1: int, t2: int, name: str, total: int,
                  w1: int, w2: int, cin_pad: int = 0) -> Op:
        """b0 | b1a | b2a of one Mixed block as one 1x1x1 conv: the weight matrix stacks their output channels, every sibling
        starting on a multiple of 64 columns (zero rows in between, never stored); columns [0, split1) -> the concat slice,
        [split1, split2) -> branch temporary t1 (w1 channels incl. the BRANCH_PAD zeros), [split2, ..) -> t2 (w2 channels)."""
        units = (m.b0, m.b1a, m.b2a)
        cin_nom = m.b0.conv3d.in_channels
        cin = cin_pad or cin_nom   # CONCAT_PAD: the block input carries zero channels behind the nominal ones
        b0 = m.b0.conv3d.out_channels
        split1 = (b0 + 63) // 64 * 64
        split2 = split1 + (w1 + 63) // 64 * 64
        cout = split2 + w2
        dev = m.b0.conv3d.weight.device
        w = torch.zeros(cout, cin, 1, 1, 1, dtype=torch.float32, device=dev)
        sc = torch.zeros(cout, dtype=torch.float32, device=dev)
        sh = torch.zeros(cout, dtype=torch.float32, device=dev)
        for u, start in zip(units, (0, split1, split2)):
            n = u.conv3d.out_channels
            scale, shift = fold_bn(u.bn.weight, u.bn.bias, u.bn.running_mean, u.bn.running_var, u.bn.eps)
            w[start:start + n, :cin_nom] = u.conv3d.weight.detach().float()
            sc[start:start + n] = scale
            sh[start:start + n] = shift
        w_off, s_off, b_off = pk.add_conv(w, sc, sh)
        return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, cin=cin, cout=cout, kernel=(1, 1, 1), stride=(1, 1, 1), pad=(0, 0, 0),
                  flags=_lib.VAD_FLAG_RELU | _lib.VAD_FLAG_CONV_SAME, dst_c_off=0, dst_c_total=total, w_off=w_off, scale_off=s_off,
                  shift_off=b_off, name=name + ".b0+b1a+b2a", dst1=t1, dst2=t2, split1=split1, split2=split2, seg_w=(b0, w1, w2))

    @staticmethod
    def _pool(p: MaxPool3dSamePadding, src: int, dst: int, name: str) -> Op:
        return Op(kind=_lib.VAD_OP_MAXPOOL, src=src, dst=dst, kernel=p.kernel_size, stride=p.stride, flags=_lib.VAD_FLAG_POOL_SAME,
                  name=name)

    def _build_table(self) -> Tuple[List[Op], ParamPacker, int]:
        pk = ParamPacker()
        ops: List[Op] = []
        T1, T2, T3 = 3, 4, 5  # branch temporaries; slots 1 / 2 ping-pong the block input / output
        ops.append(self._unit(pk, self.Conv3d_1a_7x7, 0, 1, "Conv3d_1a_7x7", fold_w=True))
        ops.append(self._pool(self.MaxPool3d_2a_3x3, 1, 2, "MaxPool3d_2a_3x3"))
        ops.append(self._unit(pk, self.Conv3d_2b_1x1, 2, 1, "Conv3d_2b_1x1"))
        ops.append(self._unit(pk, self.Conv3d_2c_3x3, 1, 2, "Conv3d_2c_3x3"))
        ops.append(self._pool(self.MaxPool3d_3a_3x3, 2, 1, "MaxPool3d_3a_3x3"))
        cur = 1
        cur_pad = 0   # width of the current block input when CONCAT_PAD widened it, else 0
        widen = self.pad_branches and self.precision != "tf32"
        for name, _, outs in MIXED:
            if name in POOL_BEFORE:
                nxt = 2 if cur == 1 else 1
                ops.append(self._pool(getattr(self, POOL_BEFORE[name][0]), cur, nxt, POOL_BEFORE[name][0]))
                cur = nxt
            m: InceptionModule = getattr(self, name)
            nxt = 2 if cur == 1 else 1
            nominal = outs[0] + outs[2] + outs[4] + outs[5]
            total = CONCAT_PAD.get(nominal, nominal) if widen else nominal
            # torch.cat([b0, b1, b2, b3], dim=1): every branch writes its channel slice of the output in place
            p1 = BRANCH_PAD.get(outs[1], 0) if self.pad_branches else 0
            p2 = BRANCH_PAD.get(outs[3], 0) if self.pad_branches else 0
            if self.fuse_siblings and self.precision != "tf32" and not self.force_gather:
                # b0, b1a and b2a are 1x1x1 convs over the same block input, each bound by reading it: ONE launch reads it once and
                # routes its output columns to the concat slice (b0) and the two branch temporaries (vad_op_desc.dst1 / dst2)
                ops.append(self._siblings(pk, m, cur, nxt, T1, T2, name, total, p1 or outs[1], p2 or outs[3], cin_pad=cur_pad))
            else:
                ops.append(self._unit(pk, m.b0, cur, nxt, name + ".b0", off=0, total=total, cin_pad=cur_pad))
                ops.append(self._unit(pk, m.b1a, cur, T1, name + ".b1a", cin_pad=cur_pad, cout_pad=p1))
                ops.append(self._unit(pk, m.b2a, cur, T2, name + ".b2a", cin_pad=cur_pad, cout_pad=p2))
            # the pooling branch right behind the 1x1x1 convs: it reads the same block input, part of which is still in L2
            ops.append(self._pool(m.b3a, cur, T3, name + ".b3a"))
            # the last branch also writes the CONCAT_PAD zero channels behind its own (zero weights, zero scale and shift)
            ops.append(self._unit(pk, m.b3b, T3, nxt, name + ".b3b", off=outs[0] + outs[2] + outs[4], total=total, cin_pad=cur_pad,
                                  cout_pad=outs[5] + total - nominal if total != nominal else 0))
            ops.append(self._unit(pk, m.b1b, T1, nxt, name + ".b1b", off=outs[0], total=total, cin_pad=p1))
            ops.append(self._unit(pk, m.b2b, T2, nxt, name + ".b2b", off=outs[0] + outs[2], total=total, cin_pad=p2))
            cur = nxt
            cur_pad = total if total != nominal else 0
        ops.append(Op(kind=_lib.VAD_OP_AVGPOOL, src=cur, name="avg_pool"))
        return ops, pk, 6

    def extract_features(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward(x)

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        if batch.dim() == 5 and (batch.shape[2] // 8 != 2 or (batch.shape[3] + 31) // 32 != 7 or (batch.shape[4] + 31) // 32 != 7):
            raise ValueError("InceptionI3d.extract_features ends in AvgPool3d([2, 7, 7]); it is built for inputs whose final "
                             "feature map is exactly 2 x 7 x 7 (16 frames of 193..224 pixels)")
        return super().forward(batch)


__all__ = ["InceptionI3d", "InceptionModule", "Unit3D", "MaxPool3dSamePadding", "MIXED"]
