"""CPU restatement of InceptionV1-3D feature extraction -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED BY THE REFERENCE: jinmang2/anomaly_detection_on_video does not contain this backbone (SURVEY.md
section 0 / Appendix B); `north_star` names it.  This file restates the public I3D architecture (Carreira &
Zisserman 2017, as in the common PyTorch port of the Kinetics checkpoint) in fp32 torch functional ops on a plain
state_dict, so the native op table (anomaly_detection_on_video_b200/inception.py) has an independent check:
  Unit3D               conv3d(bias=False) with TF-"SAME" padding (out = ceil(in / stride), front = total // 2)
                       + BatchNorm3d(eps = 1e-3) in eval mode + ReLU
  MaxPool3dSamePadding zero padding to SAME, then max-pool
  InceptionModule      cat([b0, b1b(b1a), b2b(b2a), b3b(pool)], dim = 1)
  extract_features     ... Mixed_5c -> AvgPool3d([2, 7, 7], stride 1)
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

MIXED = (
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


def _same_pad(x: torch.Tensor, k: Tuple[int, int, int], s: Tuple[int, int, int]) -> torch.Tensor:
    pads = []
    for dim in (4, 3, 2):  # F.pad wants the last dimension first
        size, kk, ss = x.shape[dim], k[dim - 2], s[dim - 2]
        tot = max(kk - ss, 0) if size % ss == 0 else max(kk - (size % ss), 0)
        pads += [tot // 2, tot - tot // 2]
    return F.pad(x, pads)


def unit3d(x, sd, name, k=(1, 1, 1), s=(1, 1, 1), emulate_bf16=False):
    w = sd[name + ".conv3d.weight"]
    if emulate_bf16:
        x, w = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    y = F.conv3d(_same_pad(x, k, s), w, None, s)
    y = F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"], sd[name + ".bn.bias"],
                     training=False, eps=1e-3)
    y = F.relu(y)
    return y.to(torch.bfloat16).float() if emulate_bf16 else y


def maxpool_same(x, k, s):
    return F.max_pool3d(_same_pad(x, k, s), k, s)


def seeded_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def unit(name, cin, cout, k=(1, 1, 1)):
        fan = cin * k[0] * k[1] * k[2]
        sd[name + ".conv3d.weight"] = torch.randn(cout, cin, *k, generator=g) * (2.0 / fan) ** 0.5
        sd[name + ".bn.weight"] = 0.8 + 0.4 * torch.rand(cout, generator=g)
        sd[name + ".bn.bias"] = 0.05 * torch.randn(cout, generator=g)
        sd[name + ".bn.running_mean"] = 0.05 * torch.randn(cout, generator=g)
        sd[name + ".bn.running_var"] = 0.7 + 0.6 * torch.rand(cout, generator=g)
        sd[name + ".bn.num_batches_tracked"] = torch.tensor(0)

    unit("Conv3d_1a_7x7", 3, 64, (7, 7, 7))
    unit("Conv3d_2b_1x1", 64, 64)
    unit("Conv3d_2c_3x3", 64, 192, (3, 3, 3))
    for name, cin, o in MIXED:
        unit(name + ".b0", cin, o[0])
        unit(name + ".b1a", cin, o[1])
        unit(name + ".b1b", o[1], o[2], (3, 3, 3))
        unit(name + ".b2a", cin, o[3])
        unit(name + ".b2b", o[3], o[4], (3, 3, 3))
        unit(name + ".b3b", cin, o[5])
    sd["logits.weight"] = torch.randn(400, 1024, 1, 1, 1, generator=g) * 0.03
    sd["logits.bias"] = torch.zeros(400)
    return sd


@torch.no_grad()
def extract_features(x: torch.Tensor, sd: Dict[str, torch.Tensor], emulate_bf16: bool = False) -> torch.Tensor:
    """[B, 3, 16, 224, 224] fp32 -> [B, 1024] (the AvgPool3d([2,7,7]) output, squeezed)."""
    e = emulate_bf16
    x = unit3d(x, sd, "Conv3d_1a_7x7", (7, 7, 7), (2, 2, 2), e)
    x = maxpool_same(x, (1, 3, 3), (1, 2, 2))
    x = unit3d(x, sd, "Conv3d_2b_1x1", emulate_bf16=e)
    x = unit3d(x, sd, "Conv3d_2c_3x3", (3, 3, 3), emulate_bf16=e)
    x = maxpool_same(x, (1, 3, 3), (1, 2, 2))
    for name, _, _ in MIXED:
        if name == "Mixed_4b":
            x = maxpool_same(x, (3, 3, 3), (2, 2, 2))
        if name == "Mixed_5b":
            x = maxpool_same(x, (2, 2, 2), (2, 2, 2))
        b0 = unit3d(x, sd, name + ".b0", emulate_bf16=e)
        b1 = unit3d(unit3d(x, sd, name + ".b1a", emulate_bf16=e), sd, name + ".b1b", (3, 3, 3), emulate_bf16=e)
        b2 = unit3d(unit3d(x, sd, name + ".b2a", emulate_bf16=e), sd, name + ".b2b", (3, 3, 3), emulate_bf16=e)
        b3 = unit3d(maxpool_same(x, (3, 3, 3), (1, 1, 1)), sd, name + ".b3b", emulate_bf16=e)
        x = torch.cat([b0, b1, b2, b3], dim=1)
    x = F.avg_pool3d(x, (2, 7, 7), 1)
    return x.reshape(x.shape[0], -1)


def conv_macs() -> int:
    """MACs of one 16 x 224 x 224 clip (27,787,569,152)."""
    total = 0

    def add(cin, cout, k, out):
        nonlocal total
        total += cin * cout * k[0] * k[1] * k[2] * out[0] * out[1] * out[2]

    add(3, 64, (7, 7, 7), (8, 112, 112))
    add(64, 64, (1, 1, 1), (8, 56, 56))
    add(64, 192, (3, 3, 3), (8, 56, 56))
    size = {"3": (8, 28, 28), "4": (4, 14, 14), "5": (2, 7, 7)}
    for name, cin, o in MIXED:
        out = size[name[6]]
        add(cin, o[0], (1, 1, 1), out); add(cin, o[1], (1, 1, 1), out); add(o[1], o[2], (3, 3, 3), out)
        add(cin, o[3], (1, 1, 1), out); add(o[3], o[4], (3, 3, 3), out); add(cin, o[5], (1, 1, 1), out)
    return total
