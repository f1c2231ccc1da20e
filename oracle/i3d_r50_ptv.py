"""CPU restatement of the backbone the reference's CLI uses BY DEFAULT: ``create_resnet(...)`` of pytorchvideo with the
arguments at src/i3d.py:339-350 ("i3d_8x8_r50") under the reference's own head (``create_res_pooler``, src/i3d.py:21-57).
TEST INFRASTRUCTURE, NOT PRODUCT CODE.

**PARITY UNPINNED.**  pytorchvideo is a third-party dependency of the reference that is neither vendored in it nor
installed in this image (the reference pins no version; the only hint is the URL comment pointing at tag 0.1.3,
src/i3d.py:14).  What follows restates the published architecture of ``pytorchvideo.models.resnet.create_resnet`` /
``create_res_basic_stem`` / ``create_res_stage`` / ``create_bottleneck_block`` (depth 50) as fp32 torch functional ops on a
plain state_dict with pytorchvideo's parameter names (``blocks.0.conv.weight``, ``blocks.1.res_blocks.0.branch2.conv_a.weight``,
``...branch1_conv.weight`` ...), so a real I3D_8x8_R50.pyth checkpoint would load:

  stem    Conv3d(3, 64, (5,7,7), stride (1,2,2), pad (2,3,3)) + BN + ReLU + MaxPool3d((1,3,3), (1,2,2), pad (0,1,1))
  res2-5  (3, 4, 6, 3) bottlenecks, inner widths 64 .. 512, outputs 256 .. 2048; conv_a temporal kernels per the call site:
          res2 all (3,1,1); res3 / res4 alternate (3,1,1), (1,1,1); res5 alternates (1,1,1), (3,1,1); conv_b (1,3,3) with the
          spatial stride 2 in the first block of res3-5; shortcut conv 1x1x1 (same stride) + BN where shapes change
  pool    MaxPool3d((2,1,1), stride (2,1,1)) after res2 (``stage1_pool``)
  head    AvgPool3d((4,7,7), stride 1) then AdaptiveAvgPool3d(1) -> (B, 2048, 1, 1, 1)
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

DEPTHS = (3, 4, 6, 3)
# conv_a temporal kernel of block b of stage s (cycled over the stage), src/i3d.py:342-347
CONV_A_KT: Tuple[Tuple[int, ...], ...] = ((3,), (3, 1), (3, 1), (1, 3))
STAGE_BLOCK_INDEX = (1, 3, 4, 5)  # position of res2..res5 in ``blocks`` (2 is the stage1 max-pool, 6 the head)


def conv_a_kt(stage: int, block: int) -> int:
    pat = CONV_A_KT[stage]
    return pat[block % len(pat)]


def seeded_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan = cin * k[0] * k[1] * k[2]
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * (2.0 / fan) ** 0.5

    def bn(name, c):
        sd[name + ".weight"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("blocks.0.conv", 64, 3, (5, 7, 7))
    bn("blocks.0.norm", 64)
    dim_in = 64
    for s, depth in enumerate(DEPTHS):
        inner, out = 64 * 2 ** s, 256 * 2 ** s
        for b in range(depth):
            p = f"blocks.{STAGE_BLOCK_INDEX[s]}.res_blocks.{b}"
            if b == 0:
                conv(p + ".branch1_conv", out, dim_in, (1, 1, 1))
                bn(p + ".branch1_norm", out)
            conv(p + ".branch2.conv_a", inner, dim_in if b == 0 else out, (conv_a_kt(s, b), 1, 1))
            bn(p + ".branch2.norm_a", inner)
            conv(p + ".branch2.conv_b", inner, inner, (1, 3, 3))
            bn(p + ".branch2.norm_b", inner)
            conv(p + ".branch2.conv_c", out, inner, (1, 1, 1))
            bn(p + ".branch2.norm_c", out)
        dim_in = out
    return sd


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """x: [B, 3, T, H, W] fp32 -> [B, 2048, 1, 1, 1] fp32."""
    x = F.relu(_bn(F.conv3d(x, sd["blocks.0.conv.weight"], None, (1, 2, 2), (2, 3, 3)), sd, "blocks.0.norm"))
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    for s, depth in enumerate(DEPTHS):
        for b in range(depth):
            p = f"blocks.{STAGE_BLOCK_INDEX[s]}.res_blocks.{b}"
            stride = 2 if (b == 0 and s > 0) else 1
            kt = conv_a_kt(s, b)
            y = F.relu(_bn(F.conv3d(x, sd[p + ".branch2.conv_a.weight"], None, 1, (kt // 2, 0, 0)), sd, p + ".branch2.norm_a"))
            y = F.relu(_bn(F.conv3d(y, sd[p + ".branch2.conv_b.weight"], None, (1, stride, stride), (0, 1, 1)), sd, p + ".branch2.norm_b"))
            y = _bn(F.conv3d(y, sd[p + ".branch2.conv_c.weight"]), sd, p + ".branch2.norm_c")
            sc = x
            if b == 0:
                sc = _bn(F.conv3d(x, sd[p + ".branch1_conv.weight"], None, (1, stride, stride)), sd, p + ".branch1_norm")
            x = F.relu(y + sc)
        if s == 0:
            x = F.max_pool3d(x, (2, 1, 1), (2, 1, 1))
    kt = min(4, x.shape[2])
    x = F.avg_pool3d(x, (kt, x.shape[3], x.shape[4]), 1)
    return F.adaptive_avg_pool3d(x, 1)


def conv_macs(t: int = 16, h: int = 224, w: int = 224) -> int:
    """Multiply-accumulates of the convolutions for one [3, t, h, w] clip."""
    macs = 0
    ho, wo = h // 2, w // 2
    macs += t * ho * wo * 64 * 3 * 5 * 7 * 7
    ho, wo = (ho + 2 - 3) // 2 + 1, (wo + 2 - 3) // 2 + 1
    dim_in, tt = 64, t
    for s, depth in enumerate(DEPTHS):
        inner, out = 64 * 2 ** s, 256 * 2 ** s
        for b in range(depth):
            stride = 2 if (b == 0 and s > 0) else 1
            cin = dim_in if b == 0 else out
            macs += tt * ho * wo * inner * cin * conv_a_kt(s, b)
            h2, w2 = (ho + 2 - 3) // stride + 1, (wo + 2 - 3) // stride + 1
            macs += tt * h2 * w2 * inner * inner * 9
            macs += tt * h2 * w2 * out * inner
            if b == 0:
                macs += tt * h2 * w2 * out * cin
            ho, wo = h2, w2
        dim_in = out
        if s == 0:
            tt //= 2
    return macs
