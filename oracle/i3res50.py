"""Oracle for the I3D-ResNet50 forward (plain fp32 torch functional ops on the CPU).
TEST INFRASTRUCTURE ONLY -- the product never imports this.

Restates ``I3Res50.forward_single`` (reference src/i3d.py:302-315) and ``Bottleneck.forward``
(src/i3d.py:98-121) from a state_dict with the reference's key names, using the same ATen ops the
reference's nn.Modules dispatch to (conv3d, batch_norm(eval), relu, max_pool3d, adaptive avg pool).

``emulate_bf16=True`` additionally rounds the conv operands and every stored activation to bf16 the
way the CUDA path stores them (fp32 accumulate, fp32 scale/shift/residual/ReLU, one rounding on
store); it predicts the numerical gap of the bf16 path without a GPU and is used to set tolerances.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

# (planes, blocks, stride, temp_conv) per stage: src/i3d.py:220-243
STAGES = ((64, 3, 1, (1, 1, 1)), (128, 4, 2, (1, 0, 1, 0)), (256, 6, 2, (1, 0, 1, 0, 1, 0)), (512, 3, 2, (0, 1, 0)))
BN_EPS = 1e-5  # nn.BatchNorm3d default, src/i3d.py:75,84,88,208


def _bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def conv_bn(x: torch.Tensor, sd: Dict[str, torch.Tensor], conv: str, bn: str, stride, padding, relu: bool,
            residual: Optional[torch.Tensor] = None, emulate_bf16: bool = False) -> torch.Tensor:
    """conv3d(bias=False) -> BatchNorm3d(eval) [-> += residual] [-> ReLU]   (src/i3d.py:101-116)."""
    w = sd[conv + ".weight"].float()
    if emulate_bf16:
        y = F.conv3d(x, _bf16(w), None, stride, padding)
        scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + BN_EPS)
        shift = sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale
        y = y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    else:
        y = F.conv3d(x, w, None, stride, padding)
        y = F.batch_norm(y, sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float(), sd[bn + ".weight"].float(),
                         sd[bn + ".bias"].float(), False, 0.0, BN_EPS)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return _bf16(y) if emulate_bf16 else y


def bottleneck(x, sd, prefix: str, stride: int, temp_conv: int, has_ds: bool, emulate_bf16=False):
    out = conv_bn(x, sd, prefix + ".conv1", prefix + ".bn1", (1, 1, 1), (temp_conv, 0, 0), True, emulate_bf16=emulate_bf16)
    out = conv_bn(out, sd, prefix + ".conv2", prefix + ".bn2", (1, stride, stride), (0, 1, 1), True, emulate_bf16=emulate_bf16)
    residual = x
    if has_ds:
        residual = conv_bn(x, sd, prefix + ".downsample.0", prefix + ".downsample.1", (1, stride, stride), (0, 0, 0), False,
                           emulate_bf16=emulate_bf16)
    return conv_bn(out, sd, prefix + ".conv3", prefix + ".bn3", (1, 1, 1), (0, 0, 0), True, residual=residual,
                   emulate_bf16=emulate_bf16)


@torch.no_grad()
def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], emulate_bf16: bool = False,
            taps: Optional[List[str]] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """x: [B, 3, T, H, W] fp32 -> ([B, 2048, 1, 1, 1] fp32, {tap name: NCTHW activation}).

    tap names: "conv1", "maxpool1", "layer1" .. "layer4", "maxpool2".
    """
    taps = taps or []
    seen: Dict[str, torch.Tensor] = {}

    def tap(name, t):
        if name in taps:
            seen[name] = t.clone()

    x = x.float()
    if emulate_bf16:
        x = _bf16(x)
    x = conv_bn(x, sd, "conv1", "bn1", (2, 2, 2), (2, 3, 3), True, emulate_bf16=emulate_bf16)  # src/i3d.py:303-305
    tap("conv1", x)
    x = F.max_pool3d(x, (2, 3, 3), (2, 2, 2), 0)                                                  # :306
    tap("maxpool1", x)
    inplanes = 64
    for li, (planes, blocks, stride, temp_conv) in enumerate(STAGES, start=1):
        for b in range(blocks):
            first = b == 0
            has_ds = first and (stride != 1 or inplanes != planes * 4)                            # :256-260
            x = bottleneck(x, sd, f"layer{li}.{b}", stride if first else 1, temp_conv[b], has_ds, emulate_bf16)
            inplanes = planes * 4
        tap(f"layer{li}", x)
        if li == 1:
            x = F.max_pool3d(x, (2, 1, 1), (2, 1, 1), 0)                                          # :309
            tap("maxpool2", x)
    x = F.adaptive_avg_pool3d(x, 1)                                                               # :314
    return x, seen


def seeded_state_dict(seed: int = 0, randomize_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights with the reference's key names and shapes.

    Conv weights follow the reference constructor's kaiming-normal(fan_out) scale (src/i3d.py:246-248);
    BatchNorm statistics are randomised (the constructor's gamma=1, beta=0, mean=0, var=1 would hide
    BN-folding bugs).  The last BN of every block gets a small gamma so activations stay O(1) through
    16 residual blocks.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        fan_out = cout * k[0] * k[1] * k[2]
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g) * (2.0 / fan_out) ** 0.5

    def bn(name, c, gamma_scale=1.0):
        if randomize_bn:
            sd[name + ".weight"] = (0.5 + torch.rand(c, generator=g)) * gamma_scale
            sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
            sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[name + ".weight"] = torch.ones(c)
            sd[name + ".bias"] = torch.zeros(c)
            sd[name + ".running_mean"] = torch.zeros(c)
            sd[name + ".running_var"] = torch.ones(c)
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    conv("conv1", 64, 3, (5, 7, 7))
    bn("bn1", 64)
    inplanes = 64
    for li, (planes, blocks, stride, temp_conv) in enumerate(STAGES, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            conv(p + ".conv1", planes, inplanes, (1 + 2 * temp_conv[b], 1, 1))
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, (1, 3, 3))
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, (1, 1, 1))
            bn(p + ".bn3", planes * 4, gamma_scale=0.5)
            if b == 0 and (stride != 1 or inplanes != planes * 4):
                conv(p + ".downsample.0", planes * 4, inplanes, (1, 1, 1))
                bn(p + ".downsample.1", planes * 4, gamma_scale=0.5)
            inplanes = planes * 4
    return sd


def conv_macs(t: int = 16, h: int = 224, w: int = 224) -> int:
    """Multiply-accumulates of all 53 convs for one clip (16,414,572,544 at 16x224x224)."""
    def out(n, k, s, p):
        return (n + 2 * p - k) // s + 1

    total = 0
    T, H, W = out(t, 5, 2, 2), out(h, 7, 2, 3), out(w, 7, 2, 3)
    total += T * H * W * 64 * 3 * 5 * 7 * 7
    T, H, W = out(T, 2, 2, 0), out(H, 3, 2, 0), out(W, 3, 2, 0)
    inplanes = 64
    for li, (planes, blocks, stride, temp_conv) in enumerate(STAGES, start=1):
        for b in range(blocks):
            s = stride if b == 0 else 1
            total += T * H * W * planes * inplanes * (1 + 2 * temp_conv[b])
            Ho, Wo = out(H, 3, s, 1), out(W, 3, s, 1)
            total += T * Ho * Wo * planes * planes * 9
            total += T * Ho * Wo * planes * 4 * planes
            if b == 0 and (stride != 1 or inplanes != planes * 4):
                total += T * Ho * Wo * planes * 4 * inplanes
            H, W = Ho, Wo
            inplanes = planes * 4
        if li == 1:
            T = out(T, 2, 2, 0)
    return total
