"""Drop-in for the reference backbone module ``src/i3d.py``.

``I3Res50`` keeps the reference's parameter names (``conv1``, ``bn1``, ``layerL.B.convK``,
``layerL.B.bnK``, ``layerL.0.downsample.{0,1}``; reference src/i3d.py:60-121,198-300) so its
checkpoints load with ``load_state_dict``, and the reference's call contract
``model(crop: [B,3,T,H,W] fp32 cuda) -> [B,2048,1,1,1] fp32`` (extract_features.py:86-89).
The forward itself is not torch: the modules only *hold* parameters; they are folded
(BatchNorm -> per-channel scale/shift), packed to bf16 and executed by the sm_100a kernels behind
``libvad_b200.so``.  There is no CPU forward.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from .engine import BackbonePlan, Op, ParamPacker, Tf32Plan, fold_bn, ingest_ncthw, ingest_ncthw_tf32

# (planes, blocks, spatial stride, temporal-conv flag per block): reference src/i3d.py:220-243
I3RES50_STAGES: Tuple[Tuple[int, int, int, Tuple[int, ...]], ...] = (
    (64, 3, 1, (1, 1, 1)),
    (128, 4, 2, (1, 0, 1, 0)),
    (256, 6, 2, (1, 0, 1, 0, 1, 0)),
    (512, 3, 2, (0, 1, 0)),
)
STEM_PAD_LEFT = 3  # == stem pw, so the folded window of output column wo starts at padded pixel 2*wo


class Bottleneck(nn.Module):
    """Parameter container for one residual block (reference src/i3d.py:60-121)."""

    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int, temp_conv: int, with_downsample: bool) -> None:
        super().__init__()
        kt = 1 + 2 * temp_conv
        self.conv1 = nn.Conv3d(inplanes, planes, (kt, 1, 1), stride=1, padding=(temp_conv, 0, 0), bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, (1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.downsample: Optional[nn.Sequential] = None
        if with_downsample:
            self.downsample = nn.Sequential(
                nn.Conv3d(inplanes, planes * 4, 1, stride=(1, stride, stride), bias=False),
                nn.BatchNorm3d(planes * 4),
            )
        self.stride = stride


def _conv_geom(conv: nn.Conv3d):
    return tuple(conv.kernel_size), tuple(conv.stride), tuple(conv.padding)


class _NativeBackbone(nn.Module):
    """Shared machinery: build the op table from the module tree, run it through the C ABI."""

    feature_dim = 0
    pad_left = STEM_PAD_LEFT  # zero pixels left of every row in the stem layout this backbone reads

    def __init__(self) -> None:
        super().__init__()
        self._plan: Optional[BackbonePlan] = None
        self._plan_key = None
        self.force_gather = False  # debug: feed every conv through the cp.async gather producer
        self.fuse_stem_pool = True  # temporal half of maxpool1 in the stem epilogue (VAD_FLAG_POOL_T2)
        self.fuse_pool2 = False     # maxpool2 in layer1's last conv3 epilogue; set per forward from the clip length
        # "bf16" (production: bf16 activations / weights, fp32 accumulate) or "tf32" (fp32 activations / weights, tcgen05
        # kind::tf32 MMAs; dedicated stem + CTA-pair kernels, otherwise general ones: features within 1e-3 of the reference's fp32 path)
        self.precision = "bf16"
        self.tf32_stem_planes = True  # tf32 mode: dedicated stem kernel on the column-parity plane layout (False: gather stem)

    # subclasses return (ops, packer, n_slots)
    def _build_table(self) -> Tuple[List[Op], ParamPacker, int]:
        raise NotImplementedError

    def _conv_op(self, packer: ParamPacker, conv: nn.Conv3d, bn: nn.BatchNorm3d, src: int, dst: int, relu: bool,
                 res: int = -1, fold_w: bool = False, name: str = "", dst_c_off: int = 0, dst_c_total: int = 0) -> Op:
        scale, shift = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
        if conv.bias is not None:
            shift = shift + conv.bias.detach().float() * scale
        k, s, p = _conv_geom(conv)
        if self.precision == "tf32":
            # fp32 weights, K padded to 32; the RGB stem reads the 4-channel (zero-padded) fp32 input, no folded window
            cin = (conv.in_channels + 3) // 4 * 4
            fold = fold_w and cin == 4 and k[2] <= 8  # RGB stem: contract 8-pixel x 4-channel windows (contiguous 128 B)
            # the dedicated TF32 stem kernel (column-parity plane input): stride 2 in h and w, 64 output channels,
            # pad 3 in w (window pixel j of output column w' is padded pixel 2 w' + j) and at most 36 (dt, dh) taps (144 KB of
            # fp32 weights per half of the output channels resident in shared memory): the I3Res50 stem, not Inception's 7x7x7
            planes = (fold and self.tf32_stem_planes and src == 0 and s[1] == 2 and s[2] == 2 and conv.out_channels == 64
                      and p[2] == 3 and k[0] * k[1] <= 36 and k[1] > 1 and res < 0 and not dst_c_total)
            w_off, s_off, b_off = packer.add_conv(conv.weight, scale, shift, cin_pad=cin, tf32=True, fold_w=fold, planes=planes)
            return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, res=res, cin=cin, cout=conv.out_channels, kernel=k, stride=s, pad=p,
                      flags=(_lib.VAD_FLAG_RELU if relu else 0) | (_lib.VAD_FLAG_STEM_FOLD_W if fold else 0) |
                            (_lib.VAD_FLAG_STEM_PLANES if planes else 0), dst_c_off=dst_c_off,
                      dst_c_total=dst_c_total, w_off=w_off, scale_off=s_off, shift_off=b_off, name=name)
        w_off, s_off, b_off = packer.add_conv(conv.weight, scale, shift, fold_w=fold_w)
        flags = (_lib.VAD_FLAG_RELU if relu else 0) | (_lib.VAD_FLAG_STEM_FOLD_W if fold_w else 0)
        if self.force_gather:
            flags |= _lib.VAD_FLAG_FORCE_GATHER
        return Op(kind=_lib.VAD_OP_CONV, src=src, dst=dst, res=res, cin=4 if fold_w else conv.in_channels,
                  cout=conv.out_channels, kernel=k, stride=s, pad=p, flags=flags, dst_c_off=dst_c_off,
                  dst_c_total=dst_c_total, w_off=w_off, scale_off=s_off, shift_off=b_off, name=name)

    def _param_key(self):
        # rebuild the packed blob whenever a parameter / buffer tensor was replaced or modified in place
        return tuple((id(t), t._version, t.device) for t in list(self.parameters()) + list(self.buffers()))

    def plan(self, device: torch.device):
        if self.precision not in ("bf16", "tf32"):
            raise ValueError(f"precision must be 'bf16' or 'tf32', not {self.precision!r}")
        key = (self._param_key(), str(device), self.force_gather, self.fuse_stem_pool, self.fuse_pool2, self.precision, self.tf32_stem_planes,
               getattr(self, "fuse_siblings", None), getattr(self, "pad_branches", None))
        if self._plan is None or self._plan_key != key:
            ops, packer, n_slots = self._build_table()
            if self.precision == "tf32":
                self._plan = Tf32Plan(ops, packer.blob(), n_slots, device, in_channels=4)
            else:
                self._plan = BackbonePlan(ops, packer.blob(), n_slots, self.pad_left, device)
            self._plan_key = key
        return self._plan

    @property
    def _fused(self) -> bool:
        """Fused-pool / folded-stem op-table variants exist for the bf16 kernels only."""
        return self.precision == "bf16" and not self.force_gather

    def op_table(self) -> List[Op]:
        return self._build_table()[0]

    def forward_stem_layout(self, x_stem: torch.Tensor) -> torch.Tensor:
        """bf16 stem-layout clips [B, T, H, W+8, 4] (what ``Preprocessor`` emits) -> [B, C] fp32."""
        if self.training:
            raise RuntimeError("the native backbone is inference-only: call .eval() first (extract_features.py:36)")
        if self.precision != "bf16":
            raise RuntimeError("the stem layout is the bf16 mode's input; in tf32 mode call forward() with fp32 NCTHW clips")
        self._select_fusions(int(x_stem.shape[1]))
        return self.plan(x_stem.device).forward(x_stem)

    def _select_fusions(self, clip_len: int) -> None:
        """Hook: pick the op-table variant for this clip length (plans are cached per variant)."""

    def forward(self, batch: torch.Tensor) -> torch.Tensor:
        """[B, 3, T, H, W] fp32 on the GPU -> [B, C, 1, 1, 1] fp32 (reference src/i3d.py:302-318)."""
        if not batch.is_cuda:
            raise RuntimeError("I3D features are computed by sm_100a kernels only; move the input (and the model) "
                               "to a CUDA device. There is no CPU fallback.")
        if self.precision == "tf32":
            if self.training:
                raise RuntimeError("the native backbone is inference-only: call .eval() first (extract_features.py:36)")
            self._select_fusions(int(batch.shape[2]))
            plan = self.plan(batch.device)
            planes = bool(plan.ops[0].flags & _lib.VAD_FLAG_STEM_PLANES) and batch.shape[-1] % 2 == 0
            if not planes and plan.ops[0].flags & _lib.VAD_FLAG_STEM_PLANES:
                raise ValueError("the TF32 stem kernel needs an even frame width; set model.tf32_stem_planes = False")
            feats = plan.forward(ingest_ncthw_tf32(batch.float(), planes=planes))
        else:
            feats = self.forward_stem_layout(ingest_ncthw(batch.float(), self.pad_left))
        return feats.view(feats.shape[0], feats.shape[1], 1, 1, 1)


class I3Res50(_NativeBackbone):
    """I3D-ResNet50 feature extractor (reference src/i3d.py:198-318), 2048-d output."""

    feature_dim = 2048

    def __init__(self, layers: Sequence[int] = (3, 4, 6, 3), use_nl: bool = False) -> None:
        super().__init__()
        if use_nl:
            raise NotImplementedError("non-local blocks are dead code in every shipped reference configuration "
                                      "(src/i3d.py:338 always passes use_nl=False) and are not built")
        self.conv1 = nn.Conv3d(3, 64, (5, 7, 7), stride=(2, 2, 2), padding=(2, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        inplanes = 64
        for li, ((planes, _, stride, temp_conv), nblocks) in enumerate(zip(I3RES50_STAGES, layers), start=1):
            blocks = []
            for b in range(nblocks):
                first = b == 0
                need_ds = first and (stride != 1 or inplanes != planes * 4)
                blocks.append(Bottleneck(inplanes, planes, stride if first else 1, temp_conv[b % len(temp_conv)], need_ds))
                inplanes = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        # same initialisation as the reference constructor (src/i3d.py:246-251)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    allow_fuse_pool2 = True

    def _select_fusions(self, clip_len: int) -> None:
        # frames entering layer1: stem (k5, s2, p2) then maxpool1's temporal half (k2, s2); the fused maxpool2 epilogue
        # is built for exactly four of them (16-frame clips, the reference's frames_per_clip)
        t1 = ((clip_len + 4 - 5) // 2 + 1) // 2
        self.fuse_pool2 = bool(self.allow_fuse_pool2 and self.fuse_stem_pool and t1 == 4)

    def _build_table(self) -> Tuple[List[Op], ParamPacker, int]:
        pk = ParamPacker()
        ops: List[Op] = []
        T1, T2, DS = 3, 4, 5  # bottleneck temporaries; slots 1/2 ping-pong the block input/output
        stem = self._conv_op(pk, self.conv1, self.bn1, src=0, dst=1, relu=True, fold_w=True, name="conv1")
        if self.fuse_stem_pool and self._fused:
            # maxpool1 = max over (2,3,3) / stride (2,2,2), pad 0 (reference src/i3d.py:212-214) separates exactly into
            # a max over frame pairs -- done in the stem kernel's epilogue, so the full-rate stem output never
            # reaches HBM -- followed by a spatial (1,3,3) / (1,2,2) max-pool.
            stem.flags |= _lib.VAD_FLAG_POOL_T2
            ops.append(stem)
            ops.append(Op(kind=_lib.VAD_OP_MAXPOOL, src=1, dst=2, kernel=(1, 3, 3), stride=(1, 2, 2), name="maxpool1"))
        else:
            ops.append(stem)
            ops.append(Op(kind=_lib.VAD_OP_MAXPOOL, src=1, dst=2, kernel=(2, 3, 3), stride=(2, 2, 2), name="maxpool1"))
        cur = 2
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(self, f"layer{li}")):
                nxt = 1 if cur == 2 else 2
                n = f"layer{li}.{bi}"
                ops.append(self._conv_op(pk, blk.conv1, blk.bn1, cur, T1, relu=True, name=n + ".conv1"))
                ops.append(self._conv_op(pk, blk.conv2, blk.bn2, T1, T2, relu=True, name=n + ".conv2"))
                res = cur
                if blk.downsample is not None:
                    ops.append(self._conv_op(pk, blk.downsample[0], blk.downsample[1], cur, DS, relu=False, name=n + ".downsample"))
                    res = DS
                conv3 = self._conv_op(pk, blk.conv3, blk.bn3, T2, nxt, relu=True, res=res, name=n + ".conv3")
                last_of_layer1 = li == 1 and bi == len(getattr(self, "layer1")) - 1
                if last_of_layer1 and self.fuse_pool2 and self._fused:
                    # maxpool2 = max over frame pairs (reference src/i3d.py:215-217,309) in conv3's staged epilogue: the
                    # unpooled block output (991 MB per 160 clips) is never written and no pool kernel runs.  Needs the
                    # 4-frame feature map of a 16-frame clip; other clip lengths take the separate pool below.
                    conv3.flags |= _lib.VAD_FLAG_POOL_T2
                    conv3.name = n + ".conv3+maxpool2"
                ops.append(conv3)
                cur = nxt
            if li == 1 and not (self.fuse_pool2 and self._fused):
                nxt = 1 if cur == 2 else 2
                ops.append(Op(kind=_lib.VAD_OP_MAXPOOL, src=cur, dst=nxt, kernel=(2, 1, 1), stride=(2, 1, 1), name="maxpool2"))
                cur = nxt
        ops.append(Op(kind=_lib.VAD_OP_AVGPOOL, src=cur, name="avgpool"))
        return ops, pk, 6


MODEL_ZOO: Dict[str, str] = {
    # file names of the reference's hub checkpoints (src/i3d.py:12-18); loaded from a local path here
    "i3d_8x8_r50": "I3D_8x8_R50.pyth",
    "tushar-n-baseline": "converted_ref_i3d.pt",
}


def print_model_size(model: nn.Module) -> None:
    bits = sum(p.numel() * (torch.finfo(p.dtype).bits if p.is_floating_point() else torch.iinfo(p.dtype).bits)
               for p in model.parameters())
    print(f"model size: {bits} / bit | {bits / 8e6:.2f} / MB")


def build_i3d_feature_extractor(model_name: str = "tushar-n-baseline", check_model_size: bool = True,
                                strict: bool = False, state_dict_path: Optional[str] = None) -> nn.Module:
    """Same signature as the reference factory (src/i3d.py:332-364) plus ``state_dict_path``.

    The reference downloads the checkpoint from the HF hub; this environment has no network, so
    weights come from ``state_dict_path`` (or stay at the constructor's random init).
    ``i3d_8x8_r50`` (the CLI's default in the reference) is pytorchvideo's third-party backbone: built from its published
    architecture (``ptv_resnet.I3D8x8R50``), parity unpinned.
    """
    if model_name == "tushar-n-baseline":
        model = I3Res50(use_nl=False)
    elif model_name == "i3d_8x8_r50":
        # pytorchvideo's create_resnet with the reference's arguments (src/i3d.py:339-350) as an op table over the native
        # kernels; third-party architecture, parity unpinned (oracle/i3d_r50_ptv.py restates it)
        from .ptv_resnet import I3D8x8R50

        model = I3D8x8R50()
    else:
        raise AttributeError(model_name)
    if state_dict_path is not None:
        sd = torch.load(state_dict_path, map_location="cpu")
        missing, unexpected = model.load_state_dict(sd, strict=strict)
        if missing or unexpected:  # the reference ignores these silently; say so
            print(f"load_state_dict: {len(missing)} missing, {len(unexpected)} unexpected keys")
    if check_model_size:
        print_model_size(model)
    return model


__all__ = ["I3Res50", "Bottleneck", "build_i3d_feature_extractor", "print_model_size", "MODEL_ZOO", "STEM_PAD_LEFT"]
