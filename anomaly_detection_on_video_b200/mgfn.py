"""Drop-in for the reference scoring head ``src/models/mgfn`` (+ the losses of ``src/loss``): inference and training.

``MGFNForVideoAnomalyDetection`` keeps the reference's module tree and parameter names
(``backbone.amplifier.to_tokens``, ``backbone.layers.S.B.{scc,attention,ffn}``, ``layer_norm``, ``fc``;
reference src/models/mgfn/modeling_mgfn.py:36-300), so its checkpoints load with ``load_state_dict``,
and the reference's call contract (modeling_mgfn.py:376-427)

    model(video [bs, ncrops, T, 2049] fp32 cuda, abnormal_labels=None, normal_labels=None)
        -> MGFNVideoAnomalyDetectionOutput(loss, abnormal_scores, normal_scores,
                                           a_feat_magnitude, n_feat_magnitude, scores)

The modules only *hold* parameters; the forward runs in ``libvad_b200.so``: every Conv1d is a tcgen05
kind::tf32 GEMM, the rest small fp32 kernels (csrc/head_kernels.cuh).  There is no CPU forward.

Training (src/runner.py:29-39,53-59).  In ``.train()`` mode ``forward`` runs ``vad_head_train_step``: the train-mode
forward (BatchNorm1d of the Focus blocks on batch statistics, dropout on the magnitude-selection mask,
modeling_mgfn.py:341-344), every loss term and the whole backward pass in one native call -- weight / input gradients
are the same tcgen05 kind::tf32 GEMM with re-arranged operands (csrc/head_train_kernels.cuh).  The returned ``loss`` is
attached to the autograd graph through a custom Function whose backward hands out the gradients the native call already
computed, so the reference's ``loss.backward(); optimizer.step()`` works with ``torch.optim.Adam`` unchanged;
``NativeAdam`` is the fused native optimizer over the same flat blobs (with the data-parallel gradient all-reduce).
While training, every parameter is a view into one flat fp32 blob in the native layout (state_dict round trips are
unaffected) and the BatchNorm running statistics are views into a second one.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import HeadConfig, check


class MGFNConfig:
    """Same fields and defaults as the reference ``MGFNConfig`` (configuration_mgfn.py:4-36)."""

    def __init__(self, classes=0, dims=(64, 128, 1024), depths=(3, 3, 2), mgfn_types=("gb", "fb", "fb"), lokernel=5,
                 channels=2048, ff_repe=4, dim_head=64, local_aggr_kernel=5, dropout=0.0, attention_dropout=0.0,
                 dropout_rate=0.7, mag_ratio=0.1, k=3):
        self.classes = classes
        self.dims = tuple(dims)
        self.depths = tuple(depths)
        self.mgfn_types = tuple(mgfn_types)
        self.lokernel = lokernel
        self.channels = channels
        self.ff_repe = ff_repe
        self.dim_head = dim_head
        self.local_aggr_kernel = local_aggr_kernel
        self.dropout = dropout
        self.attention_dropout = attention_dropout
        self.dropout_rate = dropout_rate
        self.mag_ratio = mag_ratio
        self.k = k


@dataclass
class MGFNVideoAnomalyDetectionOutput:
    """modeling_mgfn.py:26-33 (a plain dataclass instead of a transformers ``ModelOutput``)."""

    loss: Optional[torch.Tensor] = None
    abnormal_scores: Optional[torch.Tensor] = None
    normal_scores: Optional[torch.Tensor] = None
    a_feat_magnitude: Optional[torch.Tensor] = None
    n_feat_magnitude: Optional[torch.Tensor] = None
    scores: Optional[torch.Tensor] = None
    loss_terms: Optional[torch.Tensor] = None  # [total, smooth, sparsity, bce, con, con_n, con_a]


# ----------------------------------------------------------------------------- parameter containers
class MGFNLayerNorm(nn.Module):  # modeling_mgfn.py:36-47
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1))


class MGFNFeedForward(nn.Module):  # modeling_mgfn.py:50-64
    def __init__(self, dim: int, repe: int = 4):
        super().__init__()
        self.layer_norm = MGFNLayerNorm(dim)
        self.in_conv = nn.Conv1d(dim, dim * repe, 1)
        self.out_conv = nn.Conv1d(dim * repe, dim, 1)


class MGFNFeatureAmplifier(nn.Module):  # modeling_mgfn.py:67-94
    def __init__(self, config: MGFNConfig):
        super().__init__()
        self.to_tokens = nn.Conv1d(config.channels, config.dims[0], kernel_size=3, stride=1, padding=1)
        self.to_mag = nn.Conv1d(1, config.dims[0], kernel_size=3, stride=1, padding=1)


class GlanceAttention(nn.Module):  # modeling_mgfn.py:97-127
    def __init__(self, dim: int, heads: int, dim_head: int):
        super().__init__()
        self.heads = heads
        self.norm = MGFNLayerNorm(dim)
        self.to_qkv = nn.Conv1d(dim, dim_head * heads * 3, 1, bias=False)
        self.to_out = nn.Conv1d(dim_head * heads, dim, 1)


class FocusAttention(nn.Module):  # modeling_mgfn.py:152-186
    def __init__(self, dim: int, heads: int, dim_head: int, local_aggr_kernel: int):
        super().__init__()
        self.heads = heads
        self.norm = nn.BatchNorm1d(dim)
        self.to_v = nn.Conv1d(dim, dim_head * heads, 1, bias=False)
        self.rel_pos = nn.Conv1d(heads, heads, local_aggr_kernel, padding=local_aggr_kernel // 2, groups=heads)
        self.to_out = nn.Conv1d(dim_head * heads, dim, 1)


class GlanceBlock(nn.Module):  # modeling_mgfn.py:130-149
    def __init__(self, config: MGFNConfig, dim: int, heads: int):
        super().__init__()
        self.scc = nn.Conv1d(dim, dim, 3, padding=1)
        self.attention = GlanceAttention(dim=dim, heads=heads, dim_head=config.dim_head)
        self.ffn = MGFNFeedForward(dim, repe=config.ff_repe)


class FocusBlock(nn.Module):  # modeling_mgfn.py:189-212
    def __init__(self, config: MGFNConfig, dim: int, heads: int):
        super().__init__()
        self.scc = nn.Conv1d(dim, dim, 3, padding=1)
        self.attention = FocusAttention(dim=dim, heads=heads, dim_head=config.dim_head,
                                        local_aggr_kernel=config.local_aggr_kernel)
        self.ffn = MGFNFeedForward(dim, repe=config.ff_repe)


class MGFNIntermediate(nn.Module):  # modeling_mgfn.py:215-223
    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.layer_norm = MGFNLayerNorm(in_dim)
        self.conv = nn.Conv1d(in_dim, out_dim, 1, stride=1)


class MGFNModel(nn.Module):  # modeling_mgfn.py:241-283
    def __init__(self, config: MGFNConfig):
        super().__init__()
        self.amplifier = MGFNFeatureAmplifier(config)
        layers = []
        for ind, (depth, mgfn_type) in enumerate(zip(config.depths, config.mgfn_types)):
            stage_dim = config.dims[ind]
            heads = stage_dim // config.dim_head
            if mgfn_type == "gb":
                block_cls = GlanceBlock
            elif mgfn_type == "fb":
                block_cls = FocusBlock
            else:
                raise AttributeError("The type of mgfn block must be either `gb` or `fb`.")
            blocks: List[nn.Module] = [block_cls(config, dim=stage_dim, heads=heads) for _ in range(depth)]
            if ind != len(config.depths) - 1:
                blocks.append(MGFNIntermediate(stage_dim, config.dims[ind + 1]))
            layers.append(nn.Sequential(*blocks))
        self.layers = nn.Sequential(*layers)


def _pad64(t: torch.Tensor) -> torch.Tensor:
    t = t.detach().double().reshape(-1)
    pad = (-t.numel()) % 64
    return torch.cat([t, t.new_zeros(pad)]) if pad else t


class MGFNForVideoAnomalyDetection(nn.Module):
    """modeling_mgfn.py:286-427, eval-mode forward on the GPU."""

    def __init__(self, config: Optional[MGFNConfig] = None):
        super().__init__()
        self.config = config if config is not None else MGFNConfig()
        self.k = self.config.k
        last_dim = self.config.dims[-1]
        self.backbone = MGFNModel(self.config)
        self.layer_norm = nn.LayerNorm(last_dim)
        self.fc = nn.Linear(last_dim, 1)
        self._force_split = False
        self._handle: Optional[ctypes.c_void_p] = None
        self._blob: Optional[torch.Tensor] = None
        self._key = None
        self._ws: Optional[torch.Tensor] = None

    @property
    def force_split(self) -> bool:
        """Whether to separate the batch into a normal and an abnormal half in evaluation (modeling_mgfn.py:290-300)."""
        return self._force_split

    @force_split.setter
    def force_split(self, val: bool) -> None:
        self._force_split = val

    # ------------------------------------------------------------------------------ native plumbing
    def pack_parameters(self) -> torch.Tensor:
        """fp32 blob in the order ``vad_head_create`` documents (every tensor on a 64-float boundary).
        Conv1d(k=3) weights go tap-major ([cout][3][cin]); the Focus BatchNorm1d (eval statistics) is folded
        into ``to_v``:  to_v(bn(x)) = (W diag(a)) x + W c  with a = gamma / sqrt(var + eps), c = beta - mean * a."""
        cfg = self.config
        parts: List[torch.Tensor] = []
        amp = self.backbone.amplifier
        parts += [_pad64(amp.to_tokens.weight.permute(0, 2, 1)), _pad64(amp.to_tokens.bias),
                  _pad64(amp.to_mag.weight.reshape(cfg.dims[0], 3)), _pad64(amp.to_mag.bias)]
        for si, stage in enumerate(self.backbone.layers):
            for blk in stage:
                if isinstance(blk, MGFNIntermediate):
                    parts += [_pad64(blk.layer_norm.g), _pad64(blk.layer_norm.b), _pad64(blk.conv.weight[:, :, 0]),
                              _pad64(blk.conv.bias)]
                    continue
                parts += [_pad64(blk.scc.weight.permute(0, 2, 1)), _pad64(blk.scc.bias)]
                att = blk.attention
                if isinstance(att, GlanceAttention):
                    parts += [_pad64(att.norm.g), _pad64(att.norm.b), _pad64(att.to_qkv.weight[:, :, 0])]
                else:
                    bn = att.norm
                    a = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
                    c0 = bn.bias.detach().double() - bn.running_mean.detach().double() * a
                    w = att.to_v.weight.detach().double()[:, :, 0]
                    parts += [_pad64(w * a.view(1, -1)), _pad64(w @ c0), _pad64(att.rel_pos.weight[:, 0, :]),
                              _pad64(att.rel_pos.bias)]
                parts += [_pad64(att.to_out.weight[:, :, 0]), _pad64(att.to_out.bias)]
                ffn = blk.ffn
                parts += [_pad64(ffn.layer_norm.g), _pad64(ffn.layer_norm.b), _pad64(ffn.in_conv.weight[:, :, 0]),
                          _pad64(ffn.in_conv.bias), _pad64(ffn.out_conv.weight[:, :, 0]), _pad64(ffn.out_conv.bias)]
        parts += [_pad64(self.layer_norm.weight), _pad64(self.layer_norm.bias), _pad64(self.fc.weight), _pad64(self.fc.bias)]
        return torch.cat(parts).float()

    def _native(self, device: torch.device):
        steps = getattr(self, "_train", None)["steps"] if getattr(self, "_train", None) is not None else 0
        key = (tuple((id(t), t._version, t.data_ptr()) for t in list(self.parameters()) + list(self.buffers())), str(device), steps)
        if self._handle is not None and self._key == key:
            return self._handle
        if self._handle is not None:
            _lib.load().vad_head_destroy(self._handle)
            self._handle = None
        lib = _lib.load()
        cfg = self.config
        hc = HeadConfig()
        hc.channels, hc.n_stages = cfg.channels, len(cfg.dims)
        for i, (d, n, ty) in enumerate(zip(cfg.dims, cfg.depths, cfg.mgfn_types)):
            hc.dims[i], hc.depths[i] = d, n
            hc.types[i] = _lib.VAD_HEAD_GLANCE if ty == "gb" else _lib.VAD_HEAD_FOCUS
        hc.dim_head, hc.ff_repe, hc.local_aggr_kernel, hc.k = cfg.dim_head, cfg.ff_repe, cfg.local_aggr_kernel, cfg.k
        hc.mag_ratio, hc.ln_eps = cfg.mag_ratio, 1e-5
        self._blob = self.pack_parameters().to(device).contiguous()
        h = ctypes.c_void_p()
        check(lib.vad_head_create(ctypes.byref(h), ctypes.byref(hc), self._blob.data_ptr(), self._blob.numel() * 4,
                                  device.index if device.index is not None else torch.cuda.current_device()),
              "vad_head_create")
        self._handle, self._key = h, key
        return h

    def _release(self) -> None:
        if self._handle is not None:
            _lib.load().vad_head_destroy(self._handle)
            self._handle = None
        st = getattr(self, "_train", None)
        if st is not None:
            _lib.load().vad_head_train_destroy(st["handle"])
            self._train = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    @property
    def num_launches(self) -> int:
        return int(_lib.load().vad_head_num_launches(self._handle)) if self._handle is not None else 0

    def flops(self, n_seq: int, t: int) -> float:
        return float(_lib.load().vad_head_flops(self._handle, n_seq, t)) if self._handle is not None else 0.0

    # ------------------------------------------------------------------------------ training plumbing
    def _train_entries(self):
        """(tensor, blob shape, permutation blob -> module layout or None) in ``vad_head_train_create``'s order; every
        entry starts on a 64-float boundary.  Conv1d(k=3) weights are tap-major in the blob ([cout][3][cin])."""
        cfg = self.config
        ent = []

        def conv3(conv):
            co, ci, _ = conv.weight.shape
            ent.append((conv.weight, (co, 3, ci), (0, 2, 1)))
            ent.append((conv.bias, (co,), None))

        def conv1(conv, bias=True):
            co, ci, _ = conv.weight.shape
            ent.append((conv.weight, (co, ci), None))
            if bias:
                ent.append((conv.bias, (co,), None))

        def mln(ln):
            ent.append((ln.g, (ln.g.shape[1],), None))
            ent.append((ln.b, (ln.b.shape[1],), None))

        amp = self.backbone.amplifier
        conv3(amp.to_tokens)
        ent.append((amp.to_mag.weight, (cfg.dims[0], 3), None))
        ent.append((amp.to_mag.bias, (cfg.dims[0],), None))
        for stage in self.backbone.layers:
            for blk in stage:
                if isinstance(blk, MGFNIntermediate):
                    mln(blk.layer_norm)
                    conv1(blk.conv)
                    continue
                conv3(blk.scc)
                att = blk.attention
                if isinstance(att, GlanceAttention):
                    mln(att.norm)
                    conv1(att.to_qkv, bias=False)
                else:
                    conv1(att.to_v, bias=False)
                    ent.append((att.norm.weight, tuple(att.norm.weight.shape), None))
                    ent.append((att.norm.bias, tuple(att.norm.bias.shape), None))
                    ent.append((att.rel_pos.weight, (att.rel_pos.weight.shape[0], att.rel_pos.weight.shape[2]), None))
                    ent.append((att.rel_pos.bias, tuple(att.rel_pos.bias.shape), None))
                conv1(att.to_out)
                mln(blk.ffn.layer_norm)
                conv1(blk.ffn.in_conv)
                conv1(blk.ffn.out_conv)
        ent.append((self.layer_norm.weight, tuple(self.layer_norm.weight.shape), None))
        ent.append((self.layer_norm.bias, tuple(self.layer_norm.bias.shape), None))
        ent.append((self.fc.weight, (self.fc.weight.shape[1],), None))
        ent.append((self.fc.bias, (1,), None))
        return ent

    def _bn_modules(self):
        return [blk.attention.norm for stage in self.backbone.layers for blk in stage
                if not isinstance(blk, MGFNIntermediate) and isinstance(blk.attention, FocusAttention)]

    @staticmethod
    def _blob_view(flat: torch.Tensor, off: int, shape, perm, like: torch.Tensor) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        v = flat[off:off + n].view(*shape)
        if perm is not None:
            v = v.permute(*perm)
        return v.reshape(like.shape) if perm is None else v

    def _train_native(self, device: torch.device):
        """Handle + flat blobs for training; parameters / BatchNorm buffers are (re)pointed into the blobs whenever they do
        not alias them any more (first call, after ``.to()``, after ``load_state_dict`` into fresh tensors)."""
        lib = _lib.load()
        st = getattr(self, "_train", None)
        entries = self._train_entries()
        if st is not None and st["device"] == device:
            lo, hi = st["flat"].data_ptr(), st["flat"].data_ptr() + st["flat"].numel() * 4
            if all(lo <= t.data_ptr() < hi for t, _, _ in entries):
                return st
        if st is not None:
            lib.vad_head_train_destroy(st["handle"])
            self._train = None
        cfg = self.config
        hc = HeadConfig()
        hc.channels, hc.n_stages = cfg.channels, len(cfg.dims)
        for i, (d, n, ty) in enumerate(zip(cfg.dims, cfg.depths, cfg.mgfn_types)):
            hc.dims[i], hc.depths[i] = d, n
            hc.types[i] = _lib.VAD_HEAD_GLANCE if ty == "gb" else _lib.VAD_HEAD_FOCUS
        hc.dim_head, hc.ff_repe, hc.local_aggr_kernel, hc.k = cfg.dim_head, cfg.ff_repe, cfg.local_aggr_kernel, cfg.k
        hc.mag_ratio, hc.ln_eps = cfg.mag_ratio, 1e-5
        h = ctypes.c_void_p()
        check(lib.vad_head_train_create(ctypes.byref(h), ctypes.byref(hc), device.index if device.index is not None else torch.cuda.current_device()),
              "vad_head_train_create")
        n_floats = int(lib.vad_head_train_param_floats(h))
        flat = torch.zeros(n_floats, dtype=torch.float32, device=device)
        grad = torch.zeros(n_floats, dtype=torch.float32, device=device)
        off = 0
        grad_views = []
        with torch.no_grad():
            for t, shape, perm in entries:
                n = t.numel()
                view = self._blob_view(flat, off, shape, perm, t)
                view.copy_(t.detach().to(device))
                t.data = view                                   # the module tensor now lives in the blob
                grad_views.append(self._blob_view(grad, off, shape, perm, t))
                off += (n + 63) // 64 * 64
            if off != n_floats:
                raise RuntimeError(f"training layout mismatch: python packs {off} floats, the library expects {n_floats}")
            bns = self._bn_modules()
            bn_flat = torch.zeros(max(1, int(lib.vad_head_train_bn_floats(h))), dtype=torch.float32, device=device)
            boff = 0
            for bn in bns:
                d = bn.num_features
                for buf in (bn.running_mean, bn.running_var):
                    view = bn_flat[boff:boff + d]
                    view.copy_(buf.detach().to(device))
                    buf.data = view
                    boff += d
        self._train = {"handle": h, "device": device, "flat": flat, "grad": grad, "grad_views": grad_views, "bn": bn_flat,
                       "params": [t for t, _, _ in entries], "ws": None, "steps": 0}
        return self._train

    def _forward_train(self, video: torch.Tensor, abnormal_labels: torch.Tensor, normal_labels: torch.Tensor,
                       select_mask: Optional[torch.Tensor] = None) -> MGFNVideoAnomalyDetectionOutput:
        if abnormal_labels is None or normal_labels is None:
            raise ValueError("train mode needs abnormal_labels and normal_labels (src/runner.py:34-36)")
        dev = video.device
        st = self._train_native(dev)
        lib = _lib.load()
        video = video.float().contiguous()
        bs, ncrops, T, _ = video.shape
        if bs % 2:
            raise ValueError("a training batch is a normal half followed by an abnormal half (src/runner.py:31); enable drop_last")
        half, k = bs // 2, self.config.k
        labels = torch.cat([normal_labels.reshape(-1), abnormal_labels.reshape(-1)]).float().to(dev).contiguous()
        if labels.numel() != bs:
            raise ValueError("need one label per video")
        if select_mask is None and self.config.dropout_rate > 0.0:
            # the reference draws the abnormal half's mask first, then the normal half's (modeling_mgfn.py:358-366)
            ones = torch.ones(half, T, device=dev)
            mask_a = torch.nn.functional.dropout(ones, self.config.dropout_rate, training=True)
            mask_n = torch.nn.functional.dropout(ones, self.config.dropout_rate, training=True)
            select_mask = torch.cat([mask_n, mask_a])
        if select_mask is not None:
            select_mask = select_mask.float().to(dev).contiguous()
            if tuple(select_mask.shape) != (bs, T):
                raise ValueError(f"select_mask must be [bs, T] = {(bs, T)} (normal rows then abnormal rows)")
        need = ctypes.c_uint64()
        check(lib.vad_head_train_workspace_bytes(st["handle"], bs, ncrops, T, ctypes.byref(need)), "vad_head_train_workspace_bytes")
        if st["ws"] is None or st["ws"].numel() < need.value + 1024:
            st["ws"] = None
            st["ws"] = torch.empty(int(need.value) + 1024, dtype=torch.uint8, device=dev)
        ws_ptr = (st["ws"].data_ptr() + 1023) // 1024 * 1024
        terms = torch.empty(7, dtype=torch.float32, device=dev)
        scores = torch.empty(bs, T, dtype=torch.float32, device=dev)
        idx = torch.empty(bs, k, dtype=torch.int32, device=dev)
        model = self
        lw = getattr(self, "loss_weights", None)  # None: the reference's constants; else (smooth, sparsity, alpha, margin)
        loss_cfg = (ctypes.c_float * 4)(*[float(v) for v in lw]) if lw is not None else None

        class _Step(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *params):
                check(lib.vad_head_train_step(st["handle"], st["flat"].data_ptr(), st["grad"].data_ptr(), st["bn"].data_ptr(),
                                              video.data_ptr(), bs, ncrops, T, labels.data_ptr(),
                                              select_mask.data_ptr() if select_mask is not None else None, loss_cfg, ws_ptr, int(need.value),
                                              terms.data_ptr(), scores.data_ptr(), idx.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream), "vad_head_train_step")
                st["steps"] += 1
                model._key = None  # the eval-mode blob (BatchNorm folded into to_v) is stale from here on
                return terms[0].clone()

            @staticmethod
            def backward(ctx, gout):
                st["grad"].mul_(gout)  # one kernel; d loss / d params were computed by the native step
                return tuple(st["grad_views"])

        loss = _Step.apply(*st["params"])
        vid = torch.gather(scores, 1, idx.long()).mean(dim=1, keepdim=True)
        self._last_idx = idx
        return MGFNVideoAnomalyDetectionOutput(loss=loss, abnormal_scores=vid[half:], normal_scores=vid[:half], a_feat_magnitude=None,
                                               n_feat_magnitude=None, scores=scores.view(bs, T, 1), loss_terms=terms)

    @property
    def train_launches(self) -> int:
        st = getattr(self, "_train", None)
        return int(_lib.load().vad_head_train_num_launches(st["handle"])) if st is not None else 0

    # ------------------------------------------------------------------------------ forward
    def forward(self, video: torch.Tensor, abnormal_labels: Optional[torch.Tensor] = None,
                normal_labels: Optional[torch.Tensor] = None, select_mask: Optional[torch.Tensor] = None) -> MGFNVideoAnomalyDetectionOutput:
        if not video.is_cuda:
            raise RuntimeError("MGFN scores are computed by sm_100a kernels only; move the input (and the model) to a CUDA "
                               "device. There is no CPU fallback.")
        if video.dim() != 4 or video.shape[-1] != self.config.channels + 1:
            raise ValueError(f"video must be [bs, ncrops, T, {self.config.channels + 1}], got {tuple(video.shape)}")
        if self.training:
            return self._forward_train(video, abnormal_labels, normal_labels, select_mask)
        with torch.no_grad():
            return self._forward_eval(video, abnormal_labels, normal_labels)

    def _forward_eval(self, video: torch.Tensor, abnormal_labels: Optional[torch.Tensor] = None,
                      normal_labels: Optional[torch.Tensor] = None) -> MGFNVideoAnomalyDetectionOutput:
        if not video.is_cuda:
            raise RuntimeError("MGFN scores are computed by sm_100a kernels only; move the input (and the model) to a CUDA "
                               "device. There is no CPU fallback.")
        dev = video.device
        lib = _lib.load()
        h = self._native(dev)
        video = video.float().contiguous()
        bs, ncrops, T, _ = video.shape
        n_seq, dl, k = bs * ncrops, self.config.dims[-1], self.config.k
        need = ctypes.c_uint64()
        check(lib.vad_head_workspace_bytes(h, n_seq, T, ctypes.byref(need)), "vad_head_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value + 1024 or self._ws.device != dev:
            self._ws = torch.empty(int(need.value) + 1024, dtype=torch.uint8, device=dev)
        ws_ptr = (self._ws.data_ptr() + 1023) // 1024 * 1024  # the C ABI wants a 1024-byte aligned workspace
        stream = torch.cuda.current_stream(dev).cuda_stream
        xln = torch.empty(n_seq, T, dl, dtype=torch.float32, device=dev)
        score_tok = torch.empty(n_seq, T, dtype=torch.float32, device=dev)
        fmag_tok = torch.empty(n_seq, T, dtype=torch.float32, device=dev)
        check(lib.vad_head_forward(h, video.data_ptr(), bs, ncrops, T, ws_ptr, int(need.value), xln.data_ptr(),
                                   score_tok.data_ptr(), fmag_tok.data_ptr(), stream), "vad_head_forward")
        scores = torch.empty(bs, T, dtype=torch.float32, device=dev)
        vid_score = torch.empty(bs, dtype=torch.float32, device=dev)
        idx = torch.empty(bs, k, dtype=torch.int32, device=dev)

        def select(off: int, n: int) -> torch.Tensor:
            sel = torch.empty(ncrops, n, k, dl, dtype=torch.float32, device=dev)
            check(lib.vad_head_select(h, xln.data_ptr(), score_tok.data_ptr(), fmag_tok.data_ptr(), bs, ncrops, T, off, n,
                                      scores.data_ptr(), vid_score.data_ptr(), idx.data_ptr(), sel.data_ptr(), stream),
                  "vad_head_select")
            return sel.view(ncrops * n, k, dl)

        split = self.force_split
        if split:
            if bs % 2:
                raise ValueError("force_split needs an even batch: normal half then abnormal half (modeling_mgfn.py:325-334)")
            half = bs // 2
            n_feat, a_feat = select(0, half), select(half, half)
            normal_scores, abnormal_scores = vid_score[:half].view(half, 1), vid_score[half:].view(half, 1)
        else:
            n_feat = a_feat = select(0, bs)
            normal_scores = abnormal_scores = vid_score.view(bs, 1)
        loss = terms = None
        if abnormal_labels is not None and normal_labels is not None:
            if not split:
                raise NotImplementedError("the loss is defined for a split batch (normal half + abnormal half): set "
                                          "model.force_split = True, as the reference's training batches are built "
                                          "(src/runner.py:29-37)")
            labels = torch.cat([normal_labels.reshape(-1), abnormal_labels.reshape(-1)]).float().to(dev).contiguous()
            if labels.numel() != bs:
                raise ValueError("need one label per video")
            scratch = torch.empty(2 * ncrops * (bs // 2) * k, dtype=torch.float32, device=dev)
            terms = torch.empty(7, dtype=torch.float32, device=dev)
            check(lib.vad_head_loss(h, scores.data_ptr(), vid_score.data_ptr(), labels.data_ptr(), n_feat.data_ptr(),
                                    a_feat.data_ptr(), bs // 2, ncrops, T, scratch.data_ptr(), terms.data_ptr(), stream),
                  "vad_head_loss")
            loss = terms[0]
        self._last_idx = idx
        return MGFNVideoAnomalyDetectionOutput(loss=loss, abnormal_scores=abnormal_scores, normal_scores=normal_scores,
                                               a_feat_magnitude=a_feat, n_feat_magnitude=n_feat,
                                               scores=scores.view(bs, T, 1), loss_terms=terms)


class NativeAdam:
    """``torch.optim.Adam(model.parameters(), lr, weight_decay)`` (configure_optimizers, src/runner.py:53-59) as ONE fused native
    kernel over the model's flat parameter / gradient blobs.  With ``torch.distributed`` initialised the gradient blob is
    summed over the ranks first (one NCCL all-reduce of the whole head, 115 MB) and averaged inside the kernel."""

    def __init__(self, model: MGFNForVideoAnomalyDetection, lr: float = 1e-3, weight_decay: float = 5e-4,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        self.model, self.lr, self.weight_decay, self.betas, self.eps = model, lr, weight_decay, betas, eps
        self.step_count = 0
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self.last_allreduce_ms: Optional[float] = None

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.model.parameters():
            p.grad = None  # the native step overwrites the gradient blob; nothing accumulates

    def step(self) -> None:
        st = getattr(self.model, "_train", None)
        if st is None or st["steps"] == 0:
            raise RuntimeError("NativeAdam.step() follows a train-mode forward (the native step that fills the gradient blob)")
        flat, grad = st["flat"], st["grad"]
        if self._m is None or self._m.data_ptr() == 0 or self._m.numel() != flat.numel() or self._m.device != flat.device:
            self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        import torch.distributed as dist

        scale = 1.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(grad)  # sum over the data-parallel ranks
            scale = 1.0 / dist.get_world_size()
        self.step_count += 1
        check(_lib.load().vad_adam_step(flat.data_ptr(), grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), flat.numel(), self.lr,
                                        self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, scale,
                                        torch.cuda.current_stream(flat.device).cuda_stream), "vad_adam_step")
        self.model._key = None


__all__ = ["MGFNConfig", "MGFNForVideoAnomalyDetection", "MGFNVideoAnomalyDetectionOutput", "MGFNModel", "NativeAdam"]
