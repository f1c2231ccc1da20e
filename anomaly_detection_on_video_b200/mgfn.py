"""Drop-in for the reference scoring head ``src/models/mgfn`` (+ the losses of ``src/loss``), inference side.

``MGFNForVideoAnomalyDetection`` keeps the reference's module tree and parameter names
(``backbone.amplifier.to_tokens``, ``backbone.layers.S.B.{scc,attention,ffn}``, ``layer_norm``, ``fc``;
reference src/models/mgfn/modeling_mgfn.py:36-300), so its checkpoints load with ``load_state_dict``,
and the reference's call contract (modeling_mgfn.py:376-427)

    model(video [bs, ncrops, T, 2049] fp32 cuda, abnormal_labels=None, normal_labels=None)
        -> MGFNVideoAnomalyDetectionOutput(loss, abnormal_scores, normal_scores,
                                           a_feat_magnitude, n_feat_magnitude, scores)

The modules only *hold* parameters; the forward runs in ``libvad_b200.so``: every Conv1d is a tcgen05
kind::tf32 GEMM, the rest small fp32 kernels (csrc/head_kernels.cuh).  There is no CPU forward.

Not built: training.  The reference trains this head under Lightning (src/runner.py:29-59) with dropout on
the selection mask (modeling_mgfn.py:342-343); that needs a backward pass through every kernel, which
does not exist here, so ``forward`` refuses to run in ``.train()`` mode instead of silently scoring
without gradients.
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
        key = (tuple((id(t), t._version) for t in list(self.parameters()) + list(self.buffers())), str(device))
        if self._handle is not None and self._key == key:
            return self._handle
        self._release()
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

    # ------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, video: torch.Tensor, abnormal_labels: Optional[torch.Tensor] = None,
                normal_labels: Optional[torch.Tensor] = None) -> MGFNVideoAnomalyDetectionOutput:
        if self.training:
            raise RuntimeError("the native MGFN head is inference-only (no backward kernels): call .eval() first")
        if not video.is_cuda:
            raise RuntimeError("MGFN scores are computed by sm_100a kernels only; move the input (and the model) to a CUDA "
                               "device. There is no CPU fallback.")
        if video.dim() != 4 or video.shape[-1] != self.config.channels + 1:
            raise ValueError(f"video must be [bs, ncrops, T, {self.config.channels + 1}], got {tuple(video.shape)}")
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


__all__ = ["MGFNConfig", "MGFNForVideoAnomalyDetection", "MGFNVideoAnomalyDetectionOutput", "MGFNModel"]
