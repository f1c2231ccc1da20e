"""CPU restatement of the MGFN scoring head and its losses -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Follows, in fp32 torch functional ops on a plain state_dict,
    src/models/mgfn/modeling_mgfn.py:36-47    MGFNLayerNorm (divides by std + eps)
    src/models/mgfn/modeling_mgfn.py:50-94    feed-forward, feature amplifier
    src/models/mgfn/modeling_mgfn.py:97-212   Glance / Focus attention and blocks
    src/models/mgfn/modeling_mgfn.py:215-283  intermediate, stage stack
    src/models/mgfn/modeling_mgfn.py:302-427  magnitude selection, score prediction, forward (eval mode)
    src/loss/base.py:7-48, src/loss/mgfn.py:7-47  smoothness, sparsity, contrastive, MGFN loss
Pinned by tests/golden/mgfn.npz (outputs of the unmodified reference, see oracle/make_golden.py) and, when
/root/reference is mounted, by a direct comparison with the live reference (tests/test_oracle.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

DEFAULT = dict(dims=(64, 128, 1024), depths=(3, 3, 2), mgfn_types=("gb", "fb", "fb"), channels=2048, ff_repe=4, dim_head=64,
               local_aggr_kernel=5, mag_ratio=0.1, k=3)


def seeded_state_dict(seed: int = 0, cfg: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights with the reference's parameter names and shapes: fan-in scaled
    convolutions, non-trivial LayerNorm affine terms and BatchNorm running statistics."""
    c = dict(DEFAULT, **(cfg or {}))
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, bias=True, groups=1):
        fan = cin // groups * k
        sd[name + ".weight"] = torch.randn(cout, cin // groups, k, generator=g) * (1.0 / fan) ** 0.5
        if bias:
            sd[name + ".bias"] = 0.1 * torch.randn(cout, generator=g)

    def mln(name, dim):
        sd[name + ".g"] = 1.0 + 0.1 * torch.randn(1, dim, 1, generator=g)
        sd[name + ".b"] = 0.1 * torch.randn(1, dim, 1, generator=g)

    d0 = c["dims"][0]
    conv("backbone.amplifier.to_tokens", d0, c["channels"], 3)
    conv("backbone.amplifier.to_mag", d0, 1, 3)
    for si, (dim, depth, ty) in enumerate(zip(c["dims"], c["depths"], c["mgfn_types"])):
        heads = dim // c["dim_head"]
        inner = heads * c["dim_head"]
        for bi in range(depth):
            p = f"backbone.layers.{si}.{bi}"
            conv(p + ".scc", dim, dim, 3)
            if ty == "gb":
                mln(p + ".attention.norm", dim)
                conv(p + ".attention.to_qkv", 3 * inner, dim, 1, bias=False)
            else:
                sd[p + ".attention.norm.weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)
                sd[p + ".attention.norm.bias"] = 0.1 * torch.randn(dim, generator=g)
                sd[p + ".attention.norm.running_mean"] = 0.1 * torch.randn(dim, generator=g)
                sd[p + ".attention.norm.running_var"] = 0.5 + torch.rand(dim, generator=g)
                sd[p + ".attention.norm.num_batches_tracked"] = torch.tensor(0)
                conv(p + ".attention.to_v", inner, dim, 1, bias=False)
                conv(p + ".attention.rel_pos", heads, heads, c["local_aggr_kernel"], groups=heads)
            conv(p + ".attention.to_out", dim, inner, 1)
            mln(p + ".ffn.layer_norm", dim)
            conv(p + ".ffn.in_conv", dim * c["ff_repe"], dim, 1)
            conv(p + ".ffn.out_conv", dim, dim * c["ff_repe"], 1)
        if si != len(c["dims"]) - 1:
            p = f"backbone.layers.{si}.{depth}"
            mln(p + ".layer_norm", dim)
            conv(p + ".conv", c["dims"][si + 1], dim, 1)
    dl = c["dims"][-1]
    sd["layer_norm.weight"] = 1.0 + 0.1 * torch.randn(dl, generator=g)
    sd["layer_norm.bias"] = 0.1 * torch.randn(dl, generator=g)
    sd["fc.weight"] = torch.randn(1, dl, generator=g) * (1.0 / dl) ** 0.5
    sd["fc.bias"] = 0.1 * torch.randn(1, generator=g)
    return sd


def synthetic_video(seed: int, bs: int, ncrops: int, t: int, channels: int = 2048) -> torch.Tensor:
    """[bs, ncrops, t, channels + 1]: non-negative snippet features (post-ReLU pooled activations look like
    this) with their L2 norm appended (FeatureDataset.add_magnitude, src/dataset.py:121-124)."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(bs, ncrops, t, channels, generator=g).abs() * 0.5
    # a few high-magnitude ("abnormal looking") snippets per video so that top-k selection is not a tie-break
    boost = 1.0 + 2.0 * (torch.rand(bs, 1, t, 1, generator=g) > 0.85).float() * torch.rand(bs, 1, t, 1, generator=g)
    f = f * boost
    return torch.cat([f, f.norm(dim=3, keepdim=True)], dim=3)


class _TruncTF32(torch.autograd.Function):
    """fp32 -> TF32 the way the tensor core reads an fp32 operand: the low 13 mantissa bits are dropped.  Straight-through
    gradient.  Used by ``forward(emulate_tf32=True)`` to predict, on the CPU, what kind::tf32 contractions do to the outputs."""

    @staticmethod
    def forward(ctx, x):
        return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


_TF32 = [False]  # set by forward(emulate_tf32=True) for the duration of the call


def _conv(x, w, b=None, padding=0):
    """The Conv1d layers the product runs as tensor-core GEMMs (everything but the 1-channel to_mag and the depth-wise
    rel_pos, which are fp32 SIMT kernels there)."""
    if _TF32[0]:
        x, w = _TruncTF32.apply(x), _TruncTF32.apply(w)
    return F.conv1d(x, w, b, padding=padding)


def _mgfn_ln(x, g, b, eps=1e-5):  # modeling_mgfn.py:43-46
    std = torch.var(x, dim=1, unbiased=False, keepdim=True).sqrt()
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) / (std + eps) * g + b


def _ffn(x, sd, p):  # modeling_mgfn.py:58-64 (dropout 0)
    x = _mgfn_ln(x, sd[p + ".layer_norm.g"], sd[p + ".layer_norm.b"])
    x = F.gelu(_conv(x, sd[p + ".in_conv.weight"], sd[p + ".in_conv.bias"]))
    return _conv(x, sd[p + ".out_conv.weight"], sd[p + ".out_conv.bias"])


def _glance_attention(x, sd, p, heads, dim_head):  # modeling_mgfn.py:109-127
    x = _mgfn_ln(x, sd[p + ".norm.g"], sd[p + ".norm.b"])
    b, _, n = x.shape
    q, k, v = _conv(x, sd[p + ".to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, dim_head, n).permute(0, 1, 3, 2) for t in (q, k, v))
    q = q * dim_head ** -0.5
    attn = torch.einsum("bhid,bhjd->bhij", q, k).softmax(dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, heads * dim_head, n)
    return _conv(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def _focus_attention(x, sd, p, heads, training=False):  # modeling_mgfn.py:178-186
    # training: BatchNorm1d normalises with the batch statistics (the running buffers are updated on copies: the oracle
    # never mutates the state dict it is given)
    x = F.batch_norm(x, sd[p + ".norm.running_mean"].clone() if training else sd[p + ".norm.running_mean"],
                     sd[p + ".norm.running_var"].clone() if training else sd[p + ".norm.running_var"], sd[p + ".norm.weight"],
                     sd[p + ".norm.bias"], training=training, eps=1e-5)
    b, _, n = x.shape
    v = _conv(x, sd[p + ".to_v.weight"])
    c = v.shape[1] // heads
    v = v.reshape(b, c, heads, n).reshape(b * c, heads, n)           # "b (c h) t -> (b c) h t"
    k = sd[p + ".rel_pos.weight"].shape[-1]
    out = F.conv1d(v, sd[p + ".rel_pos.weight"], sd[p + ".rel_pos.bias"], padding=k // 2, groups=heads)
    out = out.reshape(b, c, heads, n).reshape(b, c * heads, n)       # "(b c) h t -> b (c h) t"
    return _conv(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def backbone(video: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: Optional[dict] = None, training: bool = False) -> torch.Tensor:
    """MGFNModel.forward (modeling_mgfn.py:67-94,241-283): [bs, ncrops, t, C+1] -> [bs*ncrops, d_last, t]."""
    c = dict(DEFAULT, **(cfg or {}))
    bs, ncrops, t, ch = video.shape
    x = video.reshape(bs * ncrops, t, ch).permute(0, 2, 1)
    x_f, x_m = x[:, :c["channels"], :], x[:, c["channels"]:, :]
    x = _conv(x_f, sd["backbone.amplifier.to_tokens.weight"], sd["backbone.amplifier.to_tokens.bias"], padding=1) + \
        c["mag_ratio"] * F.conv1d(x_m, sd["backbone.amplifier.to_mag.weight"], sd["backbone.amplifier.to_mag.bias"], padding=1)
    for si, (dim, depth, ty) in enumerate(zip(c["dims"], c["depths"], c["mgfn_types"])):
        heads = dim // c["dim_head"]
        for bi in range(depth):
            p = f"backbone.layers.{si}.{bi}"
            x = _conv(x, sd[p + ".scc.weight"], sd[p + ".scc.bias"], padding=1) + x
            if ty == "gb":
                x = _glance_attention(x, sd, p + ".attention", heads, c["dim_head"]) + x
            else:
                x = _focus_attention(x, sd, p + ".attention", heads, training) + x
            x = _ffn(x, sd, p + ".ffn") + x
        if si != len(c["dims"]) - 1:
            p = f"backbone.layers.{si}.{depth}"
            x = _conv(_mgfn_ln(x, sd[p + ".layer_norm.g"], sd[p + ".layer_norm.b"]), sd[p + ".conv.weight"], sd[p + ".conv.bias"])
    return x


def contrastive(o1, o2, label, margin=200.0):  # src/loss/base.py:36-48
    d = F.pairwise_distance(o1, o2, keepdim=True)
    return torch.mean((1 - label) * d.pow(2) + label * torch.clamp(margin - d, min=0.0).pow(2))


def forward(video: torch.Tensor, sd: Dict[str, torch.Tensor], cfg: Optional[dict] = None, split: bool = False,
            normal_labels: Optional[torch.Tensor] = None, abnormal_labels: Optional[torch.Tensor] = None,
            training: bool = False, select_masks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
            emulate_tf32: bool = False, loss_weights: Sequence[float] = (8e-4, 8e-3, 0.001, 200.0)) -> dict:
    """MGFNForVideoAnomalyDetection.forward, modeling_mgfn.py:302-427.  Default: eval mode (dropout inactive).
    ``training=True`` is the train-mode forward with every dropout probability at 0 (``MGFNConfig(dropout_rate=0.0)``;
    the shipped 0.7 on the selection mask makes the step stochastic and is not restated): BatchNorm1d of the Focus blocks
    uses batch statistics, inputs are always split into the normal and the abnormal half (modeling_mgfn.py:320-331).
    Everything is differentiable torch, so ``torch.autograd`` over this function is the gradient oracle for the training
    step (src/runner.py:29-39) -- pinned against the live reference's ``loss.backward()`` by tests/golden/mgfn_train.npz.
    ``select_masks = (mask_normal, mask_abnormal)``, each [n_size, t] with entries 0 or 1 / (1 - p): the dropout masks the
    reference multiplies onto the magnitudes before the top-k (modeling_mgfn.py:341-345), injected so that the stochastic
    step is reproducible (the reference draws the abnormal mask first)."""
    c = dict(DEFAULT, **(cfg or {}))
    bs, ncrops = video.shape[:2]
    split = split or training
    # emulate_tf32: the operands of every tensor-core contraction are truncated to TF32 first (what tcgen05 kind::tf32 does
    # with fp32 operands) -- the product's numerics predicted on the CPU.  The MGFN loss squares differences of L1 norms of
    # 1024-d features (~800 each), so 1e-3 noise on the features moves the contrastive terms' gradients by ~10 %.
    _TF32[0] = bool(emulate_tf32)
    try:
        x = backbone(video.float(), sd, c, training).permute(0, 2, 1)
    finally:
        _TF32[0] = False
    x = F.layer_norm(x, (x.shape[-1],), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
    scores_tok = torch.sigmoid(F.linear(x, sd["fc.weight"], sd["fc.bias"]))
    _, t, f = x.shape
    fm = x.norm(p=2, dim=2).view(bs, ncrops, -1).mean(dim=1)
    scores = scores_tok.view(bs, ncrops, -1).mean(dim=1).unsqueeze(2)
    if split:
        h = bs // 2
        nf, af, ns, as_, nm, am = x[:h * ncrops], x[h * ncrops:], scores[:h], scores[h:], fm[:h], fm[h:]
    else:
        nf = af = x
        ns = as_ = scores
        nm = am = fm
    n_size = nm.shape[0]

    def select(mag, feats, mask=None):
        idx = torch.topk(mag if mask is None else mag * mask, c["k"], dim=1)[1]
        idx_feat = idx.unsqueeze(2).expand(-1, -1, f)
        feats = feats.view(n_size, ncrops, t, f).permute(1, 0, 2, 3)
        return idx, torch.cat([torch.gather(fe, 1, idx_feat) for fe in feats])

    def predict(idx, sc):
        return torch.mean(torch.gather(sc, 1, idx.unsqueeze(2).expand(-1, -1, sc.shape[2])), dim=1)

    idx_a, a_feat = select(am, af, None if select_masks is None else select_masks[1])
    idx_n, n_feat = select(nm, nf, None if select_masks is None else select_masks[0])
    out = dict(scores=scores, abnormal_scores=predict(idx_a, as_), normal_scores=predict(idx_n, ns), a_feat_magnitude=a_feat,
               n_feat_magnitude=n_feat, idx_abnormal=idx_a, idx_normal=idx_n, xln=x, scores_tok=scores_tok.squeeze(-1), loss=None)
    if normal_labels is not None and abnormal_labels is not None:
        w_smooth, w_sparse, alpha, margin = loss_weights  # the reference's constants unless a test overrides them
        smooth = w_smooth * torch.sum((scores[:, 1:, :] - scores[:, :-1, :]) ** 2)
        sparsity = w_sparse * torch.mean(torch.norm(scores[: bs // 2].reshape(-1), dim=0))
        labels = torch.cat((normal_labels, abnormal_labels), 0)
        sc = torch.cat((out["normal_scores"], out["abnormal_scores"]), 0).squeeze()
        bce = F.binary_cross_entropy(sc, labels)
        a1, n1 = a_feat.norm(p=1, dim=2), n_feat.norm(p=1, dim=2)
        sep = int(len(n_feat) / 2)
        con = contrastive(a1, n1, 1, margin)
        con_n = contrastive(n1[sep:], n1[:sep], 0)
        con_a = contrastive(a1[sep:], a1[:sep], 0)
        mg = bce + alpha * (alpha * con + con_a + con_n)
        out["loss"] = mg + smooth + sparsity
        out["loss_terms"] = torch.stack([out["loss"], smooth, sparsity, bce, con, con_n, con_a])
    return out
