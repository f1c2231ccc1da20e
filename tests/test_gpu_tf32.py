"""-m gpu: the TF32 precision mode (vad_tf32_*: fp32 activations / weights, tcgen05 kind::tf32 MMAs).

Tolerance (BASELINE.json north_star: "1e-3 for the TF32 mode"): as for bf16, relative error is measured against the
scale of the tensor -- max|a-b| <= 1e-3 * max|b| and ||a-b|| <= 1e-3 * ||b|| -- with cosine >= 0.99999.
Reference = the reference's own fp32 PyTorch path (golden features made by the unmodified reference, tests/golden) and
torch fp32 conv3d / max_pool3d on the CPU for the single-op cases.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import i3res50 as O

pytestmark = pytest.mark.gpu

TOL = 1e-3


def close_tf32(got: torch.Tensor, ref: torch.Tensor, tol: float = TOL):
    got, ref = got.double(), ref.double()
    assert got.shape == ref.shape
    max_norm = ((got - ref).abs().max() / ref.abs().max()).item()
    rel_l2 = ((got - ref).norm() / ref.norm()).item()
    print(f"max-normalised err {max_norm:.2e}, rel L2 {rel_l2:.2e}")
    assert max_norm <= tol and rel_l2 <= tol, (max_norm, rel_l2)


CONV_CASES = [
    # name, cin, cout, k, s, p, B, T, H, W, res, relu
    ("stem 5x7x7/2 4->64", 4, 64, (5, 7, 7), (2, 2, 2), (2, 3, 3), 1, 8, 32, 32, False, True),
    ("1x1 64->256 +res", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 9, 9, True, True),
    ("t3 256->64", 256, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 7, False, True),
    ("s3x3 stride 2 128->128", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 14, 14, False, True),
    ("1x1 stride 2 256->512 ds, linear", 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 1, 2, 14, 14, False, False),
    ("3x3x3 24->40 (cin % 32 != 0, N tail)", 24, 40, (3, 3, 3), (1, 1, 1), (1, 1, 1), 2, 3, 6, 5, False, True),
    ("1x1 2048->512 long K", 2048, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 7, 7, False, True),
    # CTA-pair form (Cout >= 128 with a TMA activation tile): odd number of m tiles (the odd CTA of the last pair has no rows),
    # N tail inside a 128-wide pair tile, more items than CTA pairs, residual on a two-n-tile layer
    ("1x1 256->192 odd m tiles, N tail", 256, 192, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 1, 13, 13, False, True),
    ("3x3x3 64->128 many pair items", 64, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), 4, 8, 28, 28, False, True),
    ("1x1 128->512 +res, 490 m tiles", 512, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 4, 4, 28, 28, True, True),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_tf32_matches_fp32_conv3d(cuda_device, case, monkeypatch):
    """Every case runs three times: the default (activation tile by TMA im2col where Cin % 32 == 0; 256 x BN tiles on CTA pairs,
    conv_pair_kernel<.., F32>, where also Cout >= 128), the single-CTA TMA kernel (VAD_TF32_NO_PAIR=1) and the cp.async gather
    (VAD_TF32_GATHER=1); same k order per accumulator, so all three must agree bit for bit."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    _, cin, cout, k, s, p, B, T, H, W, res, relu = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, T, H, W, generator=g)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    scale, shift = 0.5 + torch.rand(cout, generator=g), 0.2 * torch.randn(cout, generator=g)
    ref = F.conv3d(x, w, None, s, p) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if res:
        ref = ref + x[:, :cout]
    if relu:
        ref = F.relu(ref)
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, tf32=True)
    ops = []
    if res:  # residual = the conv input, copied into slot 1 by an identity max-pool
        ops.append(eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=(1, 1, 1), stride=(1, 1, 1)))
    ops.append(eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=2, res=1 if res else -1, cin=cin, cout=cout, kernel=k, stride=s, pad=p,
                      flags=lib.VAD_FLAG_RELU if relu else 0, w_off=w_off, scale_off=s_off, shift_off=b_off))
    outs = []
    for gather, no_pair in (("0", "0"), ("0", "1"), ("1", "0")):
        monkeypatch.setenv("VAD_TF32_GATHER", gather)
        monkeypatch.setenv("VAD_TF32_NO_PAIR", no_pair)
        plan = eng.Tf32Plan(ops, pk.blob(), 3, cuda_device, in_channels=cin)
        plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
        torch.cuda.synchronize()
        outs.append(plan.slot_tensor(2).cpu().permute(0, 4, 1, 2, 3).clone())
    close_tf32(outs[0], ref)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("cin,cout,k,s,T,H,W", [(64, 64, (3, 3, 3), (2, 2, 2), 6, 15, 20), (64, 128, (7, 7, 7), (2, 2, 2), 8, 16, 16),
                                                (48, 64, (3, 3, 3), (1, 1, 1), 4, 7, 7), (4, 64, (7, 7, 7), (2, 2, 2), 8, 32, 32)])
def test_conv_tf32_same_padding(cuda_device, cin, cout, k, s, T, H, W):
    """VAD_FLAG_CONV_SAME (asymmetric TF-style padding of the Inception port) through both activation producers."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from oracle.inception import _same_pad

    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, cin, T, H, W, generator=g)
    w = torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5
    scale, shift = 0.5 + torch.rand(cout, generator=g), 0.2 * torch.randn(cout, generator=g)
    ref = F.relu(F.conv3d(_same_pad(x, k, s), w, None, s) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, tf32=True)
    ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=cin, cout=cout, kernel=k, stride=s, flags=lib.VAD_FLAG_RELU | lib.VAD_FLAG_CONV_SAME,
                  w_off=w_off, scale_off=s_off, shift_off=b_off)]
    plan = eng.Tf32Plan(ops, pk.blob(), 2, cuda_device, in_channels=cin)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    close_tf32(plan.slot_tensor(1).cpu().permute(0, 4, 1, 2, 3), ref)


@pytest.mark.parametrize("k,s,same", [((5, 7, 7), (2, 2, 2), False), ((7, 7, 7), (2, 2, 2), True), ((3, 5, 8), (1, 1, 2), False)])
def test_stem_folded_window_gather(cuda_device, k, s, same):
    """VAD_FLAG_STEM_FOLD_W in the TF32 mode: the RGB stem contracts (dt, dh) taps x 8-pixel x 4-channel windows (128
    contiguous bytes per tile row and k-block) instead of gathering 16 bytes per tap."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from oracle.inception import _same_pad

    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 9, 30, 34, generator=g)
    w = torch.randn(64, 3, *k, generator=g) * (2.0 / (3 * k[0] * k[1] * k[2])) ** 0.5
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    pad = (0, 0, 0) if same else (k[0] // 2, k[1] // 2, k[2] // 2)
    xin = _same_pad(x, k, s) if same else x
    ref = F.relu(F.conv3d(xin, w, None, s, pad) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True, tf32=True)
    flags = lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W | (lib.VAD_FLAG_CONV_SAME if same else 0)
    ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=k, stride=s, pad=pad, flags=flags, w_off=w_off,
                  scale_off=s_off, shift_off=b_off)]
    plan = eng.Tf32Plan(ops, pk.blob(), 2, cuda_device, in_channels=4)
    plan.forward(eng.ingest_ncthw_tf32(x.to(cuda_device)))
    torch.cuda.synchronize()
    close_tf32(plan.slot_tensor(1).cpu().permute(0, 4, 1, 2, 3), ref)


@pytest.mark.parametrize("k,st,B,T,H,W,dst", [((5, 7, 7), 2, 2, 9, 30, 34, (0, 0)), ((5, 7, 7), 2, 1, 16, 224, 224, (0, 0)),
                                              ((1, 7, 7), 1, 2, 4, 50, 38, (0, 0)), ((3, 5, 8), 2, 1, 6, 33, 70, (32, 128)),
                                              ((5, 7, 7), 2, 3, 8, 16, 16, (0, 0))])
def test_stem_plane_kernel(cuda_device, k, st, B, T, H, W, dst, monkeypatch):
    """VAD_FLAG_STEM_PLANES: the dedicated TF32 stem kernel (stem_tf32_kernel) on the column-parity plane layout --
    sliding 8-pixel windows read by the tensor core out of raw row segments, two CTAs per tile (one per half of the 64 output
    channels), resident fp32 weights -- against fp32 conv3d, incl. ragged tiles (Ho / Wo not multiples of 16 / 8), a
    destination channel slice, and the gather stem on the plain layout as a second opinion."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, 3, T, H, W, generator=g)
    w = torch.randn(64, 3, *k, generator=g) * (2.0 / (3 * k[0] * k[1] * k[2])) ** 0.5
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    s, pad = (st, 2, 2), (k[0] // 2, k[1] // 2, 3)
    ref = F.relu(F.conv3d(x, w, None, s, pad) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    outs = []
    # the CTA-pair form (two tiles per cluster, N = 64 MMAs reading the weight halves of both CTAs), the single-CTA form
    # (VAD_TF32_NO_PAIR=1: two CTAs per tile, N = 32), and the gather stem on the plain layout
    for planes, no_pair in ((True, "0"), (True, "1"), (False, "0")):
        monkeypatch.setenv("VAD_TF32_NO_PAIR", no_pair)
        pk = eng.ParamPacker()
        w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True, tf32=True, planes=planes)
        flags = lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W | (lib.VAD_FLAG_STEM_PLANES if planes else 0)
        ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=k, stride=s, pad=pad, flags=flags, w_off=w_off,
                      scale_off=s_off, shift_off=b_off, dst_c_off=dst[0], dst_c_total=dst[1])]
        plan = eng.Tf32Plan(ops, pk.blob(), 2, cuda_device, in_channels=4)
        xin = eng.ingest_ncthw_tf32(x.to(cuda_device), planes=planes)
        if planes:
            # the plane layout is a pure relayout: padded pixel xp = x + 3 -> plane xp & 1, position xp >> 1
            want = torch.zeros(B, T, H, W + 8, 4)
            want[:, :, :, 3:3 + W, :3] = x.permute(0, 2, 3, 4, 1)
            want = want.view(B, T, H, (W + 8) // 2, 2, 4).permute(0, 1, 2, 4, 3, 5)
            got = xin.cpu()
            assert got.shape == want.shape and (got - want).abs().max() <= 2.0 ** -11 * want.abs().max()   # TF32 rounding only
        if dst[1]:
            plan.configure(B, T, H, W)
            plan.slot_tensor(1).fill_(-7.0)
        plan.forward(xin)
        torch.cuda.synchronize()
        out = plan.slot_tensor(1).cpu()
        if dst[1]:
            assert (out[..., :dst[0]] == -7.0).all() and (out[..., dst[0] + 64:] == -7.0).all(), "neighbouring channel slices untouched"
            out = out[..., dst[0]:dst[0] + 64]
        outs.append(out.permute(0, 4, 1, 2, 3))
    close_tf32(outs[0], ref)
    close_tf32(outs[2], ref)
    close_tf32(outs[0], outs[2], 5e-4)
    assert torch.equal(outs[0], outs[1]), "pair and single-CTA stem kernels contract in the same order"


def test_pools_and_channel_slices_are_exact(cuda_device):
    """fp32 max-pools (torch semantics and the SAME-padding port) are exact; two ops write disjoint channel slices of
    one destination (the Inception concat)."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from oracle.inception import _same_pad

    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 5, 11, 9, generator=g)
    pk = eng.ParamPacker()
    pk.add_conv(torch.zeros(4, 4, 1, 1, 1), torch.ones(4), torch.zeros(4), tf32=True)  # a non-empty blob
    ops = [eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=(2, 3, 3), stride=(2, 2, 2)),
           eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=2, kernel=(3, 3, 3), stride=(1, 1, 1), flags=lib.VAD_FLAG_POOL_SAME, dst_c_off=0,
                  dst_c_total=64),
           eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=2, kernel=(1, 1, 1), stride=(1, 1, 1), dst_c_off=32, dst_c_total=64)]
    plan = eng.Tf32Plan(ops, pk.blob(), 3, cuda_device, in_channels=32)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    a = plan.slot_tensor(1).cpu().permute(0, 4, 1, 2, 3)
    assert torch.equal(a, F.max_pool3d(x, (2, 3, 3), (2, 2, 2)))
    b = plan.slot_tensor(2).cpu().permute(0, 4, 1, 2, 3)
    assert torch.equal(b[:, :32], F.max_pool3d(_same_pad(x, (3, 3, 3), (1, 1, 1)), (3, 3, 3), (1, 1, 1)))
    assert torch.equal(b[:, 32:], x)


@pytest.fixture(scope="module")
def model(cuda_device):
    from anomaly_detection_on_video_b200.i3d import I3Res50

    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0), strict=True)
    m.precision = "tf32"
    return m.eval().to(cuda_device)


@pytest.mark.parametrize("tag", ["small", "odd", "full"])
def test_i3res50_tf32_features_match_reference_golden(model, cuda_device, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "i3res50.npz"))
    shape = tuple(int(v) for v in g[f"{tag}/shape"])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
    y = model(x.to(cuda_device))
    assert tuple(y.shape) == (shape[0], 2048, 1, 1, 1) and y.dtype == torch.float32
    got, want = y.cpu().view(shape[0], -1), torch.from_numpy(g[f"{tag}/features"])
    for b in range(shape[0]):
        close_tf32(got[b], want[b])
        assert F.cosine_similarity(got[b].double(), want[b].double(), dim=0).item() >= 0.99999


def test_tf32_is_tighter_than_bf16_and_modes_switch(model, cuda_device):
    """Same module, both modes: tf32 lands >= 5x closer to the fp32 oracle than bf16, and switching back works."""
    x = torch.randn(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(4)).clamp(-2.0, 2.4444)
    y32, _ = O.forward(x, O.seeded_state_dict(0))
    scale = y32.abs().max()
    e_tf32 = ((model(x.to(cuda_device)).cpu().view(2, -1) - y32.view(2, -1)).abs().max() / scale).item()
    model.precision = "bf16"
    try:
        e_bf16 = ((model(x.to(cuda_device)).cpu().view(2, -1) - y32.view(2, -1)).abs().max() / scale).item()
    finally:
        model.precision = "tf32"
    print(f"max-normalised error: tf32 {e_tf32:.2e}, bf16 {e_bf16:.2e}")
    assert e_tf32 <= TOL and e_tf32 * 5 <= e_bf16
    with pytest.raises(RuntimeError):
        model.forward_stem_layout(torch.zeros(1, 8, 64, 72, 4, dtype=torch.bfloat16, device=cuda_device))


def test_inception_tf32_features_match_fp32_oracle(cuda_device):
    from anomaly_detection_on_video_b200.inception import InceptionI3d
    from oracle import inception as OI

    m = InceptionI3d()
    m.load_state_dict(OI.seeded_state_dict(0), strict=True)
    m.precision = "tf32"
    m.eval().to(cuda_device)
    x = torch.randn(1, 3, 16, 224, 224, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
    got = m(x.to(cuda_device)).cpu().view(1, -1)
    ref = OI.extract_features(x, OI.seeded_state_dict(0)).view(1, -1)
    close_tf32(got[0], ref[0])


def test_extract_clip_features_in_tf32_mode(model, cuda_device):
    """extract_clip_features with a tf32-mode backbone takes the reference's call pattern (fp32 NCTHW crops, one forward per
    crop index) and lands within the TF32 tolerance of the fp32 oracle run on the same preprocessed clips."""
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
    from anomaly_detection_on_video_b200.extract_features import extract_clip_features

    frames = np.random.default_rng(3).integers(0, 256, size=(20, 120, 160, 3), dtype=np.uint8)
    ds = TenCropVideoFrameDataset(frames, device=cuda_device)
    feats = extract_clip_features(ds, model, cuda_device, strict_compat=False, as_numpy=False).cpu()   # (2, 10, 2048)
    assert feats.shape == (2, 10, 2048)
    clip = ds[1].cpu()                                                                                  # (10, 16, 3, 224, 224)
    ref, _ = O.forward(clip[[0, 4, 9]].permute(0, 2, 1, 3, 4).contiguous(), O.seeded_state_dict(0))
    for j, c in enumerate([0, 4, 9]):
        close_tf32(feats[1, c], ref[j].view(-1))
