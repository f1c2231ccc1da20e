"""Pin the CPU oracle (oracle/*.py) to the golden vectors produced by the unmodified reference
(oracle/make_golden.py), and -- when /root/reference is mounted -- to the live reference itself."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import _refload
from oracle import i3res50 as O
from oracle import preprocess as P
from oracle import segment as S


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_frames(seed, n, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def u8_to_f32(u):
    return ((u.astype(np.float32) - np.float32(114.75)) / np.float32(57.375)).astype(np.float32)


# ----------------------------------------------------------------------------- preprocessing
def test_preprocess_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess_small.npz"))
    frames, want = g["frames"], u8_to_f32(g["clips_u8"])
    fpc, rs, cr = int(g["frames_per_clip"]), int(g["resize"]), int(g["crop"])
    assert P.n_clips(len(frames), fpc) == want.shape[0] == 2
    for ci in range(want.shape[0]):
        got = P.clip_tensor(frames, ci, fpc, rs, cr)
        assert got.dtype == np.float32 and got.shape == want[ci].shape
        assert np.array_equal(got, want[ci]), f"clip {ci} differs from the reference"
    # clip 1 holds a single real frame: LoopPad repeats it (src/gtransforms.py:115-132)
    last = P.clip_tensor(frames, 1, fpc, rs, cr)
    assert all(np.array_equal(last[:, 0], last[:, t]) for t in range(1, fpc))


@pytest.mark.parametrize("tag,clips", [("ucf_240x320", (0, 2)), ("down_360x480", (0,)), ("portrait_300x256", (0,))])
def test_preprocess_fullsize_digests(golden_dir, tag, clips):
    g = np.load(os.path.join(golden_dir, "preprocess_digests.npz"))
    seed, n, h, w = eval(str(g[f"{tag}/spec"]))
    frames = synth_frames(seed, n, h, w)
    for ci in clips:
        got = P.clip_tensor(frames, ci)
        assert got.shape == (10, 16, 3, 224, 224)
        assert sha(got) == str(g[f"{tag}/clip{ci}"])


def test_preprocess_geometry_ucf():
    assert P.resized_size(240, 320, 256) == (256, 341)
    boxes = P.ten_crop_boxes(256, 341, 224)
    assert boxes[4] == (16, 58, False)  # python round(58.5) == 58
    assert [b[:2] for b in boxes[:4]] == [(0, 0), (0, 117), (32, 0), (32, 117)]
    assert all(b[2] for b in boxes[5:]) and not any(b[2] for b in boxes[:5])
    rng = P.standardize(np.array([0, 255], dtype=np.uint8))
    assert abs(rng[0] + 2.0) < 1e-6 and abs(rng[1] - 2.44444) < 1e-4


def test_center_crop_is_tencrop_index_4():
    frames = synth_frames(5, 3, 48, 64)
    ten = P.clip_tensor(frames, 0, 4, 64, 56, ncrops=10)
    one = P.clip_tensor(frames, 0, 4, 64, 56, ncrops=1)
    assert np.array_equal(one[0], ten[4])


def test_resize_matches_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(9)
    for (h, w, oh, ow) in [(240, 320, 256, 341), (360, 480, 256, 341), (50, 70, 256, 358), (300, 256, 300, 256)]:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(P.pil_resize_bilinear(img, oh, ow), want)


# ----------------------------------------------------------------------------- backbone
def test_seeded_weights_reproduce(golden_dir):
    g = np.load(os.path.join(golden_dir, "i3res50.npz"))
    sd = O.seeded_state_dict(0)
    assert len(sd) == 318
    assert sha(sd["conv1.weight"].numpy()) == str(g["conv1_sha"])
    assert abs(sum(float(v.double().sum()) for v in sd.values()) - float(g["weights_sum"])) < 1e-6


@pytest.mark.parametrize("tag", ["small", "odd", "full"])
def test_i3res50_oracle_matches_reference_features(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "i3res50.npz"))
    shape = tuple(int(v) for v in g[f"{tag}/shape"])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
    assert sha(x.numpy()) == str(g[f"{tag}/x_sha"])
    y, _ = O.forward(x, O.seeded_state_dict(0))
    want = g[f"{tag}/features"]
    assert y.shape == (shape[0], 2048, 1, 1, 1)
    # same ATen ops in the same order as the reference modules: bit-identical on the same machine
    np.testing.assert_allclose(y.numpy().reshape(shape[0], -1), want, rtol=1e-5, atol=1e-6)


def test_conv_macs_matches_survey_table():
    assert O.conv_macs(16, 224, 224) == 16_414_572_544


def test_bf16_emulation_is_within_the_stated_tolerance():
    sd = O.seeded_state_dict(0)
    x = torch.randn(1, 3, 8, 64, 64, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
    yr, _ = O.forward(x, sd)
    yb, _ = O.forward(x, sd, emulate_bf16=True)
    yr, yb = yr.flatten(), yb.flatten()
    assert ((yb - yr).abs().max() / yr.abs().max()).item() < 1e-2
    assert torch.nn.functional.cosine_similarity(yb, yr, dim=0).item() > 0.999


# ----------------------------------------------------------------------------- segment / stacking
@pytest.mark.parametrize("n", [2, 5, 31, 32, 33, 47, 125, 188])
def test_segment_golden(golden_dir, n):
    g = np.load(os.path.join(golden_dir, "segment.npz"))
    f, want = g[f"n{n}/in"], g[f"n{n}/out"]
    assert np.array_equal(S.segment_edges(n), g[f"n{n}/edges"])
    got = S.segment_features(f)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert np.array_equal(S.segment_features_sequential(f), want), "serial fp32 add + one divide must equal np.mean"


def test_segment_edges_are_integer_arithmetic():
    for n in list(range(1, 400)) + [1999, 2000, 3008, 20000]:
        e = S.segment_edges(n)
        assert all(int(e[i]) == (i * n) // 32 for i in range(33)), n
        if n < 32:  # empty bins fall back to f[r[i]]: the index must exist
            assert max(int(e[i]) for i in range(32)) <= n - 1


def test_stacking_and_squeeze_quirk():
    rng = np.random.default_rng(0)
    batches = [[rng.standard_normal((16, 32, 1, 1, 1)).astype(np.float32) for _ in range(10)],
               [rng.standard_normal((5, 32, 1, 1, 1)).astype(np.float32) for _ in range(10)]]
    out = S.stack_clip_features(batches)
    assert out.shape == (21, 10, 32)
    assert np.array_equal(out[16 + 2, 7], batches[1][7][2, :, 0, 0, 0])
    one = [[rng.standard_normal((1, 32, 1, 1, 1)).astype(np.float32) for _ in range(10)]]
    assert S.stack_clip_features(one, strict_compat=True).shape == (10, 32)
    assert S.stack_clip_features(one, strict_compat=False).shape == (1, 10, 32)


def test_extract_golden_layout(golden_dir):
    g = np.load(os.path.join(golden_dir, "extract.npz"))
    assert list(g["listing"]) == ["Abuse001_x264_i3d.npy", "Big777_x264/Big777_x264_0.npy", "Big777_x264_i3d.npy",
                                  "Normal_Videos_003_x264_i3d.npy"]
    assert g["Abuse001_x264/features"].shape == (3, 10, 32)          # 37 frames -> 3 clips
    assert g["Normal_Videos_003_x264/features"].shape == (10, 32)   # 10 frames -> 1 clip, squeezed
    assert g["Big777_x264/features"].shape == (2, 10, 32)


def test_add_magnitude():
    f = np.random.default_rng(0).standard_normal((10, 32, 64)).astype(np.float32)
    out = S.add_magnitude(f)
    assert out.shape == (10, 32, 65)
    np.testing.assert_allclose(out[..., 64], np.sqrt((f.astype(np.float64) ** 2).sum(-1)), rtol=1e-6)


# ----------------------------------------------------------------------------- live reference (build box only)
@pytest.mark.skipif(not _refload.available(), reason="reference checkout not mounted (GPU box)")
def test_oracle_equals_live_reference():
    from PIL import Image

    ref = _refload.load()
    frames = synth_frames(21, 6, 72, 100)
    ds = ref.dataset.TenCropVideoFrameDataset([Image.fromarray(f) for f in frames], frames_per_clip=4, resize=80, cropsize=64)
    for ci in range(len(ds)):
        assert np.array_equal(ds[ci].numpy(), P.clip_tensor(frames, ci, 4, 80, 64))
    sd = O.seeded_state_dict(3)
    m = ref.i3d.I3Res50(use_nl=False).eval()
    m.load_state_dict(sd, strict=True)
    x = torch.randn(1, 3, 8, 64, 64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = m(x)
    got, _ = O.forward(x, sd)
    assert torch.equal(got, want)


# ----------------------------------------------------------------------------- InceptionV1-3D (parity unpinned by the reference)
def test_inception_oracle_layer_table_matches_survey_appendix_b():
    from oracle import inception as OI

    assert OI.conv_macs() == 27_787_569_152
    sd = OI.seeded_state_dict(0)
    from anomaly_detection_on_video_b200.inception import InceptionI3d

    m = InceptionI3d()
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    m.fuse_siblings = False   # the three-launch table: one op per Unit3D of Appendix B
    ops = m.op_table()
    assert len(ops) == 71 and sum(1 for o in ops if o.cout) == 57
    # default table: b0 | b1a | b2a of each of the nine Mixed blocks is one op with routed output columns
    m.fuse_siblings = True
    fused = m.op_table()
    assert len(fused) == 71 - 2 * 9 and sum(1 for o in fused if o.dst1) == 9
    for o in fused:
        if o.dst1:
            assert o.split1 % 64 == 0 and o.split2 % 64 == 0 and o.seg_w[0] <= o.split1 and o.seg_w[1] <= o.split2 - o.split1
            assert o.seg_w[2] == o.cout - o.split2 and len({o.dst, o.dst1, o.dst2, o.src}) == 4
    # every Inception branch writes a channel slice of the concatenated tensor: slices tile [0, total) exactly
    for name in ("Mixed_3b", "Mixed_4f", "Mixed_5c"):
        sl = sorted((o.dst_c_off, o.cout, o.dst_c_total) for o in ops if o.name.startswith(name) and o.dst_c_total)
        assert sl[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(sl, sl[1:])) and sl[-1][0] + sl[-1][1] == sl[-1][2]
