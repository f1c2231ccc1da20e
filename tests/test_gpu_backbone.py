"""-m gpu: the full I3Res50 forward and the extract()/segment() drop-ins against the reference's
golden outputs (tests/golden, produced by the unmodified reference) and the CPU oracle.

Tolerance (bf16 operands + bf16 stored activations, fp32 accumulate; BASELINE.json north_star:
"1e-2 max relative error and cosine >= 0.999"): relative error is measured against the feature
vector's scale -- max|a-b| <= 1e-2 * max|b| and ||a-b|| <= 1e-2 * ||b|| per clip -- because
post-ReLU pooled features contain values arbitrarily close to 0 for which an elementwise ratio is
meaningless.  The literal element-wise reading of north_star, floored as SURVEY section 7 proposes --
|a-b| / max(|b|, 1e-3 max|b|) -- is NOT held to 1e-2 by any pipeline that stores activations in bf16 (the
oracle's own bf16 emulation shows the same gap on the CPU): it is measured here, printed, and pinned by the
regression bounds ELEM_* below so that it cannot silently get worse; the TF32 mode (tests/test_gpu_tf32.py)
is the mode that holds 1e-3.
"""
import os

import numpy as np
import pytest
import torch

from oracle import i3res50 as O
from oracle import segment as S

pytestmark = pytest.mark.gpu

MAX_NORM_TOL = 1e-2
REL_L2_TOL = 1e-2
COS_MIN = 0.999
# floored element-wise relative error (see the module docstring): regression bounds, ~1.5-2x what a B200 measures
# (round 2: max 0.09-0.68, p99 0.04-0.12, fraction > 1e-2 0.14-0.19 over the golden cases)
ELEM_P99_MAX = 0.20
ELEM_FRAC_GT_1E2_MAX = 0.30
ELEM_MAX_MAX = 1.5


def check_features(got: np.ndarray, want: np.ndarray):
    assert got.shape == want.shape
    for b in range(want.shape[0]):
        a, r = got[b].astype(np.float64), want[b].astype(np.float64)
        max_norm = np.abs(a - r).max() / np.abs(r).max()
        rel_l2 = np.linalg.norm(a - r) / np.linalg.norm(r)
        cos = float(a @ r / (np.linalg.norm(a) * np.linalg.norm(r)))
        el = np.abs(a - r) / np.maximum(np.abs(r), 1e-3 * np.abs(r).max())
        el_max, el_p99, el_frac = float(el.max()), float(np.quantile(el, 0.99)), float((el > 1e-2).mean())
        print(f"clip {b}: max-normalised err {max_norm:.2e}, rel L2 {rel_l2:.2e}, cos {cos:.6f}; floored element-wise: "
              f"max {el_max:.3f}, p99 {el_p99:.3f}, fraction > 1e-2 {el_frac:.3f}")
        assert max_norm <= MAX_NORM_TOL and rel_l2 <= REL_L2_TOL and cos >= COS_MIN
        assert el_p99 <= ELEM_P99_MAX and el_frac <= ELEM_FRAC_GT_1E2_MAX and el_max <= ELEM_MAX_MAX


@pytest.fixture(scope="module")
def model(cuda_device):
    from anomaly_detection_on_video_b200.i3d import I3Res50

    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0), strict=True)
    return m.eval().to(cuda_device)


@pytest.mark.parametrize("tag", ["small", "odd", "full"])
def test_features_match_reference_golden(model, cuda_device, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "i3res50.npz"))
    shape = tuple(int(v) for v in g[f"{tag}/shape"])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
    y = model(x.to(cuda_device))
    assert tuple(y.shape) == (shape[0], 2048, 1, 1, 1) and y.dtype == torch.float32
    check_features(y.cpu().numpy().reshape(shape[0], -1), g[f"{tag}/features"])


def test_features_track_the_bf16_emulating_oracle(model, cuda_device):
    """Against the oracle run with the same roundings the gap shrinks ~3x: what remains is
    accumulation order, i.e. the kernels compute what the design says they compute."""
    x = torch.randn(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(4)).clamp(-2.0, 2.4444)
    sd = O.seeded_state_dict(0)
    y32, _ = O.forward(x, sd)
    y16, _ = O.forward(x, sd, emulate_bf16=True)
    y = model(x.to(cuda_device)).cpu()
    scale = y32.abs().max()
    assert ((y - y16).abs().max() / scale).item() < 8e-3
    assert ((y - y32).abs().max() / scale).item() < MAX_NORM_TOL


def test_gather_and_tma_paths_agree(model, cuda_device):
    """Every conv through the cp.async gather producer vs the TMA producers.  Layer by layer the operand tiles are
    identical (tests/test_gpu_kernels.py asserts bit-equality per layer); end to end the (3,1,1) layers run through the
    temporal-halo kernel, whose (channel block, tap) contraction order differs from the gather path's (tap, channel
    block) in fp32 summation order, so the features agree to bf16 resolution rather than bit for bit."""
    x = torch.randn(2, 3, 8, 64, 64, generator=torch.Generator().manual_seed(7)).clamp(-2.0, 2.4444).to(cuda_device)
    a = model(x).clone()
    model.force_gather = True
    try:
        b = model(x).clone()
    finally:
        model.force_gather = False
    assert float((a - b).abs().max()) <= 5e-3 * float(b.abs().max())
    assert float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)) >= 0.99999


def test_batch_invariance_and_determinism(model, cuda_device):
    x = torch.randn(5, 3, 8, 64, 64, generator=torch.Generator().manual_seed(9)).clamp(-2.0, 2.4444).to(cuda_device)
    full = model(x).clone()
    again = model(x).clone()
    assert torch.equal(full, again), "same input, same bits"
    single = torch.cat([model(x[i:i + 1]).clone() for i in range(5)])
    assert torch.equal(full, single), "a clip's features do not depend on what else is in the batch"


def test_weights_are_repacked_after_load_state_dict(model, cuda_device):
    from anomaly_detection_on_video_b200.i3d import I3Res50

    m = I3Res50().eval().to(cuda_device)
    x = torch.randn(1, 3, 8, 64, 64, generator=torch.Generator().manual_seed(2)).clamp(-2.0, 2.4444).to(cuda_device)
    before = m(x).clone()
    m.load_state_dict(O.seeded_state_dict(0), strict=True)
    after = m(x).clone()
    assert not torch.equal(before, after)
    assert torch.equal(after, model(x))


def test_training_mode_and_cpu_input_are_refused(model, cuda_device):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 8, 64, 64))
    model.train()
    try:
        with pytest.raises(RuntimeError, match="inference-only"):
            model(torch.zeros(1, 3, 8, 64, 64, device=cuda_device))
    finally:
        model.eval()


# ----------------------------------------------------------------------------- extract() / segment() drop-ins
class FakeModel(torch.nn.Module):
    """Same stand-in backbone oracle/make_golden.py drove the reference's extract() with."""

    def __init__(self, dim: int = 32):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        self.register_buffer("proj", torch.randn(dim, 12, generator=g))

    def forward(self, x):
        b = x.shape[0]
        q = x.reshape(b, 3, 4, -1).double().mean(dim=-1).float().reshape(b, 12)
        return (q @ self.proj.t()).reshape(b, -1, 1, 1, 1)


def _write_videos(tmp_path, g):
    rows = []
    for name in ("Abuse001_x264", "Normal_Videos_003_x264", "Big777_x264"):
        seed, n, h, w = (int(v) for v in g[f"{name}/spec"])
        path = os.path.join(tmp_path, name + ".npy")
        np.save(path, np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8))
        rows.append({"video_path": path, "size": 2 * 1024 ** 2 if name.startswith("Big") else 1000})
    return rows


def test_extract_reproduces_reference_files(cuda_device, golden_dir, tmp_path):
    from anomaly_detection_on_video_b200 import extract_features as E

    g = np.load(os.path.join(golden_dir, "extract.npz"))
    rows = _write_videos(str(tmp_path), g)
    outpath = os.path.join(str(tmp_path), "anomaly_features", "train")
    fake = FakeModel().eval().to(cuda_device)
    written = E.extract(rows, fake, cuda_device, outpath)
    listing = sorted(os.path.relpath(os.path.join(r, f), outpath) for r, _, fs in os.walk(outpath) for f in fs)
    assert listing == list(g["listing"])
    for name in ("Abuse001_x264", "Normal_Videos_003_x264", "Big777_x264"):
        got = np.load(os.path.join(outpath, name + "_i3d.npy"))
        want = g[f"{name}/features"]
        assert got.shape == want.shape and got.dtype == np.float32  # incl. the (10, C) squeeze of a 1-clip video
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)  # inputs are bit-identical; the fake model's
        # fp64 mean / fp32 matmul run on another device
    assert len(written) == 3
    assert E.extract(rows, fake, cuda_device, outpath) == [], "second run skips everything (idempotent resume)"
    # DatasetDict-style recursion: one sub-directory per split
    E.extract({"test": rows[:1]}, fake, cuda_device, os.path.join(str(tmp_path), "dd"))
    assert os.path.exists(os.path.join(str(tmp_path), "dd", "test", "Abuse001_x264_i3d.npy"))
    # segment(): creates its directory, handles the squeezed 1-clip file, bit-identical to the oracle
    seg_out = os.path.join(str(tmp_path), "segment_features_32")
    E.segment(outpath, seg_out, 32)
    for name in ("Abuse001_x264", "Normal_Videos_003_x264", "Big777_x264"):
        feats = np.load(os.path.join(outpath, name + "_i3d.npy"))
        seg = np.load(os.path.join(seg_out, name + "_i3d.npy"))
        feats3 = feats[None] if feats.ndim == 2 else feats
        assert seg.shape == (10, 32, 32) and np.array_equal(seg, S.segment_features(feats3, 32))


def test_extract_native_backbone_fast_path(model, cuda_device, tmp_path):
    """Native model: 10 crops batched into one forward from the bf16 stem layout == per-crop fp32 calls."""
    from anomaly_detection_on_video_b200 import extract_features as E
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    frames = np.random.default_rng(5).integers(0, 256, size=(40, 120, 160, 3), dtype=np.uint8)
    ds = TenCropVideoFrameDataset(frames, device=cuda_device)
    feats = E.extract_clip_features(ds, model, cuda_device, clips_per_batch=2)
    assert feats.shape == (3, 10, 2048) and feats.dtype == np.float32
    # the reference's call pattern: model(inputs[:, crop_idx]) on the fp32 NCTHW tensor
    clip1 = ds[1].permute(0, 2, 1, 3, 4)  # (10, 3, 16, 224, 224)
    ref = model(clip1.contiguous()).reshape(10, 2048).cpu().numpy()
    assert np.array_equal(feats[1], ref), "same bf16 inputs, same kernels -> same bits"
    path = os.path.join(str(tmp_path), "v.npy")
    np.save(path, frames)
    out = E.extract([{"video_path": path, "size": 10}], model, cuda_device, os.path.join(str(tmp_path), "o"))
    assert np.array_equal(np.load(out[0]), feats)


def test_extract_stream_matches_per_video_extraction(model, cuda_device):
    """extract_stream (uploads / read-backs of neighbouring videos overlapped) returns, video by video, exactly what the
    sequential calls return -- including for videos of different lengths and geometries back to back -- and datasets of
    one frame geometry share a single preprocessing handle."""
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
    from anomaly_detection_on_video_b200.engine import segment_mean
    from anomaly_detection_on_video_b200.extract_features import extract_clip_features, extract_stream

    rng = np.random.default_rng(5)
    videos = [rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8) for n, h, w in [(70, 240, 320), (33, 240, 320), (16, 120, 160), (90, 240, 320)]]
    want = []
    for v in videos:
        ds = TenCropVideoFrameDataset(v, device=cuda_device)
        f = extract_clip_features(ds, model, cuda_device, strict_compat=False, as_numpy=False)
        want.append((f.cpu(), segment_mean(f, 32).cpu()))
    a = TenCropVideoFrameDataset(videos[0], device=cuda_device)
    b = TenCropVideoFrameDataset(videos[1], device=cuda_device)
    c = TenCropVideoFrameDataset(videos[2], device=cuda_device)
    assert a._pp is b._pp and a._pp is not c._pp
    got = list(extract_stream(iter(videos), model, cuda_device))
    assert len(got) == len(want)
    for (gf, gs), (wf, ws) in zip(got, want):
        assert torch.equal(gf, wf) and torch.equal(gs, ws)


def test_cuda_graph_replay_is_bit_identical_to_direct_launches(cuda_device, monkeypatch):
    """vad_plan_forward captures the op table of a small batch into a CUDA graph at the second forward of a binding and
    replays it afterwards (VAD_GRAPH: 0 never, 1 any batch, unset batch <= 32).  Same kernels, same arguments: same bits,
    for the capturing forward, for replays, for new input values in the bound buffer, and after a re-bind."""
    from anomaly_detection_on_video_b200.i3d import I3Res50

    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0), strict=True)
    m = m.eval().to(cuda_device)
    gen = torch.Generator().manual_seed(11)
    xs = [torch.randn(2, 16, 224, 232, 4, generator=gen).to(torch.bfloat16).to(cuda_device) for _ in range(2)]
    monkeypatch.setenv("VAD_GRAPH", "0")
    want = [m.forward_stem_layout(x).clone() for x in xs]
    torch.cuda.synchronize()
    monkeypatch.setenv("VAD_GRAPH", "1")
    buf = xs[0].clone()
    outs = [m.forward_stem_layout(buf).clone() for _ in range(4)]      # direct, capture + launch, replay, replay
    buf.copy_(xs[1])
    outs2 = [m.forward_stem_layout(buf).clone() for _ in range(2)]     # replay on new values in the same buffer
    other = xs[0].clone()
    outs3 = [m.forward_stem_layout(other).clone() for _ in range(3)]   # re-bind: direct, capture, replay
    torch.cuda.synchronize()
    assert all(torch.equal(o, want[0]) for o in outs)
    assert all(torch.equal(o, want[1]) for o in outs2)
    assert all(torch.equal(o, want[0]) for o in outs3)
    monkeypatch.delenv("VAD_GRAPH")
    assert torch.equal(m.forward_stem_layout(other), want[0])          # default policy: batch 2 <= 32 -> graph
