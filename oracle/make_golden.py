"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE on seeded synthetic inputs.
TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference ships no tests, fixtures or golden vectors, so these files are the parity pins:
  preprocess_small.npz   TenCropVideoFrameDataset (src/dataset.py:145-195) on 5 frames of 60x80
  preprocess_digests.npz sha256 of the fp32 clip tensors for full-size frames (240x320 etc.)
  i3res50.npz            I3Res50.forward (src/i3d.py:302-318) with oracle.i3res50.seeded_state_dict(0)
  segment.npz            segment() (extract_features.py:159-185) for several clip counts
  extract.npz            extract() (extract_features.py:55-156) driven end to end with a fake decoder
                         and a cheap deterministic model: file names, shapes, stacking, chunk cache
  mgfn.npz               MGFNForVideoAnomalyDetection.forward in eval mode (modeling_mgfn.py:376-427) with
                         oracle.mgfn.seeded_state_dict(0): a split training-shaped batch with labels (scores,
                         selection, every loss term) and an unsplit validation-shaped video (T = 47)
                         (`python -m oracle.make_golden mgfn` regenerates only this file)
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

from . import _refload
from . import i3res50 as O
from . import mgfn as MG

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_frames(seed: int, n: int, h: int, w: int) -> np.ndarray:
    """The synthetic video used everywhere: uniform noise (worst case for +-1 LSB resize errors)."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def to_u8(clip_f32: np.ndarray) -> np.ndarray:
    """Invert the standardisation exactly (values are a 256-entry LUT) to store fixtures compactly."""
    u = np.rint(clip_f32.astype(np.float64) * 57.375 + 114.75).astype(np.uint8)
    back = ((u.astype(np.float32) - np.float32(114.75)) / np.float32(57.375)).astype(np.float32)
    assert np.array_equal(back, clip_f32)
    return u


class FakeModel(torch.nn.Module):
    """Cheap deterministic stand-in for the backbone: [B,3,T,H,W] -> [B,32,1,1,1]."""

    def __init__(self, dim: int = 32):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        self.register_buffer("proj", torch.randn(dim, 3 * 4, generator=g))

    def forward(self, x):
        b = x.shape[0]
        # per-channel means over 4 temporal quarters -> 12 numbers -> projection
        q = x.reshape(b, 3, 4, -1).double().mean(dim=-1).float().reshape(b, 12)
        return (q @ self.proj.t()).reshape(b, -1, 1, 1, 1)


def golden_mgfn(ref) -> None:
    """P5: the scoring head in eval() mode (dropout on the selection mask inactive -> deterministic)."""
    import importlib

    modeling = importlib.import_module("src.models.mgfn.modeling_mgfn")
    configuration = importlib.import_module("src.models.mgfn.configuration_mgfn")
    sd = MG.seeded_state_dict(0)
    model = modeling.MGFNForVideoAnomalyDetection(configuration.MGFNConfig())
    model.load_state_dict(sd, strict=True)
    model.eval()
    out = {"weights_sum": np.array(sum(float(v.double().sum()) for v in sd.values()))}

    def record(tag, r, video):
        out[f"{tag}/video_sha"] = np.array(sha(video.numpy()))
        out[f"{tag}/scores"] = r.scores.numpy()
        out[f"{tag}/abnormal_scores"] = r.abnormal_scores.numpy()
        out[f"{tag}/normal_scores"] = r.normal_scores.numpy()
        out[f"{tag}/a_feat_l1"] = r.a_feat_magnitude.norm(p=1, dim=2).numpy()
        out[f"{tag}/n_feat_l1"] = r.n_feat_magnitude.norm(p=1, dim=2).numpy()
        out[f"{tag}/a_feat_sample"] = r.a_feat_magnitude[:, :, ::16].numpy()
        out[f"{tag}/n_feat_sample"] = r.n_feat_magnitude[:, :, ::16].numpy()

    video = MG.synthetic_video(1, 4, 10, 32)
    model.force_split = True
    with torch.no_grad():
        r = model(video, abnormal_labels=torch.ones(2), normal_labels=torch.zeros(2))
    record("split", r, video)
    out["split/loss"] = r.loss.numpy()
    video = MG.synthetic_video(2, 1, 10, 47)
    model.force_split = False
    with torch.no_grad():
        r = model(video)
    record("valid", r, video)
    np.savez_compressed(os.path.join(GOLDEN, "mgfn.npz"), **out)


def golden_mgfn_train(ref) -> None:
    """a23 (src/runner.py:29-39,53-59): one training step of the live reference -- train() mode, every dropout at 0 so that
    it is deterministic -- on a batch of 2 normal + 2 abnormal bags: loss, per-parameter gradient digests (L2 norm, sum,
    first 6 entries), the BatchNorm running statistics after the step and the digests of the parameters after one Adam
    step (lr 1e-3, weight_decay 5e-4: configs/runner/default.yaml:5-7)."""
    import importlib

    modeling = importlib.import_module("src.models.mgfn.modeling_mgfn")
    configuration = importlib.import_module("src.models.mgfn.configuration_mgfn")
    sd = MG.seeded_state_dict(0)
    model = modeling.MGFNForVideoAnomalyDetection(configuration.MGFNConfig(dropout_rate=0.0))
    model.load_state_dict(sd, strict=True)
    model.train()
    video = MG.synthetic_video(3, 4, 10, 32)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    r = model(video, abnormal_labels=torch.ones(2), normal_labels=torch.zeros(2))
    opt.zero_grad()
    r.loss.backward()
    out = {"video_sha": np.array(sha(video.numpy())), "loss": r.loss.detach().numpy(), "scores": r.scores.detach().numpy()}
    names = []
    for name, p_ in model.named_parameters():
        g = p_.grad.detach().double().reshape(-1)
        names.append(name)
        out[f"grad/{name}"] = np.array([float(g.norm()), float(g.sum())] + [float(v) for v in g[:6]], dtype=np.float64)
    opt.step()
    for name, p_ in model.named_parameters():
        w = p_.detach().double().reshape(-1)
        out[f"adam/{name}"] = np.array([float(w.norm()), float(w.sum())] + [float(v) for v in w[:6]], dtype=np.float64)
    for name, b in model.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            out[f"buf/{name}"] = b.detach().numpy()
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, "mgfn_train.npz"), **out)


def main() -> None:
    ref = _refload.load()
    os.makedirs(GOLDEN, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "mgfn":
        golden_mgfn(ref)
        print("mgfn.npz", os.path.getsize(os.path.join(GOLDEN, "mgfn.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "mgfn_train":
        golden_mgfn_train(ref)
        print("mgfn_train.npz", os.path.getsize(os.path.join(GOLDEN, "mgfn_train.npz")))
        return
    golden_mgfn(ref)
    golden_mgfn_train(ref)

    # ---------------------------------------------------------------- P1 preprocessing
    frames = synth_frames(0, 5, 60, 80)
    ds = ref.dataset.TenCropVideoFrameDataset([Image.fromarray(f) for f in frames], frames_per_clip=4, resize=64, cropsize=56)
    clips = np.stack([ds[i].numpy() for i in range(len(ds))])  # (2, 10, 4, 3, 56, 56); clip 1 is loop-padded
    np.savez_compressed(os.path.join(GOLDEN, "preprocess_small.npz"), frames=frames, clips_u8=to_u8(clips),
                        frames_per_clip=4, resize=64, crop=56)
    digests = {}
    for tag, (seed, n, h, w, which) in {"ucf_240x320": (0, 37, 240, 320, (0, 2)), "down_360x480": (1, 5, 360, 480, (0,)),
                                        "portrait_300x256": (2, 3, 300, 256, (0,))}.items():
        fr = synth_frames(seed, n, h, w)
        d = ref.dataset.TenCropVideoFrameDataset([Image.fromarray(f) for f in fr])
        for ci in which:
            t = d[ci].numpy()
            digests[f"{tag}/clip{ci}"] = sha(t)
            digests[f"{tag}/clip{ci}/sum"] = repr(float(t.astype(np.float64).sum()))
        digests[f"{tag}/spec"] = repr((seed, n, h, w))
    np.savez(os.path.join(GOLDEN, "preprocess_digests.npz"), **{k: np.array(v) for k, v in digests.items()})

    # ---------------------------------------------------------------- P2 backbone
    sd = O.seeded_state_dict(0)
    model = ref.i3d.I3Res50(use_nl=False).eval()
    model.load_state_dict(sd, strict=True)
    out = {"weights_sum": np.array(sum(float(v.double().sum()) for v in sd.values())),
           "conv1_sha": np.array(sha(sd["conv1.weight"].numpy()))}
    for tag, shape in {"small": (1, 3, 8, 64, 64), "odd": (2, 3, 12, 96, 80), "full": (2, 3, 16, 224, 224)}.items():
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).clamp(-2.0, 2.4444)
        with torch.no_grad():
            y = model(x)
        out[f"{tag}/shape"] = np.array(shape)
        out[f"{tag}/x_sha"] = np.array(sha(x.numpy()))
        out[f"{tag}/features"] = y.numpy().reshape(shape[0], -1)
    np.savez_compressed(os.path.join(GOLDEN, "i3res50.npz"), **out)

    # ---------------------------------------------------------------- P3 segment
    seg = {}
    rng = np.random.default_rng(3)
    for n in (2, 5, 31, 32, 33, 47, 125, 188):
        f = (rng.standard_normal((n, 10, 8)) * 3).astype(np.float32)
        d = tempfile.mkdtemp()
        os.makedirs(os.path.join(d, "in"))
        os.makedirs(os.path.join(d, "out"))
        np.save(os.path.join(d, "in", "v_i3d.npy"), f)
        ref.extract_features.segment(os.path.join(d, "in"), os.path.join(d, "out"), 32)
        seg[f"n{n}/in"] = f
        seg[f"n{n}/out"] = np.load(os.path.join(d, "out", "v_i3d.npy"))
        seg[f"n{n}/edges"] = np.linspace(0, n, 33, dtype=int)
    np.savez_compressed(os.path.join(GOLDEN, "segment.npz"), **seg)

    # ---------------------------------------------------------------- P4 extract() end to end
    import datasets

    class FakeReader:  # stands in for decord.VideoReader: the "video" is an .npy of frames
        def __init__(self, uri):
            self.frames = np.load(uri)

        def __len__(self):
            return len(self.frames)

        def __getitem__(self, i):
            a = self.frames[i]

            class _F:
                def asnumpy(self_inner):
                    return a

            return _F()

    sys.modules["decord"].VideoReader = FakeReader
    work = tempfile.mkdtemp()
    specs = {"Abuse001_x264": (11, 37), "Normal_Videos_003_x264": (12, 10), "Big777_x264": (13, 20)}
    rows = {"video_path": [], "size": []}
    for name, (seed, n) in specs.items():
        path = os.path.join(work, name + ".npy")  # "<stem>.<ext>": the reference splits on the first '.'
        np.save(path, synth_frames(seed, n, 64, 96))
        rows["video_path"].append(path)
        rows["size"].append(2 * 1024 ** 2 if name.startswith("Big") else 1000)  # > 1 GB (in KB) -> chunk path
    dset = datasets.Dataset.from_dict(rows)
    outpath = os.path.join(work, "anomaly_features", "train")
    ref.extract_features.extract(dset, FakeModel().eval(), torch.device("cpu"), outpath)
    ext = {"listing": np.array(sorted(os.path.relpath(os.path.join(r, f), outpath) for r, _, fs in os.walk(outpath) for f in fs))}
    for name, (seed, n) in specs.items():
        ext[f"{name}/spec"] = np.array([seed, n, 64, 96])
        ext[f"{name}/features"] = np.load(os.path.join(outpath, name + "_i3d.npy"))
    np.savez_compressed(os.path.join(GOLDEN, "extract.npz"), **ext)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
