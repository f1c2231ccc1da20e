"""-m gpu: the native MGFN scoring head (tcgen05 kind::tf32 GEMMs + fp32 kernels, through the C ABI) against the CPU
oracle and the golden vectors of the unmodified reference.

Tolerance: operands are rounded to TF32 (10 mantissa bits) inside the tensor core, accumulation is fp32; the
reference is plain fp32.  Scores are sigmoids in [0, 1]: |a - b| <= 2e-3.  Normalised features (after ~50 chained
TF32 GEMMs and 19 LayerNorms): relative L2 error <= 5e-3 (measured 2.3e-3).  The contrastive terms are squared
differences of L1 norms that agree to 4 digits, so they carry ~1% error; they enter the loss with weight 1e-3.  Selection indices are integers and must match exactly on these inputs
(the synthetic videos have well-separated top-k magnitudes)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SCORE_ATOL = 2e-3
FEAT_RTOL = 5e-3


@pytest.fixture(scope="module")
def head(cuda_device):
    from anomaly_detection_on_video_b200.mgfn import MGFNConfig, MGFNForVideoAnomalyDetection
    from oracle import mgfn as M

    m = MGFNForVideoAnomalyDetection(MGFNConfig())
    m.load_state_dict(M.seeded_state_dict(0), strict=True)
    return m.eval().to(cuda_device)


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_split_batch_scores_selection_and_losses(head, cuda_device, golden_dir):
    from oracle import mgfn as M

    g = np.load(os.path.join(golden_dir, "mgfn.npz"))
    video = M.synthetic_video(1, 4, 10, 32)
    head.force_split = True
    out = head(video.to(cuda_device), abnormal_labels=torch.ones(2), normal_labels=torch.zeros(2))
    torch.cuda.synchronize()
    ref = M.forward(video, M.seeded_state_dict(0), split=True, normal_labels=torch.zeros(2), abnormal_labels=torch.ones(2))
    assert out.scores.shape == (4, 32, 1) and out.abnormal_scores.shape == (2, 1) and out.a_feat_magnitude.shape == (20, 3, 1024)
    # vs the reference's own outputs (golden) and vs the oracle
    assert np.abs(out.scores.cpu().numpy() - g["split/scores"]).max() <= SCORE_ATOL
    assert np.abs(out.abnormal_scores.cpu().numpy() - g["split/abnormal_scores"]).max() <= SCORE_ATOL
    assert np.abs(out.normal_scores.cpu().numpy() - g["split/normal_scores"]).max() <= SCORE_ATOL
    idx = head._last_idx.cpu().long()
    assert torch.equal(idx[:2], ref["idx_normal"]) and torch.equal(idx[2:], ref["idx_abnormal"])
    assert _rel(out.a_feat_magnitude.cpu(), ref["a_feat_magnitude"]) <= FEAT_RTOL
    assert _rel(out.n_feat_magnitude.cpu(), ref["n_feat_magnitude"]) <= FEAT_RTOL
    np.testing.assert_allclose(out.a_feat_magnitude.norm(p=1, dim=2).cpu().numpy(), g["split/a_feat_l1"], rtol=FEAT_RTOL)
    # every loss term: smooth, sparsity, bce, con, con_n, con_a, and the total
    got, want = out.loss_terms.cpu().numpy(), ref["loss_terms"].numpy()
    np.testing.assert_allclose(got[:4], want[:4], rtol=5e-3, atol=1e-5)   # total, smooth, sparsity, bce
    np.testing.assert_allclose(got[4:], want[4:], rtol=3e-2)              # con, con_n, con_a
    np.testing.assert_allclose(float(out.loss), float(g["split/loss"]), rtol=5e-3)


def test_validation_video_variable_length(head, cuda_device, golden_dir):
    from oracle import mgfn as M

    g = np.load(os.path.join(golden_dir, "mgfn.npz"))
    video = M.synthetic_video(2, 1, 10, 47)  # T = 47: the 64-token tile path with masked tail rows
    head.force_split = False
    out = head(video.to(cuda_device))
    torch.cuda.synchronize()
    assert out.loss is None and out.scores.shape == (1, 47, 1)
    assert np.abs(out.scores.cpu().numpy() - g["valid/scores"]).max() <= SCORE_ATOL
    assert torch.equal(out.abnormal_scores, out.normal_scores)
    np.testing.assert_allclose(out.n_feat_magnitude[:, :, ::16].cpu().numpy(), g["valid/n_feat_sample"], rtol=0, atol=2e-2)


@pytest.mark.parametrize("bs,T", [(2, 5), (1, 130), (3, 32)])
def test_per_snippet_outputs_match_oracle(head, cuda_device, bs, T):
    """Short, long (two 128-token tiles per sequence) and multi-video batches: crop-mean scores and the layer-normed
    features of every snippet."""
    from oracle import mgfn as M

    video = M.synthetic_video(10 + T, bs, 10, T)
    head.force_split = False
    out = head(video.to(cuda_device))
    torch.cuda.synchronize()
    ref = M.forward(video, M.seeded_state_dict(0), split=False)
    assert (out.scores.cpu() - ref["scores"]).abs().max() <= SCORE_ATOL
    assert _rel(out.n_feat_magnitude.cpu(), ref["n_feat_magnitude"]) <= FEAT_RTOL


def test_head_rejects_bad_inputs(head, cuda_device):
    with pytest.raises(ValueError):
        head(torch.zeros(2, 10, 32, 2048, device=cuda_device))
    head.force_split = True
    with pytest.raises(ValueError):
        head(torch.zeros(3, 10, 32, 2049, device=cuda_device))
    head.force_split = False


def test_validation_loop_over_a_feature_archive(head, cuda_device, tmp_path):
    """FeatureDataset (local zip, GPU magnitude) -> runner.validate == the reference's validation_step +
    on_validation_epoch_end formulas evaluated on the oracle's scores (src/runner.py:42-79)."""
    import json
    import zipfile

    from anomaly_detection_on_video_b200.dataset import build_feature_dataset
    from anomaly_detection_on_video_b200.runner import frame_level_metrics, validate, validation_step
    from oracle import mgfn as M

    rng = np.random.default_rng(3)
    gt, feats = {}, {}
    with zipfile.ZipFile(tmp_path / "test.zip", "w") as zf:
        for name, t in (("Abuse028_x264_i3d.npy", 7), ("Normal_Videos_050_x264_i3d.npy", 12)):
            f = np.abs(rng.standard_normal((t, 10, 2048))).astype(np.float32) * 0.5
            np.save(tmp_path / name, f)
            zf.write(tmp_path / name, arcname="test/" + name)
            feats[name] = f
            gt[name] = (rng.random(t * 16) > 0.5).astype(float).tolist()
    json.dump(gt, open(tmp_path / "gt.json", "w"))
    ds = build_feature_dataset("test", local_path=str(tmp_path), filename="test.zip", ground_truth=str(tmp_path / "gt.json"),
                               device=cuda_device)
    item = ds[0]
    want = np.concatenate([feats[ds.get_filename(0)], np.linalg.norm(feats[ds.get_filename(0)], axis=2)[:, :, None]], axis=2)
    np.testing.assert_allclose(item["feature"], want, rtol=1e-6, atol=1e-6)
    assert item["anomaly"] == 1.0 and item["label"].shape == (7 * 16,)
    head.force_split = False
    got = validate(head, ds, frames_per_clip=16, device=cuda_device)
    sd = M.seeded_state_dict(0)
    ref_outs = []
    for i in range(len(ds)):
        it = ds[i]
        video = torch.from_numpy(it["feature"]).unsqueeze(0).permute(0, 2, 1, 3)
        ref_outs.append({"preds": M.forward(video, sd, split=False)["scores"].squeeze(0).squeeze(-1).numpy(), "labels": it["label"]})
        step = validation_step(head, it, cuda_device)
        assert np.abs(step["preds"] - ref_outs[-1]["preds"]).max() <= SCORE_ATOL
    ref = frame_level_metrics(ref_outs, 16)
    assert abs(got["valid/rec_auc"] - ref["valid/rec_auc"]) < 0.02 and abs(got["valid/pr_auc"] - ref["valid/pr_auc"]) < 0.02
