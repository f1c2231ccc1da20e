"""Pin oracle/mgfn.py (CPU restatement of the MGFN scoring head + losses) to tests/golden/mgfn.npz, which the
unmodified reference produced (oracle/make_golden.py), and to the live reference when /root/reference is mounted."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import _refload
from oracle import mgfn as M


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "mgfn.npz"))


@pytest.fixture(scope="module")
def sd():
    return M.seeded_state_dict(0)


def test_seeded_weights_are_reproducible(golden, sd):
    assert abs(sum(float(v.double().sum()) for v in sd.values()) - float(golden["weights_sum"])) < 1e-6


def test_split_batch_with_losses_matches_reference_golden(golden, sd):
    video = M.synthetic_video(1, 4, 10, 32)
    assert sha(video.numpy()) == str(golden["split/video_sha"])
    o = M.forward(video, sd, split=True, normal_labels=torch.zeros(2), abnormal_labels=torch.ones(2))
    for key in ("scores", "abnormal_scores", "normal_scores"):
        np.testing.assert_allclose(o[key].numpy(), golden[f"split/{key}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o["a_feat_magnitude"].norm(p=1, dim=2).numpy(), golden["split/a_feat_l1"], rtol=1e-5)
    np.testing.assert_allclose(o["n_feat_magnitude"].norm(p=1, dim=2).numpy(), golden["split/n_feat_l1"], rtol=1e-5)
    np.testing.assert_allclose(o["a_feat_magnitude"][:, :, ::16].numpy(), golden["split/a_feat_sample"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(o["loss"]), float(golden["split/loss"]), rtol=1e-5)


def test_validation_video_matches_reference_golden(golden, sd):
    video = M.synthetic_video(2, 1, 10, 47)
    assert sha(video.numpy()) == str(golden["valid/video_sha"])
    o = M.forward(video, sd, split=False)
    np.testing.assert_allclose(o["scores"].numpy(), golden["valid/scores"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o["abnormal_scores"].numpy(), golden["valid/abnormal_scores"], rtol=1e-5, atol=1e-6)
    assert torch.equal(o["abnormal_scores"], o["normal_scores"])  # unsplit eval: both views of the same batch
    np.testing.assert_allclose(o["n_feat_magnitude"][:, :, ::16].numpy(), golden["valid/n_feat_sample"], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not _refload.available(), reason="reference checkout not mounted (GPU box)")
def test_oracle_equals_live_reference(sd):
    import importlib

    _refload.load()
    modeling = importlib.import_module("src.models.mgfn.modeling_mgfn")
    configuration = importlib.import_module("src.models.mgfn.configuration_mgfn")
    model = modeling.MGFNForVideoAnomalyDetection(configuration.MGFNConfig())
    model.load_state_dict(sd, strict=True)
    model.eval()
    model.force_split = True
    video = M.synthetic_video(5, 2, 10, 32)
    with torch.no_grad():
        r = model(video, abnormal_labels=torch.ones(1), normal_labels=torch.zeros(1))
    o = M.forward(video, sd, split=True, normal_labels=torch.zeros(1), abnormal_labels=torch.ones(1))
    for key in ("scores", "abnormal_scores", "normal_scores", "a_feat_magnitude", "n_feat_magnitude", "loss"):
        assert torch.equal(getattr(r, key), o[key]), key


def test_product_module_keeps_the_reference_parameter_names(sd):
    """The native head's module tree loads a reference state_dict strictly (drop-in checkpoints)."""
    from anomaly_detection_on_video_b200.mgfn import MGFNConfig, MGFNForVideoAnomalyDetection

    m = MGFNForVideoAnomalyDetection(MGFNConfig())
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    blob = m.pack_parameters()
    assert blob.dtype == torch.float32 and blob.numel() % 64 == 0
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m.eval()(M.synthetic_video(0, 2, 10, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):  # training runs on the native kernels too
        m.train()(M.synthetic_video(0, 2, 10, 8), abnormal_labels=torch.ones(1), normal_labels=torch.zeros(1))


def test_training_step_gradients_match_the_live_reference(golden_dir):
    """a23 groundwork: torch.autograd over the oracle's train-mode forward (dropout probabilities 0, BatchNorm on batch
    statistics) reproduces the live reference's ``loss.backward()`` -- per-parameter gradient digests from
    tests/golden/mgfn_train.npz -- and one Adam step (lr 1e-3, weight_decay 5e-4, src/runner.py:53-59) lands on the
    reference's updated parameters.  This is the oracle the backward kernels of the scoring head will be held to."""
    import os

    import numpy as np
    import torch

    from oracle import mgfn as MG

    g = np.load(os.path.join(golden_dir, "mgfn_train.npz"))
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running_" not in k and "num_batches" not in k)
          for k, v in MG.seeded_state_dict(0).items()}
    video = MG.synthetic_video(3, 4, 10, 32)
    out = MG.forward(video, sd, normal_labels=torch.zeros(2), abnormal_labels=torch.ones(2), training=True)
    np.testing.assert_allclose(out["loss"].detach().numpy(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(out["scores"].detach().numpy(), g["scores"], rtol=1e-4, atol=1e-6)
    names = [str(n) for n in g["param_names"]]
    params = [sd[n] for n in names]
    grads = torch.autograd.grad(out["loss"], params, allow_unused=False)
    worst = 0.0
    for n, p, gr in zip(names, params, grads):
        d = gr.detach().double().reshape(-1)
        got = np.array([float(d.norm()), float(d.sum())] + [float(v) for v in d[:6]])
        want = g[f"grad/{n}"]
        scale = max(want[0], 1e-12)
        err = np.abs(got - want).max() / scale
        worst = max(worst, err)
        assert err < 2e-3, (n, got, want)
    # one Adam step with L2 weight decay folded into the gradient (torch.optim.Adam semantics), first step: m = (1-b1) g,
    # v = (1-b2) g^2, bias-corrected -> delta = lr * g / (|g| + eps)
    lr, wd, eps = 1e-3, 5e-4, 1e-8
    for n, p, gr in zip(names, params, grads):
        gt = gr.detach().double() + wd * p.detach().double()
        new = (p.detach().double() - lr * gt / (gt.abs() + eps)).reshape(-1)
        got = np.array([float(new.norm()), float(new.sum())])
        want = g[f"adam/{n}"][:2]
        assert abs(got[0] - want[0]) <= 1e-4 * max(want[0], 1e-6) + 1e-7, (n, got, want)
    print(f"worst gradient digest error (relative to the gradient norm): {worst:.2e}")
