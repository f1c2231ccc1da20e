"""-m gpu: ``i3d_8x8_r50`` (pytorchvideo's I3D-R50 as the reference configures it, src/i3d.py:339-350) on the native kernels
against ``oracle/i3d_r50_ptv.py``.  PARITY UNPINNED BY THE REFERENCE: pytorchvideo is third-party and absent; the oracle restates
its published architecture.  Tolerance as for I3Res50 (bf16 operands / activations, fp32 accumulate): error relative to the
feature vector's scale <= 1e-2, cosine >= 0.999."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(cuda_device):
    from anomaly_detection_on_video_b200.i3d import build_i3d_feature_extractor
    from oracle import i3d_r50_ptv as R

    m = build_i3d_feature_extractor("i3d_8x8_r50", check_model_size=False)
    m.load_state_dict(R.seeded_state_dict(0), strict=True)
    return m.eval().to(cuda_device)


@pytest.mark.parametrize("shape", [(2, 3, 8, 64, 64), (1, 3, 16, 112, 96), (1, 3, 16, 224, 224)], ids=lambda s: "x".join(map(str, s)))
def test_features_match_the_restated_architecture(model, cuda_device, shape):
    from oracle import i3d_r50_ptv as R

    x = torch.randn(*shape, generator=torch.Generator().manual_seed(3)).clamp(-2.0, 2.4444)
    want = R.forward(x, R.seeded_state_dict(0)).reshape(shape[0], -1).numpy().astype(np.float64)
    got = model(x.to(cuda_device))
    assert tuple(got.shape) == (shape[0], 2048, 1, 1, 1)
    got = got.reshape(shape[0], -1).cpu().numpy().astype(np.float64)
    for a, r in zip(got, want):
        err = np.abs(a - r).max() / np.abs(r).max()
        cos = float(a @ r / (np.linalg.norm(a) * np.linalg.norm(r)))
        print(f"max-normalised err {err:.2e}, cos {cos:.6f}")
        assert err <= 1e-2 and cos >= 0.999


def test_weighted_temporal_head_differs_from_a_plain_mean(model, cuda_device):
    """16 input frames -> 8 frames into the head: AvgPool3d((4,7,7), stride 1) + global mean weights the frames 1,2,3,4,4,3,2,1."""
    ops = model.op_table()
    assert ops[-1].kernel[0] == 4
    plan = model.plan(torch.device(cuda_device))
    x = torch.randn(1, 3, 16, 64, 64, generator=torch.Generator().manual_seed(5)).clamp(-2.0, 2.4444).to(cuda_device)
    feats = model(x).reshape(-1)
    last = plan.slot_tensor(ops[-1].src).float()          # [1, 8, 2, 2, 2048]
    assert last.shape[1] == 8
    w = torch.tensor([1, 2, 3, 4, 4, 3, 2, 1], dtype=torch.float32, device=last.device) / 20.0
    want = (last.mean(dim=(2, 3)) * w.view(1, 8, 1)).sum(dim=1).reshape(-1)
    torch.testing.assert_close(feats, want, rtol=1e-5, atol=1e-5)
    assert not torch.allclose(feats, last.mean(dim=(1, 2, 3)).reshape(-1), rtol=1e-3, atol=1e-4)
