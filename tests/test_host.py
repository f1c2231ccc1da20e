"""Host-side logic that needs no GPU: op table, BN folding, weight packing, work queue (gloo, world 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from anomaly_detection_on_video_b200 import _lib
from anomaly_detection_on_video_b200.engine import ParamPacker, fold_bn
from anomaly_detection_on_video_b200.i3d import I3Res50, build_i3d_feature_extractor
from oracle import i3res50 as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_match_reference_checkpoint_format():
    m = I3Res50()
    sd, ref = m.state_dict(), O.seeded_state_dict(0)
    assert set(sd) == set(ref)
    assert all(sd[k].shape == ref[k].shape for k in ref)
    missing, unexpected = m.load_state_dict(ref, strict=True)
    assert not missing and not unexpected
    assert sum(p.numel() for p in m.parameters()) == 27_223_872  # 27.22 M


def test_op_table_matches_appendix_a():
    ops = I3Res50().op_table()
    convs = [o for o in ops if o.kind == _lib.VAD_OP_CONV]
    assert len(ops) == 56 and len(convs) == 53
    assert [o.name for o in ops if o.kind == _lib.VAD_OP_MAXPOOL] == ["maxpool1", "maxpool2"]
    assert ops[-1].kind == _lib.VAD_OP_AVGPOOL
    stem = convs[0]
    assert stem.kernel == (5, 7, 7) and stem.stride == (2, 2, 2) and stem.pad == (2, 3, 3)
    assert stem.flags & _lib.VAD_FLAG_STEM_FOLD_W and stem.cin == 4 and stem.cout == 64
    by = {o.name: o for o in convs}
    assert by["layer2.0.conv2"].stride == (1, 2, 2) and by["layer2.0.conv2"].kernel == (1, 3, 3)
    assert by["layer2.0.downsample"].stride == (1, 2, 2) and not by["layer2.0.downsample"].flags & _lib.VAD_FLAG_RELU
    # temporal-conv pattern per stage (src/i3d.py:221,228,237,242)
    tk = lambda n: by[n].kernel[0]
    assert [tk(f"layer1.{b}.conv1") for b in range(3)] == [3, 3, 3]
    assert [tk(f"layer2.{b}.conv1") for b in range(4)] == [3, 1, 3, 1]
    assert [tk(f"layer3.{b}.conv1") for b in range(6)] == [3, 1, 3, 1, 3, 1]
    assert [tk(f"layer4.{b}.conv1") for b in range(3)] == [1, 3, 1]
    # every conv3 adds a residual and applies the ReLU after it (src/i3d.py:115-116)
    for o in convs:
        if o.name.endswith("conv3"):
            assert o.res >= 0 and o.flags & _lib.VAD_FLAG_RELU
        assert o.src != o.dst and o.dst != 0 and o.res != o.dst
        assert o.w_off % 128 == 0 and o.scale_off % 16 == 0 and o.shift_off % 16 == 0


def test_fold_bn_equals_batchnorm_eval():
    g = torch.Generator().manual_seed(0)
    bn = torch.nn.BatchNorm3d(16).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(16, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(16, generator=g))
        bn.running_mean.copy_(torch.randn(16, generator=g))
        bn.running_var.copy_(torch.rand(16, generator=g) + 0.5)
    x = torch.randn(2, 16, 3, 5, 5, generator=g)
    scale, shift = fold_bn(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    got = x * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    torch.testing.assert_close(got, bn(x), rtol=1e-5, atol=1e-5)


def test_param_packer_layout():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 16, 3, 1, 1, generator=g)  # K = 48 -> padded to 64
    pk = ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, torch.ones(8), torch.zeros(8))
    blob = pk.blob()
    assert w_off == 0 and s_off % 16 == 0 and b_off % 16 == 0
    wk = blob[w_off:w_off + 8 * 64 * 2].view(torch.bfloat16).view(8, 64).float()
    want = w.permute(0, 2, 3, 4, 1).reshape(8, 48).to(torch.bfloat16).float()  # K order (kt, kh, kw, cin)
    assert torch.equal(wk[:, :48], want) and torch.count_nonzero(wk[:, 48:]) == 0
    # folded stem: (kt, kh, 8 px, 4 ch) with zero weights for pixel 7 and channel 3
    ws = torch.randn(64, 3, 5, 7, 7, generator=g)
    pk2 = ParamPacker()
    off, _, _ = pk2.add_conv(ws, torch.ones(64), torch.zeros(64), fold_w=True)
    kpad = (5 * 7 * 32 + 63) // 64 * 64
    wf = pk2.blob()[off:off + 64 * kpad * 2].view(torch.bfloat16).view(64, kpad).float()
    wf = wf[:, :5 * 7 * 32].view(64, 5, 7, 8, 4)
    assert torch.equal(wf[:, :, :, :7, :3], ws.permute(0, 2, 3, 4, 1).to(torch.bfloat16).float())
    assert torch.count_nonzero(wf[:, :, :, 7]) == 0 and torch.count_nonzero(wf[..., 3]) == 0


def test_factory_builds_the_cli_default_backbone_with_pytorchvideo_names():
    """a16: ``i3d_8x8_r50`` (extract_features.py:34,46; src/i3d.py:339-350).  pytorchvideo is absent, so parity is unpinned; what
    can be held on the CPU: the parameter names / shapes a hub checkpoint carries, and the op table's MAC count."""
    from anomaly_detection_on_video_b200 import _lib
    from anomaly_detection_on_video_b200.ptv_resnet import I3D8x8R50
    from oracle import i3d_r50_ptv as R

    m = build_i3d_feature_extractor("i3d_8x8_r50", check_model_size=False)
    assert isinstance(m, I3D8x8R50) and m.feature_dim == 2048
    want = R.seeded_state_dict(0)
    got = m.state_dict()
    assert set(got) == set(want)
    assert all(tuple(got[k].shape) == tuple(want[k].shape) for k in want)
    m.load_state_dict(want, strict=True)
    ops = m.op_table()
    convs = [o for o in ops if o.kind == _lib.VAD_OP_CONV]
    assert len(convs) == 1 + 3 * 16 + 4 and ops[-1].kind == _lib.VAD_OP_AVGPOOL and ops[-1].kernel[0] == 4
    assert R.conv_macs() == 56_813_682_688


def test_factory_signature():
    with pytest.raises(AttributeError):
        build_i3d_feature_extractor("nope")
    m = build_i3d_feature_extractor("tushar-n-baseline", check_model_size=False)
    assert isinstance(m, I3Res50) and m.feature_dim == 2048


# ----------------------------------------------------------------------------- work queue, world_size 2, gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _wq_worker(rank, world, port, dynamic, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from anomaly_detection_on_video_b200.workqueue import WorkQueue

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = WorkQueue.from_process_group(dynamic=dynamic)
    mine = []
    for n_items in (37, 5):  # two phases, like extract() then segment()
        got = list(q.claim(n_items, tag=f"phase{n_items}"))
        # the "work": a deterministic function of the item only -> identical bytes whoever computes it
        for i in got:
            np.save(os.path.join(outdir, f"item_{n_items}_{i}.npy"), np.full(4, i * 3 + n_items, dtype=np.float32))
        mine.append(got)
        q.barrier()
    np.save(os.path.join(outdir, f"claims_{rank}.npy"), np.array([len(m) for m in mine]))
    dist.destroy_process_group()


@pytest.mark.parametrize("dynamic", [True, False])
def test_workqueue_two_ranks_cover_every_item_once(tmp_path, dynamic):
    world, port = 2, _free_port()
    mp.spawn(_wq_worker, args=(world, port, dynamic, str(tmp_path)), nprocs=world, join=True)
    for n_items in (37, 5):
        for i in range(n_items):
            a = np.load(tmp_path / f"item_{n_items}_{i}.npy")
            assert np.array_equal(a, np.full(4, i * 3 + n_items, dtype=np.float32))
    claims = sum(np.load(tmp_path / f"claims_{r}.npy") for r in range(world))
    assert list(claims) == [37, 5]  # nothing claimed twice


def _wq_env_worker(rank, world, port, outdir):
    """No process group, no trailing barrier: WorkQueue.from_env() hosts its own TCP store inside rank 0, rank 0 runs out
    of work long before rank 1 does, and close() is the only thing that keeps the store alive for rank 1's last claims."""
    sys.path.insert(0, ROOT)
    import time

    from anomaly_detection_on_video_b200.workqueue import WorkQueue

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    q = WorkQueue.from_env()
    assert q is not None and q.owns_store == (rank == 0)
    got = []
    try:
        for i in q.claim(12, tag="videos"):
            got.append(i)
            if rank == 1:
                time.sleep(0.15)  # the slow rank: still claiming after rank 0 has seen the end of the queue
    finally:
        q.close()
    np.save(os.path.join(outdir, f"env_claims_{rank}.npy"), np.array(got, dtype=np.int64))


def test_workqueue_store_owner_outlives_the_slow_rank(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_wq_env_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)  # raises if a rank dies
    claimed = np.concatenate([np.load(tmp_path / f"env_claims_{r}.npy") for r in range(world)])
    assert sorted(claimed.tolist()) == list(range(12))


def test_atomic_save_leaves_no_npy_temporaries(tmp_path):
    from anomaly_detection_on_video_b200.extract_features import _atomic_save

    path = str(tmp_path / "video_i3d.npy")
    _atomic_save(path, np.arange(6, dtype=np.float32))
    assert os.listdir(tmp_path) == ["video_i3d.npy"] and np.array_equal(np.load(path), np.arange(6, dtype=np.float32))
    # a temporary left behind by a killed rank is neither a .npy file nor visible to segment()'s listing
    open(tmp_path / ".video2_i3d.npy.123.tmp", "wb").close()
    assert [f for f in os.listdir(tmp_path) if f.endswith(".npy")] == ["video_i3d.npy"]


def test_workqueue_single_process_is_identity():
    from anomaly_detection_on_video_b200.workqueue import WorkQueue

    assert list(WorkQueue().claim(7)) == list(range(7))
    assert WorkQueue.from_env() is None or int(os.environ.get("WORLD_SIZE", "1")) > 1


def test_frame_level_metrics_match_the_reference_formula():
    """src/runner.py:62-73: np.repeat(preds, 16) against per-frame labels, ROC-AUC / PR-AUC via sklearn."""
    from sklearn.metrics import auc, roc_curve

    from anomaly_detection_on_video_b200.runner import frame_level_metrics

    rng = np.random.default_rng(0)
    outs = []
    for t in (5, 9):
        preds = rng.random(t).astype(np.float32)
        labels = (np.repeat(preds, 16) + 0.3 * rng.standard_normal(t * 16) > 0.6).astype(np.float32)
        outs.append({"preds": preds, "labels": labels})
    m = frame_level_metrics(outs, 16)
    p = np.repeat(np.concatenate([o["preds"] for o in outs]), 16)
    l = np.concatenate([o["labels"] for o in outs])
    fpr, tpr, _ = roc_curve(l.tolist(), p)
    assert abs(m["valid/rec_auc"] - auc(fpr, tpr)) < 1e-12 and 0.5 < m["valid/rec_auc"] <= 1.0
    with pytest.raises(ValueError):
        frame_level_metrics([{"preds": np.zeros(3, np.float32), "labels": np.zeros(40, np.float32)}], 16)


def test_feature_dataset_archive_layout(tmp_path):
    """build_feature_dataset on a local zip: normal / abnormal split by file name, lazily opened members."""
    import zipfile

    from anomaly_detection_on_video_b200.dataset import build_feature_dataset

    rng = np.random.default_rng(1)
    z = tmp_path / "train.zip"
    with zipfile.ZipFile(z, "w") as zf:
        for name in ("Abuse001_x264_i3d.npy", "Normal_Videos_003_x264_i3d.npy", "Normal_Videos_010_x264_i3d.npy"):
            path = tmp_path / name
            np.save(path, rng.standard_normal((10, 32, 8)).astype(np.float32))
            zf.write(path, arcname="train/" + name)
    ds = build_feature_dataset("train", local_path=str(tmp_path), filename="train.zip", device="cpu")
    assert len(ds["normal"]) == 2 and len(ds["abnormal"]) == 1
    assert ds["abnormal"].get_filename(0).startswith("Abuse") and ds["normal"].open(ds["normal"].values[ds["normal"].get_filename(0)]).shape == (10, 32, 8)
    with pytest.raises(RuntimeError, match="GPU only|no CPU fallback"):
        ds["normal"][0]
    with pytest.raises(RuntimeError, match="no network"):
        build_feature_dataset("train")


def test_bind_to_gpu_is_best_effort_without_nvml(monkeypatch):
    """hostaffinity.bind_to_gpu never raises and never changes the cpuset when it cannot learn the GPU's local CPUs."""
    import os

    from anomaly_detection_on_video_b200.hostaffinity import bind_to_gpu

    before = os.sched_getaffinity(0)
    info = bind_to_gpu(0)
    assert isinstance(info, dict) and "bound" in info and info["allowed_cpus"] == len(before)
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
    monkeypatch.setenv("VAD_NO_NUMA_BIND", "1")
    assert bind_to_gpu(0)["bound"] is False
    assert os.sched_getaffinity(0) == before


def test_tf32_weight_packing_rounds_to_tf32_and_pads_k_to_32():
    """ParamPacker.add_conv(tf32=True): fp32 [cout, K_pad32] rows, every value rounded to the nearest TF32 (10-bit mantissa,
    ties away from zero -- what cvt.rna.tf32.f32 does), (kt, kh, kw, cin) order with cin padded."""
    import torch

    from anomaly_detection_on_video_b200.engine import ParamPacker

    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 3, 1, 2, 2, generator=g)
    pk = ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, torch.ones(8), torch.zeros(8), cin_pad=4, tf32=True)
    blob = pk.blob()
    k, k_pad = 1 * 2 * 2 * 4, 32
    rows = blob[w_off:w_off + 8 * k_pad * 4].view(torch.float32).view(8, k_pad)
    assert s_off >= w_off + 8 * k_pad * 4 and b_off > s_off
    assert torch.all(rows[:, k:] == 0)
    packed = rows[:, :k].view(8, 1, 2, 2, 4)
    assert torch.all(packed[..., 3] == 0)
    want = w.permute(0, 2, 3, 4, 1)
    got = packed[..., :3]
    assert torch.all((got.contiguous().view(torch.int32) & 0x1FFF) == 0)          # low 13 mantissa bits clear
    assert torch.all((got - want).abs() <= want.abs() * 2.0 ** -11 + 1e-30)       # nearest, not truncated


def test_cli_precision_and_model_arguments():
    """--precision is validated before anything touches a device; without a GPU the loader still fails loudly."""
    import pytest
    import torch

    from anomaly_detection_on_video_b200 import extract_features as E

    with pytest.raises(ValueError, match="precision"):
        E.load_feature_extraction_model(precision="fp8")
    with pytest.raises(SystemExit):
        E.cli(["--outdir", "/tmp/unused", "--precision", "int8"])
    if not torch.cuda.is_available():
        for name in ("tushar-n-baseline", "inception-i3d"):
            with pytest.raises(RuntimeError, match="sm_100a"):
                E.load_feature_extraction_model(name, precision="tf32")
