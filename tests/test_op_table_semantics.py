"""CPU: every backbone's op table + packed parameter blob, executed row by row in fp32 torch (tests/table_interp.py), against
the oracle's forward with the same bf16-rounded conv weights.  Catches what no GPU is needed to catch: a wrong slot, channel
slice, zero-widened temporary (BRANCH_PAD / CONCAT_PAD), fused-sibling split, fused pool flag, padding or folded BatchNorm."""
import pytest
import torch

from table_interp import run_table

TOL = 2e-5   # fp32 summation order + BatchNorm folded into scale / shift; a structural error is orders of magnitude above


def _bf16_weights(sd):
    return {k: (v.to(torch.bfloat16).float() if v.dim() == 5 else v) for k, v in sd.items()}


def _err(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _clip(seed, t=16, h=224, w=224):
    return torch.randn(1, 3, t, h, w, generator=torch.Generator().manual_seed(seed)).clamp(-2.0, 2.4444)


@pytest.mark.parametrize("variant", [{}, {"pad_branches": False}, {"fuse_siblings": False},
                                     {"pad_branches": False, "fuse_siblings": False}],
                         ids=["default", "nominal-widths", "three-launch-siblings", "plain"])
def test_inception_table_executes_to_the_oracle_features(variant):
    from anomaly_detection_on_video_b200.inception import BRANCH_PAD, CONCAT_PAD, InceptionI3d
    from oracle import inception as OI

    sd = OI.seeded_state_dict(0)
    m = InceptionI3d()
    m.load_state_dict(sd, strict=True)
    m.eval()
    for k, v in variant.items():
        setattr(m, k, v)
    ops, pk, _ = m._build_table()
    x = _clip(11)
    got = run_table(ops, pk.blob(), x)
    ref = OI.extract_features(x, _bf16_weights(sd))
    assert got.shape == ref.shape == (1, 1024) and not torch.isnan(got).any()
    assert _err(got, ref) <= TOL
    widths = {op.name: (op.cin, op.cout, op.dst_c_total) for op in ops if op.cout}
    if variant.get("pad_branches", True):
        # the widened tensors really are in this table (else the check above would be vacuous for them)
        assert widths["Mixed_4f.b3b"][0] == CONCAT_PAD[528] and widths["Mixed_4e.b3b"][2] == CONCAT_PAD[528]
        assert widths["Mixed_3b.b2b"][0] == BRANCH_PAD[16] and widths["Mixed_4e.b1b"][0] == BRANCH_PAD[144]
    else:
        assert widths["Mixed_4f.b3b"][0] == 528 and widths["Mixed_3b.b2b"][0] == 16 and widths["Mixed_4e.b1b"][0] == 144


@pytest.mark.parametrize("fuse_stem_pool,fuse_pool2", [(True, True), (True, False), (False, False)],
                         ids=["fused-pools", "stem-pool-only", "unfused"])
def test_i3res50_table_executes_to_the_oracle_features(fuse_stem_pool, fuse_pool2):
    from anomaly_detection_on_video_b200 import _lib
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from oracle import i3res50 as O

    sd = O.seeded_state_dict(0)
    m = I3Res50()
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.fuse_stem_pool, m.fuse_pool2 = fuse_stem_pool, fuse_pool2
    ops, pk, _ = m._build_table()
    n_t2 = sum(1 for op in ops if op.flags & _lib.VAD_FLAG_POOL_T2)
    assert n_t2 == int(fuse_stem_pool) + int(fuse_stem_pool and fuse_pool2)
    x = _clip(12, h=160, w=192)   # any size the strides divide; smaller than a crop to keep the CPU suite short
    got = run_table(ops, pk.blob(), x)
    ref, _ = O.forward(x, _bf16_weights(sd))
    assert _err(got, ref.reshape(1, -1)) <= TOL


def _tf32_weights(sd):
    def rna(v):   # round to nearest TF32, ties away from zero (cvt.rna.tf32.f32), like ParamPacker
        return ((v.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return {k: (rna(v.float()) if v.dim() == 5 else v) for k, v in sd.items()}


@pytest.mark.parametrize("backbone,planes", [("i3res50", True), ("i3res50", False), ("inception", False)],
                         ids=["i3res50-plane-stem", "i3res50-gather-stem", "inception"])
def test_tf32_mode_tables_execute_to_the_oracle_features(backbone, planes):
    """The TF32 precision mode builds its own tables (fp32 weights rounded to TF32, K padded to 32, unfused pools, the stem's
    folded-window weights in plane order when VAD_FLAG_STEM_PLANES)."""
    from anomaly_detection_on_video_b200 import _lib

    if backbone == "i3res50":
        from anomaly_detection_on_video_b200.i3d import I3Res50 as Model
        from oracle import i3res50 as O

        sd = O.seeded_state_dict(0)
        x = _clip(14, h=160, w=192)
        ref = lambda: O.forward(x, _tf32_weights(sd))[0].reshape(1, -1)   # noqa: E731
    else:
        from anomaly_detection_on_video_b200.inception import InceptionI3d as Model
        from oracle import inception as O

        sd = O.seeded_state_dict(0)
        x = _clip(15)
        ref = lambda: O.extract_features(x, _tf32_weights(sd))   # noqa: E731
    m = Model()
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.precision = "tf32"
    m.tf32_stem_planes = planes
    ops, pk, _ = m._build_table()
    assert not any(op.flags & _lib.VAD_FLAG_POOL_T2 for op in ops) and not any(op.dst1 for op in ops)
    assert bool(ops[0].flags & _lib.VAD_FLAG_STEM_PLANES) == (planes and backbone == "i3res50")
    got = run_table(ops, pk.blob(), x, tf32=True)
    assert _err(got, ref()) <= TOL


def test_ptv_i3d_r50_table_executes_to_the_oracle_features():
    from anomaly_detection_on_video_b200.ptv_resnet import I3D8x8R50
    from oracle import i3d_r50_ptv as R

    sd = R.seeded_state_dict(0)
    m = I3D8x8R50()
    m.load_state_dict(sd, strict=True)
    m.eval()
    ops, pk, _ = m._build_table()
    x = _clip(13, t=16, h=128, w=160)
    got = run_table(ops, pk.blob(), x)
    ref = R.forward(x, _bf16_weights(sd))
    assert _err(got, ref.reshape(1, -1)) <= TOL


def test_the_interpreter_notices_a_broken_table():
    """Negative control: one concat slice shifted by a channel leaves a NaN gap / overwrites a neighbour -> features differ."""
    from anomaly_detection_on_video_b200.inception import InceptionI3d
    from oracle import inception as OI

    sd = OI.seeded_state_dict(0)
    m = InceptionI3d()
    m.load_state_dict(sd, strict=True)
    m.eval()
    ops, pk, _ = m._build_table()
    x = _clip(11)
    ref = OI.extract_features(x, _bf16_weights(sd))
    victim = next(op for op in ops if op.name == "Mixed_4c.b2b")
    victim.dst_c_off += 1
    got = run_table(ops, pk.blob(), x)
    assert torch.isnan(got).any() or _err(got, ref) > 100 * TOL
