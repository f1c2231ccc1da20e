"""-m gpu: every CUDA kernel against the oracle / fp32 torch ops, called through the C ABI."""
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess as P
from oracle import segment as S

pytestmark = pytest.mark.gpu

# (name, cin, cout, kernel, stride, pad, B, T, H, W, residual, relu)
CONV_CASES = [
    ("1x1 64->256 layer1.conv3", 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 13, 13, False, False),
    ("1x1 256->64", 256, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 13, 13, False, True),
    ("1x1 1024->512 two n-tiles", 1024, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 7, 7, False, False),
    ("1x1 tiny M=5", 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 1, 1, 5, False, False),
    ("1x1 n-tail 136", 128, 136, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 6, 6, False, False),
    ("1x1 256->256 +res", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 9, 9, True, True),
    ("t3 64->64 pad1", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 9, False, True),
    ("t3 64->64 pad1 +res", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 9, True, True),
    ("t3 256->64 K=768", 256, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 4, 9, 9, False, True),
    ("s3x3 64->64", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 3, 11, 13, False, True),
    ("s3x3 row 55", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 1, 2, 55, 55, False, True),
    ("s3x3 stride 2 odd", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 15, 15, False, True),
    ("1x1 stride 2 downsample", 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 2, 2, 15, 15, False, False),
    ("3x3x3 64->192 (Inception)", 64, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 4, 10, 10, False, True),
    ("t3 T=1 all-padding edges", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 3, 1, 5, 5, False, True),
    ("t3 T=2 256->128 (layer2.0.conv1)", 256, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 2, 9, 11, False, True),
    ("t3 T=4 64->64 row 55 ragged tile", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 4, 55, 3, False, True),
    ("t3 T=2 512->128 two stages deep", 512, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 2, 28, 28, False, False),
]
GATHER_ONLY = [
    ("cin 24 3x3x3", 24, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 4, 10, 10, False, True),
    ("cin 16 1x1 cout 48", 16, 48, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 9, 9, False, True),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("force_gather", [False, True], ids=["tma", "gather"])
def test_conv_matches_fp32_reference(cuda_device, case, force_gather):
    from gpu_util import assert_bf16_close, run_conv_case

    out, ref = run_conv_case(*case[1:], force_gather=force_gather)
    assert out.shape == ref.shape
    assert_bf16_close(out, ref)


@pytest.mark.parametrize("case", CONV_CASES[6:13], ids=[c[0] for c in CONV_CASES[6:13]])
def test_tma_im2col_and_gather_producers_agree_bitwise(cuda_device, case):
    """Both A-operand producers must deliver the same tile, so the outputs are bit-identical."""
    from gpu_util import run_conv_case

    a, ref = run_conv_case(*case[1:], force_gather=False)
    b, _ = run_conv_case(*case[1:], force_gather=True)
    if case[3] == (3, 1, 1) and not case[10]:
        # (3,1,1) convs without a residual run through the temporal-halo kernel, which contracts in (channel block, tap)
        # order instead of (tap, channel block): same products, different fp32 summation order
        from gpu_util import assert_bf16_close

        assert_bf16_close(a, ref)
        assert (a - b).abs().max() <= 2.0 ** -7 * ref.abs().max()
    else:
        assert torch.equal(a, b)


HALF_K_CASES = [
    # Cin % 64 == 32: 32-wide k-blocks (64-byte rows, SWIZZLE_64B) through TMA, four per pipeline stage
    ("3x3x3 96->128 (Mixed_3b.b1b)", 96, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), 2, 4, 14, 14, False, True),
    ("3x3x3 160->320, N tail (Mixed_4f.b1b)", 160, 320, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 2, 7, 7, False, True),
    ("3x3x3 32->96 (Mixed_3c.b2b)", 32, 96, (3, 3, 3), (1, 1, 1), (1, 1, 1), 2, 3, 9, 11, False, True),
    ("1x1 480->192, 15 k-blocks (Mixed_4b.b0)", 480, 192, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 14, 14, False, True),
    ("1x1 32->64, one k-block", 32, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 9, 9, False, False),
    ("1x3x3 stride 2 96->64", 96, 64, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 13, 13, False, True),
    # Cin % 32 == 16: 16-wide k-blocks (32-byte rows, SWIZZLE_32B), eight per pipeline stage
    ("3x3x3 16->32, 27 k-blocks (Mixed_3b.b2b)", 16, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), 2, 4, 14, 14, False, True),
    ("3x3x3 144->288, N tail (Mixed_4e.b1b)", 144, 288, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 2, 7, 7, False, True),
    ("1x1 528->256, 33 k-blocks (Mixed_4f.b0)", 528, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 14, 14, False, True),
    ("3x3x3 48->128 (Mixed_5c.b2b)", 48, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), 2, 2, 7, 7, False, False),
    ("1x1 16->64, one k-block", 16, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 9, 9, False, True),
]


@pytest.mark.parametrize("case", HALF_K_CASES, ids=[c[0] for c in HALF_K_CASES])
def test_conv_half_width_k_blocks(cuda_device, case):
    """TMA operands with BK = 32 for Cin % 64 == 32 and BK = 16 for Cin % 32 == 16; the cp.async gather producer (BK = 64
    over the same K order) is the cross-check: bit-identical."""
    from gpu_util import assert_bf16_close, run_conv_case

    tma, ref = run_conv_case(*case[1:])
    gather, _ = run_conv_case(*case[1:], force_gather=True)
    assert_bf16_close(tma, ref)
    assert torch.equal(tma, gather)


@pytest.mark.parametrize("relu", [False, True])
def test_single_k_block_unit_conv_staged_epilogue_is_bit_identical(cuda_device, monkeypatch, relu):
    """Plain 64 -> 64 1x1x1 convs (Inception's Conv3d_2b_1x1) leave through the staged TMA-store epilogue by default;
    VAD_EPI_UNIT_MAXK=0 keeps them on the register-direct one.  Same accumulators, same BN / ReLU / rounding: bit-identical,
    ragged last m-tile included."""
    from gpu_util import assert_bf16_close, run_conv_case

    case = (64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 3, 11, 13, False, relu)
    monkeypatch.delenv("VAD_EPI_UNIT_MAXK", raising=False)
    staged, ref = run_conv_case(*case)
    monkeypatch.setenv("VAD_EPI_UNIT_MAXK", "0")
    direct, _ = run_conv_case(*case)
    assert_bf16_close(staged, ref)
    assert torch.equal(staged, direct)


@pytest.mark.parametrize("case", GATHER_ONLY, ids=[c[0] for c in GATHER_ONLY])
def test_conv_narrow_channels(cuda_device, case):
    from gpu_util import assert_bf16_close, run_conv_case

    out, ref = run_conv_case(*case[1:])
    assert_bf16_close(out, ref)


@pytest.mark.parametrize("shape", [(1, 4, 32, 32), (2, 8, 64, 64), (1, 16, 224, 224), (1, 6, 50, 38)])
def test_stem_conv_folded_window(cuda_device, shape):
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from gpu_util import assert_bf16_close

    B, T, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 3, T, H, W, generator=g).clamp(-2, 2.44)
    w = (torch.randn(64, 3, 5, 7, 7, generator=g) * 0.045).to(torch.bfloat16).float()
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    xb = x.to(torch.bfloat16).float()
    ref = F.relu(F.conv3d(xb, w, None, (2, 2, 2), (2, 3, 3)) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True)
    ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=(5, 7, 7), stride=(2, 2, 2), pad=(2, 3, 3),
                  flags=lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W, w_off=w_off, scale_off=s_off, shift_off=b_off)]
    plan = eng.BackbonePlan(ops, pk.blob(), 2, 3, cuda_device)
    xs = eng.ingest_ncthw(x.to(cuda_device), 3)
    want_stem = torch.zeros(B, T, H, W + 8, 4)
    want_stem[:, :, :, 3:3 + W, :3] = xb.permute(0, 2, 3, 4, 1)
    assert torch.equal(xs.float().cpu(), want_stem), "ingest (fp32 NCTHW -> bf16 stem layout) is a pure relayout"
    plan.forward(xs)
    torch.cuda.synchronize()
    out = plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3)
    assert_bf16_close(out, ref)


@pytest.mark.parametrize("k,kps", [((5, 7, 7), "0"), ((5, 7, 7), "1"), ((7, 7, 7), "0"), ((3, 7, 7), "0"), ((2, 5, 7), "0")])
def test_stem_through_the_generic_kernel(cuda_device, k, kps, monkeypatch):
    """VAD_STEM_GENERIC=1: the folded stem window view through conv_umma_kernel<64, 32, 4> -- four 32-wide k-blocks per
    pipeline stage, with 35 / 49 / 21 / 10 k-blocks leaving a last stage of 3 / 1 / 1 / 2 -- and (VAD_KPS=1) one per stage.
    This is the path InceptionI3d's 7x7x7 stem takes."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from gpu_util import assert_bf16_close

    monkeypatch.setenv("VAD_STEM_GENERIC", "1")
    if kps != "0":
        monkeypatch.setenv("VAD_KPS", kps)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 8, 40, 36, generator=g).clamp(-2, 2.44)
    w = (torch.randn(64, 3, *k, generator=g) * 0.045).to(torch.bfloat16).float()
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    pad = (k[0] // 2, k[1] // 2, k[2] // 2)
    ref = F.relu(F.conv3d(x.to(torch.bfloat16).float(), w, None, (2, 2, 2), pad) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True)
    ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=k, stride=(2, 2, 2), pad=pad,
                  flags=lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W, w_off=w_off, scale_off=s_off, shift_off=b_off)]
    plan = eng.BackbonePlan(ops, pk.blob(), 2, 3, cuda_device)
    plan.forward(eng.ingest_ncthw(x.to(cuda_device), 3))
    torch.cuda.synchronize()
    assert_bf16_close(plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3), ref)


@pytest.mark.parametrize("shape", [(1, 16, 224, 224), (2, 9, 50, 38), (1, 4, 32, 32), (3, 5, 16, 8)])
def test_inception_stem_streams_its_weights(cuda_device, shape, monkeypatch):
    """The 7x7x7 / 2 SAME-padded stem of the Inception port through the dedicated stem kernel: its 49 taps (196 KB) do not
    fit in shared memory next to the pipeline, so the 7 taps of each frame tap travel with that frame's stage
    (StemParams::w_stream); padding is asymmetric (2 in front, 3 behind).  Cross-checked against the generic kernel
    (VAD_STEM_GENERIC=1), which accumulates in the same k order: bit-identical."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from gpu_util import assert_bf16_close
    from oracle.inception import _same_pad

    B, T, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 3, T, H, W, generator=g).clamp(-2, 2.44)
    k, s_ = (7, 7, 7), (2, 2, 2)
    w = (torch.randn(64, 3, *k, generator=g) * 0.04).to(torch.bfloat16).float()
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    ref = F.relu(F.conv3d(_same_pad(x.to(torch.bfloat16).float(), k, s_), w, None, s_) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True)
    ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=k, stride=s_,
                  flags=lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W | lib.VAD_FLAG_CONV_SAME, w_off=w_off, scale_off=s_off, shift_off=b_off)]
    outs = []
    # default: CTA pairs, each CTA keeping half of the output channels' taps resident (stem_umma_pair_kernel); VAD_STEM_NO_PAIR=1:
    # one CTA per tile with the weights streamed per stage; VAD_STEM_GENERIC=1: the generic implicit-GEMM kernel
    for generic, no_pair in (("0", "0"), ("0", "1"), ("1", "0")):
        monkeypatch.setenv("VAD_STEM_GENERIC", generic)
        monkeypatch.setenv("VAD_STEM_NO_PAIR", no_pair)
        plan = eng.BackbonePlan(ops, pk.blob(), 2, 2, cuda_device)
        plan.forward(eng.ingest_ncthw(x.to(cuda_device), 2))
        torch.cuda.synchronize()
        outs.append(plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3).clone())
    assert outs[0].shape == ref.shape
    assert_bf16_close(outs[0], ref)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("shape,slice_", [((1, 16, 224, 224), (0, 0)), ((2, 10, 64, 48), (0, 0)), ((1, 8, 50, 38), (64, 192))])
def test_stem_fused_temporal_pool_is_exact(cuda_device, shape, slice_):
    """VAD_FLAG_POOL_T2 (max over output frame pairs in the stem epilogue) == unfused stem followed by a
    (2,1,1)/(2,1,1) max-pool, bit for bit; an odd trailing frame is dropped like MaxPool3d's floor mode."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    B, T, H, W = shape
    c_off, c_total = slice_
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, T, H, W, generator=g).clamp(-2, 2.44)
    w = (torch.randn(64, 3, 5, 7, 7, generator=g) * 0.045).to(torch.bfloat16).float()
    scale, shift = 0.5 + torch.rand(64, generator=g), 0.2 * torch.randn(64, generator=g)
    scale[::7] *= -1.0  # negative BN scales: the max must come after scale/shift, not before
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True)
    xs = eng.ingest_ncthw(x.to(cuda_device), 3)

    def run(fused):
        flags = lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W | (lib.VAD_FLAG_POOL_T2 if fused else 0)
        ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=(5, 7, 7), stride=(2, 2, 2), pad=(2, 3, 3),
                      flags=flags, w_off=w_off, scale_off=s_off, shift_off=b_off, dst_c_off=c_off, dst_c_total=c_total)]
        plan = eng.BackbonePlan(ops, pk.blob(), 2, 3, cuda_device)
        plan.configure(B, T, H, W)
        plan.slot_tensor(1).fill_(-7.0)
        plan.forward(xs)
        torch.cuda.synchronize()
        return plan.slot_tensor(1).float().cpu()

    full, fused = run(False), run(True)
    To = full.shape[1]
    want = torch.maximum(full[:, 0:2 * (To // 2):2], full[:, 1:2 * (To // 2):2])
    assert fused.shape == want.shape
    assert torch.equal(fused, want)
    if c_total:
        assert (fused[..., :c_off] == -7.0).all() and (fused[..., c_off + 64:] == -7.0).all()


@pytest.mark.parametrize("name,C,k,s,T,H,W", [("maxpool1", 64, (2, 3, 3), (2, 2, 2), 8, 30, 30),
                                              ("maxpool1 odd", 64, (2, 3, 3), (2, 2, 2), 8, 112, 112),
                                              ("maxpool2", 256, (2, 1, 1), (2, 1, 1), 4, 11, 11)])
def test_maxpool_is_exact(cuda_device, name, C, k, s, T, H, W):
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    x = torch.randn(2, C, T, H, W, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    ref = F.max_pool3d(x.float(), k, s, 0)
    plan = eng.BackbonePlan([eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=k, stride=s)],
                            torch.zeros(16, dtype=torch.uint8), 2, 0, cuda_device, in_channels=C)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    assert torch.equal(plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3), ref)


def test_maxpool_same_padding_into_channel_slice(cuda_device):
    """MaxPool3dSamePadding (3x3x3, stride 1, zero pad) written into channels [64, 128) of a 192-wide slot."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    x = torch.randn(1, 64, 4, 7, 7, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    ref = F.max_pool3d(F.pad(x.float(), (1, 1, 1, 1, 1, 1)), (3, 3, 3), (1, 1, 1), 0)
    op = eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=(3, 3, 3), stride=(1, 1, 1), flags=lib.VAD_FLAG_POOL_SAME,
                dst_c_off=64, dst_c_total=192)
    plan = eng.BackbonePlan([op], torch.zeros(16, dtype=torch.uint8), 2, 0, cuda_device, in_channels=64)
    plan.configure(1, 4, 7, 7)
    plan.slot_tensor(1).fill_(-7.0)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    out = plan.slot_tensor(1).float().cpu()
    assert torch.equal(out[..., 64:128].permute(0, 4, 1, 2, 3), ref)
    assert (out[..., :64] == -7.0).all() and (out[..., 128:] == -7.0).all(), "neighbouring channel slices untouched"


@pytest.mark.parametrize("k,s,C,T,H,W", [((3, 3, 3), (1, 1, 1), 256, 8, 28, 28), ((3, 3, 3), (1, 1, 1), 64, 1, 5, 9), ((3, 3, 3), (1, 1, 1), 832, 2, 7, 7),
                                         ((1, 3, 3), (1, 2, 2), 64, 4, 28, 28), ((1, 3, 3), (1, 2, 2), 192, 3, 13, 15),
                                         ((3, 3, 3), (2, 2, 2), 480, 8, 28, 28), ((3, 3, 3), (2, 2, 2), 64, 5, 9, 11),
                                         ((2, 2, 2), (2, 2, 2), 832, 4, 14, 14), ((2, 2, 2), (2, 2, 2), 64, 3, 7, 5)])
@pytest.mark.parametrize("negative", [False, True], ids=["randn", "all-negative"])
@pytest.mark.parametrize("rows", [None, "0", "1"], ids=["default", "grid-stride", "row-per-block"])
def test_inception_same_padding_pools_are_exact(cuda_device, monkeypatch, k, s, C, T, H, W, negative, rows):
    """Every MaxPool3dSamePadding geometry of the Inception port through its specialised kernels (sliding 3-frame window for
    the 3x3x3 / 1 branch pools; for the strided ones both the grid-stride kernel with clamped taps and the row-per-block
    kernel, whichever the plan picks by default).  SAME padding pads with ZEROS, which only shows on all-negative inputs:
    border outputs are then 0, interior ones negative."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from oracle.inception import _same_pad

    if rows is None:
        monkeypatch.delenv("VAD_POOL_ROWS", raising=False)
    else:
        monkeypatch.setenv("VAD_POOL_ROWS", rows)

    x = torch.randn(2, C, T, H, W, generator=torch.Generator().manual_seed(8))
    if negative:
        x = -x.abs() - 0.25
    x = x.to(torch.bfloat16)
    ref = F.max_pool3d(_same_pad(x.float(), k, s), k, s, 0)
    op = eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=k, stride=s, flags=lib.VAD_FLAG_POOL_SAME)
    plan = eng.BackbonePlan([op], torch.zeros(16, dtype=torch.uint8), 2, 0, cuda_device, in_channels=C)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device))
    torch.cuda.synchronize()
    out = plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3)
    assert out.shape == ref.shape
    assert torch.equal(out, ref)


def test_avgpool(cuda_device):
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    x = torch.randn(3, 2048, 2, 7, 7, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
    plan = eng.BackbonePlan([eng.Op(kind=lib.VAD_OP_AVGPOOL, src=0)], torch.zeros(16, dtype=torch.uint8), 1, 0, cuda_device,
                            in_channels=2048)
    out = plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device)).cpu()
    torch.testing.assert_close(out, x.float().mean(dim=(2, 3, 4)), rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- preprocessing (bit-exact)
def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _frames(seed, n, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def test_preprocess_small_golden_bit_exact(cuda_device, golden_dir):
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    g = np.load(os.path.join(golden_dir, "preprocess_small.npz"))
    want = ((g["clips_u8"].astype(np.float32) - np.float32(114.75)) / np.float32(57.375)).astype(np.float32)
    ds = TenCropVideoFrameDataset(g["frames"], frames_per_clip=int(g["frames_per_clip"]), resize=int(g["resize"]),
                                  cropsize=int(g["crop"]), device=cuda_device)
    assert len(ds) == 2
    for ci in range(2):
        got = ds[ci]
        assert got.dtype == torch.float32 and tuple(got.shape) == want[ci].shape
        assert np.array_equal(got.cpu().numpy(), want[ci])


@pytest.mark.parametrize("tag,clips", [("ucf_240x320", (0, 2)), ("down_360x480", (0,)), ("portrait_300x256", (0,))])
def test_preprocess_fullsize_matches_reference_digest(cuda_device, golden_dir, tag, clips):
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    g = np.load(os.path.join(golden_dir, "preprocess_digests.npz"))
    seed, n, h, w = eval(str(g[f"{tag}/spec"]))
    ds = TenCropVideoFrameDataset(_frames(seed, n, h, w), device=cuda_device)
    for ci in clips:
        assert _sha(ds[ci].cpu().numpy()) == str(g[f"{tag}/clip{ci}"])


def test_preprocess_stem_layout_and_center_crop(cuda_device):
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    frames = _frames(4, 21, 120, 160)
    ds = TenCropVideoFrameDataset(frames, device=cuda_device)
    stem = ds.clips_stem(0, len(ds)).float().cpu().numpy()          # (2*10, 16, 224, 232, 4)
    for ci in range(len(ds)):
        ref = P.clip_tensor(frames, ci)
        want = torch.from_numpy(P.to_stem_layout(ref)).to(torch.bfloat16).float().numpy()
        assert np.array_equal(stem[ci * 10:(ci + 1) * 10], want)
    one = TenCropVideoFrameDataset(frames, device=cuda_device, ncrops=1)
    assert torch.equal(one[1][0], ds[1][4]), "center crop == TenCrop index 4"


@pytest.mark.parametrize("n,h,w,resize,crop", [(21, 240, 320, 256, 224), (17, 360, 480, 256, 224), (16, 300, 256, 256, 224),
                                               (19, 480, 854, 256, 224), (16, 256, 341, 256, 224), (5, 600, 800, 256, 224),
                                               (20, 60, 80, 64, 56), (9, 40, 30, 24, 16)],
                         ids=["ucf_240x320", "down_360x480", "portrait_300x256", "down_480x854", "identity_256x341", "generic_600x800",
                              "smoke_60x80_to_64_crop56", "tiny_40x30_to_24_crop16"])
def test_preprocess_stem_kernel_equals_the_dataset_path(cuda_device, n, h, w, resize, crop, monkeypatch):
    """The column-per-thread stem-layout kernel (3- and 5-tap resampling filters; 600 x 800 falls back to the generic kernel)
    == the fp32 dataset path -- pinned by the reference digests above -- rounded to bf16 and re-laid out, bit for bit,
    including the LoopPad tail clip, the zero pad columns and the zero fourth channel."""
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    frames = _frames(h * 7 + w, n, h, w)
    ds = TenCropVideoFrameDataset(frames, resize=resize, cropsize=crop, device=cuda_device)
    stem = ds.clips_stem(0, len(ds))                                  # (clips * 10, 16, crop, crop + 8, 4) bf16
    f32 = ds.clips_f32(0, len(ds))                                    # (clips, 10, 16, 3, crop, crop) fp32
    assert tuple(stem.shape) == (len(ds) * 10, 16, crop, crop + 8, 4)
    want = torch.zeros(stem.shape, dtype=torch.bfloat16, device=cuda_device)
    want[:, :, :, 3:3 + crop, :3] = f32.reshape(-1, 16, 3, crop, crop).permute(0, 1, 3, 4, 2).to(torch.bfloat16)
    assert torch.equal(stem, want)
    monkeypatch.setenv("VAD_K1_GENERIC", "1")
    assert torch.equal(ds.clips_stem(0, len(ds)), want)


def test_preprocess_properties_at_ucf_size(cuda_device):
    """Size-independent properties on a 2,000-frame UCF-Crime-shaped video (too big for the CPU oracle)."""
    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset

    frames = _frames(11, 2000 - 7, 240, 320)  # 1993 frames: last clip holds 9 frames
    ds = TenCropVideoFrameDataset(frames, device=cuda_device)
    assert len(ds) == 125
    # clip i of the video == clip 0 of its own 16 frames (frames are independent)
    sub = TenCropVideoFrameDataset(frames[16 * 77:16 * 78], device=cuda_device)
    assert torch.equal(ds[77], sub[0])
    # LoopPad: the short last clip == a clip built from its frames repeated cyclically
    tail = frames[16 * 124:]
    cyc = tail[np.arange(16) % len(tail)]
    assert torch.equal(ds[124], TenCropVideoFrameDataset(cyc, device=cuda_device)[0])
    # batched and per-clip paths agree; flipped crops are mirror images of the un-flipped block
    batch = ds.clips_f32(3, 4)
    assert torch.equal(batch[2], ds[5])
    c = ds[0]
    for flipped, plain in ((5, 1), (6, 0), (7, 3), (8, 2)):  # corner crops of the h-flipped image
        assert torch.equal(c[flipped], torch.flip(c[plain], dims=[-1]))


# ----------------------------------------------------------------------------- segment / magnitude
@pytest.mark.parametrize("n", [2, 5, 31, 32, 33, 47, 125, 188])
def test_segment_golden_bit_exact(cuda_device, golden_dir, n):
    from anomaly_detection_on_video_b200.engine import segment_mean

    g = np.load(os.path.join(golden_dir, "segment.npz"))
    out = segment_mean(torch.from_numpy(g[f"n{n}/in"]).to(cuda_device), 32).cpu().numpy()
    assert out.dtype == np.float32 and np.array_equal(out, g[f"n{n}/out"])


@pytest.mark.parametrize("n", [1, 2000, 12345])
def test_segment_large_matches_oracle_bit_exact(cuda_device, n):
    from anomaly_detection_on_video_b200.engine import segment_mean

    f = (np.random.default_rng(n).standard_normal((n, 10, 2048)) * 3).astype(np.float32)
    out = segment_mean(torch.from_numpy(f).to(cuda_device), 32).cpu().numpy()
    assert np.array_equal(out, S.segment_features(f, 32))


def test_add_magnitude(cuda_device):
    from anomaly_detection_on_video_b200.dataset import add_magnitude

    f = (np.random.default_rng(2).standard_normal((10, 32, 2048)) * 3).astype(np.float32)
    out = add_magnitude(torch.from_numpy(f).to(cuda_device)).cpu().numpy()
    ref = S.add_magnitude(f)
    assert out.shape == (10, 32, 2049) and np.array_equal(out[..., :2048], f)
    np.testing.assert_allclose(out[..., 2048], ref[..., 2048], rtol=1e-6)


@pytest.mark.parametrize("H,W,cin,cout", [(9, 7, 256, 256), (55, 55, 256, 256), (5, 6, 256, 128)])
def test_conv3_fused_temporal_pool_is_exact(cuda_device, H, W, cin, cout):
    """VAD_FLAG_POOL_T2 on a 1x1x1 residual conv (maxpool2 fused into layer1's last conv3): bit-identical to the unfused
    conv followed by a (2,1,1)/(2,1,1) max-pool, including ragged 32-pixel tiles at the end of a frame.  The residual is
    the conv input itself (copied into slot 1 by an identity 1x1x1 max-pool), so cout <= cin."""
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    g = torch.Generator().manual_seed(21)
    B, T = 2, 4
    x = torch.randn(B, cin, T, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 1, 1, 1, generator=g) * (2.0 / cin) ** 0.5).to(torch.bfloat16)
    scale, shift = 0.5 + torch.rand(cout, generator=g), 0.2 * torch.randn(cout, generator=g)
    scale[::5] *= -1.0
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w.float(), scale, shift)
    xg = x.permute(0, 2, 3, 4, 1).contiguous().to(cuda_device)
    outs = []
    for fused in (False, True):
        ops = [eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=(1, 1, 1), stride=(1, 1, 1)),
               eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=2, res=1, cin=cin, cout=cout, kernel=(1, 1, 1), stride=(1, 1, 1), pad=(0, 0, 0),
                      flags=lib.VAD_FLAG_RELU | (lib.VAD_FLAG_POOL_T2 if fused else 0), w_off=w_off, scale_off=s_off, shift_off=b_off)]
        if not fused:
            ops.append(eng.Op(kind=lib.VAD_OP_MAXPOOL, src=2, dst=3, kernel=(2, 1, 1), stride=(2, 1, 1)))
        plan = eng.BackbonePlan(ops, pk.blob(), 4, 0, cuda_device, in_channels=cin)
        plan.forward(xg)
        torch.cuda.synchronize()
        outs.append(plan.slot_tensor(2 if fused else 3).float().cpu())
    assert outs[0].shape == outs[1].shape == (B, 2, H, W, cout)
    assert torch.equal(outs[0], outs[1])
    ref = torch.relu(torch.nn.functional.conv3d(x.float(), w.float()) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1) + x.float()[:, :cout])
    ref = torch.nn.functional.max_pool3d(ref, (2, 1, 1), (2, 1, 1)).permute(0, 2, 3, 4, 1)
    from gpu_util import assert_bf16_close

    assert_bf16_close(outs[1], ref)


PAIR_BASE_CASES = [
    ("1x1 512->256 odd m-tiles", 512, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 14, 14),   # M = 392: 4 m-tiles -> 2 pairs
    ("1x1 256->512 tail pair", 256, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 13, 13),      # M = 338: 3 m-tiles -> odd CTA idles
    ("s3x3 256->256 im2col", 256, 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 2, 14, 14),
    ("t3 512->256 K=1536", 512, 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), 3, 2, 10, 13),
    ("1x1 stride 2 512->1024 ds", 512, 1024, (1, 1, 1), (1, 2, 2), (0, 0, 0), 2, 2, 28, 28),
]


PAIR_CASES = PAIR_BASE_CASES + [
    ("1x1 1024->512 many items", 1024, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 4, 4, 28, 28),   # 98 m-tiles x 2 n-tiles: 98 items on 74 pairs
    ("s3x3 stride 2 256->256", 256, 256, (1, 3, 3), (1, 2, 2), (0, 1, 1), 3, 4, 28, 28),
    ("1x1 2048->512 K=2048 single pair", 2048, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 1, 7, 7),  # M = 49: the odd CTA is all padding
    ("s3x3 128->128 (256 x 128 tiles)", 128, 128, (1, 3, 3), (1, 1, 1), (0, 1, 1), 4, 2, 28, 28),
    ("1x1 1088->128 odd k-blocks", 1088, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 13, 13),      # 17 k-blocks, two per stage
    ("s3x3 stride 2 128->128", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 3, 2, 28, 28),
]


SPLIT_CASES = [
    # cout == 256 (one n tile), 98 pair-tiles on 74 CTA pairs: 74 full-width items + 24 tiles as 48 half-width items
    ("1x1 256->256, 98 pair tiles", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 8, 4, 28, 28),
    ("s3x3 128->256 im2col, 98 pair tiles", 128, 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), 8, 4, 28, 28),
    ("1x1 512->256, odd m-tiles (197 -> 99 pairs)", 512, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 1, 158, 159),
]


@pytest.mark.parametrize("case", SPLIT_CASES, ids=[c[0] for c in SPLIT_CASES])
def test_cta_pair_tail_round_at_half_width(cuda_device, case, monkeypatch):
    """When the last round of 256 x 256 pair tiles fills at most half of the CTA pairs, its tiles run as two 128-column
    halves each (M = 256, N = 128 MMAs into the same accumulator columns); bit-identical to the unsplit schedule and to
    the single-CTA kernel."""
    from gpu_util import assert_bf16_close, run_conv_case

    monkeypatch.setenv("VAD_PAIR_MIN_KB", "1")
    monkeypatch.setenv("VAD_PAIR", "0")
    plain, ref = run_conv_case(*case[1:], False, True)
    monkeypatch.setenv("VAD_PAIR", "1")
    split, _ = run_conv_case(*case[1:], False, True)
    monkeypatch.setenv("VAD_NO_PAIR_SPLIT", "1")
    unsplit, _ = run_conv_case(*case[1:], False, True)
    assert_bf16_close(split, ref)
    assert torch.equal(split, plain) and torch.equal(unsplit, plain)


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_cta_pair_kernel_matches_the_plain_kernel(cuda_device, case, monkeypatch):
    """conv_pair_kernel (tcgen05 cta_group::2: two CTAs share one 256 x 256 tile, each loading half of the weight rows) is
    the default for the 128 x 256-tile layers; VAD_PAIR=0 routes them through the single-CTA kernel.  Same operand
    order per accumulator element, so the two are bit-identical."""
    from gpu_util import assert_bf16_close, run_conv_case

    monkeypatch.setenv("VAD_PAIR", "0")
    plain, ref = run_conv_case(*case[1:], False, True)
    monkeypatch.setenv("VAD_PAIR", "1")
    monkeypatch.setenv("VAD_PAIR_MIN_KB", "1")  # default: only layers with >= 12 k-blocks
    pair, _ = run_conv_case(*case[1:], False, True)
    assert_bf16_close(pair, ref)
    assert torch.equal(pair, plain)


PAIR_EPI_CASES = [
    ("1x1 256->256 +res, 4 k-blocks", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 14, 14),
    ("1x1 512->512 +res, odd m-tiles", 512, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 13, 13),      # 3 m-tiles: odd CTA of pair 2 is padding
    ("1x1 512->256 +res (wider residual)", 512, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 3, 2, 14, 14),
    ("1x1 256->512... many items", 512, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 4, 4, 28, 28),           # 98 m-tiles x 2 n-tiles on 74 pairs
    ("1x1 256->256 +res, M = 49", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 1, 7, 7),
]


@pytest.mark.parametrize("relu", [True, False], ids=["relu", "linear"])
@pytest.mark.parametrize("case", PAIR_EPI_CASES, ids=[c[0] for c in PAIR_EPI_CASES])
def test_cta_pair_residual_kernel_matches_the_plain_kernel(cuda_device, case, relu, monkeypatch):
    """conv_pair_kernel<256, 1, true>: residual layers with cout % 256 == 0 and >= 4 k-blocks run as 256 x 256 CTA-pair
    tiles whose halves pass through TMA-prefetched residual staging tiles; bit-identical to the single-CTA staged epilogue."""
    from gpu_util import assert_bf16_close, run_conv_case

    monkeypatch.setenv("VAD_PAIR", "0")
    plain, ref = run_conv_case(*case[1:], True, relu)
    monkeypatch.setenv("VAD_PAIR", "1")
    monkeypatch.setenv("VAD_PAIR_EPI_MIN_KB", "1")  # default: only layers with >= 8 k-blocks
    pair, _ = run_conv_case(*case[1:], True, relu)
    assert_bf16_close(pair, ref)
    assert torch.equal(pair, plain)


# ---------------------------------------------------------------------------------- bottleneck-tail fusion (conv_tail.cuh)
def _tail_plan(mode, B, T, H, W, seed, device):
    """layer1 bottleneck tail as its own op table: conv2 (1,3,3) 64->64 + BN + ReLU, conv3 1x1x1 64->256 + BN + residual +
    ReLU; the residual is a 256-channel tensor (mode 'res': produced by an earlier op) or the block's 1x1x1 downsample of
    the 64-channel block input (mode 'ds': reference src/i3d.py:262-272)."""
    from anomaly_detection_on_video_b200 import _lib as lib
    from anomaly_detection_on_video_b200 import engine as eng
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 64, T, H, W, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(64, 64, 1, 3, 3, generator=g) * (2.0 / 576) ** 0.5).to(torch.bfloat16)
    w3 = (torch.randn(256, 64, 1, 1, 1, generator=g) * (2.0 / 64) ** 0.5).to(torch.bfloat16)
    wd = (torch.randn(256, 64, 1, 1, 1, generator=g) * (2.0 / 64) ** 0.5).to(torch.bfloat16)
    sc = [0.5 + torch.rand(c, generator=g) for c in (64, 256, 256)]
    sh = [0.2 * torch.randn(c, generator=g) for c in (64, 256, 256)]
    pk = eng.ParamPacker()
    o2 = pk.add_conv(w2.float(), sc[0], sh[0])
    o3 = pk.add_conv(w3.float(), sc[1], sh[1])
    od = pk.add_conv(wd.float(), sc[2], sh[2])
    C = lib.VAD_OP_CONV
    one, zero = (1, 1, 1), (0, 0, 0)
    proj = eng.Op(kind=C, src=0, dst=1, cin=64, cout=256, kernel=one, stride=one, pad=zero, flags=0, w_off=od[0], scale_off=od[1], shift_off=od[2])
    conv2 = eng.Op(kind=C, src=0, dst=2, cin=64, cout=64, kernel=(1, 3, 3), stride=one, pad=(0, 1, 1), flags=lib.VAD_FLAG_RELU,
                   w_off=o2[0], scale_off=o2[1], shift_off=o2[2])
    conv3 = eng.Op(kind=C, src=2, dst=3, res=1, cin=64, cout=256, kernel=one, stride=one, pad=zero, flags=lib.VAD_FLAG_RELU,
                   w_off=o3[0], scale_off=o3[1], shift_off=o3[2])
    ops = [proj, conv2, conv3] if mode == "res" else [conv2, proj, conv3]
    plan = eng.BackbonePlan(ops, pk.blob(), 4, 0, torch.device(device), in_channels=64)
    xf = x.float()
    a2 = F.relu(F.conv3d(xf, w2.float(), None, 1, (0, 1, 1)) * sc[0].view(1, -1, 1, 1, 1) + sh[0].view(1, -1, 1, 1, 1)).to(torch.bfloat16).float()
    r = F.conv3d(xf, wd.float()) * sc[2].view(1, -1, 1, 1, 1) + sh[2].view(1, -1, 1, 1, 1)
    y = F.relu(F.conv3d(a2, w3.float()) * sc[1].view(1, -1, 1, 1, 1) + sh[1].view(1, -1, 1, 1, 1) + r)
    return plan, x.permute(0, 2, 3, 4, 1).contiguous().to(device), y


TAIL_SHAPES = [(2, 2, 20, 20), (1, 1, 5, 3), (3, 4, 55, 55), (9, 4, 55, 55)]


@pytest.mark.parametrize("shape", TAIL_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_bottleneck_tail_residual_fusion_is_bit_identical(cuda_device, shape, monkeypatch):
    """conv2 -> conv3 + residual in one launch == conv_s3x3 followed by the generic staged-residual conv3, bit for bit
    (same MMA order per accumulator, same epilogue arithmetic); both within bf16 resolution of fp32 torch."""
    from gpu_util import assert_bf16_close
    outs = []
    for no_tail in ("1", "0"):
        monkeypatch.setenv("VAD_NO_TAIL", no_tail)
        plan, x, ref = _tail_plan("res", *shape, seed=3, device=cuda_device)
        plan.forward(x)
        assert plan.num_launches == (3 if no_tail == "1" else 2)
        torch.cuda.synchronize()
        outs.append(plan.slot_tensor(3).clone())
    assert torch.equal(outs[0], outs[1])
    assert_bf16_close(outs[1].float().cpu().permute(0, 4, 1, 2, 3), ref)


@pytest.mark.parametrize("shape", TAIL_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_bottleneck_tail_with_folded_downsample(cuda_device, shape, monkeypatch):
    """First block of layer1: conv2 -> [conv3 | downsample] as one K = 128 contraction with the BN scales folded into
    bf16 weights.  Not bit-identical to the unfused path (the residual is no longer rounded to bf16, the weights carry
    the scale): both must sit within bf16 resolution of the fp32 torch result."""
    from gpu_util import assert_bf16_close
    errs = []
    for no_tail in ("1", "0"):
        monkeypatch.setenv("VAD_NO_TAIL", no_tail)
        plan, x, ref = _tail_plan("ds", *shape, seed=5, device=cuda_device)
        plan.forward(x)
        plan.forward(x)  # second forward: weights already folded, maps bound
        assert plan.num_launches == (3 if no_tail == "1" else 1)
        torch.cuda.synchronize()
        out = plan.slot_tensor(3).float().cpu().permute(0, 4, 1, 2, 3)
        assert_bf16_close(out, ref)
        errs.append((out - ref).abs().max().item())
    assert errs[1] <= 2.0 * errs[0] + 1e-3
