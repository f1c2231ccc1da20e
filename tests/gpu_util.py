"""Helpers shared by the -m gpu tests: run one conv / pool op through the C ABI and compare it with
fp32 torch ops on the CPU (operands rounded to bf16 first, so the only differences are accumulation
order and the final bf16 store)."""
import torch
import torch.nn.functional as F

from anomaly_detection_on_video_b200 import _lib as lib
from anomaly_detection_on_video_b200 import engine as eng


def run_conv_case(cin, cout, k, s, p, B, T, H, W, res=False, relu=True, force_gather=False, seed=0, device="cuda:0"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, T, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5).to(torch.bfloat16)
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.2 * torch.randn(cout, generator=g)
    y = F.conv3d(x.float(), w.float(), None, s, p) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if res:  # residual = the conv input (copied into slot 1 by an identity 1x1x1 max-pool)
        assert tuple(y.shape[2:]) == (T, H, W) and cout <= cin
        y = y + x.float()[:, :cout]
    if relu:
        y = F.relu(y)
    pk = eng.ParamPacker()
    w_off, s_off, b_off = pk.add_conv(w.float(), scale, shift)
    flags = (lib.VAD_FLAG_RELU if relu else 0) | (lib.VAD_FLAG_FORCE_GATHER if force_gather else 0)
    ops = []
    if res:
        ops.append(eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=(1, 1, 1), stride=(1, 1, 1)))
    ops.append(eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=2, res=1 if res else -1, cin=cin, cout=cout, kernel=k, stride=s,
                      pad=p, flags=flags, w_off=w_off, scale_off=s_off, shift_off=b_off))
    dev = torch.device(device)
    plan = eng.BackbonePlan(ops, pk.blob(), 3, 0, dev, in_channels=cin)
    plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(dev))
    torch.cuda.synchronize()
    out = plan.slot_tensor(2).float().cpu().permute(0, 4, 1, 2, 3).contiguous()
    return out, y


def assert_bf16_close(out, ref):
    """|out - ref| <= one bf16 ulp of the value (2^-7 relative) + fp32 accumulation-order slack."""
    ref_max = ref.abs().max().item()
    err = (out - ref).abs()
    bound = 2.0 ** -7 * ref.abs() + 2e-3 * ref_max
    bad = err > bound
    assert not bad.any(), (f"{int(bad.sum())} of {err.numel()} elements off; max err {err.max().item():.4g} "
                           f"(ref max {ref_max:.4g}); first {torch.nonzero(bad)[:4].tolist()}")
