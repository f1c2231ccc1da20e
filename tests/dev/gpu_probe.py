"""GPU bring-up probe: runs each section in its own subprocess (a CUDA fault in one section must not
poison the next) and writes a log per section under gpurun_out/probe/.

    python tools/gpu_probe.py            # all sections
    python tools/gpu_probe.py conv_unit  # one section, in-process
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SECTIONS = ["conv_tma2d", "conv_im2col", "conv_gather", "conv_stem", "pools", "preproc", "segment", "net_small",
            "net_full", "timing"]


def _conv_case(torch, F, eng, lib, name, cin, cout, k, s, p, B, T, H, W, res, relu, force_gather, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, T, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5).to(torch.bfloat16)
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = 0.2 * torch.randn(cout, generator=g)
    y = F.conv3d(x.float(), w.float(), None, s, p) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if res:
        # residual = the conv input itself (identity 1x1x1 max-pool copies slot 0 into slot 1);
        # needs an output of the input's shape and cout <= cin
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
    dev = torch.device("cuda")
    plan = eng.BackbonePlan(ops, pk.blob(), 3, 0, dev, in_channels=cin)
    xg = x.permute(0, 2, 3, 4, 1).contiguous().to(dev)
    plan.forward(xg)
    torch.cuda.synchronize()
    out = plan.slot_tensor(2).float().cpu().permute(0, 4, 1, 2, 3)
    err = (out - y).abs()
    ref_max = y.abs().max().item()
    tol = 2 ** -8 * ref_max + 1e-3
    bad = (err > (2 ** -7 * y.abs() + 2e-3 * ref_max)).sum().item()
    rec = {"name": name, "max_err": err.max().item(), "ref_max": ref_max, "n_bad": bad, "n": err.numel(),
           "ok": bool(err.max().item() <= tol * 2 and bad == 0)}
    if not rec["ok"]:
        idx = torch.nonzero(err > (2 ** -7 * y.abs() + 2e-3 * ref_max))[:8].tolist()
        rec["first_bad"] = [(i, out[tuple(i)].item(), y[tuple(i)].item()) for i in idx]
    return rec


def section_conv(mode: str):
    import torch
    import torch.nn.functional as F
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    cases = []
    if mode == "conv_tma2d":
        cases = [
            ("1x1 64->256 M=2*4*13*13", 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 13, 13, False, False),
            ("1x1 256->64 relu", 256, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 13, 13, False, True),
            ("1x1 512->128", 512, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), 3, 2, 9, 9, False, True),
            ("1x1 1024->512 multi-ntile", 1024, 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 7, 7, False, False),
            ("1x1 64->64 tiny M=5", 64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 1, 1, 5, False, False),
            ("1x1 128->136 n-tail", 128, 136, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 6, 6, False, False),
            ("1x1 256->256 res relu", 256, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 2, 9, 9, True, True),
        ]
        fg = False
    elif mode == "conv_im2col":
        cases = [
            ("t3 64->64 pad1", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 9, False, True),
            ("s3x3 64->64 pad1", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 3, 11, 13, False, True),
            ("s3x3 128->128 stride2 odd", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 15, 15, False, True),
            ("1x1 stride2 256->512", 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 2, 2, 15, 15, False, False),
            ("t3 256->64 K=768", 256, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 1, 4, 9, 9, False, True),
            ("3x3x3 64->192", 64, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 4, 10, 10, False, True),
            ("s3x3 64->64 wide row 55", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 1, 2, 55, 55, False, True),
            ("t3 64->64 pad1 res", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 9, True, True),
        ]
        fg = False
    else:
        cases = [
            ("g 1x1 64->256", 64, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), 2, 4, 13, 13, False, False),
            ("g t3 64->64", 64, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 7, 9, False, True),
            ("g s3x3 128->128 s2", 128, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), 2, 2, 15, 15, False, True),
            ("g 1x1 s2 256->512", 256, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 2, 2, 15, 15, False, False),
            ("g cin24 3x3x3 24->64", 24, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), 1, 4, 10, 10, False, True),
            ("g cin16 1x1 16->48", 16, 48, (1, 1, 1), (1, 1, 1), (0, 0, 0), 1, 2, 9, 9, False, True),
        ]
        fg = True
    results = []
    for c in cases:
        name = c[0]
        try:
            rec = _conv_case(torch, F, eng, lib, name, *c[1:], force_gather=fg)
        except Exception as e:  # noqa: BLE001
            rec = {"name": name, "error": repr(e)}
        print(json.dumps(rec), flush=True)
        results.append(rec)
    return results


def section_conv_stem():
    import torch
    import torch.nn.functional as F
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    res = []
    for (B, T, H, W) in [(1, 4, 32, 32), (2, 8, 64, 64), (1, 16, 224, 224)]:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, 3, T, H, W, generator=g).clamp(-2, 2.44)
        w = (torch.randn(64, 3, 5, 7, 7, generator=g) * (2.0 / (64 * 245)) ** 0.5 * 4).to(torch.bfloat16).float()
        scale = 0.5 + torch.rand(64, generator=g)
        shift = 0.2 * torch.randn(64, generator=g)
        xb = x.to(torch.bfloat16).float()
        y = F.relu(F.conv3d(xb, w, None, (2, 2, 2), (2, 3, 3)) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
        pk = eng.ParamPacker()
        w_off, s_off, b_off = pk.add_conv(w, scale, shift, fold_w=True)
        ops = [eng.Op(kind=lib.VAD_OP_CONV, src=0, dst=1, cin=4, cout=64, kernel=(5, 7, 7), stride=(2, 2, 2), pad=(2, 3, 3),
                      flags=lib.VAD_FLAG_RELU | lib.VAD_FLAG_STEM_FOLD_W, w_off=w_off, scale_off=s_off, shift_off=b_off)]
        plan = eng.BackbonePlan(ops, pk.blob(), 2, 3, torch.device("cuda"))
        xs = eng.ingest_ncthw(x.cuda(), 3)
        # check the ingest layout itself
        xs_ref = torch.zeros(B, T, H, W + 8, 4)
        xs_ref[:, :, :, 3:3 + W, :3] = xb.permute(0, 2, 3, 4, 1)
        ing_ok = bool(torch.equal(xs.float().cpu(), xs_ref))
        plan.forward(xs)
        torch.cuda.synchronize()
        out = plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3)
        err = (out - y).abs()
        rec = {"name": f"stem {B}x{T}x{H}x{W}", "ingest_ok": ing_ok, "max_err": err.max().item(), "ref_max": y.abs().max().item(),
               "ok": bool(err.max().item() <= 2 ** -7 * y.abs().max().item() + 1e-3)}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    return res


def section_pools():
    import torch
    import torch.nn.functional as F
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng

    res = []
    dev = torch.device("cuda")
    for name, C, k, s, T, H, W in [("maxpool1", 64, (2, 3, 3), (2, 2, 2), 8, 30, 30), ("maxpool2", 256, (2, 1, 1), (2, 1, 1), 4, 11, 11)]:
        x = torch.randn(2, C, T, H, W, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
        y = F.max_pool3d(x.float(), k, s, 0)
        ops = [eng.Op(kind=lib.VAD_OP_MAXPOOL, src=0, dst=1, kernel=k, stride=s)]
        plan = eng.BackbonePlan(ops, torch.zeros(16, dtype=torch.uint8), 2, 0, dev, in_channels=C)
        plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(dev))
        torch.cuda.synchronize()
        out = plan.slot_tensor(1).float().cpu().permute(0, 4, 1, 2, 3)
        rec = {"name": name, "exact": bool(torch.equal(out, y))}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    # avgpool
    x = torch.randn(3, 2048, 2, 7, 7, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
    y = x.float().mean(dim=(2, 3, 4))
    ops = [eng.Op(kind=lib.VAD_OP_AVGPOOL, src=0)]
    plan = eng.BackbonePlan(ops, torch.zeros(16, dtype=torch.uint8), 1, 0, dev, in_channels=2048)
    out = plan.forward(x.permute(0, 2, 3, 4, 1).contiguous().to(dev)).cpu()
    rec = {"name": "avgpool", "max_err": (out - y).abs().max().item(), "ok": bool((out - y).abs().max().item() < 1e-5)}
    print(json.dumps(rec), flush=True)
    res.append(rec)
    return res


def section_preproc():
    import numpy as np
    import torch
    from anomaly_detection_on_video_b200 import _lib as lib, engine as eng
    from oracle import preprocess as P

    res = []
    rng = np.random.default_rng(0)
    for (n, h, w) in [(37, 240, 320), (5, 360, 480), (3, 300, 256)]:
        frames = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        pp = eng.Preprocessor(h, w, 256, 224, 10, torch.device("cuda"))
        fr = torch.from_numpy(frames).cuda()
        nclips = P.n_clips(n)
        out = pp.run(fr, 0, nclips, 16, lib.VAD_OUT_DATASET_F32).cpu().numpy()
        stem = pp.run(fr, 0, nclips, 16, lib.VAD_OUT_STEM_BF16, 3).float().cpu().numpy()
        ok_all, ok_stem = True, True
        for ci in sorted({0, nclips - 1}):
            ref = P.clip_tensor(frames, ci)
            ok_all &= bool(np.array_equal(ref, out[ci]))
            sref = torch.from_numpy(P.to_stem_layout(ref)).to(torch.bfloat16).float().numpy()
            ok_stem &= bool(np.array_equal(sref, stem[ci * 10:(ci + 1) * 10]))
        rec = {"name": f"preproc {n}x{h}x{w}", "info": pp.info(), "dataset_f32_bit_exact": ok_all, "stem_bf16_exact": ok_stem}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    return res


def section_segment():
    import numpy as np
    import torch
    from anomaly_detection_on_video_b200 import engine as eng
    from oracle import segment as S

    res = []
    rng = np.random.default_rng(1)
    for n in [1, 5, 31, 32, 33, 47, 125, 188, 2000]:
        f = (rng.standard_normal((n, 10, 2048)) * 3).astype(np.float32)
        out = eng.segment_mean(torch.from_numpy(f).cuda(), 32).cpu().numpy()
        ref = S.segment_features(f, 32)
        rec = {"name": f"segment n={n}", "bit_exact": bool(np.array_equal(out, ref))}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    f = (rng.standard_normal((10, 32, 2048)) * 3).astype(np.float32)
    out = eng.add_magnitude(torch.from_numpy(f).cuda()).cpu().numpy()
    ref = S.add_magnitude(f)
    rec = {"name": "add_magnitude", "max_rel": float(np.abs(out - ref).max() / np.abs(ref).max())}
    print(json.dumps(rec), flush=True)
    res.append(rec)
    return res


def _net(shape, force_gather=False):
    import torch
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from oracle import i3res50 as O

    sd = O.seeded_state_dict(0)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).clamp(-2, 2.44)
    taps = ["conv1", "maxpool1", "layer1", "maxpool2", "layer2", "layer3", "layer4"]
    yr, tr = O.forward(x, sd, taps=taps)
    yb, tb = O.forward(x, sd, emulate_bf16=True, taps=taps)
    m = I3Res50()
    m.load_state_dict(sd)
    m.eval().cuda()
    m.force_gather = force_gather
    y = m(x.cuda())
    torch.cuda.synchronize()
    y = y.float().cpu()
    rec = {"name": f"net {shape} gather={force_gather}", "shape": list(y.shape)}
    yr2, yb2, y2 = yr.flatten(1), yb.flatten(1), y.flatten(1)
    rec["max_norm_err_vs_fp32"] = ((y2 - yr2).abs().max() / yr2.abs().max()).item()
    rec["max_norm_err_vs_bf16emu"] = ((y2 - yb2).abs().max() / yr2.abs().max()).item()
    rec["rel_l2_vs_fp32"] = ((y2 - yr2).norm() / yr2.norm()).item()
    rec["cos_min"] = torch.nn.functional.cosine_similarity(y2, yr2, dim=1).min().item()
    # per-stage activations (slots are reused, so only the last stage survives in the workspace; report the final one)
    print(json.dumps(rec), flush=True)
    return [rec]


def section_timing():
    import torch
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from oracle import i3res50 as O

    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0))
    m.eval().cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    res = []
    for B in [16, 160]:
        xs = torch.randn(B, 16, 224, 232, 4, device=dev).to(torch.bfloat16)
        for _ in range(2):
            m.forward_stem_layout(xs)
        torch.cuda.synchronize()
        plan = m.plan(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            m.forward_stem_layout(xs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec = {"name": f"forward B={B}", "ms": ms, "clips_per_s": B / ms * 1e3, "tflops": plan.flops / ms / 1e9}
        print(json.dumps(rec), flush=True)
        res.append(rec)
        if B == 160:
            plan.profile_begin()
            for _ in range(3):
                m.forward_stem_layout(xs)
            prof = plan.profile_end()
            for p in prof:
                if not p["calls"]:
                    continue
                t = p["ms"] / p["calls"]
                print("%-22s %8.3f ms  %7.1f TFLOP/s  %7.1f GB/s(alg)" % (p["name"], t, p["flops"] / p["calls"] / t / 1e9,
                                                                      p["bytes"] / p["calls"] / t / 1e6), flush=True)
    # host link
    h = torch.empty(460_800_000, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        t0 = time.perf_counter()
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(json.dumps({"name": "h2d pinned 460.8MB", "ms": dt * 1e3, "GB/s": 0.4608 / dt}), flush=True)
    return res


def run_section(name: str):
    if name.startswith("conv_") and name != "conv_stem":
        return section_conv(name)
    if name == "conv_stem":
        return section_conv_stem()
    if name == "pools":
        return section_pools()
    if name == "preproc":
        return section_preproc()
    if name == "segment":
        return section_segment()
    if name == "net_small":
        return _net((1, 3, 8, 64, 64)) + _net((2, 3, 8, 64, 64), force_gather=True)
    if name == "net_full":
        return _net((2, 3, 16, 224, 224))
    if name == "timing":
        return section_timing()
    if name == "e2e":
        return section_e2e_breakdown()
    raise SystemExit(f"unknown section {name}")


def main():
    if len(sys.argv) > 1:
        run_section(sys.argv[1])
        return
    outdir = os.path.join(ROOT, "gpurun_out", "probe")
    os.makedirs(outdir, exist_ok=True)
    summary = {}
    for s in SECTIONS:
        t0 = time.time()
        log = os.path.join(outdir, s + ".log")
        with open(log, "w") as f:
            try:
                rc = subprocess.run([sys.executable, os.path.abspath(__file__), s], stdout=f, stderr=subprocess.STDOUT,
                                    timeout=420, cwd=ROOT).returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        summary[s] = {"rc": rc, "sec": round(time.time() - t0, 1)}
        print(s, summary[s], flush=True)
        with open(log) as f:
            tail = f.read()[-3000:]
        print(tail, flush=True)
    with open(os.path.join(outdir, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)



def section_e2e_breakdown():
    """Where does the end-to-end (host frames -> host features) step spend its time?"""
    import numpy as np
    import torch

    from anomaly_detection_on_video_b200.dataset import TenCropVideoFrameDataset
    from anomaly_detection_on_video_b200.engine import segment_mean
    from anomaly_detection_on_video_b200.extract_features import extract_clip_features
    from anomaly_detection_on_video_b200.i3d import I3Res50
    from oracle import i3res50 as O

    dev = torch.device("cuda", 0)
    m = I3Res50()
    m.load_state_dict(O.seeded_state_dict(0))
    m.eval().to(dev)
    frames = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2000, 240, 320, 3), dtype=np.uint8)).pin_memory()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(3):
        t0 = sync()
        ds = TenCropVideoFrameDataset(frames, device=dev)
        t1 = sync()
        feats = extract_clip_features(ds, m, dev, strict_compat=False, as_numpy=False)
        t2 = sync()
        seg = segment_mean(feats, 32)
        t3 = sync()
        fh, sh = feats.cpu(), seg.cpu()
        t4 = sync()
        del ds
        t5 = sync()
        print(json.dumps({"iter": it, "dataset_ctor_h2d_ms": (t1 - t0) * 1e3, "extract_ms": (t2 - t1) * 1e3,
                          "segment_ms": (t3 - t2) * 1e3, "d2h_ms": (t4 - t3) * 1e3, "free_ms": (t5 - t4) * 1e3}), flush=True)

if __name__ == "__main__":
    main()
