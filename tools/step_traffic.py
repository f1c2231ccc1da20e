"""Summarise tools/ncu_step_traffic.sh: DRAM bytes read + written by all kernels of one timed bench step -> profiles/step_dram_traffic.json
(keyed by a hash of the kernel sources, so that bench.py only reports it for the build it was measured on).

    python tools/step_traffic.py <tag>
"""
import csv, hashlib, json, os, re, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# the sources every kernel of the I3Res50 bf16 step is built from (the TF32 mode, the scoring head and the Inception-only stem
# live in other files and do not change this step's traffic)
STEP_SOURCES = ("ptx_sm100.cuh", "aux_kernels.cuh", "aux_api.cuh", "conv_umma.cuh", "stem_umma.cuh", "conv_thalo.cuh", "conv_s3x3.cuh",
                "conv_pair.cuh", "conv_tail.cuh", "plan_configure.cuh", "plan_bind.cuh", "plan_run.cuh")


def _code_only(text: str) -> str:
    """The source without comments and with runs of white space collapsed: editing a comment does not change what was measured."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return " ".join(text.split())


def source_hash(code_only: bool = True) -> str:
    h = hashlib.sha256()
    d = os.path.join(ROOT, "anomaly_detection_on_video_b200", "csrc")
    for f in STEP_SOURCES:
        raw = open(os.path.join(d, f), "rb").read()
        h.update(f.encode()); h.update(_code_only(raw.decode()).encode() if code_only else raw)
    return h.hexdigest()[:16]


if __name__ == "__main__":
    tag = sys.argv[1]
    rows, hdr = [], None
    for r in csv.reader(open(f"gpurun_out/{tag}_step_traffic.csv")):
        if r and r[0] == "ID":
            hdr = r
        elif r and r[0].isdigit():
            rows.append(r)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = OrderedDict()
    for r in rows:
        launches.setdefault(int(r[0]), {"name": r[ki]})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
    ids = list(launches)
    seg = [i for i in ids if "segment_mean" in launches[i]["name"]]
    # one whole step = the launches after one segment mean up to and including the next; the last such pair is the timed step
    lo, hi = (seg[-2] + 1, seg[-1] + 1) if len(seg) >= 2 else (ids[0], ids[-1] + 1)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = wr = ns = 0.0
    per = OrderedDict()
    for i in ids:
        if lo <= i < hi:
            L = launches[i]
            a = L["dram__bytes_read.sum"]; b = L["dram__bytes_write.sum"]; t = L["gpu__time_duration.sum"]
            r_, w_ = a[0] * scale[a[1]], b[0] * scale[b[1]]
            rd += r_; wr += w_; ns += t[0] * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(t[1], 1.0)
            k = L["name"].split("(")[0].replace("void ", "").replace("vad::", "")
            p = per.setdefault(k, [0, 0.0])
            p[0] += 1; p[1] += r_ + w_
    out = {"source_hash": source_hash(), "capture": f"{tag}_step_traffic (ncu --metrics dram__bytes_read/write.sum, one bench step, launches {lo}..{hi - 1})",
           "launches": sum(p[0] for p in per.values()), "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_step": rd + wr,
           "kernel_time_ms_serialised": ns / 1e6,
           "per_kernel_gb": {k: {"launches": v[0], "gb": round(v[1] / 1e9, 3)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}}
    json.dump(out, open(os.path.join(ROOT, "profiles", "step_dram_traffic.json"), "w"), indent=1)
    print(json.dumps(out)[:600])
