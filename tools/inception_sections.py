"""Summarise tools/ncu_inception.sh: one row per launch of an InceptionI3d forward (op order) -> profiles/<tag>_ncu_sections.md.

    python tools/inception_sections.py <tag>        # reads gpurun_out/<tag>_raw.csv; runs on CPU (the op table needs no GPU)
"""
import csv, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from anomaly_detection_on_video_b200.inception import InceptionI3d


def short(name: str) -> str:
    m = re.match(r"(?:void )?(?:vad::)?(\w+?)(?:_kernel)?(<.*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


if __name__ == "__main__":
    tag = sys.argv[1]
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_raw.csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    ops = [op.name for op in InceptionI3d().op_table() if op.name]
    # the avg-pool head is one more launch after the op table's convs and pools
    names = ops + ["avgpool"] * max(0, len(data) - len(ops))
    out = [f"# {tag}: ncu section pass over one InceptionI3d forward (160 clip-crops), `tools/ncu_inception.sh`\n"]
    tot = sum(float(r[col["gpu__time_duration.sum"]]) for r in data)
    out.append(f"{len(data)} launches, {tot:.2f} ms of kernel time (serialised under ncu, `--clock-control none`).\n")
    out.append("| op | kernel | duration [us] | tensor pipe active % | DRAM TB/s | DRAM % | L2 % |\n|---|---|---|---|---|---|---|")
    for n, r in zip(names, data):
        tbs = float(r[col["dram__bytes.sum.per_second"]])
        if units[col["dram__bytes.sum.per_second"]] == "Gbyte/s":
            tbs /= 1e3
        out.append("| %s | `%s` | %.1f | %.1f | %.2f | %.1f | %.1f |" % (
            n, short(r[col["Kernel Name"]]), float(r[col["gpu__time_duration.sum"]]) * 1e3,
            float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]), tbs,
            float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
            float(r[col["lts__throughput.avg.pct_of_peak_sustained_elapsed"]])))
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_sections.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:8]))
