"""Turn the raw ncu CSV pages written by tools/ncu_round.sh into a small markdown summary under profiles/.

    python tools/ncu_summarize.py <tag> [out.md]
"""
import csv
import json
import os
import sys
from collections import OrderedDict

tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else f"profiles/{tag}_ncu_summary.md"
G = "gpurun_out"
lines = [f"# {tag}: ncu evidence (tools/ncu_round.sh)", ""]

# ---- launch list: one timed bench step
path = f"{G}/{tag}_launches.csv"
if os.path.exists(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    hdr = None
    for r in csv.reader(open(path)):
        if r and r[0] == "ID":
            hdr = r
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi = hdr.index("Grid Size")
    names = [r[ki] for r in rows]
    # one step = the launches between two consecutive segment_mean kernels; take the step before the last (timed) one
    seg = [i for i, n in enumerate(names) if "segment_mean" in n]
    lo, hi = (seg[2] + 1, seg[3] + 1) if len(seg) >= 4 else (0, len(rows))
    step = rows[lo:hi]
    agg = OrderedDict()
    for r in step:
        short = r[ki].split("(")[0].replace("void ", "").replace("vad::", "")
        k = short
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3  # ns -> us
    tot = sum(a[1] for a in agg.values())
    lines += [f"## Launch list of one bench step (`python bench.py --steps 1 --warmup 3 --no-cpu-baseline`, "
              f"launches {lo}..{hi - 1} of {len(rows)} captured)", "",
              f"{len(step)} launches, {tot / 1e3:.2f} ms of kernel time (serialised, cold-cache: compare shares).", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% |")
    lines.append("")

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem throughput %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__cycles_active.avg", "SMSP active cycles"), ("sm__cycles_elapsed.max", "SM cycles elapsed")]
for part in ("stem", "conv_l1", "conv_l3", "aux"):
    path = f"{G}/{tag}_{part}_raw.csv"
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    lines += [f"## `--set full` capture: {part} (`python tools/ncu_target.py`, 160 clip-crops)", ""]
    cols = [(h, lab) for h, lab in WANT if h in hdr]
    lines.append("| kernel | " + " | ".join(f"{lab} [{units[hdr.index(h)]}]" for h, lab in cols) + " |")
    lines.append("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("vad::", "")
        lines.append(f"| `{name}` #{d['ID']} | " + " | ".join(d[h] for h, _ in cols) + " |")
    lines.append("")
# per-launch DRAM traffic of the dominant kernel (the stem) for bench.py's roofline.traffic
path = f"{G}/{tag}_stem_raw.csv"
if os.path.exists(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    def gb(key):
        v, u = float(d[key]), units[hdr.index(key)].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}[u]
    traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    json.dump({"kernel": "stem_umma_mf_kernel", "capture": f"{tag}_stem (ncu --set full, tools/ncu_target.py, 160 clip-crops)",
               "dram_bytes_per_launch": traffic, "dram_read_bytes": gb("dram__bytes_read.sum"),
               "dram_write_bytes": gb("dram__bytes_write.sum"), "clip_crops_per_launch": 160},
              open("profiles/stem_dram_traffic.json", "w"), indent=1)
    lines += [f"Stem DRAM traffic per launch: {traffic / 1e9:.3f} GB for 160 clip-crops "
              f"(algorithmic: 1.064 GB stem-layout input read once + 1.028 GB pooled bf16 output written once).", ""]
open(out, "w").write("\n".join(lines) + "\n")
print(out)
