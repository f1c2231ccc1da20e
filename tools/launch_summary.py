"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total us, share.
    python tools/launch_summary.py gpurun_out/head_train_launches.csv [first_launch [last_launch]]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows, hdr = [], None
for r in csv.reader(open(path)):
    if r and r[0] == "ID":
        hdr = r
    elif hdr and r and r[0].isdigit():
        rows.append(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
agg = OrderedDict()
for r in rows[lo:hi]:
    k = r[ki].split("(")[0].replace("void ", "").replace("vad::", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{hi - lo} launches, {tot / 1e3:.3f} ms of kernel time")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us:.0f} | {100 * us / tot:.1f}% |")
