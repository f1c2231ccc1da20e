"""Top warp-stall lines of an `ncu --page source --csv` export:  python tools/ncu_top_stalls.py <source.csv> [n]"""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        ks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
seen = set()
for k in ks:
    key = (k["name"], len(k["rows"]))
    if key in seen or "hdr" not in k:
        continue
    seen.add(key)
    hdr = k["hdr"]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ei = hdr.index("Instructions Executed")
    tot = sum(int(r[si] or 0) for r in k["rows"]) or 1
    print(k["name"][:100], "samples", tot)
    top = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][si] or 0))[:n]
    for j in sorted(top):
        r = k["rows"][j]
        print(f"{j:5d} {r[1][:80].strip():80s} {int(r[si] or 0):6d} {100 * int(r[si] or 0) / tot:5.1f}%  exec {r[ei]}")
