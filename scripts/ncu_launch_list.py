"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total us, us/launch, share.
usage: python scripts/ncu_launch_list.py launches.csv [out.md] [title] [name-prefix]
name-prefix (e.g. k_): only kernels whose name starts with it -- the library's own (the synthetic inputs are made by torch kernels)"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0].replace("void ", "").replace("is::", "")
    if len(sys.argv) > 4 and not name.startswith(sys.argv[4]): continue
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[mu], 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
lines = []
if len(sys.argv) > 3: lines += [sys.argv[3], ""]
lines += ["| kernel | launches | total us | us/launch | share |", "|---|---|---|---|---|"]
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {name} | {n} | {us:.1f} | {us / n:.1f} | {us / tot:.3f} |")
text = "\n".join(lines)
print(text)
if len(sys.argv) > 2: open(sys.argv[2], "w").write(text + "\n")
