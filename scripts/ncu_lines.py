"""Per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
import os
rep = sys.argv[1]
KSEL = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []   # NCU_KERNEL=regex selects the kernel
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
lines = [r for r in rows[hi + 1:] if len(r) > ie and r[2] == "-"]      # source-line summary rows
ti = sum(float(r[ie] or 0) for r in lines); ts = sum(float(r[isamp] or 0) for r in lines)
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for r in lines:
    a, b = float(r[ie] or 0) / ti * 100, float(r[isamp] or 0) / ts * 100
    if a >= thr or b >= thr: print(f"{r[0]:>5} {a:5.1f}% inst {b:5.1f}% smp  {r[1].strip()[:120]}")
