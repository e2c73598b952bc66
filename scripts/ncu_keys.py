"""Key metrics + stall / opcode histograms of the first kernel in an ncu report.  usage: python scripts/ncu_keys.py report.ncu-rep"""
import csv, subprocess, sys
from collections import Counter
import os
rep = sys.argv[1]
KSEL = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []   # NCU_KERNEL=regex selects the kernel
raw = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size', 'smsp__thread_inst_executed.sum']
for k in keys:
    if k in hdr: print(f"{k:70s} {vals[hdr.index(k)]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
data = []
for r in rows[hi + 1:]:            # first kernel of the report only
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(h):
        data.append(r)
ti = sum(float(r[ix['Instructions Executed']] or 0) for r in data); ts = sum(float(r[ix['# Samples']] or 0) for r in data)
c = Counter(); s = Counter()
for r in data:
    t = r[ix['Source']].split()
    if not t: continue
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += float(r[ix['Instructions Executed']] or 0); s[op] += float(r[ix['# Samples']] or 0)
print(f"SASS instructions {len(data)}, executed warp instructions {ti:.0f}")
print("  ".join(f"{op} {v / ti * 100:.1f}/{s[op] / ts * 100:.1f}" for op, v in c.most_common(18)), "(inst% / sample%)")
st = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = {k: sum(float(r[ix[k]] or 0) for r in data) for k in st}
print("  ".join(f"{k[6:]} {v / ts * 100:.1f}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
