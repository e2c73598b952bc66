"""Summarises an .ncu-rep (raw page, csv) into a small table: per kernel launch duration, DRAM bytes, DRAM
throughput %, achieved occupancy, registers, top stall hints.  Usage: ncu_summary.py report.ncu-rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
lines = ["| " + " | ".join(n for _, n in want) + " |", "|" + "---|" * len(want)]
for r in data:
    vals = []
    for key, _ in want:
        v = g(r, key)
        u = units[col[key]] if key in col else ""
        if key == "Kernel Name":
            v = v.split("(")[0][:48]
        elif v:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f} {u}" if u not in ("", "%") else f"{f:.1f}"
            except ValueError:
                pass
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
text = "\n".join(lines)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
